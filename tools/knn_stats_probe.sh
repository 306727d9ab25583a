#!/bin/bash
# builds the library with the k-NN search counters (WC_KNN_STATS) and runs one C3 pass
export WC_NVCC_EXTRA=-DWC_KNN_STATS
python -c "from wildcat_slam_b200 import build; build.build(force=True)" > /dev/null 2>&1
python tools/profile_pass.py C3 1 2>&1 | grep -E "knn stats|grid cells" | head -6
unset WC_NVCC_EXTRA
python -c "from wildcat_slam_b200 import build; build.build(force=True)" > /dev/null 2>&1
