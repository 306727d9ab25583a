#!/bin/bash
# builds the library with the k-NN search counters (WC_KNN_STATS) and runs one C3 pass
export WC_NVCC_EXTRA=-DWC_KNN_STATS
python -c "from wildcat_slam_b200 import build; build.build(force=True)" > /dev/null 2>&1
python tools/profile_pass.py C3 1 > gpurun_out/knn_stats.log 2>&1; grep -a -E "knn|grid|pass" gpurun_out/knn_stats.log | head -12
unset WC_NVCC_EXTRA
python -c "from wildcat_slam_b200 import build; build.build(force=True)" > /dev/null 2>&1
