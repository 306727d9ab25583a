#!/bin/bash
# builds the library with the in-kernel LM phase clocks (WC_LM_TIMING) into a scratch copy and runs a reduced C5 solve
set -e
export WC_NVCC_EXTRA=-DWC_LM_TIMING
python -c "from wildcat_slam_b200 import build; build.build(force=True)" > /dev/null 2>&1
python tools/c5_precision_sweep.py ${1:-200000} 2>&1 | grep -E "cycles|f64" | head -6
unset WC_NVCC_EXTRA
python -c "from wildcat_slam_b200 import build; build.build(force=True)" > /dev/null 2>&1
