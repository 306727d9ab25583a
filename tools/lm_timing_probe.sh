#!/bin/bash
# builds with the in-kernel LM clocks (WC_LM_TIMING) and prints the lm_step start / end globaltimer stamps of one C3 solve:
# consecutive differences = step duration, linearisation + the two kernel boundaries, step duration, ...
export WC_NVCC_EXTRA=-DWC_LM_TIMING
python -c "from wildcat_slam_b200 import build; build.build(force=True)" > /dev/null 2>&1
python tools/profile_pass.py C3 3 2>&1 | grep -a -E "stamps|cycles" | tail -3 | cut -c1-1500
unset WC_NVCC_EXTRA
python -c "from wildcat_slam_b200 import build; build.build(force=True)" > /dev/null 2>&1
