#!/bin/bash
# round 2, run ac: carry-over of the last rays of a work unit into the next one (RTW_CARRY) — A/B and the GPU suite on the variant
mkdir -p gpurun_out; L=gpurun_out/ac_carry.log; : > $L
V=$PWD/rust-ray-tracing-in-a-weekend_b200/variants
for cfg in "X=1" "RTW_LIB_PATH=$V/carry.so"; do
  echo "== $cfg" | tee -a $L
  env $cfg timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -2 | tee -a $L
  env $cfg RTW_EMULATE_RANKS=8 timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -3 | tee -a $L
  env $cfg RTW_UNITS_PER_WARP=192 timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -1 | tee -a $L
  env $cfg RTW_TAG=ac timeout 600 python tools/exp_time2.py 2>&1 | tee -a $L
done
RTW_LIB_PATH=$V/carry.so timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee -a $L
