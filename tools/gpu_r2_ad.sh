#!/bin/bash
# round 2, run ad: carry-over enabled per variant (default build): timings, unit-size retune, GPU suite
mkdir -p gpurun_out; L=gpurun_out/ad_carry_units.log; : > $L; rm -f gpurun_out/parity_measured.jsonl
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee -a $L
RTW_TAG=ad timeout 600 python tools/exp_time2.py 2>&1 | tee -a $L
for cfg in "X=1" "RTW_UNITS_PER_WARP=32" "RTW_UNITS_PER_WARP=48" "RTW_UNITS_PER_WARP=64" "RTW_UNITS_PER_WARP=96" "RTW_UNITS_PER_WARP=48 RTW_B_SHARE=30" "RTW_UNITS_PER_WARP=48 RTW_ONE_PHASE=1"; do
  echo "== N=1 $cfg" | tee -a $L; env $cfg timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -2 | tee -a $L
done
for cfg in "X=1" "RTW_B_SHARE=30" "RTW_B_SHARE=40" "RTW_A_MIN=16" "RTW_A_MIN=16 RTW_B_SHARE=30" "RTW_B_SPP=4" "RTW_A_MIN=24 RTW_B_SHARE=30"; do
  echo "== emulate 8: $cfg" | tee -a $L; env $cfg RTW_EMULATE_RANKS=8 timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -3 | tee -a $L
done
