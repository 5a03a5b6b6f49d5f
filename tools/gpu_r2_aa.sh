#!/bin/bash
# round 2, run aa: out-of-line variants (instruction-cache pressure of the big kernel variants)
mkdir -p gpurun_out; L=gpurun_out/aa_outline.log; : > $L
V=$PWD/rust-ray-tracing-in-a-weekend_b200/variants
for cfg in "X=1" "RTW_LIB_PATH=$V/ol1.so" "RTW_LIB_PATH=$V/ol3.so" "RTW_LIB_PATH=$V/ol7.so" "X=2"; do
  echo "== $cfg" | tee -a $L
  env $cfg RTW_TAG=aa timeout 600 python tools/exp_time2.py 2>&1 | tee -a $L
  env $cfg timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -1 | tee -a $L
done
