#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/ag_logic.log; : > $L
V=$PWD/rust-ray-tracing-in-a-weekend_b200/variants
for cfg in "X=1" "RTW_LIB_PATH=$V/wfl4.so" "X=2" "RTW_LIB_PATH=$V/wfl4.so"; do echo "== $cfg" | tee -a $L; env $cfg timeout 600 python tools/sweep.py 1 4 --spp 32 2>&1 | cut -c1-200 | tee -a $L; done
