#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/al_small_notile.log; : > $L
V=$PWD/rust-ray-tracing-in-a-weekend_b200/variants
for cfg in "X=1" "RTW_LIB_PATH=$V/smallnotile.so" "X=2" "RTW_LIB_PATH=$V/smallnotile.so"; do echo "== $cfg" | tee -a $L; env $cfg RTW_TAG=al timeout 600 python tools/exp_time2.py 2>&1 | grep -v final_scene | tee -a $L; done
