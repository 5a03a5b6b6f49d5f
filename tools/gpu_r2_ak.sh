#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/ak_c1_scan_ab.log; : > $L
V=$PWD/rust-ray-tracing-in-a-weekend_b200/variants
for cfg in "X=1" "RTW_LIB_PATH=$V/noscan.so" "X=2" "RTW_LIB_PATH=$V/noscan.so"; do echo "== $cfg" | tee -a $L; env $cfg timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -3 | tee -a $L; done
