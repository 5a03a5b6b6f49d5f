#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/at_box_face.log; : > $L
V=$PWD/rust-ray-tracing-in-a-weekend_b200/variants
for cfg in "X=1" "RTW_LIB_PATH=$V/head.so" "X=2" "RTW_LIB_PATH=$V/head.so"; do echo "== $cfg" | tee -a $L; env $cfg RTW_TAG=at timeout 600 python tools/exp_time2.py 2>&1 | grep -E "final_scene|cornell_box " | tee -a $L; done
timeout 900 python -m pytest tests -m gpu -q -x -k "box or hittable or paths or cornell" 2>&1 | tail -2 | tee -a $L
