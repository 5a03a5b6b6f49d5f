#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/as_medium_direct.log; : > $L
V=$PWD/rust-ray-tracing-in-a-weekend_b200/variants
for cfg in "X=1" "RTW_LIB_PATH=$V/medold.so" "X=2" "RTW_LIB_PATH=$V/medold.so"; do echo "== $cfg" | tee -a $L; env $cfg RTW_TAG=as timeout 600 python tools/exp_time2.py 2>&1 | grep -E "final_scene|smoke" | tee -a $L; done
timeout 900 python -m pytest tests -m gpu -q -x -k "medium or final_scene or smoke or paths or psnr" 2>&1 | tail -2 | tee -a $L
