#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/ai_check.log; : > $L
for i in 1 2; do RTW_TAG=ai timeout 600 python tools/exp_time2.py 2>&1 | tee -a $L; done
timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -2 | tee -a $L
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee -a $L
