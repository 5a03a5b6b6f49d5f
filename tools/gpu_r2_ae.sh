#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/ae_final_check.log; : > $L; rm -f gpurun_out/parity_measured.jsonl
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee -a $L
RTW_TAG=ae timeout 600 python tools/exp_time2.py 2>&1 | tee -a $L
timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -2 | tee -a $L
RTW_EMULATE_RANKS=8 timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -3 | tee -a $L
RTW_EMULATE_RANKS=4 timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -3 | tee -a $L
RTW_EMULATE_RANKS=2 timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -3 | tee -a $L
