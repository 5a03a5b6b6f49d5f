#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1" | tee -a gpurun_out/k_units.log; env $1 python tools/profile_one.py random_scene 500 2>&1 | tail -3 | tee -a gpurun_out/k_units.log; }
run "RTW_EMULATE_RANKS=8"
run "RTW_EMULATE_RANKS=8 RTW_A_MIN=48"
run "RTW_EMULATE_RANKS=8 RTW_A_MIN=64"
run "RTW_EMULATE_RANKS=8 RTW_A_MIN=64 RTW_B_SPP=16"
run "RTW_EMULATE_RANKS=8 RTW_A_MIN=64 RTW_B_SHARE=30"
run "RTW_EMULATE_RANKS=8 RTW_A_MIN=100"
run "RTW_EMULATE_RANKS=8 RTW_A_MIN=100 RTW_B_SHARE=30"
run "RTW_EMULATE_RANKS=8 RTW_A_MIN=134 RTW_B_SHARE=33"
run "RTW_EMULATE_RANKS=4 RTW_A_MIN=64"
run "RTW_EMULATE_RANKS=4 RTW_A_MIN=100"
run "RTW_A_MIN=134"
run "RTW_UNITS_PER_WARP=12"
