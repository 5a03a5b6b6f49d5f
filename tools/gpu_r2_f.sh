#!/bin/bash
# Round-2 session F: 8-GPU unit sizing emulated on one GPU (RTW_EMULATE_RANKS=8: every 8th unit, unit sizes of the 8-GPU run).
mkdir -p gpurun_out
run() { echo "== $1" | tee -a gpurun_out/f_units.log; env $1 python tools/profile_one.py random_scene 500 2>&1 | tail -2 | tee -a gpurun_out/f_units.log; }
run "RTW_NOP=1"
run "RTW_EMULATE_RANKS=8"
run "RTW_EMULATE_RANKS=8 RTW_B_SPP=8"
run "RTW_EMULATE_RANKS=8 RTW_B_SPP=8 RTW_B_SHARE=10"
run "RTW_EMULATE_RANKS=8 RTW_B_SPP=16 RTW_B_SHARE=10"
run "RTW_EMULATE_RANKS=8 RTW_A_MIN=16 RTW_B_SPP=8"
run "RTW_EMULATE_RANKS=8 RTW_UNITS_PER_WARP=16"
run "RTW_EMULATE_RANKS=8 RTW_UNITS_PER_WARP=16 RTW_B_SPP=8"
run "RTW_EMULATE_RANKS=8 RTW_UNITS_PER_WARP=12 RTW_B_SPP=8 RTW_B_SHARE=25"
run "RTW_EMULATE_RANKS=8 RTW_ONE_PHASE=1"
run "RTW_EMULATE_RANKS=4"
run "RTW_EMULATE_RANKS=4 RTW_B_SPP=16"
python tools/profile_one.py random_scene 50 > gpurun_out/plain_c1.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -f -o gpurun_out/prof_c1_r2 python tools/profile_one.py random_scene 50 > gpurun_out/ncu_c1.log 2>&1
tail -2 gpurun_out/ncu_c1.log
