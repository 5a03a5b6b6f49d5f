#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/aj_c1_list.log; : > $L
for cfg in "RTW_LIST_MAX=0" "RTW_LIST_MAX=8" "RTW_LIST_MAX=0" "RTW_LIST_MAX=8"; do echo "== $cfg" | tee -a $L; env $cfg timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -3 | tee -a $L; done
