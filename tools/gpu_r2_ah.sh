#!/bin/bash
# round 2, run ah: candidate scan instead of the BVH walk for scenes of a handful of primitives (RTW_LIST_MAX)
mkdir -p gpurun_out; L=gpurun_out/ah_list_scan.log; : > $L
for cfg in "RTW_LIST_MAX=0" "RTW_LIST_MAX=8" "RTW_LIST_MAX=3" "RTW_LIST_MAX=0" "RTW_LIST_MAX=8"; do
  echo "== $cfg" | tee -a $L; env $cfg RTW_TAG=ah timeout 600 python tools/exp_time2.py 2>&1 | tee -a $L
done
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee -a $L
