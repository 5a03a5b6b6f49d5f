#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/au_staged_upload.log; : > $L
RTW_TIMING=1 timeout 300 python tools/exp_commit.py 16 4 2>&1 | grep -E "commit\]|device build\] (upload|1|4)" | tee -a $L
echo "== RTW_SCRATCH_POOL=0" | tee -a $L
RTW_SCRATCH_POOL=0 timeout 300 python tools/exp_commit.py 16 2>&1 | grep -E "commit\]" | tee -a $L
timeout 600 python -m pytest tests -m gpu -q -x -k "staged_upload or device_bvh_build or sweep or wavefront" 2>&1 | tail -3 | tee -a $L
