#!/bin/bash
# Round-2 session B: the 8-wide compressed BVH.  Parity tests forced onto it, then timings binary vs wide on every config.
mkdir -p gpurun_out; rm -f gpurun_out/parity_measured.jsonl
RTW_BVH=8 timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 | tee gpurun_out/pytest_gpu_wide.log
cp gpurun_out/parity_measured.jsonl gpurun_out/parity_measured_wide.jsonl
for B in 2 8; do
  RTW_BVH=$B RTW_TAG=bvh$B timeout 600 python tools/exp_time2.py 2>&1 | tee -a gpurun_out/b_configs.log
  RTW_BVH=$B timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -2 | sed "s/^/[bvh$B] /" | tee -a gpurun_out/b_configs.log
  RTW_BVH=$B timeout 900 python tools/sweep.py 1 4 --spp 64 2>&1 | sed "s/^/[bvh$B] /" | tee -a gpurun_out/b_configs.log
done
