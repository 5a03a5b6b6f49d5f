#!/bin/bash
# round 2, run w: box-as-one-leaf (PRIM_BOX) parity + timing A/B, default policy for big scenes (device build + wide + wavefront)
mkdir -p gpurun_out; L=gpurun_out/w_box.log; : > $L; rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_w.log
RTW_TAG=box1 timeout 600 python tools/exp_time2.py 2>&1 | tee -a $L
RTW_BOX_PRIM=0 RTW_TAG=box0 timeout 600 python tools/exp_time2.py 2>&1 | tee -a $L
timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -2 | tee -a $L
echo "== sweep, default policy" | tee -a $L
RTW_TIMING=1 timeout 900 python tools/sweep.py 1 4 16 --spp 32 2>&1 | grep -v "^\[flatten\]\|^\[build\]" | cut -c1-230 | tee -a $L
echo "== sweep 0.25 0.5 (below the threshold: megakernel binary; then forced wavefront wide)" | tee -a $L
timeout 600 python tools/sweep.py 0.25 0.5 --spp 32 2>&1 | cut -c1-230 | tee -a $L
RTW_BIG_MIN=1000 timeout 600 python tools/sweep.py 0.25 0.5 --spp 32 2>&1 | cut -c1-230 | tee -a $L
