#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_u.log
timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -2 | tee gpurun_out/u_slack.log
RTW_TAG=peraxis timeout 600 python tools/exp_time2.py 2>&1 | tee -a gpurun_out/u_slack.log
export RTW_DEVICE_BUILD=0
for cfg in "mega 2" "wavefront 8" "wavefront 2"; do
  set -- $cfg
  echo "== RTW_KERNEL=$1 RTW_BVH=$2" | tee -a gpurun_out/u_slack.log
  RTW_KERNEL=$1 RTW_BVH=$2 timeout 900 python tools/sweep.py 1 4 --spp 32 2>&1 | cut -c1-200 | tee -a gpurun_out/u_slack.log
done
RTW_WF_TIMELINE=1 RTW_TIMING=1 RTW_KERNEL=wavefront RTW_BVH=8 RTW_LIB_PATH=$PWD/rust-ray-tracing-in-a-weekend_b200/variants/instr.so python tools/profile_sweep.py 1 8 2>&1 | grep -E "wf timeline|wavefront\]|^10485" | head -6 | tee -a gpurun_out/u_slack.log
