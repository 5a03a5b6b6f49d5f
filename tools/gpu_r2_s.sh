#!/bin/bash
mkdir -p gpurun_out
export RTW_KERNEL=wavefront RTW_DEVICE_BUILD=0 RTW_BVH=8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wf_trace2w -s 2 -c 1 -f -o gpurun_out/prof_wf_trace2w python tools/profile_sweep.py 1 8 > gpurun_out/ncu_wf_trace2w.log 2>&1
ls -la gpurun_out/prof_wf_trace2w*
