#!/bin/bash
mkdir -p gpurun_out
export RTW_KERNEL=wavefront RTW_DEVICE_BUILD=0
for B in 2 8; do
RTW_BVH=$B python tools/profile_sweep.py 1 8 > gpurun_out/plain_wf$B.log 2>&1; tail -1 gpurun_out/plain_wf$B.log
RTW_BVH=$B timeout 600 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 90 --csv --log-file gpurun_out/wf_launches_bvh$B.csv python tools/profile_sweep.py 1 8 > gpurun_out/ncu_wf_list$B.log 2>&1
RTW_BVH=$B timeout 900 ncu --set full --clock-control none --import-source on -k regex:wf_trace -s 6 -c 1 -f -o gpurun_out/prof_wf_trace_bvh$B python tools/profile_sweep.py 1 8 > gpurun_out/ncu_wf_trace$B.log 2>&1
done
ls -la gpurun_out/prof_wf*
