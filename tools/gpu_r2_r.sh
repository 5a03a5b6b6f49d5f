#!/bin/bash
mkdir -p gpurun_out
RTW_BVH=8 timeout 900 python -m pytest tests -m gpu -q -x -k "wavefront" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r.log
export RTW_KERNEL=wavefront RTW_DEVICE_BUILD=0
for B in 8 2; do
  echo "== lockstep RTW_BVH=$B" | tee -a gpurun_out/r_lockstep.log
  RTW_BVH=$B timeout 900 python tools/sweep.py 1 4 --spp 32 2>&1 | cut -c1-200 | tee -a gpurun_out/r_lockstep.log
done
RTW_BVH=8 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 45 --csv --log-file gpurun_out/wf2w_launches.csv python tools/profile_sweep.py 1 8 > gpurun_out/ncu_wf2w_list.log 2>&1
