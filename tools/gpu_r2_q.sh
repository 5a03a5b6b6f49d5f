#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "wavefront" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_q.log
export RTW_KERNEL=wavefront RTW_DEVICE_BUILD=0 RTW_BVH=2
for T in lockstep whilewhile; do
  echo "== RTW_WF_TRACE=$T" | tee -a gpurun_out/q_lockstep.log
  RTW_WF_TRACE=$T timeout 900 python tools/sweep.py 1 4 --spp 32 2>&1 | cut -c1-200 | tee -a gpurun_out/q_lockstep.log
done
RTW_WF_TRACE=lockstep timeout 600 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 60 --csv --log-file gpurun_out/wf2_launches.csv python tools/profile_sweep.py 1 8 > gpurun_out/ncu_wf2_list.log 2>&1
