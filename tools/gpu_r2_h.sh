#!/bin/bash
# Round-2 session H: the warp-pool kernel (dynamic ray fetch) over the 8-wide BVH on the sweep; equality with the megakernel.
mkdir -p gpurun_out
RTW_BVH=8 timeout 600 python -m pytest tests -m gpu -q -x -k "pool or work_units" 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_h.log
for K in mega pool64 pool128 pool256; do
 for B in 2 8; do
  echo "== RTW_KERNEL=$K RTW_BVH=$B (host SAH build)" | tee -a gpurun_out/h_pool.log
  RTW_KERNEL=$K RTW_BVH=$B RTW_DEVICE_BUILD=0 timeout 900 python tools/sweep.py 1 4 --spp 32 2>&1 | cut -c1-215 | tee -a gpurun_out/h_pool.log
 done
done
RTW_KERNEL=pool128 RTW_BVH=8 RTW_DEVICE_BUILD=0 python tools/profile_sweep.py 1 8 > gpurun_out/plain_sweep_pool.log 2>&1 &&
RTW_KERNEL=pool128 RTW_BVH=8 RTW_DEVICE_BUILD=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_pool -s 1 -c 1 -f -o gpurun_out/prof_sweep1m_pool128_bvh8 python tools/profile_sweep.py 1 8 > gpurun_out/ncu_sweep_pool.log 2>&1
tail -2 gpurun_out/ncu_sweep_pool.log
