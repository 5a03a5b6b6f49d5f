#!/bin/bash
# Round-2 session C: device-side BVH build (tests, commit times), binary vs wide on the sweep, ncu of both traversals.
mkdir -p gpurun_out; rm -f gpurun_out/parity_measured.jsonl
timeout 900 python -m pytest tests -m gpu -q -x -k "device_bvh_build or sweep" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_devbuild.log
for cfg in "2 0" "2 1" "8 1"; do
  set -- $cfg
  echo "== RTW_BVH=$1 RTW_DEVICE_BUILD=$2" | tee -a gpurun_out/c_sweep.log
  RTW_TIMING=1 RTW_BVH=$1 RTW_DEVICE_BUILD=$2 timeout 900 python tools/sweep.py 1 4 16 --spp 32 2>&1 | grep -v "^\[flatten\]\|^\[build\]" | tee -a gpurun_out/c_sweep.log
done
echo "== pool kernel, binary, host build" | tee -a gpurun_out/c_sweep.log
RTW_KERNEL=pool128 RTW_BVH=2 RTW_DEVICE_BUILD=0 timeout 600 python tools/sweep.py 1 --spp 32 2>&1 | tee -a gpurun_out/c_sweep.log
for B in 2 8; do
  RTW_BVH=$B RTW_DEVICE_BUILD=0 python tools/profile_sweep.py 1 8 > gpurun_out/plain_sweep_bvh$B.log 2>&1 &&
  RTW_BVH=$B RTW_DEVICE_BUILD=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/prof_sweep1m_bvh$B python tools/profile_sweep.py 1 8 > gpurun_out/ncu_sweep_bvh$B.log 2>&1
  tail -2 gpurun_out/ncu_sweep_bvh$B.log
done
RTW_BVH=8 python tools/profile_one.py random_scene 50 > gpurun_out/plain_c1_bvh8.log 2>&1 &&
RTW_BVH=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -f -o gpurun_out/prof_c1_bvh8 python tools/profile_one.py random_scene 50 > gpurun_out/ncu_c1_bvh8.log 2>&1
ls -la gpurun_out/*.ncu-rep
