#!/bin/bash
# Round-2 session N: the global-memory wavefront pipeline — equality with the megakernel, then the sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "wavefront" 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_n.log
for B in 2 8; do
  echo "== RTW_KERNEL=wavefront RTW_BVH=$B (host SAH build)" | tee -a gpurun_out/n_wavefront.log
  RTW_KERNEL=wavefront RTW_BVH=$B RTW_DEVICE_BUILD=0 timeout 900 python tools/sweep.py 1 4 --spp 32 2>&1 | cut -c1-215 | tee -a gpurun_out/n_wavefront.log
done
echo "== C1 wavefront" | tee -a gpurun_out/n_wavefront.log
RTW_KERNEL=wavefront timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -2 | tee -a gpurun_out/n_wavefront.log
