#!/bin/bash
mkdir -p gpurun_out
for P in 262144 1048576 2097152 4194304; do
 for B in 2 8; do
  echo "== RTW_WF_POOL=$P RTW_BVH=$B" | tee -a gpurun_out/o_wfpool.log
  RTW_WF_POOL=$P RTW_KERNEL=wavefront RTW_BVH=$B RTW_DEVICE_BUILD=0 timeout 900 python tools/sweep.py 1 --spp 16 2>&1 | cut -c1-215 | tee -a gpurun_out/o_wfpool.log
 done
done
