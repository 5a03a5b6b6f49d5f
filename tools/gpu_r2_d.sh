#!/bin/bash
# Round-2 session D: stream layout v2 + warp-cooperative unit-ball sampling (parity, A/B), FMA-pipe slab test, occupancy on the sweep.
mkdir -p gpurun_out; rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_d.log
V=rust-ray-tracing-in-a-weekend_b200/variants
for L in rust-ray-tracing-in-a-weekend_b200/librtw.so $V/nocoop.so $V/slabfma.so $V/occ8.so; do
  T=$(basename $L .so)
  RTW_LIB_PATH=$PWD/$L timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -3 | sed "s/^/[$T] /" | tee -a gpurun_out/d_ab.log
  RTW_LIB_PATH=$PWD/$L RTW_TAG=$T timeout 600 python tools/exp_time2.py 2>&1 | tee -a gpurun_out/d_ab.log
done
for L in rust-ray-tracing-in-a-weekend_b200/librtw.so $V/occ8.so $V/occ9.so; do
  T=$(basename $L .so)
  RTW_LIB_PATH=$PWD/$L RTW_BVH=2 RTW_DEVICE_BUILD=0 timeout 900 python tools/sweep.py 1 4 --spp 32 2>&1 | sed "s/^/[$T] /" | tee -a gpurun_out/d_ab.log
done
RTW_LIB_PATH=$PWD/$V/instr.so timeout 600 python tools/flop_model_device.py 2>&1 | tail -4 | tee gpurun_out/flop_model_device.log
cp profiles/flop_model.json gpurun_out/flop_model.json
RTW_LIB_PATH=$PWD/$V/instr.so timeout 600 python tools/diag_div.py 2>&1 | tail -4 | tee gpurun_out/diag_div.log
