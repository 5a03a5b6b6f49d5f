#!/bin/bash
# Round-2 session G: prefetch of pushed / postponed nodes on scenes beyond the caches (A/B), C1 unchanged?
mkdir -p gpurun_out
V=rust-ray-tracing-in-a-weekend_b200/variants
for L in rust-ray-tracing-in-a-weekend_b200/librtw.so $V/nopf.so; do
  T=$(basename $L .so)
  RTW_LIB_PATH=$PWD/$L timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -2 | sed "s/^/[$T] /" | tee -a gpurun_out/g_pf.log
  RTW_LIB_PATH=$PWD/$L timeout 900 python tools/sweep.py 1 4 16 --spp 32 2>&1 | sed "s/^/[$T devbuild] /" | tee -a gpurun_out/g_pf.log
  RTW_LIB_PATH=$PWD/$L RTW_DEVICE_BUILD=0 timeout 900 python tools/sweep.py 1 4 --spp 32 2>&1 | sed "s/^/[$T hostbuild] /" | tee -a gpurun_out/g_pf.log
done
