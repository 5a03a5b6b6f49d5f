#!/bin/bash
mkdir -p gpurun_out
V=rust-ray-tracing-in-a-weekend_b200/variants
for L in rust-ray-tracing-in-a-weekend_b200/librtw.so $V/occ6.so; do
  T=$(basename $L .so)
  RTW_LIB_PATH=$PWD/$L timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -2 | sed "s/^/[$T] /" | tee -a gpurun_out/l_occ6.log
  RTW_LIB_PATH=$PWD/$L RTW_TAG=$T timeout 600 python tools/exp_time2.py 2>&1 | tee -a gpurun_out/l_occ6.log
done
