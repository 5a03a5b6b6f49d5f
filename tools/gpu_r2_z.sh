#!/bin/bash
# round 2, run z: A/B of the slab slack form in the megakernel (per axis vs one shared term)
mkdir -p gpurun_out; L=gpurun_out/z_slab.log; : > $L
V=$PWD/rust-ray-tracing-in-a-weekend_b200/variants
for rep in 1 2; do
for cfg in "X=1" "RTW_LIB_PATH=$V/slabshared.so"; do
  echo "== $cfg" | tee -a $L
  env $cfg timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -2 | tee -a $L
  env $cfg RTW_TAG=z timeout 600 python tools/exp_time2.py 2>&1 | tee -a $L
done; done
RTW_LIB_PATH=$V/slabshared.so timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee -a $L
