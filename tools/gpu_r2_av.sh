#!/bin/bash
# ncu --set full of the kernels of cornell_box (small-scene variant) and final_scene on the final commit
mkdir -p gpurun_out
for sc in cornell_box final_scene; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -f -o gpurun_out/prof_v5_$sc python tools/profile_one.py $sc 50 > gpurun_out/v5_ncu_$sc.log 2>&1; tail -1 gpurun_out/v5_ncu_$sc.log
done
ls -la gpurun_out/prof_v5_*
