#!/bin/bash
# final_scene variant smaller than the instruction caches' reach: rolled perlin_noise / no tile list in the big media variants
mkdir -p gpurun_out; L=gpurun_out/ax_final_scene_code_size.log; : > $L
V=$PWD/rust-ray-tracing-in-a-weekend_b200/variants
for lib in "" roll notile both "" both; do
  echo "== ${lib:-head}" | tee -a $L
  if [ -z "$lib" ]; then RTW_ONLY=final_scene,two_perlin_spheres,simple_light RTW_TAG=ax timeout 300 python tools/exp_time2.py 2>&1 | tee -a $L
  else RTW_LIB_PATH=$V/$lib.so RTW_ONLY=final_scene,two_perlin_spheres,simple_light RTW_TAG=ax timeout 300 python tools/exp_time2.py 2>&1 | tee -a $L; fi
done
