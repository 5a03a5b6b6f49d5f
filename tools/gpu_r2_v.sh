#!/bin/bash
# round 2, run v: wavefront + wide tree variants on the 1 M sweep (fill collapse, K pools, thresholds, occupancy), device build, mid-size scenes, ncu at steady state
mkdir -p gpurun_out; L=gpurun_out/v_wf.log; : > $L
RTW_BVH=8 timeout 900 python -m pytest tests -m gpu -q -x -k "wavefront or wide or device_bvh" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_v.log
export RTW_KERNEL=wavefront RTW_BVH=8 RTW_DEVICE_BUILD=0
V=$PWD/rust-ray-tracing-in-a-weekend_b200/variants
run() { echo "== $1" | tee -a $L; shift; env "$@" timeout 600 python tools/sweep.py 1 --spp 32 2>&1 | cut -c1-220 | tee -a $L; }
run "base fill=1 K=1" RTW_TIMING=1
run "fill=0" RTW_WIDE_FILL=0
run "K=2" RTW_WF_POOLS=2
run "K=3" RTW_WF_POOLS=3
run "K=2 pool16M" RTW_WF_POOLS=2 RTW_WF_POOL=16777216
run "pool 4M" RTW_WF_POOL=4194304
for v in wf_lb8 wf_t4 wf_l4r8 wf_l12r8 wf_l8r16; do run "variant $v" RTW_LIB_PATH=$V/$v.so; done
echo "== device build wide 1M 4M" | tee -a $L
RTW_DEVICE_BUILD=1 timeout 600 python tools/sweep.py 1 4 --spp 32 2>&1 | cut -c1-220 | tee -a $L
echo "== host build wide 4M" | tee -a $L
timeout 600 python tools/sweep.py 4 --spp 32 2>&1 | cut -c1-220 | tee -a $L
echo "== mid-size scenes, wavefront bvh8 / bvh2" | tee -a $L
RTW_TAG=wf8 timeout 600 python tools/exp_time2.py 2>&1 | head -3 | tee -a $L
RTW_BVH=2 RTW_TAG=wf2 timeout 600 python tools/exp_time2.py 2>&1 | head -3 | tee -a $L
# ncu: steady-state trace launch (pool full) of the 1 M sweep
timeout 900 ncu --set full --import-source on --clock-control none -k regex:wf_trace2w --launch-skip 14 --launch-count 1 -o gpurun_out/prof_wf_trace2w_steady -f python tools/profile_sweep.py 1 32 > gpurun_out/ncu_v.log 2>&1
tail -3 gpurun_out/ncu_v.log
