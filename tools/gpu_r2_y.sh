#!/bin/bash
# round 2, run y (1 GPU): full suite, smoke, bench both arms, launch list, unit-size probes at full frame size, ncu of C1 / cornell_box / final_scene kernels (box leaf)
mkdir -p gpurun_out; rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_y.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke_y.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/y_bench_n1.json 2> gpurun_out/y_bench_n1.err; tail -c 400 gpurun_out/y_bench_n1.err; cut -c1-600 gpurun_out/y_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/y_bench_ref.json 2>> gpurun_out/y_bench_n1.err; cut -c1-400 gpurun_out/y_bench_ref.json
L=gpurun_out/y_units.log; : > $L
for cfg in "X=1" "RTW_UNITS_PER_WARP=192" "RTW_UNITS_PER_WARP=96" "RTW_UNITS_PER_WARP=48" "RTW_UNITS_PER_WARP=192 RTW_ONE_PHASE=1" "RTW_FLAGS_NOCULL=1"; do
  echo "== $cfg" | tee -a $L; env $cfg timeout 300 python tools/profile_one.py random_scene 500 2>&1 | tail -2 | tee -a $L
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/y_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sweep 1 > gpurun_out/y_ncu_list.log 2>&1
for sc in random_scene cornell_box final_scene; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -f -o gpurun_out/prof_y_$sc python tools/profile_one.py $sc 50 > gpurun_out/y_ncu_$sc.log 2>&1; tail -1 gpurun_out/y_ncu_$sc.log
done
