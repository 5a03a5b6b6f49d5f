#!/bin/bash
# round 2 FINAL (1 GPU): GPU suite, smoke, bench both arms, launch list, ncu of the C1 kernel
mkdir -p gpurun_out; rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_final.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke_final.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; tail -c 300 gpurun_out/final_bench_n1.err; cut -c1-300 gpurun_out/final_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2>> gpurun_out/final_bench_n1.err; cut -c1-200 gpurun_out/final_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sweep "" > gpurun_out/final_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -f -o gpurun_out/prof_final_c1 python tools/profile_one.py random_scene 50 > gpurun_out/final_ncu_c1.log 2>&1; tail -1 gpurun_out/final_ncu_c1.log
