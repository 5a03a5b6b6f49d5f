#!/bin/bash
# Round-2 final single-GPU session: full parity suite, smoke, bench (both arms), ncu launch list + full captures (C1, final_scene).
mkdir -p gpurun_out; rm -f gpurun_out/parity_measured.jsonl
nvidia-smi -L; nproc
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_i.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke_i.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; tail -3 gpurun_out/bench_i.err; cut -c1-400 gpurun_out/bench_i.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_i.json 2>> gpurun_out/bench_i.err; cut -c1-300 gpurun_out/bench_ref_i.json
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/plain_i.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_i.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/ncu_list_i.log 2>&1
python tools/profile_one.py random_scene 50 > gpurun_out/plain_c1_i.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -f -o gpurun_out/prof_c1_final python tools/profile_one.py random_scene 50 > gpurun_out/ncu_c1_i.log 2>&1
python tools/profile_one.py final_scene 100 > gpurun_out/plain_final_i.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -f -o gpurun_out/prof_final_final python tools/profile_one.py final_scene 100 > gpurun_out/ncu_final_i.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
