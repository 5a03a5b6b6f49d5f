#!/bin/bash
# last session of round 2 (1 GPU): GPU suite, smoke, bench both arms on the final commit (ncu captures: r2_v4_*)
mkdir -p gpurun_out; rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/v5_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/v5_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/v5_bench_n1.json 2> gpurun_out/v5_bench_n1.err; tail -c 300 gpurun_out/v5_bench_n1.err; cut -c1-300 gpurun_out/v5_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/v5_bench_ref.json 2>> gpurun_out/v5_bench_n1.err; cut -c1-200 gpurun_out/v5_bench_ref.json
