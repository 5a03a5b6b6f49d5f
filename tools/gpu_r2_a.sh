#!/bin/bash
# Round-2 session A: measured FP32 peak, parity tests (with measured figures recorded), smoke, bench with every BASELINE config.
mkdir -p gpurun_out; rm -f gpurun_out/parity_measured.jsonl
nvidia-smi -L; nproc
timeout 120 tools/fp32_peak | tee gpurun_out/fp32_peak.json
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -5 gpurun_out/bench_r2a.err; cat gpurun_out/bench_r2a.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2a.json 2>> gpurun_out/bench_r2a.err; cat gpurun_out/bench_ref_r2a.json
ls -la gpurun_out | tail -15
