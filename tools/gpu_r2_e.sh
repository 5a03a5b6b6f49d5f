#!/bin/bash
# Round-2 session E (2 GPUs): in-process 2-GPU pixel test, torchrun bench at N=2 with the multi-GPU check and every config.
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -q -x -k "two_gpus or device_bvh_build_equals_host_build" 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_e.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "rc=$?"; tail -5 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/bench_n1_e.json 2> gpurun_out/bench_n1_e.err; cat gpurun_out/bench_n1_e.json | cut -c1-600
