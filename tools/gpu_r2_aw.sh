#!/bin/bash
# last commit of round 2 on a 2-GPU box: the whole GPU suite (the two-GPU tests run here), torchrun bench with every config
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/v5_pytest_gpu_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/v5_bench_n2.json 2> gpurun_out/v5_bench_n2.err
tail -c 400 gpurun_out/v5_bench_n2.err; cut -c1-260 gpurun_out/v5_bench_n2.json
