#!/bin/bash
# diagnosis: where does the fixed ~50 ms per wavefront frame go when several ranks share the frame?
mkdir -p gpurun_out
RTW_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline --sweep 1 > gpurun_out/an_bench_n2.json 2> gpurun_out/an_bench_n2.err
grep -E "wavefront\]|device build" gpurun_out/an_bench_n2.err | tail -40
