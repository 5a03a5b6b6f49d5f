#!/bin/bash
# A/B of two library builds on 2 GPUs under torchrun (one rank per GPU): tools/ab_n2.sh <libA> <libB>
mkdir -p gpurun_out
for L in "$@"; do
  for rep in 1 2; do
  RTW_LIB_PATH=$PWD/$L python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$L] N=2', round(d['ms_per_step'],2), 'ms kernel', round(d['kernel_ms_per_step'],2))"
  done
done 2>&1 | tee gpurun_out/ab_n2.log
