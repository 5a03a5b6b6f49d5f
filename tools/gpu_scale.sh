#!/bin/bash
# 1/2/4/8-GPU scaling of the headline bench, one rank per GPU under torchrun (the driver's own launch line).
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | head -8
for n in 1 2 4 8; do
  if [ $n -gt $N ]; then break; fi
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1])
    print("N=$n", round(d["value"], 1), "Mpaths/s", round(d["ms_per_step"], 2), "ms/step  e2e", round(d["e2e"]["value"], 1), "kernel ms", round(d["kernel_ms_per_step"], 2), d["clocks"])
except Exception as e:
    print("N=$n FAILED", e); print(open("gpurun_out/scale_n$n.err").read()[-1500:])
PY
done
