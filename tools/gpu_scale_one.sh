#!/bin/bash
# one point of the scaling curve (the driver's launch line, headline only): gpu_scale_one.sh N
N=$1; mkdir -p gpurun_out
if [ $N -eq 1 ]; then python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err; fi
python - <<PY
import json
d = json.loads(open("gpurun_out/scale_n$N.json").read().strip().splitlines()[-1])
print("N=$N", round(d["value"], 1), "Mpaths/s", round(d["ms_per_step"], 2), "ms/step  e2e", round(d["e2e"]["value"], 1), "kernel ms", round(d["kernel_ms_per_step"], 2), d.get("multi_gpu_check", {}).get("max_abs_diff") if d.get("multi_gpu_check") else None)
PY
