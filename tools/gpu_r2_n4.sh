#!/bin/bash
# round 2: the driver's N = 8 line (one rank per GPU under torchrun, every config) on the final commit
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29619 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/n4_bench.json 2> gpurun_out/n4_bench.err
echo "rc $?"; tail -c 800 gpurun_out/n4_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/n4_bench.json").read().strip().splitlines()[-1])
print("N=4", round(d["value"], 1), "Mpaths/s", round(d["ms_per_step"], 2), "ms/step kernel", round(d["kernel_ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), d.get("multi_gpu_check"), d["clocks"])
for k, v in d.get("configs", {}).items(): print(k, round(v["mpaths_per_s"], 1), "Mpaths/s", round(v["ms"], 1), "ms", "commit", v.get("commit_s"))
PY
