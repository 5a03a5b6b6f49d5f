#!/bin/bash
# round 2, run x (2 GPUs): full GPU suite incl. the two-GPU tests, box boundary records, torchrun bench with every config (C5 through the wavefront pipeline across ranks)
mkdir -p gpurun_out; rm -f gpurun_out/parity_measured.jsonl
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_x.log
RTW_TAG=boxb timeout 600 python tools/exp_time2.py 2>&1 | grep -E "cornell|final" | tee gpurun_out/x_box.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --sweep 1,4 > gpurun_out/x_bench_n2.json 2> gpurun_out/x_bench_n2.err
tail -c 600 gpurun_out/x_bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/x_bench_n2.json").read().strip().splitlines()[-1])
print("N=2", round(d["value"], 1), "Mpaths/s", round(d["ms_per_step"], 2), "ms/step", d.get("multi_gpu_check"))
for k, v in d.get("configs", {}).items(): print(k, round(v["mpaths_per_s"], 1), "Mpaths/s", round(v["ms"], 1), "ms", "commit", v.get("commit_s"))
PY
