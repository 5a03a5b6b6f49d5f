#!/bin/bash
# round 2, run aq (2 GPUs): the local-framebuffer merge of the wavefront pipeline, scalar reds vs four floats per red
mkdir -p gpurun_out; L=gpurun_out/aq_merge.log; : > $L
for cfg in "RTW_WF_MERGE=scalar" "RTW_WF_MERGE=v4"; do echo "== $cfg" | tee -a $L
  env $cfg RTW_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29626 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline --sweep 1 > gpurun_out/aq_bench_n2.json 2> gpurun_out/aq_bench_n2.err
  grep -E "merge of|\(done\)" gpurun_out/aq_bench_n2.err | tail -6 | tee -a $L
  python - <<'PY' | tee -a $L
import json
d = json.loads(open("gpurun_out/aq_bench_n2.json").read().strip().splitlines()[-1])
for k, v in d.get("configs", {}).items():
    if k.startswith("C5"): print(k, round(v["mpaths_per_s"], 1), "Mpaths/s", round(v["ms"], 1), "ms", "mean", v["image_mean"])
PY
done
timeout 600 python -m pytest tests -m gpu -q -x -k "two_gpus or wavefront" 2>&1 | tail -2 | tee -a $L
