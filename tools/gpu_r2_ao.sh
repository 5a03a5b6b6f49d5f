#!/bin/bash
# round 2, run ao (2 GPUs): wavefront finisher kernel (end-of-frame drain): tests, A/B on one GPU, two ranks under torchrun
mkdir -p gpurun_out; L=gpurun_out/ao_finisher.log; : > $L
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee -a $L
for cfg in "RTW_WF_FINISH=0" "RTW_WF_FINISH=1"; do echo "== $cfg" | tee -a $L; env $cfg RTW_TIMING=1 timeout 600 python tools/sweep.py 1 --spp 32 2>&1 | grep -E "done|mpaths" | cut -c1-200 | tee -a $L; done
for cfg in "RTW_WF_FINISH=0" "RTW_WF_FINISH=1"; do echo "== N=2 $cfg" | tee -a $L
  env $cfg timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29624 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline --sweep 1,16 > gpurun_out/ao_bench_n2.json 2> gpurun_out/ao_bench_n2.err
  python - <<'PY' | tee -a $L
import json
d = json.loads(open("gpurun_out/ao_bench_n2.json").read().strip().splitlines()[-1])
for k, v in d.get("configs", {}).items():
    if k.startswith("C5"): print(k, round(v["mpaths_per_s"], 1), "Mpaths/s", round(v["ms"], 1), "ms", "rays/path", round(v["rays_per_path"], 4))
PY
done
