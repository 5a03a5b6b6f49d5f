#!/bin/bash
# like ab.sh but each argument is "tag:ENV=VAL,...:lib"
mkdir -p gpurun_out
for spec in "$@"; do
  T=${spec%%:*}; rest=${spec#*:}; E=${rest%%:*}; L=${rest#*:}
  envs=$(echo $E | tr ',' ' ')
  env $envs RTW_LIB_PATH=$PWD/$L python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$T] C1', round(d['ms_per_step'],2), 'ms', round(d['value'],1), 'Mpaths/s')"
  env $envs RTW_LIB_PATH=$PWD/$L RTW_TAG=$T python tools/exp_time2.py
done 2>&1 | tee gpurun_out/ab.log
