#!/bin/bash
# timings of the config-like renders under different environment settings: tools/ab_env.sh "TAG:ENV=V,ENV2=V" ...
mkdir -p gpurun_out
for spec in "$@"; do
  T=${spec%%:*}; E=${spec#*:}
  env $(echo $E | tr ',' ' ') RTW_TAG=$T python tools/exp_time2.py
done 2>&1 | tee gpurun_out/ab_env.log
