#!/bin/bash
mkdir -p gpurun_out
V=rust-ray-tracing-in-a-weekend_b200/variants
for E in "RTW_NOP=1" "RTW_EMULATE_RANKS=8" "RTW_EMULATE_RANKS=8 RTW_B_SPP=4" "RTW_EMULATE_RANKS=8 RTW_B_SHARE=40" "RTW_EMULATE_RANKS=2"; do
  echo "== $E" | tee -a gpurun_out/j_tail.log
  env $E RTW_LIB_PATH=$PWD/$V/instr.so timeout 600 python tools/diag_tail.py 2>&1 | grep -v "^$" | tee -a gpurun_out/j_tail.log
done
