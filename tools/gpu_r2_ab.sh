#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/ab_units.log; : > $L
for u in 0 192; do for c in 0 1; do timeout 300 python tools/exp_units.py $u $c 2>&1 | tail -1 | tee -a $L; done; done
