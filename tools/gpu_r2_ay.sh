#!/bin/bash
# adopted: no tile list in the big media variants — every config, then the whole GPU suite
mkdir -p gpurun_out; L=gpurun_out/ay_notile_media_adopted.log; : > $L
RTW_TAG=ay timeout 300 python tools/exp_time2.py 2>&1 | tee -a $L
rm -f gpurun_out/parity_measured.jsonl
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee -a $L
