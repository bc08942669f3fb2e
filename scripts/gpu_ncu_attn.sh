#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention_tc' -s 8 -c 1 \
    -f -o gpurun_out/prof_attn python bench.py --only embed --steps 3 --warmup 3 > gpurun_out/ncu_attn.log 2>&1
tail -n 3 gpurun_out/ncu_attn.log
