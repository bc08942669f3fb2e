#!/bin/bash
# gpurun --timeout 1800 -- bash scripts/gpu_profiles.sh : the ncu evidence of the round (numbers printed under ncu are never bench values)
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --skip-cpu --skip-extras --skip-check \
    > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_embed.csv python bench.py --only embed --steps 3 --warmup 3 --skip-extras --skip-check \
    > gpurun_out/ncu_launches_embed.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_single.csv python bench.py --only single --steps 20 --warmup 5 \
    > gpurun_out/ncu_launches_single.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'scan_tc|rerank' -s 6 -c 3 \
    -f -o gpurun_out/prof_scan python bench.py --steps 3 --warmup 3 --skip-cpu --skip-extras --skip-check \
    > gpurun_out/ncu_scan.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc|attention' -s 20 -c 5 \
    -f -o gpurun_out/prof_enc python bench.py --only embed --steps 3 --warmup 3 --skip-extras --skip-check \
    > gpurun_out/ncu_enc.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches*.csv
