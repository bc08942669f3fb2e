#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_store_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q > gpurun_out/t_store.log 2>&1; echo "rc=$?" >> gpurun_out/t_store.log
tail -n 4 gpurun_out/t_store.log
echo "--- seeding on"; timeout 300 python scripts/diag_scan_fixed.py 2>&1 | grep "k 10\|k 26" 
echo "--- seeding off"; MX_SCAN_TC_SAMPLE=0 timeout 300 python scripts/diag_scan_fixed.py 2>&1 | grep "k 10"
python bench.py --steps 30 --warmup 5 --skip-cpu --skip-extras > gpurun_out/seed_on.json 2> gpurun_out/seed_on.err
python -c "
import json; d=json.load(open('gpurun_out/seed_on.json')); print('10M on:', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
