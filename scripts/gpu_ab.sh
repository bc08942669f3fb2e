#!/bin/bash
set -u
mkdir -p gpurun_out
python bench.py --only ingest --ingest-seconds 2 > gpurun_out/ingest.json 2> gpurun_out/ingest.err; echo "rc=$?" >> gpurun_out/ingest.err
cat gpurun_out/ingest.json; tail -5 gpurun_out/ingest.err
d=$(mktemp -d); memex_b200/_lib/test_host gpu $d > gpurun_out/t_host_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/t_host_gpu.log; tail -3 gpurun_out/t_host_gpu.log
bash scripts/gpu_sanitize.sh
