#!/bin/bash
# K2 ring-depth / query-block sweep on the headline step (10Mx384 fp16, 64 queries)
set -u
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 30 --warmup 5 --skip-cpu --skip-extras > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/sweep_{n}.json'))
    print(n, 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), flush=True)
except Exception as e:
    print(n, 'failed', e)
PY
}
run qm64_s9 MX_X=1
run qm128_s6 MX_SCAN_TC_QM=128
run qm64_s6 MX_SCAN_TC_STAGES=6
run qm64_s7 MX_SCAN_TC_STAGES=7
run qm64_s8 MX_SCAN_TC_STAGES=8
timeout 900 python -m pytest tests/test_store_gpu.py -m gpu -x -q > gpurun_out/t_store.log 2>&1; echo "rc=$?" >> gpurun_out/t_store.log
tail -n 4 gpurun_out/t_store.log
