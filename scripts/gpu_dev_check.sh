#!/bin/bash
# Development loop on one B200: store / pipeline parity tests, the scan's phase profile, one bench line without the CPU legs.
#   gpurun --timeout 600 -- bash scripts/gpu_dev_check.sh
set -u
mkdir -p gpurun_out
python -m pytest tests/test_store_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -5
python scripts/scan_tc_prof.py 2>&1 | tee gpurun_out/scan_tc_prof_dev.txt | head -18
python bench.py --skip-cpu --skip-ingest --skip-check > gpurun_out/bench_dev.json 2> gpurun_out/bench_dev.err
tail -c 600 gpurun_out/bench_dev.err
python - <<'PY'
import json
l = json.load(open("gpurun_out/bench_dev.json"))
for k in ("value", "ms_per_step", "e2e", "sustained", "step_budget_ms", "clocks"):
    print(k, l.get(k))
print("single", l["single_query"]["value"], l["single_query"]["e2e"]["value"])
e = l["embed"]
print("embed", e["value"], e["roofline"]["whole_step_frac"], e["clocks"], e["sustained"]["value"], e["default_model"]["value"], e["ragged"]["value"])
PY
