#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_store_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q > gpurun_out/t_gpu_store.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu_store.log; tail -n 3 gpurun_out/t_gpu_store.log
MX_RERANK_PROF=1 python scripts/rerank_prof.py > gpurun_out/rerank_prof.txt 2>&1; tail -18 gpurun_out/rerank_prof.txt
python bench.py --steps 20 --warmup 5 --skip-ingest --skip-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","step_budget_ms","result_digest")}, d["e2e"]["value"], d["result_check"]["ok"])
print("single", d["single_query"]["value"], d["single_query"]["e2e"]["value"], d["single_query"]["roofline"]["frac"], d["single_query"]["roofline"]["other_kernels_ms_per_step"])
print("embed", d["embed"]["value"], d["embed"]["roofline"]["whole_step_frac"])
PY
tail -n 3 gpurun_out/bench.err
