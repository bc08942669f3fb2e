#!/bin/bash
# gpurun --timeout 1500 -- bash scripts/gpu_check2.sh : GPU tests + smoke + bench + the launch list of the headline step
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -rs > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --skip-cpu --skip-extras --skip-check \
    > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_single.csv python bench.py --only single --steps 20 --warmup 5 \
    > gpurun_out/ncu_launches_single.log 2>&1
python scripts/ncu_summary.py --launches gpurun_out/launches.csv > gpurun_out/launches.txt 2>&1
python scripts/ncu_summary.py --launches gpurun_out/launches_single.csv > gpurun_out/launches_single.txt 2>&1
tail -n 8 gpurun_out/t_gpu.log; tail -n 2 gpurun_out/smoke.log; tail -n 3 gpurun_out/bench.err
head -n 12 gpurun_out/launches.txt; head -n 10 gpurun_out/launches_single.txt
python - <<PY
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","step_budget_ms","result_digest")}, d["e2e"]["value"], d["result_check"]["ok"])
print("single", d["single_query"]["value"], d["single_query"]["e2e"]["value"], d["single_query"]["roofline"]["frac"], d["single_query"]["roofline"]["other_kernels_ms_per_step"])
print("embed", d["embed"]["value"], d["embed"].get("parity"), d["embed"]["roofline"]["whole_step_frac"])
print("config1", d.get("config1"))
PY
MX_RERANK_PROF=1 python scripts/rerank_prof.py > gpurun_out/rerank_prof.txt 2>&1; cat gpurun_out/rerank_prof.txt | tail -20
