#!/bin/bash
# gpurun --gpus N -- bash scripts/gpu_final.sh N : everything the round's numbers rest on, on ONE box with N GPUs
set -u
N=${1:-1}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/final_smi_n$N.txt 2>&1
python -m pytest tests -m gpu -x -q -rs > gpurun_out/final_t_gpu_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_t_gpu_n$N.log
tail -n 8 gpurun_out/final_t_gpu_n$N.log
d=$(mktemp -d); memex_b200/_lib/test_host gpu $d > gpurun_out/final_t_host_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/final_t_host_n$N.log; tail -n 3 gpurun_out/final_t_host_n$N.log
for n in 8 4 2 1; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "rc=$?" >> gpurun_out/final_bench_n1.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
        bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/final_bench_n$n.json 2> gpurun_out/final_bench_n$n.err; echo "rc=$?" >> gpurun_out/final_bench_n$n.err
    fi
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/final_bench_n$n.json").read().strip().splitlines()[-1])
    print("N=$n", round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["step_budget_ms"], d["result_digest"], d["result_check"]["ok"], d["config"].get("exchange","")[:20])
    if "ingest" in d: print("   ingest", {k: d["ingest"][k] for k in ("search_alone","ingest_alone","concurrent","concurrent_vs_alone")})
    if "embed" in d: print("   embed", round(d["embed"]["value"]))
except Exception as e: print("N=$n FAILED", e)
PY
    tail -n 2 gpurun_out/final_bench_n$n.err
  fi
done
