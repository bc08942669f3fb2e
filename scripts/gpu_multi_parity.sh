#!/bin/bash
# gpurun --gpus N -- bash scripts/gpu_multi_parity.sh N
# Hardware multi-GPU parity: the C++ ShardedB200Store (one process, N GPUs, mx_shard_group_connect_local), the
# one-process-per-GPU ShardedStore vs the oracle over the unsharded corpus (both exchange forms), + the N-rank bench with
# its in-run answer check and result digest.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_smi_n$N.txt 2>&1
d=$(mktemp -d); memex_b200/_lib/test_host gpu $d > gpurun_out/t_host_gpu_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/t_host_gpu_n$N.log
python -m pytest tests/test_sharded_gpu.py -m gpu -q -rs > gpurun_out/t_sharded_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_sharded_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 ${BENCH_FLAGS:-} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?" >> gpurun_out/bench_n$N.err
tail -n 4 gpurun_out/t_host_gpu_n$N.log
tail -n 6 gpurun_out/t_sharded_n$N.log
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","step_budget_ms","result_digest")}, d["e2e"]["value"], d["result_check"]["ok"], d["config"].get("exchange"))
PY
tail -n 3 gpurun_out/bench_n$N.err
