#!/bin/bash
# gpurun --gpus N -- bash scripts/gpu_multi_parity.sh N
# Hardware multi-GPU parity: the one-process-per-GPU ShardedStore vs the oracle over the unsharded corpus (both exchange
# forms) + the N-rank bench with its in-run answer check and result digest (+ the 1-rank bench's end-to-end figure).
set -u
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_sharded_gpu.py -m gpu -q -rs > gpurun_out/t_sharded_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_sharded_n$N.log
tail -n 4 gpurun_out/t_sharded_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --skip-extras > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?" >> gpurun_out/bench_n$N.err
python bench.py --steps 20 --warmup 5 --skip-extras --skip-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?" >> gpurun_out/bench_n1.err
python - <<PY
import json
for n in ($N, 1):
    d=json.loads(open(f"gpurun_out/bench_n{n}.json").read().strip().splitlines()[-1])
    print("N", n, {k:d[k] for k in ("value","ms_per_step","result_digest")}, "e2e", d["e2e"]["value"], d["result_check"]["ok"], d["config"].get("exchange","")[:24])
PY
tail -n 2 gpurun_out/bench_n$N.err gpurun_out/bench_n1.err
