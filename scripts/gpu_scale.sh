#!/bin/bash
# gpurun --gpus N -- bash scripts/gpu_scale.sh N : the N-rank bench exactly as the driver launches it
set -u
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?" >> gpurun_out/bench_n$N.err
MX_BENCH_HNSW_ROWS=${REF_ROWS:-200000} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
echo "rc=$?" >> gpurun_out/bench_ref_n$N.err
tail -n 3 gpurun_out/bench_n$N.json gpurun_out/bench_n$N.err gpurun_out/bench_ref_n$N.json
