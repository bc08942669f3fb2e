#!/bin/bash
# gpurun --gpus N -- bash scripts/gpu_multi_parity.sh N
# Hardware multi-GPU parity (tests/test_sharded_gpu.py: ShardedStore vs the oracle over the unsharded corpus, both exchange
# forms) + the N-rank bench with its in-run answer check and result digest, + the 1-rank bench for the digest to compare with.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_smi_n$N.txt 2>&1
python -m pytest tests/test_sharded_gpu.py -m gpu -q -rs > gpurun_out/t_sharded_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_sharded_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?" >> gpurun_out/bench_n$N.err
python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > gpurun_out/bench_n1_digest.json 2> gpurun_out/bench_n1_digest.err
echo "rc=$?" >> gpurun_out/bench_n1_digest.err
tail -n 6 gpurun_out/t_sharded_n$N.log
tail -c 1500 gpurun_out/bench_n$N.json; tail -n 3 gpurun_out/bench_n$N.err
tail -c 600 gpurun_out/bench_n1_digest.json; tail -n 3 gpurun_out/bench_n1_digest.err
