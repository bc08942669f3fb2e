#!/bin/bash
# attention kernel bring-up: parity of the three K6 implementations, then the encoder A/B (tcgen05 vs mma.sync attention)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py -m gpu -x -q -k "attention" > gpurun_out/t_attn.log 2>&1; echo "rc=$?" >> gpurun_out/t_attn.log
tail -n 15 gpurun_out/t_attn.log
timeout 300 python -m pytest tests/test_encoder_gpu.py -m gpu -x -q > gpurun_out/t_enc.log 2>&1; echo "rc=$?" >> gpurun_out/t_enc.log
tail -n 5 gpurun_out/t_enc.log
timeout 300 python bench.py --only embed --steps 20 --warmup 5 > gpurun_out/embed_tc.json 2> gpurun_out/embed_tc.err
MX_ATTENTION_MMA=1 timeout 300 python bench.py --only embed --steps 20 --warmup 5 > gpurun_out/embed_mma.json 2> gpurun_out/embed_mma.err
cat gpurun_out/embed_tc.json gpurun_out/embed_mma.json
