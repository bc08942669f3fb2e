#!/bin/bash
# attention kernel bring-up: parity of the K6 implementations, then the encoder A/B (tc4 vs tc vs mma.sync attention)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py -m gpu -x -q -k "attention" > gpurun_out/t_attn.log 2>&1; echo "rc=$?" >> gpurun_out/t_attn.log
tail -n 15 gpurun_out/t_attn.log
MX_ATTENTION_TC4=1 timeout 300 python bench.py --only embed --steps 20 --warmup 5 > gpurun_out/embed_tc4.json 2> gpurun_out/embed_tc4.err
timeout 300 python bench.py --only embed --steps 20 --warmup 5 > gpurun_out/embed_tc.json 2> gpurun_out/embed_tc.err
tail -n 3 gpurun_out/embed_tc4.err
