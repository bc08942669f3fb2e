#!/bin/bash
# One gpurun call after an encoder-only change: GPU parity tests, smoke, the bench, and the encoder's ncu evidence (the scan
# captures of scripts/gpu_check.sh stay valid).   gpurun --timeout 900 -- bash scripts/gpu_final_encoder.sh
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_embed.csv python bench.py --only embed --steps 3 --warmup 3 --skip-extras \
    > gpurun_out/ncu_launches_embed.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc|attention' -s 20 -c 5 \
    -f -o gpurun_out/prof_enc python bench.py --only embed --steps 3 --warmup 3 --skip-extras \
    > gpurun_out/ncu_enc.log 2>&1
tail -n 4 gpurun_out/t_gpu.log gpurun_out/smoke.log
cut -c1-400 gpurun_out/bench.json
