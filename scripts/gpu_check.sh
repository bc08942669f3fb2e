#!/bin/bash
# One gpurun call: GPU parity tests, smoke, the bench (both arms), the ncu launch lists and full captures of the
# dominant kernels.  Outputs land in gpurun_out/.
#   gpurun --timeout 1800 -- bash scripts/gpu_check.sh [quick]
set -u
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
nvidia-smi > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
if [ "${1:-}" != "quick" ]; then
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
# launch lists (per-launch device time, cold-cache + serialised): the headline command, then the two sub-benches
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --skip-cpu --skip-extras \
    > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_embed.csv python bench.py --only embed --steps 3 --warmup 3 \
    > gpurun_out/ncu_launches_embed.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_single.csv python bench.py --only single --steps 20 --warmup 5 \
    > gpurun_out/ncu_launches_single.log 2>&1
# the dominant kernels, full set
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'scan_tc' -s 4 -c 2 \
    -f -o gpurun_out/prof_scan python bench.py --steps 3 --warmup 3 --skip-cpu --skip-extras \
    > gpurun_out/ncu_scan.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'scan_stream' -s 10 -c 2 \
    -f -o gpurun_out/prof_stream python bench.py --only single --steps 20 --warmup 5 \
    > gpurun_out/ncu_stream.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc|attention' -s 20 -c 5 \
    -f -o gpurun_out/prof_enc python bench.py --only embed --steps 3 --warmup 3 \
    > gpurun_out/ncu_enc.log 2>&1
fi
tail -n 5 gpurun_out/t_gpu.log gpurun_out/smoke.log gpurun_out/bench.json gpurun_out/bench.err
