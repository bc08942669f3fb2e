#!/bin/bash
# Last GPU call of round 2 (tight budget): the encoder tests that were not re-run after the late changes, the bench exactly
# as the driver runs it, the launch list of the headline command and one full capture of the scan + rerank.
#   gpurun --timeout 600 -- bash scripts/gpu_final_r2b.sh
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_encoder_gpu.py -m gpu -x -q \
    -k "attention_kernels or tensor_core_paths or packed_layout or encode_errors or sentence_t5 or bert_base_shape" \
    > gpurun_out/t_encoder_subset.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_encoder_subset.log
tail -n 3 gpurun_out/t_encoder_subset.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --skip-cpu --skip-extras --skip-check \
    > gpurun_out/ncu_launches.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'scan_tc|rerank' -s 6 -c 3 \
    -f -o gpurun_out/prof_scan python bench.py --steps 3 --warmup 3 --skip-cpu --skip-extras --skip-check \
    > gpurun_out/ncu_scan.log 2>&1
python - <<'PY'
import json
l = json.load(open("gpurun_out/bench_final.json"))
for k in ("value", "ms_per_step", "e2e", "sustained", "step_budget_ms", "clocks", "result_check"):
    print(k, l.get(k))
print("roofline", l["roofline"])
print("single", l["single_query"]["value"], l["single_query"]["e2e"]["value"])
e = l["embed"]
print("embed", e["value"], e["roofline"]["whole_step_frac"], e["parity"], e["sustained"]["value"])
print("ingest", l["ingest"]["concurrent"])
print("cpu", l["cpu_baseline"]["value"], l["config1"]["gpu"]["us_per_query"])
PY
ls -la gpurun_out/prof_scan.ncu-rep gpurun_out/launches.csv
