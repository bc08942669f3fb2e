#!/bin/bash
# One gpurun call: parity of the opt-in GEMM variants (fresh processes: the switches are read once) and the embedding
# sub-bench under each.   gpurun --timeout 600 -- bash scripts/gpu_gemm_variants.sh
set -u
mkdir -p gpurun_out
for sw in ${PARITY_SWITCHES:-}; do
    env $sw=1 timeout 200 python -m pytest tests/test_encoder_gpu.py -q -x -m gpu \
        -k "tcgen05_gemm_against_torch or tensor_core_paths_vs_oracle or bert_base_shape" > gpurun_out/t_$sw.log 2>&1
    echo "$sw: $(tail -1 gpurun_out/t_$sw.log)"
done
run() {  # name, env...
    local name=$1; shift
    env "$@" timeout 120 python bench.py --only embed --steps 20 --warmup 5 --skip-cpu --skip-extras \
        > gpurun_out/embed_$name.json 2> gpurun_out/embed_$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
for l in open(f"gpurun_out/embed_{name}.json"):
    if l.startswith("{"):
        d = json.loads(l)
        r = d["roofline"]
        print(f"{name:10s} {d['value']:10.0f} seg/s  step {d['ms_per_step']:.3f} ms  gemm {r['gemm_ms_per_step']:.3f} ms  other {r['other_ms_per_step']:.3f} ms")
PY
}
# A/B pairs on the same box: default, then each switch given on the command line, twice
for rep in 1 2; do
    run base$rep MX_NONE=1
    for sw in "$@"; do run ${sw#MX_GEMM_}$rep $sw=1; done
done
