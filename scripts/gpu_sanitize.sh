#!/bin/bash
# gpurun --timeout 1500 -- bash scripts/gpu_sanitize.sh
# compute-sanitizer over small shapes of every kernel family that mixes generic-proxy, async-proxy (TMA) and tcgen05
# accesses, a grid barrier or cross-GPU flag waits: the tcgen05 scan (plain and seeded), every tcgen05 GEMM epilogue
# (bias, GELU, LayerNorm whole-row and split over a cluster), the tcgen05 attention kernel, the stream scan + rerank +
# exact fallback, and the peer-memory push / merge.  Logs -> gpurun_out/sanitize_<tool>.log; summary at the end.
set -u
mkdir -p gpurun_out
SCAN='parity_tcgen05_scan and (128-384-8 or 129-64-9 or 7777-100-33 or 700-448-130 or 5000-512-9)'
SEED='seeded_scan_keeps_lowest_ids'
STORE='test_hnsw or golden_small or peer_memory_exchange or certificate_catches_a_near_tie_crowd_stream'
GEMM='gemm_against_torch and (128-128-64 or 300-1536-384 or 1000-384-384 or 129-256-128 or 1000-768-768)'
ATT='attention_kernels_against_torch and (3-64-64-2 or 5-37-384-12)'
run() {  # tool, tag, pytest -k expression, files...
  tool=$1; tag=$2; expr=$3; shift 3
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 86 --print-limit 20 \
      python -m pytest "$@" -m gpu -x -q -p no:cacheprovider -k "$expr" > gpurun_out/sanitize_${tool}_${tag}.log 2>&1
  echo "exit=$?" >> gpurun_out/sanitize_${tool}_${tag}.log
}
run memcheck scan "$SCAN" tests/test_store_gpu.py
run memcheck seeded "$SEED" tests/test_store_gpu.py
run memcheck store "$STORE" tests/test_store_gpu.py
run memcheck gemm "$GEMM" tests/test_encoder_gpu.py
run memcheck attention "$ATT" tests/test_encoder_gpu.py
run racecheck scan "$SCAN" tests/test_store_gpu.py
run racecheck store "$STORE" tests/test_store_gpu.py
run racecheck gemm "$GEMM" tests/test_encoder_gpu.py
run racecheck attention "$ATT" tests/test_encoder_gpu.py
run synccheck scan "$SCAN" tests/test_store_gpu.py
run synccheck gemm "$GEMM" tests/test_encoder_gpu.py
run synccheck attention "$ATT" tests/test_encoder_gpu.py
for f in gpurun_out/sanitize_*.log; do
  echo "== $f: $(grep -E 'passed|failed|error' $f | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $f | tail -1) | $(tail -1 $f)"
done | tee gpurun_out/sanitize_summary.txt
