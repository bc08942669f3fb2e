#!/bin/bash
set -u
mkdir -p gpurun_out
for v in "4 16" "4 8" "6 8" "8 8" "6 4"; do
  set -- $v
  echo "== MX_SCAN_TC_SAMPLE=$1 DIV=$2"; MX_SCAN_TC_SAMPLE=$1 MX_SCAN_TC_SAMPLE_DIV=$2 python scripts/diag_scan_fixed.py 2>&1 | grep "k 10"
done | tee gpurun_out/diag_scan_sample2.txt
