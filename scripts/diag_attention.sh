#!/bin/bash
# What bounds K6?  Timing-only builds of attention_tc.cu with parts removed (wrong results; MX_ATTN_DIAG bit flags:
# 1 no MUFU, 2 no max pass, 4 no P stores / proxy fence, 8 no pass-2 work, 16 no O read-out / store).
mkdir -p gpurun_out
for d in 0 1 2 4 3; do
  MX_ATTN_DIAG=$d python bench.py --only embed --steps 10 --warmup 3 2>gpurun_out/diag_err_$d.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('diag $d: other_ms_per_step', round(d['roofline']['other_ms_per_step'],4), '-> attention ~', round((d['roofline']['other_ms_per_step']-0.075)/6*1e3,1), 'us/layer')"
done | tee gpurun_out/diag_attention.txt
