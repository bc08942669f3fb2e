"""Sampled tiles per CTA (MX_SCAN_TC_SAMPLE) against shard size: scan kernel time (library event timer), 64 queries, k = 10.
Every measurement starts after a 0.4 s idle so that none of them runs power-capped."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from memex_b200 import capi  # noqa: E402
from memex_b200.sharded import ShardedStore  # noqa: E402

L = capi.lib()
dev = torch.device("cuda", 0)
q = bench.queries_device(64, dev)
sizes = [int(a) for a in sys.argv[1:]] or [625_000, 1_250_000, 2_500_000, 10_000_000]
for rows in sizes:
    st = ShardedStore(f"/tmp/mx_sweep_{rows}", 384, rows, dtype="f16", device=0)
    bench.fill_shard(st, 0, rows, dev)
    line = []
    for sample in (0, 1, 2, 3, 4, 6):
        os.environ["MX_SCAN_TC_SAMPLE"] = str(sample)
        for _ in range(3):
            st.search_device(q, 10)
        torch.cuda.synchronize()
        time.sleep(0.4)
        for _ in range(2):
            st.search_device(q, 10)
        L.mx_store_set_timing(st.local.handle, 1)
        for _ in range(12):
            st.search_device(q, 10)
        torch.cuda.synchronize()
        a, n, b, m = C.c_double(), C.c_uint64(), C.c_double(), C.c_uint64()
        L.mx_store_get_timing(st.local.handle, C.byref(a), C.byref(n), C.byref(b), C.byref(m))
        L.mx_store_set_timing(st.local.handle, 0)
        line.append(f"P={sample}: {a.value / n.value * 1e3:7.1f}")
    print(f"rows {rows:9d} (ideal {rows * 772 / 6552.3e3:7.1f} us): " + "  ".join(line), flush=True)
    st.close()
    del st
    torch.cuda.empty_cache()
os.environ.pop("MX_SCAN_TC_SAMPLE", None)
