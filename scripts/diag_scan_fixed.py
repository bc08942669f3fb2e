"""Where does the per-launch fixed cost of K2 come from?  Times the scan kernel (library event timer) over shards of
different sizes and list capacities (k = 10 -> L = 16, k = 26 -> L = 32) with 64 queries."""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from memex_b200 import capi
from memex_b200.sharded import ShardedStore
import bench

L = capi.lib()
dev = torch.device("cuda", 0)
q = bench.queries_device(64, dev)
for rows in (312_500, 625_000, 1_250_000, 2_500_000, 5_000_000):
    st = ShardedStore(f"/tmp/mx_diag_{rows}", 384, rows, dtype="f16", device=0)
    bench.fill_shard(st, 0, rows, dev)
    for k in (10, 26):
        for _ in range(5):
            st.search_device(q, k)
        torch.cuda.synchronize()
        L.mx_store_set_timing(st.local.handle, 1)
        for _ in range(20):
            st.search_device(q, k)
        torch.cuda.synchronize()
        a, n, b, m = C.c_double(), C.c_uint64(), C.c_double(), C.c_uint64()
        L.mx_store_get_timing(st.local.handle, C.byref(a), C.byref(n), C.byref(b), C.byref(m))
        L.mx_store_set_timing(st.local.handle, 0)
        per = a.value / n.value
        gb = rows * (384 * 2 + 4) / 1e9
        print(f"rows {rows:8d} k {k:2d}: scan {per*1e3:7.1f} us  ({gb/per*1e3:6.0f} GB/s)  other {b.value/20*1e3:6.1f} us/step", flush=True)
    st.close()
    del st
    torch.cuda.empty_cache()
