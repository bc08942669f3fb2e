"""scan_tc_kernel phase timestamps of CTA 0 (MX_SCAN_TC_PROF=1): where a small shard's scan loses its time.
    MX_SCAN_TC_PROF=1 python scripts/scan_tc_prof.py"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MX_SCAN_TC_PROF"] = "1"
import torch  # noqa: E402

import bench  # noqa: E402
from memex_b200 import capi  # noqa: E402
from memex_b200.sharded import ShardedStore  # noqa: E402

L = capi.lib()
L.mx_debug_scan_tc_prof.restype = C.c_int32
L.mx_debug_scan_tc_prof.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
dev = torch.device("cuda", 0)
q = bench.queries_device(64, dev)
names = ["query preparation", "sampling pass (4 tiles)", "grid barrier + tau0", "sampled tiles again (L2)", "remaining tiles", ]
for rows in (1_250_000, 5_000_000):
    st = ShardedStore(f"/tmp/mx_prof_tc_{rows}", 384, rows, dtype="f16", device=0)
    bench.fill_shard(st, 0, rows, dev)
    for _ in range(5):
        st.search_device(q, 10)
    torch.cuda.synchronize()
    acc = np.zeros(5)
    fine = np.zeros(10)
    reps = 20
    for _ in range(reps):
        st.search_device(q, 10)
        torch.cuda.synchronize()
        out = (C.c_uint64 * 32)()
        assert L.mx_debug_scan_tc_prof(st.local.handle, out) == 0
        t = np.array(list(out)[:6], dtype=np.float64)
        acc += np.diff(t)
        f = np.array([out[2], out[6], out[3]] + list(out)[8:16], dtype=np.float64)
        fine += np.diff(f)
    tiles = (rows + 127) // 128 / 148
    print(f"{rows} rows ({tiles:.0f} tiles per CTA, ideal {rows * 772 / 6552.3e3:.1f} us at the copy peak): CTA 0 phases (us, mean of {reps})")
    for nm, v in zip(names, acc / reps / 1e3):
        print(f"   {nm:30s} {v:8.2f}")
    print(f"   {'total first mark -> last tile':30s} {acc.sum() / reps / 1e3:8.2f}")
    fnames = ["sampled -> barrier passed", "tau0 from the samples", "(to the first real tile)", "tile A: tau load", "tile A: accumulator wait",
              "tile A: filter 128 columns", "(loop)", "tile B: tau load", "tile B: accumulator wait", "tile B: filter 128 columns"]
    for i, (nm, v) in enumerate(zip(fnames, fine / reps / 1e3)):
        if i < 2 or out[8] != 0:       # the per-tile marks exist in a -DMX_TC_DIAG build only
            print(f"      {nm:30s} {v:8.2f}")
    st.close()
