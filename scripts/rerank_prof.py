"""rerank_kernel phase timestamps (MX_RERANK_PROF=1, CTA 0's %globaltimer): where the fixed per-step cost goes.
    MX_RERANK_PROF=1 python scripts/rerank_prof.py"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MX_RERANK_PROF"] = "1"
import torch  # noqa: E402

from memex_b200 import capi  # noqa: E402
from memex_b200.storage import B200Store  # noqa: E402

L = capi.lib()
L.mx_debug_rerank_prof.restype = C.c_int32
L.mx_debug_rerank_prof.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
names = ["loads issued", "sorted (5-10 warp sorts + merges)", "tree merge + certificate", "zero rows / entries", "exact f64 fold", "rank sort + write"]
for dtype, n, nq in (("f16", 2_000_000, 64), ("f32", 1_000_000, 1)):
    st = B200Store.new(f"/tmp/mx_prof_{dtype}", dim=384, dtype=dtype, capacity=n)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    for c in range(n // 250_000):
        x = torch.nn.functional.normalize(torch.randn((250_000, 384), generator=g, device="cuda"), dim=1).contiguous()
        torch.cuda.synchronize()
        first = C.c_uint64()
        assert L.mx_store_add_device(st.handle, x.data_ptr(), 250_000, C.byref(first)) == 0
    q = torch.nn.functional.normalize(torch.randn((nq, 384), generator=g, device="cuda"), dim=1).cpu().numpy()
    for _ in range(5):
        st.search_matrix(q, 10)
    acc = np.zeros(6)
    reps = 20
    for _ in range(reps):
        st.search_matrix(q, 10)
        out = (C.c_uint64 * 8)()
        assert L.mx_debug_rerank_prof(st.handle, out) == 0
        t = np.array(list(out)[:7], dtype=np.float64)
        acc += np.diff(t)
    print(f"{dtype} store, {n} rows, {nq} queries: rerank CTA 0 phases (us, mean of {reps})")
    for nm, v in zip(names, acc / reps / 1e3):
        print(f"   {nm:38s} {v:7.2f}")
    print(f"   {'total inside the kernel':38s} {acc.sum() / reps / 1e3:7.2f}")
    st.close()
