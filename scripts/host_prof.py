"""Where the host-buffer search call spends its wall clock (MX_HOST_PROF=1): python scripts/host_prof.py [rows]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MX_HOST_PROF"] = "1"
import torch  # noqa: E402

import bench  # noqa: E402
from memex_b200 import capi  # noqa: E402
from memex_b200.sharded import ShardedStore  # noqa: E402

L = capi.lib()
L.mx_debug_host_prof.restype = C.c_int32
L.mx_debug_host_prof.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int32]
dev = torch.device("cuda", 0)
names = ["finiteness check", "device + staging buffers", "copy into pinned", "H2D enqueue", "kernel launches", "D2H enqueue",
         "stream synchronise", "copy out"]
for rows in [int(a) for a in sys.argv[1:]] or [1_250_000]:
    st = ShardedStore(f"/tmp/mx_hprof_{rows}", 384, rows, dtype="f16", device=0)
    bench.fill_shard(st, 0, rows, dev)
    q_dev = bench.queries_device(64, dev)
    q = q_dev.cpu().numpy()
    def dev_loop(tag):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            st.search_device(q_dev, 10)
        e1.record()
        torch.cuda.synchronize()
        clk = os.popen("nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw --format=csv,noheader -i 0").read().strip()
        print(f"{rows} rows: device-resident step {tag} {e0.elapsed_time(e1) / n * 1e3:.1f} us   [{clk}]")
    n = 200
    for _ in range(10):
        st.search_device(q_dev, 10)
    dev_loop("before")
    for _ in range(10):
        st.search(q, 10, nq=64)
    out, calls = (C.c_double * 12)(), C.c_uint64()
    L.mx_debug_host_prof(out, C.byref(calls), 1)
    t0 = time.perf_counter()
    for _ in range(n):
        st.search(q, 10, nq=64)
    wall = (time.perf_counter() - t0) / n * 1e6
    L.mx_debug_host_prof(out, C.byref(calls), 1)
    dev_loop("after")
    print(f"{rows} rows: host-buffer call {wall:.1f} us per step (python included)")
    tot = 0.0
    for nm, v in zip(names, list(out)[:8]):
        print(f"   {nm:28s} {v / calls.value:8.2f}")
        tot += v / calls.value
    print(f"   {'inside the C call':28s} {tot:8.2f}    python around it {wall - tot:8.2f}")
    st.close()
