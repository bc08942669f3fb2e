"""Where the scan's power goes: sustained loops of the 10 M-row scan with parts switched off (MX_SCAN_TC_DIAG; results are
wrong in those runs, only GB/s and the board power are read).  Needs a diagnostic build of the library:
    nvcc ... -DMX_TC_DIAG -c memex_b200/csrc/scan_tc.cu -o memex_b200/_lib/DIAG/scan_tc.o, linked with the other objects into
    memex_b200/_lib/DIAG/libmemex_b200.so, and MX_B200_LIB pointing at it (the switches are compiled out of the product)."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from memex_b200.sharded import ShardedStore  # noqa: E402

dev = torch.device("cuda", 0)
import pynvml  # noqa: E402
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)


def sustained(fn, seconds, unit_bytes, tag):
    samples = []
    stop = False

    def poll():
        while not stop:
            samples.append((pynvml.nvmlDeviceGetPowerUsage(h) / 1e3, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            time.sleep(0.02)
    th = threading.Thread(target=poll)
    th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    n = 0
    t0 = time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(10):
            fn()
        n += 10
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    stop = True
    th.join()
    ms = e0.elapsed_time(e1) / n
    tail = samples[len(samples) // 2:]
    print(f"{tag:44s} {unit_bytes / (ms * 1e-3) / 1e9:7.0f} GB/s  {ms * 1e3:8.1f} us/step   power (2nd half) "
          f"{sum(p for p, _ in tail) / len(tail):6.0f} W  max {max(p for p, _ in samples):6.0f} W   SM clock min {min(c for _, c in tail)} MHz")
    time.sleep(1.0)


rows = 10_000_000
a = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
b = torch.empty_like(a)
sustained(lambda: b.copy_(a), 1.5, 2 * (1 << 30), "copy 1 GiB (r+w)")
sustained(lambda: a.sum(), 1.5, (1 << 30), "torch sum over 1 GiB u8 (read only)")
st = ShardedStore("/tmp/mx_power", 384, rows, dtype="f16", device=0)
bench.fill_shard(st, 0, rows, dev)
for nq in (64, 8):
    q = bench.queries_device(nq, dev)
    for diag, what in ((0, "full"), (2, "no MMA"), (4, "no epilogue"), (6, "TMA ring only")):
        os.environ["MX_SCAN_TC_DIAG"] = str(diag)
        sustained(lambda: st.search_device(q, 10), 1.5, rows * 772, f"scan nq={nq} {what}")
os.environ["MX_SCAN_TC_DIAG"] = "0"
q1 = bench.queries_device(1, dev)
sustained(lambda: st.search_device(q1, 10), 1.5, rows * 772, "stream scan f16 nq=1 (CUDA cores)")
st.close()
