"""Burst vs sustained: step time of the 10 M-row scan and of a plain device copy in windows of 20, over ~1 s each."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from memex_b200.sharded import ShardedStore  # noqa: E402

dev = torch.device("cuda", 0)


def smi():
    return os.popen("nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,temperature.memory,"
                    "clocks_event_reasons.active --format=csv,noheader -i 0").read().strip()


def windows(fn, n_win, per, unit_bytes, tag):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_win + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for w in range(n_win):
        for _ in range(per):
            fn()
        ev[w + 1].record()
    torch.cuda.synchronize()
    ms = [ev[w].elapsed_time(ev[w + 1]) / per for w in range(n_win)]
    print(tag, " ".join(f"{unit_bytes / (m * 1e-3) / 1e9:.0f}" for m in ms), "GB/s  |", smi())


print("idle:", smi())
a = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
b = torch.empty_like(a)
for rep in range(2):
    windows(lambda: b.copy_(a), 24, 20, 2 * (1 << 30), "copy 1 GiB (r+w)      ")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
st = ShardedStore("/tmp/mx_sustain", 384, rows, dtype="f16", device=0)
bench.fill_shard(st, 0, rows, dev)
q = bench.queries_device(64, dev)
torch.cuda.synchronize()
import time  # noqa: E402
time.sleep(1.0)
print("idle:", smi())
for rep in range(2):
    windows(lambda: st.search_device(q, 10), 24, 20, rows * 772, f"scan {rows} x 384 f16 nq=64")
    time.sleep(0.5)
q8 = bench.queries_device(8, dev)
windows(lambda: st.search_device(q8, 10), 24, 20, rows * 772, f"scan {rows} x 384 f16 nq=8 ")
st.close()
