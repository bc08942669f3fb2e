"""A/B of the seeded scan's launch form on a small shard: cooperative (default) / plain launch / plain launch as a
programmatic dependent.  Device-resident steps per second over 40 steps, each measurement after a 0.4 s idle."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from memex_b200.sharded import ShardedStore  # noqa: E402

dev = torch.device("cuda", 0)
q = bench.queries_device(64, dev)
for rows in [int(a) for a in sys.argv[1:]] or [1_250_000]:
    st = ShardedStore(f"/tmp/mx_ab_{rows}", 384, rows, dtype="f16", device=0)
    bench.fill_shard(st, 0, rows, dev)
    for rep in range(2):
        for plain in ("0", "1", "2"):
            os.environ["MX_SCAN_TC_PLAIN"] = plain
            for _ in range(3):
                st.search_device(q, 10)
            torch.cuda.synchronize()
            time.sleep(0.4)
            for _ in range(3):
                st.search_device(q, 10)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(40):
                ids, sc, cn = st.search_device(q, 10)
            e1.record()
            torch.cuda.synchronize()
            print(f"rows {rows} MX_SCAN_TC_PLAIN={plain}: {e0.elapsed_time(e1) / 40 * 1e3:7.1f} us per step   (ids checksum {int(ids.sum().item())})", flush=True)
    st.close()
os.environ.pop("MX_SCAN_TC_PLAIN", None)
