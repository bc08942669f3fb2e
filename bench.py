#!/usr/bin/env python
"""bench.py -- headline benchmark of the memex embedding + vector-search hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): queries/sec over a 10 M x 384 corpus, top-10.  One STEP = one batch of 64
queries scanned against the whole corpus (fp16 rows, cosine, exact top-10).  At N GPUs the 10 M
rows are dealt contiguously over the ranks (row sharding): every rank scans
its shard, ONE all-gather moves the per-shard top-k, every rank merges -> "scaling": "strong".

The JSON line also carries, as sub-objects, the two other configurations BASELINE.json names for
one GPU (they do not shard, so they are reported at N = 1 by rank 0 only):
  "single_query": config 2, 1 M x 384 fp32, one query per call, top-10
  "embed":        config 3, segments embedded / sec, MiniLM-L6, B = 256, S = 256, bf16

`--impl reference` times the CPU restatement of the reference's own search path (HNSW, M = 16,
ef_construction = 200, ef = 32; oracle/hnsw_oracle.cpp) on the host cores: the reference is Rust
over un-vendored crates and cannot be built here (DESIGN.md "Oracle").
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DIM = 384
TOPK = 10
NQ = 64
IDLE_BEFORE_REGION_S = 0.5     # the GPU idles this long before a timed region so that every region starts from the same power state
SUSTAINED_S = 1.0              # length of the back-to-back run behind the `sustained` sub-object
CORPUS_SEED, QUERY_SEED = 1234, 4321
CHUNK_ROWS = 250_000           # generator granularity: chunk c is torch.Generator(seed = CORPUS_SEED + c)
# bounded sample the CPU HNSW restatement is built over (--impl reference); MX_BENCH_HNSW_ROWS shrinks it for the contract test
HNSW_SAMPLE_ROWS = int(os.environ.get("MX_BENCH_HNSW_ROWS", "200000"))
HNSW_LEG_ROWS = 20_000         # ... and for the cpu_baseline leg of the GPU arm
EXACT_SAMPLE_ROWS = 500_000    # bounded sample for the all-core exact brute force


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def ncu_traffic(kernel_key: str):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture"""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel_key)
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark(self):
        return time.perf_counter()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0: float, t1: float):
        sm, mx, reasons, power = [], 0.0, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if not (t0 - 0.15 <= t <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            try:
                power.append(float(f[2]))
            except ValueError:
                pass
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            # region shorter than the sampling period: take whatever was seen
            for t, line in self.rows[-3:]:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[0]))
                    mx = max(mx, float(f[1]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w": float(np.median(power)) if power else None}


# ------------------------------------------------------------------------------------------------
# synthetic corpus / queries (N(0,1) rows, L2-normalised; queries = rows + noise)
# ------------------------------------------------------------------------------------------------
def box_copy_bandwidth(device, nbytes: int = 1 << 30, reps: int = 5) -> float:
    """device-to-device copy bandwidth of THIS box in GB/s (bytes read + bytes written per second), measured the way
    MEASURED_PEAKS.json's figure was: context for roofline.frac, whose denominator is the pool-wide recorded number --
    individual B200s of the pool copy at 6.5-7.0 TB/s"""
    import torch
    a = torch.empty(nbytes, dtype=torch.uint8, device=device)
    b = torch.empty(nbytes, dtype=torch.uint8, device=device)
    for _ in range(2):
        b.copy_(a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        b.copy_(a)
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / reps
    del a, b
    return 2 * nbytes / (ms * 1e-3) / 1e9


def corpus_chunk_device(chunk: int, rows: int, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(CORPUS_SEED + chunk)
    x = torch.randn((rows, DIM), generator=g, device=device, dtype=torch.float32)
    return torch.nn.functional.normalize(x, dim=1)


def queries_device(nq: int, device):
    """rows 0, 97, 194, ... of chunk 0 plus N(0, 0.1^2) noise, renormalised -- identical on every rank"""
    import torch
    base = corpus_chunk_device(0, CHUNK_ROWS, device)[torch.arange(nq, device=device) * 97 % CHUNK_ROWS]
    g = torch.Generator(device=device)
    g.manual_seed(QUERY_SEED)
    q = base + 0.1 * torch.randn((nq, DIM), generator=g, device=device, dtype=torch.float32)
    return torch.nn.functional.normalize(q, dim=1).contiguous()


def fill_shard(store, start: int, count: int, device):
    """rows [start, start + count) of the global corpus -> the local shard, generated on the device"""
    import torch
    done = 0
    while done < count:
        g_row = start + done
        chunk, off = divmod(g_row, CHUNK_ROWS)
        take = min(CHUNK_ROWS - off, count - done)
        x = corpus_chunk_device(chunk, CHUNK_ROWS, device)[off:off + take].contiguous()
        torch.cuda.synchronize(device)
        store.add_local_device(x.data_ptr(), take)
        done += take
        del x


# ------------------------------------------------------------------------------------------------
# answer check (outside every timed region): the bench proves its result at every N
# ------------------------------------------------------------------------------------------------
def result_digest(ids: np.ndarray, scores: np.ndarray) -> str:
    """sha256 over the returned ids (u64) and the BITS of the returned scores: equal answers <=> equal digests, so the
    N = 1, 2, 4, 8 lines of one corpus can be compared with each other and with the checker's own answer"""
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(ids, dtype=np.uint64).tobytes())
    h.update(np.ascontiguousarray(scores, dtype=np.float32).view(np.uint32).tobytes())
    return h.hexdigest()[:32]


def verify_topk(device, rows: int, q_dev, ids: np.ndarray, scores: np.ndarray, cand: int = 32):
    """Independent recomputation of the WHOLE answer on rank 0, none of this library's kernels involved:
      1. the corpus is regenerated chunk by chunk from its seeds, rounded to fp16 as the store keeps it, scored against
         the queries with a plain torch fp32 matmul (cosine with the stored rows' own norms), and the `cand` best rows
         per query are kept across all chunks -- a superset of the top-k with a wide margin;
      2. the CHECKER (oracle.cosine, the CPU restatement of hnsw_rs DistCosine + local.rs:86 -- test infrastructure, never
         timed, never on the product path) re-scores those rows with the reference's arithmetic and ranks them by
         (distance asc, id asc);
      3. ids and score BITS of the library's answer must equal that ranking's first k.
    Returns the fields of the line's "result_check" object."""
    import torch
    from oracle import cosine
    nq, k = ids.shape
    q = q_dev.to(torch.float32)
    best_s = torch.full((nq, cand), -float("inf"), device=device)
    best_i = torch.zeros((nq, cand), dtype=torch.int64, device=device)
    keep = {}
    n_chunks = (rows + CHUNK_ROWS - 1) // CHUNK_ROWS
    for c in range(n_chunks):
        take = min(CHUNK_ROWS, rows - c * CHUNK_ROWS)
        x = corpus_chunk_device(c, CHUNK_ROWS, device)[:take].to(torch.float16).to(torch.float32)
        s = (q @ x.T) / x.norm(dim=1).clamp_min(1e-30)[None, :]
        ts, ti = torch.topk(s, min(cand, take), dim=1)
        ms = torch.cat([best_s, ts], dim=1)
        mi = torch.cat([best_i, ti + c * CHUNK_ROWS], dim=1)
        best_s, sel = torch.topk(ms, cand, dim=1)
        best_i = torch.gather(mi, 1, sel)
        del x, s
    # the candidate rows themselves (regenerated once more, chunk by chunk, only the chunks that hold candidates)
    bi = best_i.cpu().numpy()
    need = np.unique(bi)
    rows_f32 = {}
    for c in np.unique(need // CHUNK_ROWS):
        x = corpus_chunk_device(int(c), CHUNK_ROWS, device).to(torch.float16).to(torch.float32)
        sel = need[need // CHUNK_ROWS == c]
        got = x[torch.from_numpy(sel - c * CHUNK_ROWS).to(device)].cpu().numpy()
        for r, v in zip(sel, got):
            rows_f32[int(r)] = v
        del x
    qh = q.cpu().numpy()
    ok_ids = ok_bits = margin_ok = True
    chk_ids = np.zeros((nq, k), np.uint64)
    chk_scores = np.zeros((nq, k), np.float32)
    for i in range(nq):
        cand_rows = np.sort(bi[i])                                       # ascending GLOBAL row: the oracle's tie rule
        sub = np.stack([rows_f32[int(r)] for r in cand_rows])            # (distance asc, position asc) is then (d, id)
        oi, os_, oc = cosine.exact_topk(sub, qh[i:i + 1], len(cand_rows))
        n_ok = int(oc[0])
        chk_ids[i] = cand_rows[oi[0][:k].astype(np.int64) - 1] + 1       # ids are 1-based (local.rs:63)
        chk_scores[i] = os_[0][:k]
        # the cut was wide enough: the worst kept candidate is clearly below the k-th exact score
        margin_ok &= n_ok < cand or bool(os_[0][n_ok - 1] < os_[0][k - 1] - 1e-4)
    ok_ids = bool((chk_ids == ids.astype(np.uint64)).all())
    ok_bits = bool((chk_scores.view(np.uint32) == np.ascontiguousarray(scores, np.float32).view(np.uint32)).all())
    return {"checker": f"torch fp32 matmul over the regenerated corpus -> top-{cand} rows per query -> oracle.cosine "
                       f"(DistCosine f64 fold) ranking; none of the library's kernels",
            "ids_equal": ok_ids, "score_bits_equal": ok_bits, "candidate_margin_ok": bool(margin_ok),
            "checker_digest": result_digest(chk_ids, chk_scores)}


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle/ is test infrastructure: only these baseline legs and the answer check may execute it)
# ------------------------------------------------------------------------------------------------
def corpus_rows_host(n: int) -> np.ndarray:
    rng = np.random.default_rng(CORPUS_SEED)
    x = rng.standard_normal((n, DIM), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def queries_host(x: np.ndarray, nq: int) -> np.ndarray:
    rng = np.random.default_rng(QUERY_SEED)
    q = x[(np.arange(nq) * 97) % len(x)] + 0.1 * rng.standard_normal((nq, DIM), dtype=np.float32)
    return (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)


def cpu_hnsw_baseline(steps: int, warmup: int, threads: int, rows: int = HNSW_SAMPLE_ROWS):
    """The reference's search as shipped (local.rs:48,76): HNSW M=16 efC=200, ef=32, cosine.
    Built over a bounded sample with all host threads (the build -- 2-7 ms per insert on one thread -- is not
    timed); each step = one 64-query batch on `threads` threads.  Returns (queries/sec, info)."""
    from oracle import cosine
    x = corpus_rows_host(rows)
    h = cosine.HnswOracle(DIM, seed=1)
    t0 = time.perf_counter()
    h.insert(x, threads=os.cpu_count() or 1)
    build_s = time.perf_counter() - t0
    q = queries_host(x, NQ)
    for _ in range(max(1, warmup)):
        h.search(q, TOPK, 32, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        ids, _, _ = h.search(q, TOPK, 32, threads)
    dt = time.perf_counter() - t0
    e_ids, _, _ = cosine.exact_topk(x, q, TOPK)
    recall = float(np.mean([len(set(ids[i]) & set(e_ids[i])) / TOPK for i in range(NQ)]))
    return NQ * steps / dt, dict(build_s=round(build_s, 1), recall_at_10=round(recall, 3), ms_per_step=dt / steps * 1e3, rows=rows)


def cpu_config1(n_rows: int = 1000, n_queries: int = 1000):
    """BASELINE.json configs[0] at its stated size: 1,000 unit-norm rows (seed 1234), 1,000 queries
    = rows + N(0, 0.1^2) noise (seed 4321), HNSW as memex builds it (M = 16, efC = 200, ef = 32), ONE thread (the
    reference's search is single-threaded behind one mutex), one query per call -> us/query and recall@10 vs exact"""
    from oracle import cosine
    x = corpus_rows_host(n_rows)
    rng = np.random.default_rng(QUERY_SEED)
    q = x[np.arange(n_queries) % n_rows] + 0.1 * rng.standard_normal((n_queries, DIM), dtype=np.float32)
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    h = cosine.HnswOracle(DIM, seed=1)
    h.insert(x, threads=1)
    for i in range(50):
        h.search(q[i:i + 1], TOPK, 32, 1)
    got = np.zeros((n_queries, TOPK), np.uint64)
    t0 = time.perf_counter()
    for i in range(n_queries):
        got[i] = h.search(q[i:i + 1], TOPK, 32, 1)[0][0]
    dt = time.perf_counter() - t0
    e_ids, _, _ = cosine.exact_topk(x, q, TOPK)
    recall = float(np.mean([len(set(got[i]) & set(e_ids[i])) / TOPK for i in range(n_queries)]))
    return {"workload": f"{n_rows}x{DIM} fp32, HNSW (M=16, efC=200, ef=32) top-{TOPK}, one query per call, 1 thread "
                        f"(BASELINE.json configs[0] at its stated size; synthetic unit-norm rows, SURVEY.md 8(d))",
            "us_per_query": dt / n_queries * 1e6, "queries_per_s": n_queries / dt, "recall_at_10": round(recall, 4),
            "cores": 1, "kind": "port", "queries": n_queries}, x, q, e_ids


def gpu_config1(device_index: int, x: np.ndarray, q: np.ndarray, e_ids: np.ndarray):
    """the same 1,000-row corpus and queries through the drop-in store (host buffers, one query per call)"""
    from memex_b200.storage import B200Store
    st = B200Store.new("/tmp/mx_bench_config1", dim=DIM, dtype="f32", device=device_index)
    st.add_matrix(x)
    for i in range(20):
        st.search_matrix(q[i:i + 1], TOPK)
    got = np.zeros((len(q), TOPK), np.uint64)
    t0 = time.perf_counter()
    for i in range(len(q)):
        got[i] = st.search_matrix(q[i:i + 1], TOPK)[0][0]
    dt = time.perf_counter() - t0
    st.close()
    return {"us_per_query": dt / len(q) * 1e6, "queries_per_s": len(q) / dt, "ids_equal_exact_oracle": bool((got == e_ids).all()),
            "recall_at_10": 1.0 if (got == e_ids).all() else float(np.mean([len(set(got[i]) & set(e_ids[i])) / TOPK for i in range(len(q))])),
            "note": "latency-bound at this size (launch + two small kernels + one H2D / D2H pair per call)"}


def cpu_exact_baseline():
    """all-core exact brute force with the reference's distance arithmetic (oracle/cosine_oracle.c)
    over a bounded sample of the rows; a scan is linear in rows, so q/s at 10 M = q/s(sample) * sample / 10 M"""
    from oracle import cosine
    x = corpus_rows_host(EXACT_SAMPLE_ROWS)
    q = queries_host(x, NQ)
    cosine.exact_topk(x[:10000], q, TOPK)
    t0 = time.perf_counter()
    cosine.exact_topk(x, q, TOPK)
    dt = time.perf_counter() - t0
    return NQ / dt, cosine.num_threads()


def cpu_embed_baseline(batch: int = 16, seq: int = 256):
    """HF BertModel fp32 on torch-CPU (the libtorch kernels tch dispatches to), all host threads"""
    import torch
    from oracle import encoder as enc_oracle
    cfg = enc_oracle.MINILM_L6
    w = enc_oracle.make_weights(cfg, seed=3)
    ids, lens = enc_oracle.make_inputs(cfg, batch, seq, seed=7)
    enc_oracle.hf_encode(cfg, w, ids[:2], lens[:2])
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < 8.0 or reps < 2:
        enc_oracle.hf_encode(cfg, w, ids, lens)
        reps += 1
    dt = time.perf_counter() - t0
    return batch * reps / dt, torch.get_num_threads(), f"{reps} x (B={batch}, S={seq}) MiniLM-L6 fp32, HF BertModel on torch-CPU"


# ------------------------------------------------------------------------------------------------
def workload_config(rows: int, n_gpus: int):
    return {"workload": f"{rows // 1_000_000}Mx{DIM} fp16 corpus, cosine top-{TOPK}, {NQ}-query batches"
            if rows % 1_000_000 == 0 else f"{rows}x{DIM} fp16 corpus, cosine top-{TOPK}, {NQ}-query batches",
            "rows": rows, "dim": DIM, "k": TOPK, "queries_per_step": NQ, "corpus_dtype": "f16",
            "parallelism": f"row-sharded x{n_gpus}, one all-gather of per-shard top-k" if n_gpus > 1 else "single GPU",
            "l2": "inputs larger than L2 (shard >= 0.96 GB vs 126 MB L2); no flush needed"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    v, info = cpu_hnsw_baseline(args.steps, args.warmup, threads)
    sample = (f"HNSW restatement (M=16, efC=200, ef=32) built over a {HNSW_SAMPLE_ROWS}-row sample of the corpus "
              f"(build {info['build_s']} s, untimed); {args.steps} steps x {NQ} queries on {threads} threads; "
              f"recall@10 vs exact = {info['recall_at_10']} (the GPU arm is exact: 1.0); HNSW search cost grows "
              f"with log N, so this over-states the reference's q/s at 10 M rows, and memex serialises searches behind "
              f"one mutex (storage/mod.rs:85-92), which this arm does not")
    cfg = workload_config(args.rows, args.gpus)
    # the workload is the 10 M-row one; what this arm can INDEX within minutes is a bounded sample of it -- said in the
    # config itself, not only in the prose (HNSW build: ~4 k rows / s on all threads, 10 M rows = ~40 min)
    cfg["reference_rows_indexed"] = info["rows"]
    cfg["reference_note"] = (f"approximate HNSW over a {info['rows']}-row sample of the {args.rows}-row corpus, "
                             f"recall@10 {info['recall_at_10']}; the GPU arm is exact over all {args.rows} rows")
    c1, _, _, _ = cpu_config1()
    line = {"impl": "reference", "metric": "queries/sec", "value": v, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": info["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "config1": c1,
            "cpu_baseline": {"value": v, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bench_single_query(device, steps: int, warmup: int, pk):
    """config 2: 1 M x 384 fp32, one query per call, top-10"""
    import torch
    from memex_b200 import capi
    from memex_b200.sharded import ShardedStore
    rows = 1_000_000
    L = capi.lib()
    st = ShardedStore("/tmp/mx_bench_single", DIM, rows, dtype="f32", device=device.index)
    fill_shard(st, 0, rows, device)
    q = queries_device(64, device)
    for i in range(warmup):
        st.search_device(q[i % 64:i % 64 + 1], TOPK)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        st.search_device(q[i % 64:i % 64 + 1], TOPK)
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / steps
    # second pass with the library's per-kernel events on: the scan / other split (see run_ours)
    L.mx_store_set_timing(st.local.handle, 1)
    for i in range(steps):
        st.search_device(q[i % 64:i % 64 + 1], TOPK)
    torch.cuda.synchronize(device)
    scan_ms, scan_n, oth_ms, oth_n = C.c_double(), C.c_uint64(), C.c_double(), C.c_uint64()
    L.mx_store_get_timing(st.local.handle, C.byref(scan_ms), C.byref(scan_n), C.byref(oth_ms), C.byref(oth_n))
    L.mx_store_set_timing(st.local.handle, 0)
    algo = rows * DIM * 4 + rows * 4
    per = scan_ms.value / max(1, scan_n.value)
    ach = algo / (per * 1e-3) / 1e9
    # end to end with host buffers
    qh = q.cpu().numpy()
    st.local.search_matrix(qh[:1], TOPK)
    t0 = time.perf_counter()
    for i in range(steps):
        st.local.search_matrix(qh[i % 64:i % 64 + 1], TOPK)
    e2e = steps / (time.perf_counter() - t0)
    st.close()
    return {"workload": "1Mx384 fp32 corpus, cosine top-10, one query per call", "value": 1e3 / ms, "unit": "queries/s",
            "ms_per_query": ms, "e2e": {"value": e2e, "unit": "queries/s", "h2d_bytes_per_step": DIM * 4,
                                        "d2h_bytes_per_step": TOPK * 12 + 4},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                         "traffic": ncu_traffic("scan_stream_f32_q1"), "kernel": "scan_stream_kernel<float>",
                         "kernel_ms": per, "other_kernels_ms_per_step": oth_ms.value / max(1, steps)}}


def random_bert_weights(Lyr, H, F, vocab, max_pos, seed=3):
    """seeded random weights at the true shapes (no checkpoint on the box); generated here, not via oracle/"""
    rng = np.random.default_rng(seed)
    w = {}

    def lin(name, o, i):
        w[name + ".weight"] = (rng.standard_normal((o, i)) * (1.5 / np.sqrt(i))).astype(np.float32)
        w[name + ".bias"] = (0.1 * rng.standard_normal(o)).astype(np.float32)

    def ln(name):
        w[name + ".weight"] = (1 + 0.1 * rng.standard_normal(H)).astype(np.float32)
        w[name + ".bias"] = (0.1 * rng.standard_normal(H)).astype(np.float32)

    w["embeddings.word_embeddings.weight"] = (0.5 * rng.standard_normal((vocab, H))).astype(np.float32)
    w["embeddings.position_embeddings.weight"] = (0.5 * rng.standard_normal((max_pos, H))).astype(np.float32)
    w["embeddings.token_type_embeddings.weight"] = (0.5 * rng.standard_normal((2, H))).astype(np.float32)
    ln("embeddings.LayerNorm")
    for i in range(Lyr):
        p = f"encoder.layer.{i}."
        for n in ("query", "key", "value"):
            lin(p + "attention.self." + n, H, H)
        lin(p + "attention.output.dense", H, H)
        ln(p + "attention.output.LayerNorm")
        lin(p + "intermediate.dense", F, H)
        lin(p + "output.dense", H, F)
        ln(p + "output.LayerNorm")
    return w


def bench_embed(device, steps: int, warmup: int, pk, cpu: bool, extras: bool = True, parity: bool = True):
    """config 3: batch-256 segment embedding, MiniLM-L6, S = 256, bf16 activations on tcgen05"""
    import torch
    from memex_b200 import capi
    from memex_b200.embedding import Architecture, B200Encoder
    Lyr, H, heads, F, vocab, max_pos = 6, 384, 12, 1536, 30522, 512
    w = random_bert_weights(Lyr, H, F, vocab, max_pos)
    B, S = 256, 256
    arch = Architecture(Lyr, H, heads, F, vocab, max_pos)
    enc = B200Encoder(arch, w, precision="bf16", device=device.index, max_tokens=B * S)
    L = capi.lib()
    ids_h = torch.from_numpy(np.random.default_rng(7).integers(1000, 30000, size=(B, S)).astype(np.int32)).pin_memory()
    lens = np.full(B, S, dtype=np.int32)
    ids_d = ids_h.to(device)
    out_d = torch.zeros((B, H), dtype=torch.float32, device=device)
    st = torch.cuda.current_stream(device).cuda_stream or 1   # 0 would mean "the embedder's own stream"; 1 = cudaStreamLegacy

    def step():
        rc = L.mx_embedder_encode_device(enc.handle, ids_d.data_ptr(), lens.ctypes.data, B, S, out_d.data_ptr(), st)
        assert rc == 0, L.mx_last_error(enc.handle)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize(device)

    def timed(fn):
        # every timed region starts from the same power state (see IDLE_BEFORE_REGION_S): idle, one untimed step, the region
        torch.cuda.synchronize(device)
        time.sleep(IDLE_BEFORE_REGION_S)
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(device)
        return e0.elapsed_time(e1) / steps

    # two passes of `steps` steps: the first is the throughput (nothing but the kernels on the stream); the second runs with
    # the library's per-kernel CUDA events on (two event records around each of the 32 launches of a step, which cost
    # 2-3 % of the step) and gives the GEMM / other split the roofline is computed from
    l0 = L.mx_launch_count()
    ms = timed(step)
    launches = L.mx_launch_count() - l0
    # SM clock and board power under THIS kernel mix (a tensor-heavy step runs into the power cap long before the tensor pipe
    # is full: MEASURED_PEAKS.json's cuBLAS figure was taken at ~1335 MHz): ~0.4 s of back-to-back steps, sampled at 20 ms
    sampler = ClockSampler(device.index)
    sampler.start()
    time.sleep(0.2)
    t_a = sampler.mark()
    n_sus = max(1, int(400.0 / max(ms, 1e-3)))
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es0.record()
    for _ in range(n_sus):
        step()
    es1.record()
    torch.cuda.synchronize(device)
    t_b = sampler.mark()
    sampler.stop()
    clocks = sampler.summary(t_a + 0.05, t_b - 0.02)
    sustained = {"value": B * n_sus / (es0.elapsed_time(es1) * 1e-3), "unit": "segments/s", "steps": n_sus,
                 "seconds": es0.elapsed_time(es1) * 1e-3, "clocks": clocks,
                 "note": "steps back to back for ~0.4 s: the board sits at its power cap"}
    L.mx_embedder_set_timing(enc.handle, 1)
    ms_events = timed(step)
    g_ms, g_n, o_ms, o_n = C.c_double(), C.c_uint64(), C.c_double(), C.c_uint64()
    L.mx_embedder_get_timing(enc.handle, C.byref(g_ms), C.byref(g_n), C.byref(o_ms), C.byref(o_n))
    L.mx_embedder_set_timing(enc.handle, 0)
    T = B * S
    gemm_flops = Lyr * T * 24 * H * H                      # QKV + out + FFN up/down, 2 flop / MAC
    att_flops = Lyr * T * 4 * S * H
    gemm_tf = gemm_flops * steps / (g_ms.value * 1e-3) / 1e12 if g_ms.value > 0 else 0.0
    step_tf = (gemm_flops + att_flops) / (ms * 1e-3) / 1e12
    # end to end: host ids -> C ABI -> host embeddings
    ids_np = ids_h.numpy()
    enc.encode_ids(ids_np, lens)
    t0 = time.perf_counter()
    n_e2e = max(3, steps // 2)
    for _ in range(n_e2e):
        enc.encode_ids(ids_np, lens)
    e2e = B * n_e2e / (time.perf_counter() - t0)
    res = {"workload": "batch-256 segment embedding, MiniLM-L6, seq_len 256, bf16 activations (seeded random weights)",
           "value": B * 1e3 / ms, "unit": "segments/s", "ms_per_step": ms, "dtype": "bf16",
           "e2e": {"value": e2e, "unit": "segments/s", "h2d_bytes_per_step": B * S * 4 + B * 4, "d2h_bytes_per_step": B * H * 4},
           "gpu_launches_per_step": launches // max(1, steps), "clocks": clocks, "sustained": sustained,
           "idle_before_region_s": IDLE_BEFORE_REGION_S,
           # the timed region is tens of milliseconds, not seconds: the BURST bf16 figure is the apt denominator
           "roofline": {"bound": "tensor", "achieved": gemm_tf, "peak": pk["tf_burst"], "unit": "TFLOP/s",
                        "frac": gemm_tf / pk["tf_burst"], "traffic": ncu_traffic("gemm_tc"),
                        "kernel": "gemm_tc_kernel (4 launches / layer)", "gemm_ms_per_step": g_ms.value / steps,
                        "other_ms_per_step": o_ms.value / steps, "ms_per_step_with_kernel_events": ms_events,
                        "whole_step_tflops": step_tf,
                        "whole_step_frac": step_tf / pk["tf_burst"], "peak_kind": "burst bf16 (a ~25 ms region), " + pk["src"],
                        "frac_of_sustained": gemm_tf / pk["tf_sustained"],
                        "whole_step_frac_of_sustained": step_tf / pk["tf_sustained"]}}
    # parity in the run itself (outside the timed regions): 16 rows of the timed batch against the checker (HF BertModel
    # fp32 on torch-CPU over the same weights -- oracle/encoder.py, test infrastructure)
    if parity:
        from oracle import encoder as enc_oracle
        ref = enc_oracle.hf_encode(enc_oracle.EncoderConfig(layers=Lyr, hidden=H, heads=heads, ffn=F, vocab=vocab, max_pos=max_pos),
                                   w, ids_np[:16], lens[:16])
        got = enc.encode_ids(ids_np, lens)[:16]
        cos = (got * ref).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref, axis=1))
        res["parity"] = {"vs": "oracle.encoder.hf_encode (HF BertModel fp32, torch-CPU), 16 rows of the timed batch",
                         "min_cos": float(cos.min()), "max_abs": float(np.abs(got - ref).max()),
                         "gate": "cos >= 1 - 1e-4 (BASELINE.md section 3)", "ok": bool(cos.min() >= 1 - 1e-4)}
    if not extras:
        enc.close()
        return res
    # the same batch with ragged lengths ~ U[16, 256]: packed layout, the padding rows are dropped
    # before the first GEMM -- flops are counted on REAL tokens
    lens_r = np.random.default_rng(8).integers(16, S + 1, size=B).astype(np.int32)

    def step_ragged():
        rc = L.mx_embedder_encode_device(enc.handle, ids_d.data_ptr(), lens_r.ctypes.data, B, S, out_d.data_ptr(), st)
        assert rc == 0, L.mx_last_error(enc.handle)

    for _ in range(3):
        step_ragged()
    torch.cuda.synchronize(device)
    ms_r = timed(step_ragged)
    real = int(lens_r.sum())
    flops_r = Lyr * (real * 24 * H * H + 4 * H * int((lens_r.astype(np.int64) ** 2).sum()))
    res["ragged"] = {"workload": "same batch, lengths ~ U[16, 256]; packed layout (padding rows dropped before the first GEMM)", "value": B * 1e3 / ms_r, "unit": "segments/s",
                     "ms_per_step": ms_r, "real_tokens": real, "token_slots": T,
                     "real_token_tflops": flops_r / (ms_r * 1e-3) / 1e12}
    enc.close()
    # memex's DEFAULT model (embedding.rs:64-72): all-MiniLM-L12-v2, whose sentence_bert_config truncates to 128 tokens
    enc12 = B200Encoder(Architecture(12, H, heads, F, vocab, max_pos), random_bert_weights(12, H, F, vocab, max_pos, seed=4),
                        precision="bf16", device=device.index, max_tokens=B * 128)
    ids12 = ids_d[:, :128].contiguous()
    lens12 = np.full(B, 128, dtype=np.int32)

    def step12():
        rc = L.mx_embedder_encode_device(enc12.handle, ids12.data_ptr(), lens12.ctypes.data, B, 128, out_d.data_ptr(), st)
        assert rc == 0, L.mx_last_error(enc12.handle)

    for _ in range(3):
        step12()
    torch.cuda.synchronize(device)
    ms12 = timed(step12)
    f12 = 12 * B * 128 * (24 * H * H + 4 * 128 * H)
    res["default_model"] = {"workload": "all-MiniLM-L12-v2 shape (memex's default), B = 256, S = 128, bf16", "value": B * 1e3 / ms12,
                            "unit": "segments/s", "ms_per_step": ms12, "tflops": f12 / (ms12 * 1e-3) / 1e12}
    enc12.close()
    if cpu:
        v, cores, sample = cpu_embed_baseline()
        res["cpu_baseline"] = {"value": v, "unit": "segments/s", "cores": cores, "kind": "port", "sample": sample}
    return res


def bench_ingest(device, pk, rows: int, seconds: float, rank: int = 0, world: int = 1, group=None, scan_sms: int = 104):
    """config 5 (BASELINE.json configs[4]): a 50 M x 768 fp16 corpus row-sharded over 8 GPUs (6.25 M rows = 9.6 GB per GPU)
    searched with 64-query batches WHILE BERT-base-shape segments (L = 12, H = 768, S = 512) are embedded and appended to the
    local shard -- the worker's `embed + add_vectors` (reference lib/worker/src/tasks.rs:15-59) next to the API's search
    (lib/api/src/endpoints/collections/handlers.rs:61-81).

    Both roles are persistent kernels that each fill every SM when alone, so the SMs are PARTITIONED while they run
    together: the scan grid gets `scan_sms` CTAs (mx_store_set_sm_limit; it stays HBM-bound far below the full chip), the
    embedder's grids the rest (mx_embedder_set_sm_limit).  One host thread per role and per rank, each with its own
    stream; the ingester appends with mx_store_add_device_stream (the rows join the searchable range -- the committed-rows
    watermark -- when the append has completed; the store is created with the capacity it will reach, so the matrix never
    moves under a running scan).  At N ranks every rank runs both roles on its own shard; a search is the sharded one
    (peer-memory exchange of the per-shard top-k, memex_b200/sharded.py) and new rows get round-robin global ids
    (local_row * N + rank + 1), so ingest needs no exchange.  Three windows: search alone and ingest alone (whole GPU
    each), then both (partitioned).  Wall clock over `seconds` per window: throughput figures, not kernel times."""
    import torch
    import torch.distributed as dist
    from memex_b200 import capi
    from memex_b200.embedding import Architecture, B200Encoder
    from memex_b200.sharded import ShardedStore
    L = capi.lib()
    dim, Lyr, heads, F, vocab, max_pos = 768, 12, 12, 3072, 30522, 512
    B, S = 16, 512
    headroom = 400_000
    st = ShardedStore(f"/tmp/mx_bench_ingest_{rank}", dim, (rows + headroom) * world, dtype="f16", device=device.index,
                      rank=rank, world=world, group=group, id_stride=world)
    g = torch.Generator(device=device)
    done = 0
    while done < rows:
        take = min(250_000, rows - done)
        g.manual_seed(CORPUS_SEED + 1000 * rank + done // 250_000)
        x = torch.nn.functional.normalize(torch.randn((take, dim), generator=g, device=device), dim=1).contiguous()
        torch.cuda.synchronize(device)
        st.add_local_device(x.data_ptr(), take)
        done += take
        del x
    g.manual_seed(QUERY_SEED)
    q = torch.nn.functional.normalize(torch.randn((NQ, dim), generator=g, device=device), dim=1).contiguous()
    enc = B200Encoder(Architecture(Lyr, dim, heads, F, vocab, max_pos), random_bert_weights(Lyr, dim, F, vocab, max_pos),
                      precision="bf16", device=device.index, max_tokens=B * S)
    ids_d = torch.from_numpy(np.random.default_rng(7 + rank).integers(1000, 30000, size=(B, S)).astype(np.int32)).to(device)
    lens = np.full(B, S, dtype=np.int32)
    s_search, s_ingest = torch.cuda.Stream(device), torch.cuda.Stream(device)
    out_d = torch.zeros((B, dim), dtype=torch.float32, device=device)
    counts = {"queries": 0, "segments": 0, "rows_scanned": 0}
    n_sms = torch.cuda.get_device_properties(device).multi_processor_count

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier(group=group)

    def searcher(n_batches):
        # every rank issues the SAME number of batches: the sharded search is a collective step (SPMD)
        torch.cuda.set_device(device)
        with torch.cuda.stream(s_search):
            for _ in range(n_batches):
                n_now = len(st)
                st.search_device(q, TOPK)
                s_search.synchronize()
                counts["queries"] += NQ
                counts["rows_scanned"] += n_now

    diag_mode = [0]   # MX_INGEST_DIAG: which part of the ingest loop runs (measurement aid; 0 = all of it)
    mm_a = torch.randn((4096, 4096), device=device, dtype=torch.bfloat16)

    def ingester(stop):
        torch.cuda.set_device(device)
        first = C.c_uint64()
        while not stop.is_set():
            if diag_mode[0] == 3:      # no library call at all: plain torch matmuls on the ingest stream
                with torch.cuda.stream(s_ingest):
                    for _ in range(8):
                        torch.matmul(mm_a, mm_a)
                s_ingest.synchronize()
                counts["segments"] += B
                continue
            if diag_mode[0] == 2:      # only the append
                rc = L.mx_store_add_device_stream(st.local.handle, out_d.data_ptr(), B, C.byref(first), s_ingest.cuda_stream)
                assert rc == 0, L.mx_last_error(st.local.handle)
                counts["segments"] += B
                continue
            rc = L.mx_embedder_encode_device(enc.handle, ids_d.data_ptr(), lens.ctypes.data, B, S, out_d.data_ptr(),
                                             s_ingest.cuda_stream)
            assert rc == 0, L.mx_last_error(enc.handle)
            if diag_mode[0] == 1:      # only the forward pass
                s_ingest.synchronize()
                counts["segments"] += B
                continue
            rc = L.mx_store_add_device_stream(st.local.handle, out_d.data_ptr(), B, C.byref(first), s_ingest.cuda_stream)
            assert rc == 0, L.mx_last_error(st.local.handle)
            counts["segments"] += B

    def window(search: bool, ingest: bool, n_batches: int):
        for k in counts:
            counts[k] = 0
        stop = threading.Event()
        barrier()
        t0 = time.perf_counter()
        ti = threading.Thread(target=ingester, args=(stop,)) if ingest else None
        if ti:
            ti.start()
        if search:
            searcher(n_batches)          # fixed number of batches on every rank
        else:
            time.sleep(seconds)
        t_search = time.perf_counter() - t0
        stop.set()
        if ti:
            ti.join()
        torch.cuda.synchronize(device)
        dt = time.perf_counter() - t0
        res = {"queries": counts["queries"] / t_search if search else 0.0, "segments": counts["segments"] / dt,
               "rows_scanned": counts["rows_scanned"] / t_search if search else 0.0}
        t = torch.tensor([res["queries"], res["segments"], res["rows_scanned"]], dtype=torch.float64, device=device)
        if world > 1:
            # queries/s: every rank answers the same batches -> the slowest rank's rate; segments and bytes add up
            qmin = t[:1].clone()
            dist.all_reduce(qmin, op=dist.ReduceOp.MIN, group=group)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            t[0] = qmin[0]
        return {"queries": t[0].item(), "segments": t[1].item(), "rows_scanned": t[2].item()}

    for _ in range(3):
        st.search_device(q, TOPK)
    torch.cuda.synchronize(device)
    # calibrate the batch count of a window from the scan-alone rate (same number on every rank)
    barrier()
    t0 = time.perf_counter()
    searcher(10)
    per_batch = torch.tensor([(time.perf_counter() - t0) / 10], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(per_batch, op=dist.ReduceOp.MAX, group=group)
    n_alone = max(10, int(seconds / per_batch.item()))
    alone_s = window(True, False, n_alone)
    alone_i = window(False, True, 0)
    # together: partition the SMs -- green contexts confine EVERY kernel of a role to its share; the grid budgets size the
    # persistent kernels for it
    part = C.c_void_p()
    partition = "CUDA green contexts"
    if L.mx_sm_partition_create(device.index, scan_sms, C.byref(part)) == capi.OK:
        scan_sms = int(L.mx_sm_partition_sms(part, 0))
        emb_sms = int(L.mx_sm_partition_sms(part, 1))
        s_search = torch.cuda.ExternalStream(L.mx_sm_partition_stream(part, 0), device=device)
        s_ingest = torch.cuda.ExternalStream(L.mx_sm_partition_stream(part, 1), device=device)
    else:
        part = None
        emb_sms = n_sms - scan_sms
        partition = "grid budgets only (no green contexts on this driver: " + (L.mx_last_error(None) or b"").decode() + ")"
    L.mx_store_set_sm_limit(st.local.handle, scan_sms)
    L.mx_embedder_set_sm_limit(enc.handle, emb_sms)
    with torch.cuda.stream(s_search):
        for _ in range(3):
            st.search_device(q, TOPK)
    torch.cuda.synchronize(device)
    both = window(True, True, max(10, n_alone * 2 // 3))
    diag = {}
    if os.environ.get("MX_INGEST_DIAG"):
        for mode, name in ((1, "forward pass only"), (2, "append only"), (3, "torch matmuls only (no library call)")):
            diag_mode[0] = mode
            r = window(True, True, max(10, n_alone // 3))
            diag[name] = {"queries_per_s": r["queries"], "ingest_iterations_per_s": r["segments"] / B}
        diag_mode[0] = 0
    rows_end = torch.tensor([len(st)], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(rows_end, op=dist.ReduceOp.SUM, group=group)
    # the freshly ingested rows are searchable: the last appended segment must come back as its own nearest neighbour
    torch.cuda.synchronize(device)
    probe = out_d[-1:].clone()
    ids_p, scores_p, _ = st.search_device(probe.contiguous(), 1) if world == 1 else (None, None, None)   # default stream
    found = None
    if world == 1:
        torch.cuda.synchronize(device)
        found = bool(abs(float(scores_p[0, 0]) - 1.0) < 1e-3)
    seg_flops = Lyr * S * (24 * dim * dim + 4 * S * dim)
    bytes_per_row = dim * 2 + 4
    res = {"workload": f"{rows * world}x{dim} fp16 corpus over {world} GPU(s) ({rows} rows = {rows * bytes_per_row / 1e9:.1f} GB per GPU; "
                       f"config 5 is 50Mx768 over 8), {NQ}-query batches, WHILE BERT-base-shape segments (L=12, H=768, S={S}, "
                       f"B={B}) are embedded and appended to the local shards",
           "n_gpus": world, "rows_at_end": int(rows_end.item()),
           "sm_partition_when_concurrent": {"scan_sms": scan_sms, "embedder_sms": emb_sms, "how": partition},
           "search_alone": {"queries_per_s": alone_s["queries"], "hbm_gbs_summed": alone_s["rows_scanned"] * bytes_per_row / 1e9},
           "ingest_alone": {"segments_per_s": alone_i["segments"], "tflops_summed": alone_i["segments"] * seg_flops / 1e12},
           "concurrent": {"queries_per_s": both["queries"], "segments_per_s": both["segments"],
                          "hbm_gbs_summed": both["rows_scanned"] * bytes_per_row / 1e9,
                          "tflops_summed": both["segments"] * seg_flops / 1e12},
           "concurrent_vs_alone": {"search": both["queries"] / max(alone_s["queries"], 1e-9),
                                   "ingest": both["segments"] / max(alone_i["segments"], 1e-9)},
           "peaks_per_gpu": {"hbm_gbs": pk["hbm"], "tflops_burst": pk["tf_burst"]},
           "fresh_rows_searchable": found, **({"diag": diag} if diag else {}),
           "timing": f"wall clock, ~{seconds} s windows, one host thread per role and rank, stream sync per batch (throughput, not kernel time)"}
    enc.close()
    st.close()
    if part is not None:
        torch.cuda.synchronize(device)
        L.mx_sm_partition_destroy(part)
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    from memex_b200 import capi
    from memex_b200.sharded import ShardedStore

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    L = capi.lib()
    if L.mx_device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: libmemex_b200 has no CPU path")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    group = None
    saved_stdout = None
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner to fd 1 when the first communicator comes up
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=device)
        group = dist.group.WORLD
    pk = peaks()

    store = ShardedStore(f"/tmp/mx_bench_{rank}", DIM, args.rows, dtype="f16", device=local, rank=rank, world=world,
                         group=group)
    fill_shard(store, store.plan.start(rank), store.plan.count(rank), device)
    q_dev = queries_device(NQ, device)
    torch.cuda.synchronize(device)

    def barrier():
        if world > 1:
            dist.barrier(group=group)
        torch.cuda.synchronize(device)

    def cool():
        # every timed region starts from the same power state: see the note at the end-to-end region below
        barrier()
        time.sleep(IDLE_BEFORE_REGION_S)
        barrier()

    # ---- device-resident: queries already in HBM ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    cool()                      # filling the shard was seconds of full-power work
    for _ in range(args.warmup):
        store.search_device(q_dev, TOPK)
    # two passes of `steps` steps: the first is the throughput -- nothing but the kernels on the stream (the library's
    # per-kernel CUDA events sit between the launches, cost ~1 us each and keep a dependent launch from being scheduled
    # early); the second runs with those events on and gives the scan / other split the roofline is computed from
    l0 = L.mx_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = sampler.mark()
    e0.record()
    for _ in range(args.steps):
        ids_d, scores_d, counts_d = store.search_device(q_dev, TOPK)
    e1.record()
    barrier()
    t_end = sampler.mark()
    ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    launches = torch.tensor([L.mx_launch_count() - l0], dtype=torch.int64, device=device)
    cool()
    for _ in range(args.warmup):
        store.search_device(q_dev, TOPK)
    barrier()
    L.mx_store_set_timing(store.local.handle, 1)
    for _ in range(args.steps):
        store.search_device(q_dev, TOPK)
    barrier()
    scan_ms, scan_n, oth_ms, oth_n = C.c_double(), C.c_uint64(), C.c_double(), C.c_uint64()
    L.mx_store_get_timing(store.local.handle, C.byref(scan_ms), C.byref(scan_n), C.byref(oth_ms), C.byref(oth_n))
    L.mx_store_set_timing(store.local.handle, 0)
    scan_per = torch.tensor([scan_ms.value / max(1, scan_n.value)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(launches, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(scan_per, op=dist.ReduceOp.MAX, group=group)
    ms_step = ms_total.item() / args.steps
    ids_dev_result = ids_d.cpu().numpy().copy()
    scores_dev_result = scores_d.cpu().numpy().copy()

    # ---- end to end: host queries in, host results out, through the public API ----
    # The board power-caps this kernel after ~40 ms of back-to-back steps (scripts/sustain_probe.py: 6.7 TB/s for the first
    # ~40 steps, then 5.7-5.9 at ~950 W with sw_power_cap; a plain device copy draws 740 W at the same bandwidth).  The
    # device-timed region above starts after 5 warm-up steps, inside that burst window; this one must start from the same
    # power state to be comparable, so the GPU idles for IDLE_BEFORE_REGION_S first.  `sustained` below is the other regime.
    q_host = q_dev.cpu().numpy() if rank == 0 else None
    # Two forms of the host-buffer call.  (1) one search in flight: the blocking call, every step waits for the answer of
    # the one before it -- the device idles while the host stages and launches.  (2) two in flight: the call split into
    # submit / collect (mx_*_search_submit / _collect), step i + 1 is staged, copied in and enqueued while step i runs;
    # every step's H2D and D2H are still inside the timed region.  (2) is the reported e2e -- the reference arm serves its
    # queries from all host threads at once, this is the same freedom on this side -- and (1) rides along.
    cool()
    for _ in range(2):
        store.search(q_host, TOPK, nq=NQ)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ids_1, scores_1, counts_1 = store.search(q_host, TOPK, nq=NQ)
    barrier()
    e2e_one_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    cool()
    for _ in range(2):
        store.search_collect(store.search_submit(q_host, TOPK, nq=NQ))
    barrier()
    t_e2e_begin = sampler.mark()
    t0 = time.perf_counter()
    prev = None
    for _ in range(args.steps):
        tk = store.search_submit(q_host, TOPK, nq=NQ)
        if prev is not None:
            ids_h, scores_h, counts_h = store.search_collect(prev)
        prev = tk
    ids_h, scores_h, counts_h = store.search_collect(prev)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    t_e2e_end = sampler.mark()
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(e2e_one_s, op=dist.ReduceOp.MAX, group=group)
    assert (ids_1 == ids_h).all() and (scores_1.view(np.uint32) == scores_h.view(np.uint32)).all(), "blocking and split calls disagree"

    # ---- sustained: ~SUSTAINED_S of back-to-back device-resident steps, timed in windows of 20 (max over ranks) ----
    sustained = None
    if not args.skip_extras:
        per = 20
        n_win = max(3, int(np.ceil(SUSTAINED_S / (per * ms_step * 1.25e-3))))
        cool()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_win + 1)]
        barrier()
        t_s0 = sampler.mark()
        evs[0].record()
        for w in range(n_win):
            for _ in range(per):
                store.search_device(q_dev, TOPK)
            evs[w + 1].record()
            if w >= 2:
                evs[w - 1].synchronize()     # at most two windows queued ahead of the device
        barrier()
        t_s1 = sampler.mark()
        win_ms = torch.tensor([evs[w].elapsed_time(evs[w + 1]) for w in range(n_win)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(win_ms, op=dist.ReduceOp.MAX, group=group)
        win_ms = win_ms.cpu().numpy()
        tail = win_ms[-max(1, n_win // 3):]
        sustained = {"seconds": float(win_ms.sum() / 1e3), "steps": per * n_win,
                     "value": NQ * per * n_win / (win_ms.sum() / 1e3), "unit": "queries/s",
                     "first_window": NQ * per / (win_ms[0] / 1e3), "last_third": NQ * per / (tail.mean() / 1e3),
                     "note": "device-resident steps back to back; the board settles at its power cap after ~40 ms"}
        if rank == 0:
            sustained["clocks"] = sampler.summary(t_s0, t_s1)
    if rank == 0:
        sampler.stop()
    assert (ids_h.astype(np.int64) == ids_dev_result.astype(np.int64)).all(), "host and device paths disagree"
    assert (counts_h == TOPK).all() and (np.diff(scores_h, axis=1) <= 0).all()
    digest = result_digest(ids_h, scores_h)
    assert digest == result_digest(ids_dev_result, scores_dev_result), "host and device paths disagree (score bits)"
    # the answer itself, proven at every N (outside the timed regions): every rank's copy must be THE SAME answer, and
    # rank 0 recomputes it from the corpus seeds without any of this library's kernels
    result_check = None
    if world > 1:
        dg = torch.tensor(list(bytes.fromhex(digest)), dtype=torch.uint8, device=device)
        all_dg = [torch.empty_like(dg) for _ in range(world)]
        dist.all_gather(all_dg, dg, group=group)
        same_on_all_ranks = all(bool((d == dg).all()) for d in all_dg)
    else:
        same_on_all_ranks = True
    if rank == 0 and not args.skip_check:
        result_check = verify_topk(device, args.rows, q_dev, ids_h, scores_h)
        result_check["same_on_all_ranks"] = same_on_all_ranks
        result_check["ok"] = bool(result_check["ids_equal"] and result_check["score_bits_equal"] and same_on_all_ranks and
                                  result_check["checker_digest"] == digest)

    rows_local = store.plan.count(0)    # the largest shard
    elem = 2
    algo_bytes = rows_local * DIM * elem + rows_local * 4
    ach = algo_bytes / (scan_per.item() * 1e-3) / 1e9
    path = L.mx_store_scan_path(store.local.handle, NQ, TOPK, -1)
    kernel = {0: "scan_stream_kernel<float>", 1: "scan_stream_kernel<__half>", 2: "scan_tc_kernel (tcgen05)"}[path]
    line = {
        "metric": "queries/sec", "value": NQ * 1e3 / ms_step, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": dict(workload_config(args.rows, world),
                       **({"exchange": ("peer-memory push + flag wait inside the merge kernel (NVLink P2P stores, no collective call)"
                                        if store.exchange == "p2p" else "one NCCL all_gather_into_tensor of the per-shard blobs")}
                          if world > 1 else {})),
        "e2e": {"value": NQ * args.steps / e2e_s.item(), "unit": "queries/s", "h2d_bytes_per_step": NQ * DIM * 4,
                "d2h_bytes_per_step": NQ * TOPK * 12 + NQ * 4},
        "gpu_launches": int(launches.item()),
        "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                     # DRAM bytes per launch under ncu: captured at N = 1 only (profiles/roofline_traffic.json); a
                     # shard's launch moves 1/N of it and was not captured separately -> null rather than a wrong number
                     "traffic": ncu_traffic(f"scan_path{path}_nq{NQ}") if world == 1 else None, "kernel": kernel,
                     "kernel_ms": scan_per.item(), "algorithmic_bytes_per_launch": algo_bytes,
                     "other_kernels_ms_per_step": oth_ms.value / args.steps, "peak_kind": "copy bandwidth, " + pk["src"]},
    }
    other_ms = oth_ms.value / args.steps
    line["step_budget_ms"] = {"scan_kernel": scan_per.item(), "other_kernels_of_the_store": other_ms,
                              "exchange_merge_and_launch_gaps": max(0.0, ms_step - scan_per.item() - other_ms),
                              "ideal_scan_at_peak": algo_bytes / (pk["hbm"] * 1e9) * 1e3}
    if sustained is not None:
        line["sustained"] = sustained
    if rank == 0:
        line["e2e"]["clocks"] = sampler.summary(t_e2e_begin, t_e2e_end)
        line["e2e"]["idle_before_region_s"] = IDLE_BEFORE_REGION_S
    line["e2e"]["in_flight"] = 2
    line["e2e"]["api"] = "search_submit / search_collect (mx_store_search_submit, mx_shard_group_search_submit), host buffers"
    line["e2e"]["one_in_flight"] = {"value": NQ * args.steps / e2e_one_s.item(), "unit": "queries/s",
                                    "api": "the blocking call (mx_store_search / mx_shard_group_search)"}
    line["result_digest"] = digest
    line["result_check"] = result_check
    if rank == 0:
        line["clocks"] = sampler.summary(t_begin, t_end)
        # outside the timed region: what a plain 1 GiB device-to-device copy reaches on this very GPU
        line["roofline"]["copy_gbs_this_box"] = box_copy_bandwidth(device)
    store.close()
    del store
    torch.cuda.empty_cache()

    if world > 1 and not args.skip_extras:
        # the embedder does not shard: every GPU runs its own replica on its own batches, no collective.
        # All ranks run the same timed batches between two barriers; the aggregate is the sum of segments over the
        # slowest rank's time
        barrier()
        emb = bench_embed(device, max(5, args.steps // 2), max(3, args.warmup), pk, False, extras=False, parity=False)
        ms = torch.tensor([emb["ms_per_step"]], dtype=torch.float64, device=device)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=group)
        line["embed"] = {"workload": emb["workload"] + f", one replica per GPU x{world}", "value": world * 256 * 1e3 / ms.item(),
                         "unit": "segments/s", "ms_per_step": ms.item(), "dtype": "bf16", "scaling": "weak (replicas, no collective)",
                         "per_gpu_roofline": emb["roofline"]}
    if not args.skip_extras and not args.skip_ingest:
        # config 5: streamed ingest (embed + append) WHILE searching, SMs partitioned, every rank on its own shard
        barrier()
        ing = bench_ingest(device, pk, args.ingest_rows, args.ingest_seconds, rank, world, group)
        if rank == 0:
            line["ingest"] = ing
    if rank == 0 and world == 1:
        if not args.skip_extras:
            line["single_query"] = bench_single_query(device, max(50, args.steps * 5), max(20, args.warmup), pk)
            line["embed"] = bench_embed(device, max(5, args.steps // 2), max(3, args.warmup), pk, not args.skip_cpu,
                                        parity=not args.skip_check)
        if not args.skip_cpu:
            v, info = cpu_hnsw_baseline(20, 3, 1, HNSW_LEG_ROWS)
            line["cpu_baseline"] = {
                "value": v, "unit": "queries/s", "cores": 1, "kind": "port",
                "sample": (f"HNSW restatement of the reference's search (M=16, efC=200, ef=32), single thread as the "
                           f"reference's mutex-serialised search, over a {HNSW_LEG_ROWS}-row sample (build "
                           f"{info['build_s']} s untimed), 20 x {NQ} queries; recall@10 = {info['recall_at_10']}; "
                           f"approximate search, cost ~log N")}
            c1, x1, q1, e1 = cpu_config1()
            line["config1"] = {"cpu": c1, "gpu": gpu_config1(local, x1, q1, e1)}
            ev, cores = cpu_exact_baseline()
            line["cpu_exact"] = {
                "value": ev * EXACT_SAMPLE_ROWS / args.rows, "unit": "queries/s", "cores": cores, "kind": "port",
                "sample": (f"exact brute force with DistCosine arithmetic over {EXACT_SAMPLE_ROWS} of the rows x {NQ} "
                           f"queries = {ev:.1f} q/s, scaled linearly to {args.rows} rows")}
    if world > 1:
        dist.barrier(group=group)
        dist.destroy_process_group()
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if rank == 0:
        print(json.dumps(line), flush=True)
        if result_check is not None and not result_check["ok"]:
            sys.stderr.write("bench.py: THE ANSWER DIFFERS FROM THE CHECKER'S -- the numbers above are not valid\n")
            sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--ingest-rows", type=int, default=6_250_000, help="config 5: rows of the 768-d shard PER GPU (50 M / 8)")
    ap.add_argument("--ingest-seconds", type=float, default=2.0, help="config 5: length of each timed window")
    ap.add_argument("--skip-ingest", action="store_true", help="leave out the config-5 sub-bench (ingest while searching)")
    ap.add_argument("--skip-cpu", action="store_true", help="leave out the CPU baseline legs")
    ap.add_argument("--skip-extras", action="store_true", help="leave out the single_query / embed sub-benches")
    ap.add_argument("--skip-check", action="store_true", help="profiling aid: leave out the independent answer check")
    ap.add_argument("--only", default="", choices=["", "embed", "single", "ingest"],
                    help="profiling aid: run just one sub-bench on one GPU and print its object")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.only:
        import torch
        torch.cuda.set_device(0)
        dev = torch.device("cuda", 0)
        if args.only == "ingest":
            res = bench_ingest(dev, peaks(), args.ingest_rows, args.ingest_seconds)
        elif args.only == "embed":
            res = bench_embed(dev, args.steps, args.warmup, peaks(), False, not args.skip_extras, parity=not args.skip_check)
        else:
            res = bench_single_query(dev, args.steps, args.warmup, peaks())
        print(json.dumps(res), flush=True)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
