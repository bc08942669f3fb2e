"""Row-sharded multi-GPU search: one process per GPU, one all-gather of per-shard top-k candidates.

The reference is single-process; its `VectorStore::search` (reference
lib/libmemex/src/storage/mod.rs:85-92 -> storage/local.rs:71-91) answers from one index.  Here the
corpus rows are dealt contiguously over the ranks of a `torch.distributed` group; every rank scans
its own shard with the same C-ABI call the single-GPU store uses and reports GLOBAL 1-based ids
(`mx_store_cfg.id_offset`), then ONE all-gather moves each rank's `{ids | keys | counts}` blob
(mx_topk_blob_bytes) and every rank merges the G * k candidates per query on its device with
`mx_merge_topk_blobs_device` -- ordering (key asc, id asc), exactly the single-store ordering.

The exchange step has two forms.  Default: peer memory -- every rank owns an exchange buffer its peers can address
(torch symmetric memory: CUDA IPC / fabric handles over NVLink), `mx_exchange_push_device` stores the rank's blob into
all peers' buffers and release-stores an epoch flag, and `mx_merge_topk_blobs_wait_device` waits for its flags inside
the merge kernel: two of this library's own kernels, no collective library call on the query path.  Fallback (when
the symmetric-memory rendezvous is not available, or MX_EXCHANGE=nccl): one NCCL all_gather_into_tensor.

torch is used for device buffers, the stream and the collective only; all arithmetic is in
libmemex_b200.so.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import capi
from .storage import B200Store, InsertionError, SearchError, _raise


@dataclass(frozen=True)
class ShardPlan:
    """Contiguous row ranges: rank r owns global rows [start, start + count)."""
    total: int
    world: int

    def count(self, rank: int) -> int:
        base, rem = divmod(self.total, self.world)
        return base + (1 if rank < rem else 0)

    def start(self, rank: int) -> int:
        base, rem = divmod(self.total, self.world)
        return rank * base + min(rank, rem)

    def owner(self, global_row: int) -> int:
        for r in range(self.world):
            if global_row < self.start(r) + self.count(r):
                return r
        raise IndexError(global_row)


def pack_blob(ids: np.ndarray, keys: np.ndarray, counts: np.ndarray) -> np.ndarray:
    """host-side view of the blob layout (tests / gloo): {ids u64 [nq,k] | keys f32 [nq,k] | counts u32 [nq]}"""
    nq, k = ids.shape
    n = ((nq * k * 12 + nq * 4) + 15) & ~15
    out = np.zeros(n, dtype=np.uint8)
    out[: nq * k * 8] = np.ascontiguousarray(ids, dtype=np.uint64).view(np.uint8).ravel()
    out[nq * k * 8: nq * k * 12] = np.ascontiguousarray(keys, dtype=np.float32).view(np.uint8).ravel()
    out[nq * k * 12: nq * k * 12 + nq * 4] = np.ascontiguousarray(counts, dtype=np.uint32).view(np.uint8).ravel()
    return out


def unpack_blob(blob: np.ndarray, nq: int, k: int):
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    ids = blob[: nq * k * 8].view(np.uint64).reshape(nq, k)
    keys = blob[nq * k * 8: nq * k * 12].view(np.float32).reshape(nq, k)
    counts = blob[nq * k * 12: nq * k * 12 + nq * 4].view(np.uint32)
    return ids, keys, counts


class ShardedStore:
    """The local shard of a row-sharded corpus plus the exchange step.

    `group` is a torch.distributed process group (NCCL on the GPU box); with world size 1 no
    collective is issued and this is a thin wrapper over one B200Store.
    """

    def __init__(self, storage_path, dim: int, total_rows: int, dtype: str = "f16", metric: str = "cosine",
                 device: int = 0, rank: int = 0, world: int = 1, group=None, id_stride: int = 1):
        """id_stride = 1: contiguous row ranges (rank r owns global rows [start, start + count), a corpus known up front);
        id_stride = world: round-robin ids (local row i of rank r is global row i * world + r) -- a corpus that GROWS on
        every rank independently (streamed ingest), `total_rows` then only sizes the shards' capacity."""
        self.plan = ShardPlan(total_rows, world)
        self.rank, self.world, self.group = rank, world, group
        self.dim, self.metric, self.device = dim, metric, device
        self.local = B200Store.new(storage_path, dim=dim, dtype=dtype, metric=metric, device=device,
                                   capacity=self.plan.count(rank),
                                   id_offset=self.plan.start(rank) if id_stride == 1 else rank, id_stride=id_stride)
        self._bufs = {}
        self.exchange = "none" if world == 1 else "nccl"
        self._want_p2p = world > 1 and os.environ.get("MX_EXCHANGE", "p2p") != "nccl"
        self._shard_group = None          # mx_shard_group member (peer-memory exchange), made on the first search
        self._group_nq, self._group_k = 64, 16
        self._group_error = ""

    def _setup_group(self, nq: int, k: int) -> bool:
        """The peer-memory exchange below the C ABI (mx_shard_group, csrc/shard_group.cu): every rank creates its member,
        exports the 64-byte CUDA IPC handle of its exchange buffer, the handles travel over torch.distributed (plumbing:
        any transport would do) and every rank opens its peers'.  All ranks or none."""
        import torch
        import torch.distributed as dist
        L = capi.lib()
        dev = torch.device("cuda", self.device)
        ok = 1
        g = C.c_void_p()
        handle = torch.zeros(capi.IPC_HANDLE_BYTES, dtype=torch.uint8)
        rc = L.mx_shard_group_create(self.device, self.world, self.rank, self.dim, max(nq, self._group_nq), max(k, self._group_k),
                                     C.byref(g))
        if rc != capi.OK or L.mx_shard_group_export(g, handle.data_ptr()) != capi.OK:
            self._group_error = (L.mx_last_error(g if g.value else None) or b"").decode(errors="replace")
            ok = 0
        gathered = [torch.zeros(capi.IPC_HANDLE_BYTES, dtype=torch.uint8, device=dev) for _ in range(self.world)]
        dist.all_gather(gathered, handle.to(dev), group=self.group)
        if ok:
            handles = torch.stack(gathered).cpu().contiguous()
            if L.mx_shard_group_connect(g, handles.data_ptr()) != capi.OK:
                self._group_error = (L.mx_last_error(g) or b"").decode(errors="replace")
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) != 1:
            if g.value:
                L.mx_shard_group_destroy(g)
            return False
        self._shard_group = g
        self._group_nq, self._group_k = max(nq, self._group_nq), max(k, self._group_k)
        dist.barrier(group=self.group)   # every member is connected before anyone pushes
        return True

    def _ensure_group(self, nq: int, k: int):
        import torch
        import torch.distributed as dist
        if self._shard_group is not None:        # a larger batch than the group was sized for: make a new one
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            capi.lib().mx_shard_group_destroy(self._shard_group)
            self._shard_group = None
        if self._setup_group(nq, k):
            self.exchange = "p2p"
        else:
            self._want_p2p = False

    # ---- ingest: rows already on this rank's device (f32 [n, dim]); ids follow the plan ----
    def add_local_device(self, rows_dev_ptr: int, n: int) -> int:
        first = C.c_uint64()
        rc = capi.lib().mx_store_add_device(self.local.handle, rows_dev_ptr, n, C.byref(first))
        if rc != capi.OK:
            _raise(rc, self.local.handle, InsertionError)
        return first.value

    def add_local(self, rows: np.ndarray) -> int:
        return self.local.add_matrix(rows)

    def __len__(self):
        return len(self.local)

    def _buffers(self, nq: int, k: int):
        import torch
        key = (nq, k)
        if key not in self._bufs:
            dev = torch.device("cuda", self.device)
            blob = int(capi.lib().mx_topk_blob_bytes(nq, k))
            # the answer {ids i64 [nq,k] | scores f32 [nq,k] | counts i32 [nq]} lives in ONE buffer on each side, so the
            # end-to-end call reads it back with a single device-to-host copy
            a, c = nq * k * 8, nq * k * 12

            def views(buf):
                return (buf[:a].view(torch.int64).view(nq, k), buf[a:c].view(torch.float32).view(nq, k),
                        buf[c:].view(torch.int32))
            out = torch.zeros(c + nq * 4, dtype=torch.uint8, device=dev)
            out_pin = torch.zeros(c + nq * 4, dtype=torch.uint8).pin_memory()
            ids, scores, counts = views(out)
            ids_pin, scores_pin, counts_pin = views(out_pin)
            self._bufs[key] = dict(
                blob_bytes=blob,
                gathered=torch.zeros((self.world, blob), dtype=torch.uint8, device=dev),
                mine=torch.zeros(blob, dtype=torch.uint8, device=dev),
                q=torch.zeros((nq, self.dim), dtype=torch.float32, device=dev),
                out=out, ids=ids, scores=scores, counts=counts,
                q_pin=torch.zeros((nq, self.dim), dtype=torch.float32).pin_memory(),
                out_pin=out_pin, ids_pin=ids_pin, scores_pin=scores_pin, counts_pin=counts_pin,
            )
        return self._bufs[key]

    def search_device(self, q_dev, k: int, nq: int | None = None, query_root: int = -1):
        """q_dev: torch f32 [nq, dim] on this rank's device, identical on every rank (query_root = -1); or, with the
        peer-memory exchange, only on rank `query_root` (the others pass None and `nq`: the root's block arrives through
        their exchange buffer).
        -> (ids i64 [nq,k], scores f32 [nq,k], counts i32 [nq]) device tensors, identical on every
        rank, enqueued on torch's current stream (no host synchronisation)."""
        import torch
        import torch.distributed as dist
        if q_dev is not None:
            nq = q_dev.shape[0]
        b = self._buffers(nq, k)
        L = capi.lib()
        # torch's default stream has handle 0, which the C ABI reads as "the store's own stream": name the
        # legacy default stream explicitly (cudaStreamLegacy == 0x1) so the work is ordered with torch's
        st = torch.cuda.current_stream(self.device).cuda_stream or 1
        metric = capi.METRIC_DOT if self.metric == "dot" else capi.METRIC_COSINE
        if self._want_p2p and (self._shard_group is None or nq > self._group_nq or k > self._group_k):
            self._ensure_group(nq, k)
        if self._shard_group is not None:
            # scan + rerank, push to every peer, merge-with-wait: all below the C ABI (csrc/shard_group.cu)
            rc = L.mx_shard_group_search_device(self._shard_group, self.local.handle,
                                                q_dev.data_ptr() if q_dev is not None else None, query_root, nq, k,
                                                b["ids"].data_ptr(), b["scores"].data_ptr(), b["counts"].data_ptr(), st)
            if rc != capi.OK:
                _raise(rc, self._shard_group, SearchError)
            return b["ids"], b["scores"], b["counts"]
        if self.world == 1:
            # one shard: the store's own answer IS the answer (no blob, no merge launch)
            rc = L.mx_store_search_device(self.local.handle, q_dev.data_ptr(), nq, k, b["ids"].data_ptr(), b["scores"].data_ptr(),
                                          None, b["counts"].data_ptr(), st)
            if rc != capi.OK:
                _raise(rc, self.local.handle, SearchError)
            return b["ids"], b["scores"], b["counts"]
        mine = b["mine"]
        rc = L.mx_store_search_blob_device(self.local.handle, q_dev.data_ptr(), nq, k, mine.data_ptr(), st)
        if rc != capi.OK:
            _raise(rc, self.local.handle, SearchError)
        if self.world > 1:
            # the one exchange step, collective form: each rank contributes blob_bytes
            dist.all_gather_into_tensor(b["gathered"].view(-1), mine, group=self.group)
            blobs, n_shards = b["gathered"], self.world
        else:
            blobs, n_shards = mine, 1
        rc = L.mx_merge_topk_blobs_device(blobs.data_ptr(), b["blob_bytes"], n_shards, nq, k, metric,
                                          b["ids"].data_ptr(), b["scores"].data_ptr(), b["counts"].data_ptr(),
                                          self.device, st)
        if rc != capi.OK:
            _raise(rc, None, SearchError)
        return b["ids"], b["scores"], b["counts"]

    def search(self, queries: np.ndarray | None, k: int, nq: int | None = None):
        """End-to-end call with HOST buffers: rank 0 passes the queries ([nq, dim] f32), the other
        ranks pass None and `nq`; every rank returns (ids u64, scores f32, counts u32) as numpy."""
        import torch
        import torch.distributed as dist
        if queries is not None:
            queries = np.ascontiguousarray(queries, dtype=np.float32)
            nq = queries.shape[0]
            if queries.ndim != 2 or queries.shape[1] != self.dim:
                raise SearchError(f"query has dimension {queries.shape[-1]}, store has {self.dim}")
        L = capi.lib()
        if self.world == 1:
            # one shard: the store's own host-buffer call (pinned staging, one H2D, one D2H, fallback only when flagged)
            return self.local.search_matrix(queries, k)
        if self._want_p2p and (self._shard_group is None or nq > self._group_nq or k > self._group_k):
            self._ensure_group(nq, k)
        if self._shard_group is not None:
            # the whole step below the C ABI (csrc/shard_group.cu): rank 0's block is staged through pinned memory and pushed
            # to the peers' exchange buffers, every rank scans / pushes / merges, ONE device-to-host copy of the answer
            ids = np.zeros((nq, k), dtype=np.uint64)
            scores = np.zeros((nq, k), dtype=np.float32)
            counts = np.zeros(nq, dtype=np.uint32)
            rc = L.mx_shard_group_search(self._shard_group, self.local.handle,
                                         queries.ctypes.data if queries is not None else None, 0, nq, k,
                                         ids.ctypes.data, scores.ctypes.data, counts.ctypes.data)
            if rc != capi.OK:
                _raise(rc, self._shard_group, SearchError)
            return ids, scores, counts
        b = self._buffers(nq, k)
        if queries is not None:
            b["q_pin"].numpy()[...] = queries
            b["q"].copy_(b["q_pin"], non_blocking=True)
        if self.world > 1:
            dist.broadcast(b["q"], src=0, group=self.group)    # collective form of the exchange (MX_EXCHANGE=nccl)
        self.search_device(b["q"], k)                     # fills b["out"] (ids | scores | counts)
        b["out_pin"].copy_(b["out"], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return (b["ids_pin"].numpy().astype(np.uint64), b["scores_pin"].numpy().copy(),
                b["counts_pin"].numpy().astype(np.uint32))

    def search_submit(self, queries: np.ndarray | None, k: int, nq: int | None = None):
        """First half of `search`: enqueues the step and returns a ticket; two may be in flight, collected in order.  Every
        rank submits and collects the same sequence (rank 0 passes the queries, the others None and `nq`)."""
        if queries is not None:
            queries = np.ascontiguousarray(queries, dtype=np.float32)
            nq = queries.shape[0]
            if queries.ndim != 2 or queries.shape[1] != self.dim:
                raise SearchError(f"query has dimension {queries.shape[-1]}, store has {self.dim}")
        if self.world == 1:
            return ("local", self.local.search_submit(queries, k))
        if self._want_p2p and (self._shard_group is None or nq > self._group_nq or k > self._group_k):
            self._ensure_group(nq, k)
        if self._shard_group is None:
            # collective exchange (MX_EXCHANGE=nccl): no split form, the step runs here and collect hands the answer over
            return ("done", self.search(queries, k, nq=nq))
        ticket = C.c_uint64()
        rc = capi.lib().mx_shard_group_search_submit(self._shard_group, self.local.handle,
                                                     queries.ctypes.data if queries is not None else None, 0, nq, k,
                                                     C.byref(ticket))
        if rc != capi.OK:
            _raise(rc, self._shard_group, SearchError)
        return ("group", (ticket.value, nq, k))

    def search_collect(self, ticket):
        kind, t = ticket
        if kind == "local":
            return self.local.search_collect(t)
        if kind == "done":
            return t
        tv, nq, k = t
        ids = np.zeros((nq, k), dtype=np.uint64)
        scores = np.zeros((nq, k), dtype=np.float32)
        counts = np.zeros(nq, dtype=np.uint32)
        rc = capi.lib().mx_shard_group_search_collect(self._shard_group, tv, ids.ctypes.data, scores.ctypes.data, counts.ctypes.data)
        if rc != capi.OK:
            _raise(rc, self._shard_group, SearchError)
        return ids, scores, counts

    def close(self):
        if self._shard_group is not None:
            capi.lib().mx_shard_group_destroy(self._shard_group)
            self._shard_group = None
        self.local.close()
