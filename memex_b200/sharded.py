"""Row-sharded multi-GPU search: one process per GPU, one all-gather of per-shard top-k candidates.

The reference is single-process; its `VectorStore::search` (reference
lib/libmemex/src/storage/mod.rs:85-92 -> storage/local.rs:71-91) answers from one index.  Here the
corpus rows are dealt contiguously over the ranks of a `torch.distributed` group; every rank scans
its own shard with the same C-ABI call the single-GPU store uses and reports GLOBAL 1-based ids
(`mx_store_cfg.id_offset`), then ONE all-gather moves each rank's `{ids | keys | counts}` blob
(mx_topk_blob_bytes) and every rank merges the G * k candidates per query on its device with
`mx_merge_topk_blobs_device` -- ordering (key asc, id asc), exactly the single-store ordering.

torch is used for device buffers, the stream and the collective only; all arithmetic is in
libmemex_b200.so.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
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
                 device: int = 0, rank: int = 0, world: int = 1, group=None):
        self.plan = ShardPlan(total_rows, world)
        self.rank, self.world, self.group = rank, world, group
        self.dim, self.metric, self.device = dim, metric, device
        self.local = B200Store.new(storage_path, dim=dim, dtype=dtype, metric=metric, device=device,
                                   capacity=self.plan.count(rank), id_offset=self.plan.start(rank), id_stride=1)
        self._bufs = {}

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
            self._bufs[key] = dict(
                blob_bytes=blob,
                gathered=torch.zeros((self.world, blob), dtype=torch.uint8, device=dev),
                mine=torch.zeros(blob, dtype=torch.uint8, device=dev),
                q=torch.zeros((nq, self.dim), dtype=torch.float32, device=dev),
                ids=torch.zeros((nq, k), dtype=torch.int64, device=dev),
                scores=torch.zeros((nq, k), dtype=torch.float32, device=dev),
                counts=torch.zeros(nq, dtype=torch.int32, device=dev),
                q_pin=torch.zeros((nq, self.dim), dtype=torch.float32).pin_memory(),
                ids_pin=torch.zeros((nq, k), dtype=torch.int64).pin_memory(),
                scores_pin=torch.zeros((nq, k), dtype=torch.float32).pin_memory(),
                counts_pin=torch.zeros(nq, dtype=torch.int32).pin_memory(),
            )
        return self._bufs[key]

    def search_device(self, q_dev, k: int):
        """q_dev: torch f32 [nq, dim] on this rank's device, identical on every rank.
        -> (ids i64 [nq,k], scores f32 [nq,k], counts i32 [nq]) device tensors, identical on every
        rank, enqueued on torch's current stream (no host synchronisation)."""
        import torch
        import torch.distributed as dist
        nq = q_dev.shape[0]
        b = self._buffers(nq, k)
        L = capi.lib()
        # torch's default stream has handle 0, which the C ABI reads as "the store's own stream": name the
        # legacy default stream explicitly (cudaStreamLegacy == 0x1) so the work is ordered with torch's
        st = torch.cuda.current_stream(self.device).cuda_stream or 1
        mine = b["mine"]
        rc = L.mx_store_search_blob_device(self.local.handle, q_dev.data_ptr(), nq, k, mine.data_ptr(), st)
        if rc != capi.OK:
            _raise(rc, self.local.handle, SearchError)
        if self.world > 1:
            # the one exchange step: each rank contributes blob_bytes
            dist.all_gather_into_tensor(b["gathered"].view(-1), mine, group=self.group)
            blobs, n_shards = b["gathered"], self.world
        else:
            blobs, n_shards = mine, 1
        rc = L.mx_merge_topk_blobs_device(blobs.data_ptr(), b["blob_bytes"], n_shards, nq, k,
                                          capi.METRIC_DOT if self.metric == "dot" else capi.METRIC_COSINE,
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
        b = self._buffers(nq, k)
        if queries is not None:
            b["q_pin"].numpy()[...] = queries
            b["q"].copy_(b["q_pin"], non_blocking=True)
        if self.world > 1:
            dist.broadcast(b["q"], src=0, group=self.group)
        ids, scores, counts = self.search_device(b["q"], k)
        b["ids_pin"].copy_(ids, non_blocking=True)
        b["scores_pin"].copy_(scores, non_blocking=True)
        b["counts_pin"].copy_(counts, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return (b["ids_pin"].numpy().astype(np.uint64), b["scores_pin"].numpy().copy(),
                b["counts_pin"].numpy().astype(np.uint32))

    def close(self):
        self.local.close()
