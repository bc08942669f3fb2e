"""CPU coverage of the N > 1 search protocol (memex_b200/sharded.py): shard plan, blob layout and the
one all-gather, run with the `gloo` backend at world size 2.  The per-shard answers come from the
CPU oracle here (the product's come from the CUDA kernels -- tests/test_store_gpu.py checks those
and the device merge against the same oracle); what this file pins is the host-side logic that both
share: contiguous row ranges, global 1-based ids, `{ids | keys | counts}` blobs, merge order."""
import os
import socket

import numpy as np
import pytest

from memex_b200 import capi
from memex_b200.sharded import ShardPlan, pack_blob, unpack_blob
from oracle import cosine


def test_shard_plan_partitions_rows():
    for total in (0, 1, 7, 8, 1000, 10_000_000):
        for world in (1, 2, 3, 8):
            p = ShardPlan(total, world)
            counts = [p.count(r) for r in range(world)]
            assert sum(counts) == total and max(counts) - min(counts) <= 1
            assert p.start(0) == 0
            for r in range(1, world):
                assert p.start(r) == p.start(r - 1) + p.count(r - 1)
            if total:
                assert p.owner(0) == 0 and p.owner(total - 1) == max(r for r in range(world) if counts[r])


def test_blob_layout_matches_the_c_abi():
    rng = np.random.default_rng(0)
    for nq, k in ((1, 1), (3, 10), (64, 10), (5, 256)):
        ids = rng.integers(1, 1 << 40, (nq, k)).astype(np.uint64)
        keys = rng.random((nq, k), dtype=np.float32)
        counts = rng.integers(0, k + 1, nq).astype(np.uint32)
        blob = pack_blob(ids, keys, counts)
        assert blob.nbytes == capi.lib().mx_topk_blob_bytes(nq, k) and blob.nbytes % 16 == 0
        i2, k2, c2 = unpack_blob(blob, nq, k)
        np.testing.assert_array_equal(i2, ids)
        np.testing.assert_array_equal(k2, keys)
        np.testing.assert_array_equal(c2, counts)


def merge_blobs_host(blobs, nq, k):
    """numpy restatement of merge_kernel's ordering (rerank.cu): (key asc, id asc), best k"""
    out_ids = np.zeros((nq, k), np.uint64)
    out_keys = np.zeros((nq, k), np.float32)
    out_counts = np.zeros(nq, np.uint32)
    parts = [unpack_blob(b, nq, k) for b in blobs]
    for q in range(nq):
        ent = []
        for ids, keys, counts in parts:
            ent += [(float(keys[q, j]), int(ids[q, j])) for j in range(int(counts[q]))]
        ent.sort()
        ent = ent[:k]
        out_counts[q] = len(ent)
        for j, (d, i) in enumerate(ent):
            out_ids[q, j], out_keys[q, j] = i, d
    return out_ids, out_keys, out_counts


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, nq, k, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)                       # same corpus on every rank; each keeps its range
        corpus = rng.standard_normal((n, d)).astype(np.float32)
        corpus[n // 2 + 3] = corpus[5]                        # a cross-shard exact tie
        plan = ShardPlan(n, world)
        lo, cnt = plan.start(rank), plan.count(rank)
        q = torch.zeros((nq, d), dtype=torch.float32)
        if rank == 0:
            q = torch.from_numpy(corpus[[5, 17, n - 1]] + 0.05 * rng.standard_normal((nq, d)).astype(np.float32))
        dist.broadcast(q, src=0)                              # ShardedStore.search: queries come from rank 0
        queries = q.numpy()
        ids, _, counts = cosine.exact_topk(corpus[lo:lo + cnt], queries, k)
        keys = np.zeros((nq, k), np.float32)
        for i in range(nq):
            for j in range(int(counts[i])):
                keys[i, j] = cosine.dist_cosine(queries[i], corpus[lo + int(ids[i, j]) - 1])
        gids = np.where(np.arange(k)[None, :] < counts[:, None], ids + np.uint64(lo), 0).astype(np.uint64)  # id_offset
        mine = torch.from_numpy(pack_blob(gids, keys, counts))
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)                       # the one exchange step
        m_ids, m_keys, m_counts = merge_blobs_host([g.numpy() for g in gathered], nq, k)
        o_ids, o_scores, o_counts = cosine.exact_topk(corpus, queries, k)
        ok = (m_ids == o_ids).all() and (m_counts == o_counts).all()
        scores = np.array([[cosine.similarity(float(x)) for x in row] for row in m_keys], np.float32)
        ok = ok and (scores.view(np.uint32) == o_scores.view(np.uint32)).all()
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_search_protocol():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        procs = [mp.get_context("spawn").Process(target=_worker, args=(r, world, port, 3001, 48, 3, 10, ret))
                 for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(240)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert ret.get(0) is True and ret.get(1) is True
