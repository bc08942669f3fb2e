"""Worker of tests/test_sharded_gpu.py: one process per GPU (torchrun), ShardedStore.search against the CPU oracle.

Every rank builds the SAME unsharded corpus from a seed, keeps its contiguous row range in a ShardedStore on its own
GPU, and for several epochs (both slot sets of the peer-memory exchange, both query paths) compares the merged answer
with oracle.cosine.exact_topk over the UNSHARDED corpus: ids identical, scores bit-identical, on every rank.
Prints one line `rank R ok ...` per rank and exits non-zero on any mismatch.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=600_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--epochs", type=int, default=6)
    ap.add_argument("--k", type=int, default=10)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from memex_b200.sharded import ShardedStore
    from oracle import cosine

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)                                    # NCCL's banner goes to stderr
    dist.init_process_group("nccl", device_id=device)
    n, d, k = args.rows, args.dim, args.k
    rng = np.random.default_rng(2024)
    corpus = rng.standard_normal((n, d), dtype=np.float32)
    corpus *= rng.uniform(0.5, 2.0, (n, 1)).astype(np.float32)
    # exact ties across shard boundaries and inside one shard; a zero-norm row in the last shard
    plan_rows = [n * r // world for r in range(world)]
    for r in range(1, world):
        corpus[plan_rows[r] + 7] = corpus[11]
        corpus[plan_rows[r] - 3] = corpus[11]
    corpus[n - 5] = 0.0
    stored = corpus.astype(np.float16).astype(np.float32)
    store = ShardedStore(f"/tmp/mx_sharded_test_{rank}", d, n, dtype="f16", device=local, rank=rank, world=world,
                         group=dist.group.WORLD)
    lo, cnt = store.plan.start(rank), store.plan.count(rank)
    store.add_local(corpus[lo:lo + cnt])
    checked = 0
    for epoch in range(args.epochs):
        for nq in (64, 1, 9):                        # tcgen05 batch, single-query stream scan, small tcgen05 batch
            qrng = np.random.default_rng(1000 * epoch + nq)
            picks = qrng.integers(0, n, nq)
            picks[0] = 11                            # the cross-shard tie group
            queries = (corpus[picks] + 0.2 * qrng.standard_normal((nq, d))).astype(np.float32)
            if epoch == 1:
                queries[0] = corpus[11]
            ids, scores, counts = store.search(queries if rank == 0 else None, k, nq=nq)
            oi, os_, oc = cosine.exact_topk(stored, queries, k)
            if not ((ids == oi).all() and (counts == oc).all() and (scores.view(np.uint32) == os_.view(np.uint32)).all()):
                bad = np.argwhere(ids != oi)
                print(f"rank {rank} MISMATCH epoch {epoch} nq {nq} exchange {store.exchange}: first at {bad[:3].tolist()}",
                      file=sys.stderr, flush=True)
                os._exit(3)
            checked += nq
    # the split call: two searches in flight on every rank, collected in order, same answers as the oracle
    batches = []
    for i in range(5):
        qrng = np.random.default_rng(77 + i)
        batches.append((corpus[qrng.integers(0, n, 64)] + 0.2 * qrng.standard_normal((64, d))).astype(np.float32))
    got, prev = [], None
    for q in batches:
        t = store.search_submit(q if rank == 0 else None, k, nq=64)
        if prev is not None:
            got.append(store.search_collect(prev))
        prev = t
    got.append(store.search_collect(prev))
    for q, (ids, scores, counts) in zip(batches, got):
        oi, os_, oc = cosine.exact_topk(stored, q, k)
        if not ((ids == oi).all() and (counts == oc).all() and (scores.view(np.uint32) == os_.view(np.uint32)).all()):
            print(f"rank {rank} MISMATCH in the submit / collect form, exchange {store.exchange}", file=sys.stderr, flush=True)
            os._exit(4)
        checked += 64
    exchange = store.exchange
    store.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(saved, 1)
    print(f"rank {rank} ok: {checked} queries over {args.epochs} epochs, world {world}, exchange {exchange}", flush=True)


if __name__ == "__main__":
    main()
