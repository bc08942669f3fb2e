"""GPU parity tests of the vector store (through the C ABI) against the CPU oracle.

First block re-expresses the reference's own tests, lib/libmemex/src/storage/local.rs:175-242.
Bar: ids identical to the exact (distance asc, id asc) ranking, scores BIT-identical to the
oracle's DistCosine + local.rs:86 arithmetic evaluated on the stored rows.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from memex_b200 import capi
from memex_b200.storage import B200Store, VectorData, VectorStoreError, get_vector_storage
from oracle import cosine

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def reference_test_data():
    """local.rs:175-199"""
    return [
        VectorData(_id="test-one", document_id="test-one", text="", segment_id=0, vector=[0.0, 0.1, 0.2]),
        VectorData(_id="test-two", document_id="test-two", text="", segment_id=0, vector=[0.1, 0.1, 0.1]),
        VectorData(_id="test-three", document_id="test-three", text="", segment_id=0, vector=[0.3, 0.2, 0.1]),
    ]


# ---- the reference's tests, re-expressed --------------------------------------------------------

def test_hnsw(tmp_path):
    """local.rs:201-214"""
    store = B200Store.new(tmp_path)
    store.bulk_insert(reference_test_data())
    results = store.search([0.1, 0.1, 0.1], 3)
    assert len(results) == 3
    doc_id, _ = results[0]
    assert doc_id == "test-two"
    # beyond what the reference asserts: full order and scores of the golden fixture
    fx = json.load(open(os.path.join(GOLD, "search_ref_fixture.json")))
    assert [r[0] for r in results] == fx["expected_ids"]
    assert [int(np.float32(r[1]).view(np.uint32)) for r in results] == fx["expected_score_bits"]
    store.delete_all()


def test_save_load(tmp_path):
    """local.rs:216-227"""
    path = tmp_path / "vectortest"
    store = B200Store.new(path)
    store.bulk_insert(reference_test_data())
    store.save(path)
    loaded = B200Store.load(path)
    assert len(loaded._id_map) == len(store._id_map)
    assert loaded.get_nb_point() == 3
    assert loaded.search([0.1, 0.1, 0.1], 3) == store.search([0.1, 0.1, 0.1], 3)
    # vectors.meta.json stays byte-compatible with serde_json's HashMap<usize, String> dump
    meta = json.load(open(path / "vectors.meta.json"))
    assert meta == {"1": "test-one", "2": "test-two", "3": "test-three"}
    store.delete_all()


def test_delete_all(tmp_path):
    """local.rs:229-242"""
    store = B200Store.new(tmp_path)
    store.bulk_insert(reference_test_data())
    store.save(tmp_path)
    store.delete_all()
    assert len(store._id_map) == 0
    assert store.get_nb_point() == 0
    with pytest.raises(VectorStoreError):
        B200Store.load(tmp_path)


def test_delete_single_is_unsupported_not_a_panic(tmp_path):
    """local.rs:29-32 is unimplemented!(); the ABI returns Unsupported"""
    store = B200Store.new(tmp_path)
    store.bulk_insert(reference_test_data())
    with pytest.raises(VectorStoreError) as ei:
        store.delete("test-one")
    assert ei.value.variant == "Unsupported"
    assert len(store.search([0.1, 0.1, 0.1], 3)) == 3     # handle still usable


def test_factory_new_then_load(tmp_path):
    """storage/mod.rs:107-121 with the b200:// scheme"""
    vs = get_vector_storage(f"b200://{tmp_path}", "coll")
    vs.add_vectors(reference_test_data())
    assert vs.search([0.1, 0.1, 0.1], 2)[0][0] == "test-two"
    vs2 = get_vector_storage(f"b200://{tmp_path}", "coll")     # meta exists -> load
    assert [r[0] for r in vs2.search([0.1, 0.1, 0.1], 3)] == ["test-two", "test-three", "test-one"]
    vs2.delete_collection()
    assert vs2.search([0.1, 0.1, 0.1], 3) == []


def test_empty_batch_leaves_no_files_and_limit_cap(tmp_path):
    """an empty add on a fresh collection must not leave a meta file without its matrix (the directory would be
    unusable on the next open); limit > MX_MAX_K is a SearchError in every host, never a shorter list"""
    vs = get_vector_storage(f"b200://{tmp_path}", "fresh")
    vs.add_vectors([])
    assert not B200Store.has_store(tmp_path / "fresh")
    vs = get_vector_storage(f"b200://{tmp_path}", "fresh")           # opens as new again
    vs.add_vectors(reference_test_data())
    assert vs.search([0.1, 0.1, 0.1], 256)[0][0] == "test-two"
    with pytest.raises(VectorStoreError) as ei:
        vs.search([0.1, 0.1, 0.1], 257)
    assert ei.value.variant == "SearchError"
    # what an older build left behind: "{}" and no matrix -> an empty store, not a FileIOError
    orphan = tmp_path / "orphan"
    orphan.mkdir()
    (orphan / "vectors.meta.json").write_text("{}")
    vo = get_vector_storage(f"b200://{tmp_path}", "orphan")
    assert vo.search([0.1, 0.1, 0.1], 3) == []
    vo.add_vectors(reference_test_data())
    assert vo.search([0.1, 0.1, 0.1], 1)[0][0] == "test-two"


# ---- parity against the oracle -------------------------------------------------------------------

def unit_rows(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def check_parity(store, corpus_as_stored, queries, k, metric="cosine"):
    ids, scores, counts = store.search_matrix(queries, k)
    oi, os_, oc = cosine.exact_topk(corpus_as_stored, queries, k, metric=metric)
    np.testing.assert_array_equal(counts, oc)
    np.testing.assert_array_equal(ids, oi)
    np.testing.assert_array_equal(scores.view(np.uint32), os_.view(np.uint32))


def test_golden_small(tmp_path):
    g = np.load(os.path.join(GOLD, "search_small.npz"))
    store = B200Store.new(tmp_path, dim=24)
    store.add_matrix(g["corpus"])
    ids, scores, counts = store.search_matrix(g["queries"], 7)
    np.testing.assert_array_equal(ids, g["ids"])
    np.testing.assert_array_equal(scores.view(np.uint32), g["score_bits"])


@pytest.mark.parametrize("n,d,nq,k", [
    (1, 3, 1, 1), (5, 3, 2, 10), (1000, 384, 1, 10), (1000, 384, 7, 10), (20000, 384, 1, 10),
    (20000, 384, 64, 10), (3001, 17, 3, 5), (4096, 768, 2, 32), (777, 100, 5, 64), (5000, 384, 2, 256),
    (300, 1, 1, 3), (70000, 64, 9, 10),
])
def test_parity_f32(tmp_path, n, d, nq, k):
    corpus = unit_rows(n, d, 1234)
    queries = unit_rows(nq, d, 4321)
    store = B200Store.new(tmp_path, dim=d)
    store.add_matrix(corpus)
    check_parity(store, corpus, queries, k)


@pytest.mark.parametrize("n,d,nq,k", [(1000, 384, 1, 10), (20000, 384, 64, 10), (9000, 768, 3, 10), (513, 40, 2, 7)])
def test_parity_f16_store(tmp_path, n, d, nq, k):
    """fp16 corpus: ids / scores are exact w.r.t. the oracle fed the SAME fp16-rounded rows, and
    within 1e-4 of the fp32-corpus oracle (BASELINE.md parity gates)."""
    corpus = unit_rows(n, d, 1234)
    queries = unit_rows(nq, d, 4321)
    store = B200Store.new(tmp_path, dim=d, dtype="f16")
    store.add_matrix(corpus)
    stored = corpus.astype(np.float16).astype(np.float32)
    check_parity(store, stored, queries, k)
    ids, scores, _ = store.search_matrix(queries, k)
    for i in range(nq):
        ref = cosine.scores_of(corpus, queries[i], ids[i])
        assert np.abs(ref - scores[i]).max() <= 1e-4
    back = np.zeros((min(n, 100), d), dtype=np.float32)
    assert capi.lib().mx_store_get_rows(store.handle, 0, back.shape[0], back.ctypes.data) == 0
    np.testing.assert_array_equal(back, stored[:back.shape[0]])


def test_unnormalised_rows_duplicates_and_zero_rows(tmp_path):
    rng = np.random.default_rng(5)
    corpus = (rng.standard_normal((4000, 48)) * rng.uniform(0.01, 50, (4000, 1))).astype(np.float32)
    corpus[100] = corpus[7]              # duplicates: tie -> lower id first
    corpus[2500] = corpus[7]
    corpus[33] = 0.0                     # zero-norm rows have distance 0 -> score 1.0 (DistCosine)
    corpus[3999] = 0.0
    queries = np.stack([corpus[7] * 0.5, rng.standard_normal(48).astype(np.float32), corpus[1234]])
    store = B200Store.new(tmp_path, dim=48)
    store.add_matrix(corpus)
    check_parity(store, corpus, queries, 10)
    ids, scores, _ = store.search_matrix(queries[:1], 10)
    # the zero rows tie at d == 0 with the exact duplicates of the query direction; all score 1.0 or within an ulp
    top = list(ids[0][:5])
    assert 34 in top and 4000 in top and top.index(34) < top.index(4000)
    assert scores[0][top.index(34)] == 1.0 and scores[0][top.index(4000)] == 1.0


@pytest.mark.parametrize("dtype", ["f32", "f16"])
def test_exact_fold_on_rows_spanning_many_decades(tmp_path, dtype):
    """The exact re-score folds the f32 products left to right in f64, the reference's order.  On ordinary rows every
    partial sum is exactly representable and any order would give the same bits; on rows whose elements span 18 decades
    the order of the additions decides the last bits.  Both kinds, mixed in one candidate list, must give the oracle's
    bits."""
    rng = np.random.default_rng(77)
    n, d = 6000, 200
    x = rng.standard_normal((n, d)).astype(np.float32)
    wide = rng.random(n) < 0.5
    decades = (-4, 4) if dtype == "f16" else (-12, 6)          # f16 rows: stay inside the format's range
    x[wide] *= (10.0 ** rng.uniform(*decades, size=(int(wide.sum()), d))).astype(np.float32)
    x[5] = 0.0
    x[6, 1:] = 0.0                                             # a single non-zero product
    queries = rng.standard_normal((12, d)).astype(np.float32)
    queries[:6] *= (10.0 ** rng.uniform(-8, 4, size=(6, d))).astype(np.float32)
    queries[11] = x[6]
    store = B200Store.new(tmp_path, dim=d, dtype=dtype)
    store.add_matrix(x)
    stored = x if dtype == "f32" else x.astype(np.float16).astype(np.float32)
    for k in (1, 10, 40):
        check_parity(store, stored, queries, k)                # batch: every query against the oracle, bits and all
        check_parity(store, stored, queries[3:4], k)
        check_parity(store, stored, queries[8:9], k)
    store.close()


def test_zero_query_returns_first_rows(tmp_path):
    corpus = unit_rows(500, 16, 1)
    store = B200Store.new(tmp_path, dim=16)
    store.add_matrix(corpus)
    check_parity(store, corpus, np.zeros((1, 16), np.float32), 5)


def test_dot_metric(tmp_path):
    rng = np.random.default_rng(6)
    corpus = rng.standard_normal((3000, 96)).astype(np.float32)
    queries = rng.standard_normal((4, 96)).astype(np.float32)
    store = B200Store.new(tmp_path, dim=96, metric="dot")
    store.add_matrix(corpus)
    check_parity(store, corpus, queries, 10, metric="dot")


def test_incremental_inserts_match_bulk(tmp_path):
    corpus = unit_rows(300, 32, 2)
    q = unit_rows(2, 32, 3)
    a = B200Store.new(tmp_path / "a", dim=32, capacity=4)          # forces several regrowths
    for i in range(0, 300, 7):
        first = a.add_matrix(corpus[i:i + 7])
        assert first == i + 1                                      # next_id = len + 1 (local.rs:63)
    check_parity(a, corpus, q, 10)


def test_empty_store_and_bad_arguments(tmp_path):
    store = B200Store.new(tmp_path, dim=8)
    ids, scores, counts = store.search_matrix(np.ones((2, 8), np.float32), 4)
    assert (counts == 0).all() and (ids == 0).all()
    assert store.search([1.0] * 8, 4) == []
    L = capi.lib()
    assert L.mx_store_search(store.handle, None, 1, 1, None, None, None) == capi.ERR_INVALID
    q = np.ones((1, 8), np.float32)
    i = np.zeros((1, 300), np.uint64); s = np.zeros((1, 300), np.float32); c = np.zeros(1, np.uint32)
    assert L.mx_store_search(store.handle, q.ctypes.data, 1, 300, i.ctypes.data, s.ctypes.data, c.ctypes.data) == capi.ERR_INVALID
    assert L.mx_store_search(store.handle, q.ctypes.data, 1, 0, i.ctypes.data, s.ctypes.data, c.ctypes.data) == capi.ERR_INVALID
    with pytest.raises(VectorStoreError) as ei:
        store.add_matrix(np.ones((2, 9), np.float32))
    assert ei.value.variant == "InsertionError"
    bad = np.ones((3, 8), np.float32); bad[1, 2] = np.nan
    with pytest.raises(VectorStoreError) as ei:
        store.add_matrix(bad)
    assert ei.value.variant == "InsertionError" and store.get_nb_point() == 0
    qn = np.ones((1, 8), np.float32); qn[0, 0] = np.inf
    store.add_matrix(np.ones((2, 8), np.float32))
    with pytest.raises(VectorStoreError) as ei:
        store.search_matrix(qn, 1)
    assert ei.value.variant == "SearchError"
    # the vectorised finiteness check names the FIRST bad query of a batch (NaN, +-inf, any position) and leaves the
    # store usable
    qb = np.ones((5, 8), np.float32); qb[3, 7] = -np.inf; qb[4, 0] = np.nan
    with pytest.raises(VectorStoreError) as ei:
        store.search_matrix(qb, 1)
    assert ei.value.variant == "SearchError" and "query 3" in str(ei.value)
    ids, _, counts = store.search_matrix(np.ones((5, 8), np.float32), 1)
    assert (counts == 1).all() and (ids[:, 0] == 1).all()


def test_sharded_ids_and_device_merge(tmp_path):
    """row sharding: two stores with id_offset / id_stride + mx_merge_topk_device == one store"""
    import torch
    corpus = unit_rows(6000, 64, 8)
    queries = unit_rows(5, 64, 9)
    k = 10
    whole = B200Store.new(tmp_path / "w", dim=64)
    whole.add_matrix(corpus)
    wi, ws, wc = whole.search_matrix(queries, k)
    for mode in ("contiguous", "round_robin"):
        if mode == "contiguous":
            shards = [B200Store.new(tmp_path / f"c{r}", dim=64, id_offset=3000 * r) for r in range(2)]
            parts = [corpus[:3000], corpus[3000:]]
        else:
            shards = [B200Store.new(tmp_path / f"r{r}", dim=64, id_offset=r, id_stride=2) for r in range(2)]
            parts = [corpus[0::2], corpus[1::2]]
        qd = torch.from_numpy(queries).cuda()
        ids = torch.zeros((2, 5, k), dtype=torch.int64, device="cuda")
        dists = torch.zeros((2, 5, k), dtype=torch.float32, device="cuda")
        scores = torch.zeros((2, 5, k), dtype=torch.float32, device="cuda")
        counts = torch.zeros((2, 5), dtype=torch.int32, device="cuda")
        L = capi.lib()
        for r in range(2):
            shards[r].add_matrix(parts[r])
            rc = L.mx_store_search_device(shards[r].handle, qd.data_ptr(), 5, k, ids[r].data_ptr(), scores[r].data_ptr(),
                                          dists[r].data_ptr(), counts[r].data_ptr(), None)
            assert rc == 0
            assert L.mx_store_sync(shards[r].handle) == 0
        oi = torch.zeros((5, k), dtype=torch.int64, device="cuda")
        os_ = torch.zeros((5, k), dtype=torch.float32, device="cuda")
        oc = torch.zeros(5, dtype=torch.int32, device="cuda")
        rc = L.mx_merge_topk_device(ids.data_ptr(), dists.data_ptr(), counts.data_ptr(), 2, 5, k, capi.METRIC_COSINE,
                                    oi.data_ptr(), os_.data_ptr(), oc.data_ptr(), 0, None)
        assert rc == 0
        torch.cuda.synchronize()
        np.testing.assert_array_equal(oi.cpu().numpy().astype(np.uint64), wi)
        np.testing.assert_array_equal(os_.cpu().numpy().view(np.uint32), ws.view(np.uint32))
        np.testing.assert_array_equal(oc.cpu().numpy().astype(np.uint32), wc)


def test_peer_memory_exchange_kernels(tmp_path):
    """the peer-memory form of the exchange step (mx_exchange_push_device + mx_merge_topk_blobs_wait_device) with three
    "ranks" living on one GPU: every rank's exchange buffer ends up with all blobs, flags carry the epoch, and the merged
    answer equals the single-store answer over several epochs (both slot sets)"""
    import torch
    world, nq, k, d = 3, 7, 10, 48
    corpus = unit_rows(9000, d, 21)
    L = capi.lib()
    whole = B200Store.new(tmp_path / "w", dim=d)
    whole.add_matrix(corpus)
    shards = []
    for r in range(world):
        st = B200Store.new(tmp_path / f"s{r}", dim=d, id_offset=3000 * r)
        st.add_matrix(corpus[3000 * r:3000 * (r + 1)])
        shards.append(st)
    blob = int(L.mx_topk_blob_bytes(nq, k))
    stride = (blob + 15) & ~15
    xbufs = [torch.zeros(2 * world * stride + 64, dtype=torch.uint8, device="cuda") for _ in range(world)]
    peers = (C.c_uint64 * world)(*[x.data_ptr() for x in xbufs])
    mine = [torch.zeros(stride, dtype=torch.uint8, device="cuda") for _ in range(world)]
    oi = torch.zeros((nq, k), dtype=torch.int64, device="cuda")
    os_ = torch.zeros((nq, k), dtype=torch.float32, device="cuda")
    oc = torch.zeros(nq, dtype=torch.int32, device="cuda")
    for epoch in range(1, 5):
        queries = unit_rows(nq, d, 100 + epoch)
        qd = torch.from_numpy(queries).cuda()
        wi, ws, wc = whole.search_matrix(queries, k)
        half = (epoch & 1) * world * stride
        for r in range(world):
            assert L.mx_store_search_blob_device(shards[r].handle, qd.data_ptr(), nq, k, mine[r].data_ptr(), None) == 0
            assert L.mx_store_sync(shards[r].handle) == 0
            assert L.mx_exchange_push_device(mine[r].data_ptr(), stride, peers, world, r, half + r * stride, 2 * world * stride,
                                             epoch, 0, None) == 0
        for r in range(world):
            base = xbufs[r].data_ptr()
            rc = L.mx_merge_topk_blobs_wait_device(base + half, stride, world, nq, k, capi.METRIC_COSINE, oi.data_ptr(),
                                                   os_.data_ptr(), oc.data_ptr(), base + 2 * world * stride, epoch, 0, None)
            assert rc == 0
            torch.cuda.synchronize()
            np.testing.assert_array_equal(oi.cpu().numpy().astype(np.uint64), wi)
            np.testing.assert_array_equal(os_.cpu().numpy().view(np.uint32), ws.view(np.uint32))
            np.testing.assert_array_equal(oc.cpu().numpy().astype(np.uint32), wc)
            flags = xbufs[r][2 * world * stride:2 * world * stride + 4 * world].view(torch.int32).cpu().numpy()
            assert (flags == epoch).all()
    # bad shapes are status codes, not crashes
    assert L.mx_exchange_push_device(mine[0].data_ptr(), stride + 1, peers, world, 0, 0, 0, 1, 0, None) == capi.ERR_INVALID
    assert L.mx_exchange_push_device(mine[0].data_ptr(), stride, peers, world, world, 0, 0, 1, 0, None) == capi.ERR_INVALID


def test_full_size_properties_1m(tmp_path):
    """BASELINE config 2 size (1M x 384 f32): size-independent properties + spot-checked scores."""
    n, d, k = 1_000_000, 384, 10
    rng = np.random.default_rng(1234)
    corpus = rng.standard_normal((n, d), dtype=np.float32)
    corpus /= np.linalg.norm(corpus, axis=1, keepdims=True)
    store = B200Store.new(tmp_path, dim=d, capacity=n)
    for i in range(0, n, 250_000):
        store.add_matrix(corpus[i:i + 250_000])
    assert store.get_nb_point() == n
    rows = np.array([0, 17, 499_999, 999_999])
    queries = corpus[rows] + 0.05 * rng.standard_normal((4, d)).astype(np.float32)
    ids, scores, counts = store.search_matrix(queries, k)
    assert (counts == k).all()
    np.testing.assert_array_equal(ids[:, 0], rows + 1)                       # planted neighbour found
    assert (np.diff(scores, axis=1) <= 0).all()                              # sorted best first
    for i in range(4):
        assert len(set(ids[i])) == k
        np.testing.assert_array_equal(cosine.scores_of(corpus, queries[i], ids[i]).view(np.uint32),
                                      scores[i].view(np.uint32))             # bit-exact scores
    # exactness of the k-th boundary, checked with the full oracle for one query
    oi, os_, _ = cosine.exact_topk(corpus, queries[:1], k)
    np.testing.assert_array_equal(ids[:1], oi)
    np.testing.assert_array_equal(scores[:1].view(np.uint32), os_.view(np.uint32))
    # idempotence + batch == singles
    ids2, scores2, _ = store.search_matrix(queries, k)
    np.testing.assert_array_equal(ids, ids2)
    for i in range(4):
        si, ss, _ = store.search_matrix(queries[i:i + 1], k)
        np.testing.assert_array_equal(si[0], ids[i])
        np.testing.assert_array_equal(ss[0].view(np.uint32), scores[i].view(np.uint32))


# ---- K2: the tcgen05 batched scan (fp16 stores, nq >= 8) ------------------------------------------

@pytest.mark.parametrize("n,d,nq,k,metric", [
    (20000, 384, 64, 10, "cosine"), (128, 384, 8, 10, "cosine"), (129, 64, 9, 3, "cosine"),
    (50001, 384, 128, 10, "cosine"), (30000, 384, 200, 10, "cosine"), (7777, 100, 33, 20, "cosine"),
    (40000, 256, 64, 26, "cosine"), (25000, 384, 64, 10, "dot"), (100, 8, 16, 10, "cosine"),
    # dim > 384: the M = 64 form of the kernel (64 queries per pass), e5-base / BERT-base dimension 768
    (30000, 768, 64, 10, "cosine"), (9001, 768, 100, 10, "cosine"), (5000, 512, 9, 20, "dot"), (700, 448, 130, 10, "cosine"),
    # one k-block per tile: the TMA producer runs a dozen tiles ahead of the epilogue (inverse-norm slot depth), M = 128 form
    (300000, 32, 100, 10, "cosine"),
])
def test_parity_tcgen05_scan(tmp_path, n, d, nq, k, metric):
    """the batched tensor-core path gives the same ids and bit-identical scores as the oracle"""
    rng = np.random.default_rng(n + d + nq)
    corpus = (rng.standard_normal((n, d)) * rng.uniform(0.5, 2.0, (n, 1))).astype(np.float32)   # unnormalised rows
    queries = (corpus[rng.integers(0, n, nq)] + 0.3 * rng.standard_normal((nq, d))).astype(np.float32)
    store = B200Store.new(tmp_path, dim=d, dtype="f16", metric=metric)
    store.add_matrix(corpus)
    assert capi.lib().mx_store_scan_path(store.handle, nq, k, -1) == 2
    stored = corpus.astype(np.float16).astype(np.float32)
    check_parity(store, stored, queries, k, metric=metric)
    # and agrees with the CUDA-core stream path on the same store
    a = store.search_matrix(queries, k)
    assert capi.lib().mx_store_scan_path(store.handle, nq, k, 1) == 1
    b = store.search_matrix(queries, k)
    capi.lib().mx_store_scan_path(store.handle, nq, k, -1)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


def test_tcgen05_scan_duplicates_and_growth(tmp_path):
    """many exact duplicates spread over the CTAs' row ranges: ties resolve to the lowest ids; the store
    regrows between searches (tensor maps are rebuilt per launch)"""
    rng = np.random.default_rng(77)
    d, nq, k = 128, 16, 10
    corpus = rng.standard_normal((60000, d)).astype(np.float32)
    dup = rng.standard_normal(d).astype(np.float32)
    dup_rows = rng.choice(60000, 300, replace=False)
    corpus[dup_rows] = dup
    queries = np.tile(dup, (nq, 1)) + 0.01 * rng.standard_normal((nq, d)).astype(np.float32)
    store = B200Store.new(tmp_path, dim=d, dtype="f16", capacity=1000)
    stored = corpus.astype(np.float16).astype(np.float32)
    for lo in range(0, 60000, 20000):
        store.add_matrix(corpus[lo:lo + 20000])
        check_parity(store, stored[:lo + 20000], queries, k)
    ids, _, _ = store.search_matrix(queries, k)
    np.testing.assert_array_equal(ids[0], np.sort(dup_rows)[:k] + 1)


@pytest.mark.parametrize("n,d,nq,k,metric", [
    (400_000, 64, 64, 10, "cosine"),    # 21 tiles per CTA -> 4 sample tiles (min(4, tiles / 4))
    (700_000, 64, 16, 26, "cosine"),    # 37 tiles per CTA -> 4 sample tiles, L = 32 lists (tau0 = 16th of the second bests)
    (1_300_001, 32, 64, 10, "dot"),     # 68 tiles per CTA -> 4 sample tiles, ragged last tile, dot metric
    (120_000, 96, 32, 10, "cosine"),    # 6 tiles per CTA -> 1 sample tile (the smallest seeded shape)
])
def test_tcgen05_scan_threshold_seeding(tmp_path, n, d, nq, k, metric):
    """shards large enough for the seeded scan (two-best sampling pass + grid barrier + tau0): full oracle parity with
    exact ties planted at the top -- duplicates of the query rows inside and outside the sampled tiles -- and with
    near-duplicates crowding the threshold"""
    rng = np.random.default_rng(n % 1000 + d + k)
    corpus = rng.standard_normal((n, d)).astype(np.float32)
    picks = rng.integers(0, n, nq)
    queries = corpus[picks].copy()
    # exact duplicates of the first 8 query rows: some in the very first tiles (the sampled ones), some anywhere
    for j in range(8):
        where = np.concatenate([rng.integers(0, 148 * 128, 3), rng.integers(0, n, 5)])
        corpus[where] = corpus[picks[j]]
    # a crowd of near-ties for query 8: 15 rows that differ only by fp16 rounding (fewer than the 22 spare places the
    # rerank list has beyond k = 10 -- the documented limit of the approximate stage)
    crowd = rng.integers(0, n, 15)
    corpus[crowd] = corpus[picks[8]] * (1 + 1e-3 * rng.standard_normal((15, 1)).astype(np.float32))
    queries = corpus[picks].copy()
    store = B200Store.new(tmp_path, dim=d, dtype="f16", metric=metric, capacity=n)
    store.add_matrix(corpus)
    assert capi.lib().mx_store_scan_path(store.handle, nq, k, -1) == 2
    stored = corpus.astype(np.float16).astype(np.float32)
    check_parity(store, stored, queries, k, metric=metric)
    # a second, unrelated batch on the same store (the barrier counter and tau are reset per launch)
    q2 = rng.standard_normal((nq, d)).astype(np.float32)
    check_parity(store, stored, q2, k, metric=metric)


def test_full_size_properties_tcgen05_2m(tmp_path):
    """2 M x 384 fp16, 64 queries: planted neighbours found, scores bit-exact, one query checked in full"""
    n, d, k, nq = 2_000_000, 384, 10, 64
    rng = np.random.default_rng(99)
    store = B200Store.new(tmp_path, dim=d, dtype="f16", capacity=n)
    keep = []
    for i in range(0, n, 250_000):
        part = rng.standard_normal((250_000, d), dtype=np.float32)
        part /= np.linalg.norm(part, axis=1, keepdims=True)
        store.add_matrix(part)
        keep.append(part.astype(np.float16))
    stored16 = np.concatenate(keep)
    rows = rng.integers(0, n, nq)
    queries = stored16[rows].astype(np.float32) + 0.03 * rng.standard_normal((nq, d)).astype(np.float32)
    assert capi.lib().mx_store_scan_path(store.handle, nq, k, -1) == 2
    ids, scores, counts = store.search_matrix(queries, k)
    assert (counts == k).all()
    np.testing.assert_array_equal(ids[:, 0], rows + 1)
    assert (np.diff(scores, axis=1) <= 0).all()
    stored = stored16.astype(np.float32)
    for i in (0, 31, 63):
        np.testing.assert_array_equal(cosine.scores_of(stored, queries[i], ids[i]).view(np.uint32), scores[i].view(np.uint32))
    oi, os_, _ = cosine.exact_topk(stored, queries[:2], k)
    np.testing.assert_array_equal(ids[:2], oi)
    np.testing.assert_array_equal(scores[:2].view(np.uint32), os_.view(np.uint32))


# ---- seeded-scan tie order, the superset certificate and the exact fallback -----------------------

def _verify_stats(store):
    q, f = C.c_uint64(), C.c_uint64()
    assert capi.lib().mx_store_verify_stats(store.handle, C.byref(q), C.byref(f)) == 0
    return q.value, f.value


def test_tcgen05_seeded_scan_keeps_lowest_ids_of_a_tie_group(tmp_path):
    """20 bit-identical rows inside ONE CTA's stripe -- two of them in the tile the sampling pass
    scans (rows 3 and 70 of CTA 0's first tile), the rest in CTA 0's later tiles (tile 148 j) -- more than the L = 16 list
    holds.  The scan alone (certificate off) must already keep the lowest rows: a CTA visits its rows in increasing
    order, so the strict admission test never trades a lower row for an equal-score higher one."""
    n, d, nq, k = 400_000, 64, 16, 10
    rng = np.random.default_rng(5)
    corpus = rng.standard_normal((n, d)).astype(np.float32)
    dup = rng.standard_normal(d).astype(np.float32)
    dup_rows = np.array([3, 70] + [148 * 128 * j + int(rng.integers(0, 128)) for j in range(1, 19)])
    corpus[dup_rows] = dup
    queries = rng.standard_normal((nq, d)).astype(np.float32)
    queries[0] = dup
    queries[5] = dup * 3.0 + 0.001 * rng.standard_normal(d).astype(np.float32)
    store = B200Store.new(tmp_path, dim=d, dtype="f16", capacity=n)
    store.add_matrix(corpus)
    L = capi.lib()
    assert L.mx_store_scan_path(store.handle, nq, k, -1) == 2
    stored = corpus.astype(np.float16).astype(np.float32)
    want = np.sort(dup_rows)[:k] + 1
    assert L.mx_store_set_verify(store.handle, 0) == 0
    ids, scores, _ = store.search_matrix(queries, k)
    np.testing.assert_array_equal(ids[0], want)
    np.testing.assert_array_equal(ids[5], want)
    oi, os_, _ = cosine.exact_topk(stored, queries, k)
    np.testing.assert_array_equal(ids, oi)
    np.testing.assert_array_equal(scores.view(np.uint32), os_.view(np.uint32))
    # with the certificate on the answer is the same, and the tie group that spills over the lists is noticed
    assert L.mx_store_set_verify(store.handle, 1) == 0
    check_parity(store, stored, queries, k)
    _, flagged = _verify_stats(store)
    assert flagged >= 2


def _near_tie_crowd(rng, q, d, m, alpha=0.8):
    """m rows at (in real arithmetic) the SAME cosine alpha to q: alpha q + sqrt(1 - alpha^2) h_i, h_i unit, h_i . q = 0.
    Rounding the rows for storage then decides their exact order; rounding the QUERY to fp16 (tcgen05 scan) or
    accumulating in f32 (stream scan) shuffles the approximate one."""
    qn = q / np.linalg.norm(q)
    h = rng.standard_normal((m, d))
    h -= (h @ qn)[:, None] * qn[None, :]
    h /= np.linalg.norm(h, axis=1, keepdims=True)
    return (alpha * qn[None, :] + np.sqrt(1 - alpha * alpha) * h).astype(np.float32)


def test_certificate_catches_a_near_tie_crowd_tcgen05(tmp_path):
    """40 rows within ~1e-5 of each other inside one CTA's stripe, more than the 16-entry list holds and
    closer than the fp16-query approximation resolves.  The certificate must flag the query and the exact scan must
    return the reference's order; the other queries of the batch stay on the fast path."""
    d, nq, k = 384, 16, 10
    n = 148 * 128 * 6
    rng = np.random.default_rng(11)
    corpus = unit_rows(n, d, 12)
    queries = unit_rows(nq, d, 13)
    crowd = _near_tie_crowd(rng, queries[3].astype(np.float64), d, 40)
    rows = np.array([(5 + 148 * j) * 128 + i for j in range(4) for i in range(0, 120, 12)])   # tiles 5, 153, 301, 449: CTA 5
    corpus[rows] = crowd
    store = B200Store.new(tmp_path, dim=d, dtype="f16", capacity=n)
    store.add_matrix(corpus)
    L = capi.lib()
    assert L.mx_store_scan_path(store.handle, nq, k, -1) == 2
    stored = corpus.astype(np.float16).astype(np.float32)
    oi, os_, _ = cosine.exact_topk(stored, queries, k)
    assert set(oi[3] - 1) <= set(rows)          # the crowd IS the answer of query 3
    q0, f0 = _verify_stats(store)
    check_parity(store, stored, queries, k)
    q1, f1 = _verify_stats(store)
    assert q1 - q0 == nq and 1 <= f1 - f0 <= 2
    # the device-buffer call (fallback kernels always enqueued) gives the same answer
    import torch
    qd = torch.from_numpy(queries).cuda()
    ids_d = torch.zeros((nq, k), dtype=torch.int64, device="cuda")
    sc_d = torch.zeros((nq, k), dtype=torch.float32, device="cuda")
    cn_d = torch.zeros(nq, dtype=torch.int32, device="cuda")
    assert L.mx_store_search_device(store.handle, qd.data_ptr(), nq, k, ids_d.data_ptr(), sc_d.data_ptr(), None,
                                    cn_d.data_ptr(), None) == 0
    assert L.mx_store_sync(store.handle) == 0
    np.testing.assert_array_equal(ids_d.cpu().numpy().astype(np.uint64), oi)
    np.testing.assert_array_equal(sc_d.cpu().numpy().view(np.uint32), os_.view(np.uint32))
    # what the approximate stage alone would have answered: recorded, not asserted (it may get lucky)
    L.mx_store_set_verify(store.handle, 0)
    ids_off, _, _ = store.search_matrix(queries, k)
    L.mx_store_set_verify(store.handle, 1)
    print("crowd query without the certificate:", "same" if (ids_off[3] == oi[3]).all() else "DIFFERENT", "ids")
    np.testing.assert_array_equal(np.delete(ids_off, 3, 0), np.delete(oi, 3, 0))


@pytest.mark.parametrize("dtype", ["f32", "f16"])
def test_certificate_catches_a_near_tie_crowd_stream_scan(tmp_path, dtype):
    """the same for the CUDA-core stream scan (single query): 100 contiguous near-tie rows sit in one CTA's range, its
    32-entry list cannot hold them, f32 accumulation cannot order them"""
    d, k, n = 384, 10, 60_000
    rng = np.random.default_rng(21)
    corpus = unit_rows(n, d, 22)
    q = unit_rows(1, d, 23)
    corpus[1000:1100] = _near_tie_crowd(rng, q[0].astype(np.float64), d, 100)
    store = B200Store.new(tmp_path, dim=d, dtype=dtype, capacity=n)
    store.add_matrix(corpus)
    stored = corpus if dtype == "f32" else corpus.astype(np.float16).astype(np.float32)
    _, f0 = _verify_stats(store)
    check_parity(store, stored, q, k)
    _, f1 = _verify_stats(store)
    assert f1 - f0 == 1
    check_parity(store, stored, unit_rows(3, d, 24), k)       # ordinary queries are not flagged
    _, f2 = _verify_stats(store)
    assert f2 == f1


@pytest.mark.parametrize("metric", ["cosine", "dot"])
def test_mass_duplicates_answered_by_the_exact_scan(tmp_path, metric):
    """memex has no dedup: the same document ingested thousands of times.  5000 bit-identical rows tie for every place
    of the top k; whatever the lists kept, the exact scan ranks them by (distance, id)"""
    d, nq, k, n = 128, 12, 10, 120_000
    rng = np.random.default_rng(31)
    corpus = rng.standard_normal((n, d)).astype(np.float32)
    dup = rng.standard_normal(d).astype(np.float32)
    dup_rows = np.sort(rng.choice(n, 5000, replace=False))
    corpus[dup_rows] = dup
    queries = rng.standard_normal((nq, d)).astype(np.float32)
    queries[7] = dup + 0.01 * rng.standard_normal(d).astype(np.float32)
    store = B200Store.new(tmp_path, dim=d, dtype="f16", metric=metric, capacity=n)
    store.add_matrix(corpus)
    stored = corpus.astype(np.float16).astype(np.float32)
    check_parity(store, stored, queries, k, metric=metric)
    ids, _, _ = store.search_matrix(queries, k)
    np.testing.assert_array_equal(ids[7], dup_rows[:k] + 1)
    _, flagged = _verify_stats(store)
    assert flagged >= 2
    # k = 40 goes through the stream scan's 64-entry lists (E = 2) and the E = 2 exact scan
    check_parity(store, stored, queries[6:8], 40, metric=metric)


def test_ordinary_corpus_is_never_flagged(tmp_path):
    """the certificate costs nothing on well-separated data: no query of the parity cases above takes the exact scan"""
    n, d, nq, k = 200_000, 384, 64, 10
    corpus = unit_rows(n, d, 41)
    rng = np.random.default_rng(42)
    queries = corpus[rng.integers(0, n, nq)] + 0.1 * unit_rows(nq, d, 43)
    store = B200Store.new(tmp_path, dim=d, dtype="f16", capacity=n)
    store.add_matrix(corpus)
    stored = corpus.astype(np.float16).astype(np.float32)
    check_parity(store, stored, queries, k)
    check_parity(store, stored, queries[:1], k)
    qn, flagged = _verify_stats(store)
    assert qn == nq + 1 and flagged == 0


def test_submit_collect_two_in_flight_equals_the_blocking_call(tmp_path):
    """mx_store_search_submit / _collect: same answer, bit for bit, as mx_store_search, with two searches in flight; a third
    submit is refused; an empty store and a non-finite query behave as in the blocking call."""
    n, d, k = 300_000, 384, 10
    x = unit_rows(n, d, 31)
    store = B200Store.new(tmp_path, dim=d, dtype="f16")
    empty = store.search_collect(store.search_submit(x[:3], k))
    assert (empty[2] == 0).all() and (empty[0] == 0).all()
    store.add_matrix(x)
    batches = [unit_rows(nq, d, 40 + i) for i, nq in enumerate((64, 64, 8, 1, 64))]   # tcgen05 and stream paths
    want = [store.search_matrix(q, k) for q in batches]
    got, prev = [], None
    for q in batches:
        t = store.search_submit(q, k)
        if prev is not None:
            got.append(store.search_collect(prev))
        prev = t
    got.append(store.search_collect(prev))
    for (wi, ws, wc), (gi, gs, gc) in zip(want, got):
        assert (wi == gi).all() and (ws.view(np.uint32) == gs.view(np.uint32)).all() and (wc == gc).all()
    t0 = store.search_submit(batches[0], k)
    t1 = store.search_submit(batches[1], k)
    with pytest.raises(VectorStoreError):
        store.search_submit(batches[2], k)
    with pytest.raises(VectorStoreError):
        store.search_collect((t0[0] + 7, 64, k))
    a, b = store.search_collect(t0), store.search_collect(t1)
    assert (a[0] == want[0][0]).all() and (b[0] == want[1][0]).all()
    bad = batches[0].copy(); bad[5, 9] = np.nan
    with pytest.raises(VectorStoreError) as ei:
        store.search_submit(bad, k)
    assert "query 5" in str(ei.value)
    again = store.search_collect(store.search_submit(batches[3], k))
    assert (again[0] == want[3][0]).all()
