"""CPU tests of the oracle itself (the checker must be pinned before it is trusted).

Search: the reference's only results fixture for this path is `test_hnsw`
(reference lib/libmemex/src/storage/local.rs:175-214).  Encoder: the reference pins no numbers
(SURVEY.md section 8c), so the two independent restatements are checked against each other and
against the committed golden outputs.
"""
import json
import os

import numpy as np
import pytest

from oracle import cosine, encoder

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_reference_test_hnsw_fixture(oracle_lib):
    fx = json.load(open(os.path.join(GOLD, "search_ref_fixture.json")))
    corpus = np.array([d["vector"] for d in fx["data"]], dtype=np.float32)
    names = [d["_id"] for d in fx["data"]]
    ids, scores, counts = oracle_lib.exact_topk(corpus, np.array(fx["query"], dtype=np.float32), fx["limit"])
    assert counts[0] == fx["reference_asserts"]["len"]
    got = [names[int(i) - 1] for i in ids[0]]
    assert got[0] == fx["reference_asserts"]["first"]          # what local.rs:211-212 asserts
    assert got == fx["expected_ids"]
    assert [int(s.view(np.uint32)) for s in scores[0]] == fx["expected_score_bits"]


def test_c_oracle_matches_numpy_restatement(oracle_lib):
    g = np.load(os.path.join(GOLD, "search_small.npz"))
    ids, scores, counts = oracle_lib.exact_topk(g["corpus"], g["queries"], 7)
    assert (counts == 7).all()
    np.testing.assert_array_equal(ids, g["ids"])
    np.testing.assert_array_equal(scores.view(np.uint32), g["score_bits"])
    # spot checks of the scalar entry points
    a, b = g["corpus"][5], g["queries"][1]
    assert np.float32(oracle_lib.dist_cosine(b, a)) == cosine.dist_cosine_np(b, a)


def test_similarity_formula_edge_cases(oracle_lib):
    # local.rs:86: 1 - 1/(1/d); d == 0 -> 1/(inf) = 0 -> similarity exactly 1
    assert oracle_lib.similarity(0.0) == 1.0
    assert cosine.similarity_np(0.0) == np.float32(1.0)
    for d in (1e-7, 0.25, 1.0, 1.999):
        assert np.float32(oracle_lib.similarity(d)) == cosine.similarity_np(d)


def test_zero_vectors_have_distance_zero(oracle_lib):
    z = np.zeros(8, dtype=np.float32)
    v = np.arange(8, dtype=np.float32)
    assert oracle_lib.dist_cosine(z, v) == 0.0
    assert oracle_lib.dist_cosine(v, z) == 0.0


def test_dot_metric(oracle_lib):
    rng = np.random.default_rng(0)
    c = rng.standard_normal((50, 16)).astype(np.float32)
    q = rng.standard_normal((2, 16)).astype(np.float32)
    ids, scores, _ = oracle_lib.exact_topk(c, q, 5, metric="dot")
    ref = (c.astype(np.float64) @ q.astype(np.float64).T).T
    for i in range(2):
        order = np.argsort(-ref[i], kind="stable")[:5]
        np.testing.assert_array_equal(ids[i], order + 1)
        np.testing.assert_allclose(scores[i], ref[i][order], rtol=1e-6)


def test_k_larger_than_corpus(oracle_lib):
    c = np.eye(3, dtype=np.float32)
    ids, scores, counts = oracle_lib.exact_topk(c, c[1], 10)
    assert counts[0] == 3 and ids[0, 0] == 2 and (ids[0, 3:] == 0).all()


def test_hnsw_restatement_recall(oracle_lib):
    """The timed CPU baseline (M=16, efC=200, ef=32 as local.rs:48,76) must actually find neighbours."""
    rng = np.random.default_rng(1234)
    c = rng.standard_normal((1000, 64)).astype(np.float32)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    q = c[:50] + 0.1 * rng.standard_normal((50, 64)).astype(np.float32)
    h = oracle_lib.HnswOracle(64, seed=1)
    h.insert(c)
    assert len(h) == 1000
    ids_h, scores_h, counts_h = h.search(q, 10, ef=32)
    ids_e, scores_e, _ = oracle_lib.exact_topk(c, q, 10)
    recall = np.mean([len(set(ids_h[i]) & set(ids_e[i])) / 10 for i in range(50)])
    assert (counts_h == 10).all()
    assert recall > 0.8
    # scores of the hits it does find are the exact formula's
    i0 = ids_h[0, 0]
    assert scores_h[0, 0] == oracle_lib.scores_of(c, q[0], [i0])[0]


def test_hnsw_parallel_build_matches_serial_quality(oracle_lib):
    """bench.py builds its HNSW sample with several threads (the build is not timed): the graph must be as good"""
    rng = np.random.default_rng(7)
    c = rng.standard_normal((3000, 16)).astype(np.float32)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    q = c[:64] + 0.05 * rng.standard_normal((64, 16)).astype(np.float32)
    ids_e, _, _ = oracle_lib.exact_topk(c, q, 10)
    recalls = []
    for threads in (1, 4):
        h = oracle_lib.HnswOracle(16, seed=3)
        h.insert(c, threads=threads)
        assert len(h) == 3000
        ids_h, _, counts = h.search(q, 10, ef=32, threads=2)
        assert (counts == 10).all()
        recalls.append(np.mean([len(set(ids_h[i]) & set(ids_e[i])) / 10 for i in range(64)]))
    assert recalls[0] > 0.9 and recalls[1] > 0.9, recalls


@pytest.mark.parametrize("name,cfg", [("encoder_tiny", encoder.TINY), ("encoder_l6", encoder.MINILM_L6),
                                      ("encoder_tiny_roberta", encoder.TINY_ROBERTA),
                                      ("encoder_tiny_distiluse", encoder.TINY_DISTILUSE),
                                      ("encoder_tiny_albert", encoder.TINY_ALBERT)])
def test_encoder_restatements_agree_with_golden(name, cfg):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    stored = json.loads(str(g["cfg"]))   # the first fixtures predate the family fields
    assert stored == {k: cfg.to_dict()[k] for k in stored}
    w = encoder.make_weights(cfg, seed=int(g["weight_seed"]))
    out_np = encoder.np_encode(cfg, w, g["ids"], g["lens"])          # numpy, f64 accumulation
    np.testing.assert_allclose(out_np, g["out"], atol=2e-5)          # vs HF/torch fp32 golden
    cos = (out_np * g["out"]).sum(1) / (np.linalg.norm(out_np, axis=1) * np.linalg.norm(g["out"], axis=1))
    assert (cos > 1 - 1e-6).all()
    assert g["out"].shape[1] == (cfg.dense_out or cfg.hidden)
    if cfg.normalize:
        np.testing.assert_allclose(np.linalg.norm(g["out"], axis=1), 1.0, atol=1e-5)


@pytest.mark.parametrize("cfg", [encoder.TINY_ROBERTA, encoder.TINY_DISTILUSE, encoder.TINY_ALBERT])
def test_checkpoint_names_of_the_other_stacks_map_onto_the_c_abi_names(cfg):
    """memex_b200.embedding.canonical_weights (what the host applies to a RoBERTa / DistilBERT / ALBERT checkpoint)
    inverts the oracle's canonical -> HF renaming, on the real state_dict key set of the HF model."""
    from memex_b200.embedding import canonical_weights
    w = encoder.make_weights(cfg, seed=3)
    model = encoder.hf_model(cfg, w)
    state = {k: v.numpy() for k, v in model.state_dict().items()}
    if cfg.dense_out:   # sentence-transformers 2_Dense/ checkpoint names
        state["linear.weight"], state["linear.bias"] = w["dense.linear.weight"], w["dense.linear.bias"]
    back = canonical_weights(cfg.family, {"%s.%s" % (cfg.family, k) if not k.startswith("linear.") else k: v
                                          for k, v in state.items()})
    assert set(back) == set(w)
    for k in w:
        np.testing.assert_array_equal(back[k], w[k])


def test_encoder_hf_matches_golden_tiny():
    g = np.load(os.path.join(GOLD, "encoder_tiny.npz"))
    w = encoder.make_weights(encoder.TINY, seed=int(g["weight_seed"]))
    out = encoder.hf_encode(encoder.TINY, w, g["ids"], g["lens"])
    np.testing.assert_allclose(out, g["out"], atol=1e-6)


def test_encoder_padding_is_ignored():
    cfg = encoder.TINY
    w = encoder.make_weights(cfg, seed=1)
    ids, lens = encoder.make_inputs(cfg, 2, 16, seed=2, ragged=True, min_len=4)
    a = encoder.np_encode(cfg, w, ids, lens)
    ids2 = ids.copy()
    ids2[0, lens[0]:] = 7                                           # garbage in the padded tail
    b = encoder.np_encode(cfg, w, ids2, lens)
    np.testing.assert_allclose(a, b, atol=1e-7)
