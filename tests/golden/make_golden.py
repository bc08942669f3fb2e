"""Generates the committed golden fixtures under tests/golden/.

Run in the build container (CPU):  python tests/golden/make_golden.py

* search_ref_fixture.json -- the reference's only results fixture for the search path
  (lib/libmemex/src/storage/local.rs:175-214, `test_hnsw`): the three vectors, the query, the
  ranking the reference asserts (first hit "test-two") and the scores DistCosine + local.rs:86
  give, computed by the oracle's independent numpy restatement (oracle/cosine.py).
* search_small.npz -- seeded 257 x 24 corpus + 5 queries with the exact top-7 from the numpy
  restatement (ids, scores as f32 bit patterns).
* encoder_tiny.npz / encoder_l6.npz -- HF transformers BertModel (torch CPU fp32; the libtorch
  kernels tch dispatches to) outputs for seeded weights and inputs (oracle/encoder.py hf_encode).
* encoder_tiny_{roberta,distiluse,albert}.npz -- the same for HF RobertaModel, DistilBertModel +
  Dense/Tanh and AlbertModel (AllDistilrobertaV1, DistiluseBaseMultilingualCased,
  ParaphraseAlbertSmallV2 of the reference's enum).

The reference itself (Rust; rust-bert / hnsw_rs are un-vendored crates) cannot be imported or
built here, so these are outputs of independent restatements of its published algorithm.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import cosine, encoder  # noqa: E402


def main():
    data = [("test-one", [0.0, 0.1, 0.2]), ("test-two", [0.1, 0.1, 0.1]), ("test-three", [0.3, 0.2, 0.1])]
    query = [0.1, 0.1, 0.1]
    corpus = np.array([v for _, v in data], dtype=np.float32)
    top = cosine.exact_topk_np(corpus, np.array(query, dtype=np.float32), 3)
    fixture = {
        "source": "reference lib/libmemex/src/storage/local.rs:175-214",
        "data": [{"_id": i, "vector": v} for i, v in data],
        "query": query,
        "limit": 3,
        "expected_ids": [data[i - 1][0] for i, _ in top],
        "expected_scores": [float(s) for _, s in top],
        "expected_score_bits": [int(np.float32(s).view(np.uint32)) for _, s in top],
        "reference_asserts": {"len": 3, "first": "test-two"},
    }
    with open(os.path.join(HERE, "search_ref_fixture.json"), "w") as f:
        json.dump(fixture, f, indent=1)

    rng = np.random.default_rng(99)
    c = rng.standard_normal((257, 24)).astype(np.float32)
    c[17] = c[3]                      # exact duplicate -> tie broken by id
    c[40] = 0.0                       # zero-norm row -> distance 0 (DistCosine)
    q = rng.standard_normal((5, 24)).astype(np.float32)
    q[4] = c[100] * 3.0
    ids = np.zeros((5, 7), dtype=np.uint64)
    bits = np.zeros((5, 7), dtype=np.uint32)
    for i in range(5):
        t = cosine.exact_topk_np(c, q[i], 7)
        ids[i] = [a for a, _ in t]
        bits[i] = [int(np.float32(s).view(np.uint32)) for _, s in t]
    np.savez(os.path.join(HERE, "search_small.npz"), corpus=c, queries=q, ids=ids, score_bits=bits)

    for name, cfg, B, S, seed in (("encoder_tiny", encoder.TINY, 5, 24, 3), ("encoder_l6", encoder.MINILM_L6, 4, 48, 5),
                                  # the other stacks of the enum (embedding.rs:24-55): HF RobertaModel / DistilBertModel
                                  # (+ Dense, Tanh) / AlbertModel outputs
                                  ("encoder_tiny_roberta", encoder.TINY_ROBERTA, 5, 24, 13),
                                  ("encoder_tiny_distiluse", encoder.TINY_DISTILUSE, 5, 24, 14),
                                  ("encoder_tiny_albert", encoder.TINY_ALBERT, 5, 24, 15)):
        if os.path.exists(os.path.join(HERE, name + ".npz")) and name in ("encoder_tiny", "encoder_l6"):
            continue   # committed in an earlier pass; the generator is deterministic, the files stay byte-stable
        w = encoder.make_weights(cfg, seed=seed)
        ids_, lens = encoder.make_inputs(cfg, B, S, seed=seed + 100, ragged=True, min_len=3)
        out = encoder.hf_encode(cfg, w, ids_, lens)
        np.savez(os.path.join(HERE, name + ".npz"), ids=ids_, lens=lens, out=out, weight_seed=seed,
                 cfg=json.dumps(cfg.to_dict()))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
