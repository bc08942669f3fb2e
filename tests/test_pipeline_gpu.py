"""The whole hot path in one piece: segments -> embedder (GPU) -> vector store (GPU) -> search.

This is the reference's ingest + query flow (worker/src/tasks.rs:15-59: embed the segments, `add_vectors`;
api .../collections/handlers.rs:61-81: `encode_single` the query, `search`) with both halves on the device, and
BASELINE.json's last configuration in miniature: streamed ingest (embed + index) with searches running concurrently.
Checked against the CPU oracle: the embeddings against the numpy encoder, the search results BIT-exactly against
DistCosine evaluated on the very rows the store holds.
"""
import ctypes as C
import threading

import numpy as np
import pytest

from memex_b200 import capi
from memex_b200.embedding import Architecture, B200Encoder
from memex_b200.storage import B200Store, VectorData, VectorStorage
from oracle import cosine
from oracle import encoder as enc_oracle

pytestmark = pytest.mark.gpu


def arch_of(cfg):
    return Architecture(cfg.layers, cfg.hidden, cfg.heads, cfg.ffn, cfg.vocab, cfg.max_pos, cfg.type_vocab, cfg.ln_eps, cfg.normalize)


def stored_rows(store, n):
    out = np.zeros((n, store.dim), dtype=np.float32)
    rc = capi.lib().mx_store_get_rows(store.handle, 0, n, out.ctypes.data)
    assert rc == 0
    return out


@pytest.mark.parametrize("dtype", ["f32", "f16"])
def test_embed_index_search_on_device(tmp_path, dtype):
    """embeddings never visit the host between the encoder and the store: mx_embedder_encode_device -> mx_store_add_device"""
    import torch
    cfg = enc_oracle.MINILM_L6
    w = enc_oracle.make_weights(cfg, seed=11)
    B, S = 24, 96
    ids, lens = enc_oracle.make_inputs(cfg, B, S, seed=12, ragged=True, min_len=4)
    enc = B200Encoder(arch_of(cfg), w, precision="bf16", max_tokens=B * S)
    store = B200Store.new(tmp_path, dim=cfg.hidden, dtype=dtype)
    L = capi.lib()
    ids_d = torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int32)).cuda()
    out_d = torch.zeros((B, cfg.hidden), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream or 1
    lens32 = np.ascontiguousarray(lens, dtype=np.int32)
    assert L.mx_embedder_encode_device(enc.handle, ids_d.data_ptr(), lens32.ctypes.data, B, S, out_d.data_ptr(), st) == 0
    torch.cuda.synchronize()
    first = C.c_uint64()
    assert L.mx_store_add_device(store.handle, out_d.data_ptr(), B, C.byref(first)) == 0
    assert first.value == 1 and store.get_nb_point() == B
    emb = out_d.cpu().numpy()
    # the embedder half against its oracle
    ref = enc_oracle.np_encode(cfg, w, ids, lens)
    assert ((emb * ref).sum(1) >= 1 - 2e-4).all()
    assert np.abs(np.linalg.norm(emb, axis=1) - 1).max() < 1e-5
    # the store half against ITS oracle, on the rows it actually holds
    rows = stored_rows(store, B)
    if dtype == "f32":
        assert (rows.view(np.uint32) == emb.view(np.uint32)).all()
    else:
        assert (rows == emb.astype(np.float16).astype(np.float32)).all()
    queries = emb[[0, 7, 23]] + 0.01 * np.random.default_rng(1).standard_normal((3, cfg.hidden)).astype(np.float32)
    got_i, got_s, got_c = store.search_matrix(queries, 5)
    want_i, want_s, want_c = cosine.exact_topk(rows, queries, 5)
    assert (got_c == want_c).all() and (got_i == want_i).all()
    assert (got_s.view(np.uint32) == want_s.view(np.uint32)).all()
    assert list(got_i[:, 0]) == [1, 8, 24]
    enc.close()
    store.close()


def test_streamed_ingest_with_concurrent_search(tmp_path):
    """one thread embeds and indexes batch after batch, another searches all the while (both behind VectorStorage's one
    lock, as the reference's tokio Mutex, storage/mod.rs:70-92); every answer must be exact for SOME prefix of the stream"""
    cfg = enc_oracle.MINILM_L6
    w = enc_oracle.make_weights(cfg, seed=31)
    enc = B200Encoder(arch_of(cfg), w, precision="bf16", max_tokens=16 * 64)
    store = B200Store.new(tmp_path, dim=cfg.hidden, dtype="f16")
    storage = VectorStorage(store, autosave=False)
    n_batches, B, S = 12, 16, 64
    batches = [enc_oracle.make_inputs(cfg, B, S, seed=100 + i, ragged=True, min_len=8) for i in range(n_batches)]
    all_emb = []
    errors = []
    done = threading.Event()
    probe = None

    def ingest():
        try:
            for i, (ids, lens) in enumerate(batches):
                emb = enc.encode_ids(ids, lens)
                all_emb.append(emb)
                storage.add_vectors([VectorData(_id=f"doc{i}-seg{j}", vector=emb[j]) for j in range(B)])
        except Exception as e:  # noqa: BLE001
            errors.append(e)
        finally:
            done.set()

    answers = []

    def search():
        try:
            while probe is None:
                pass
            while not done.is_set() or len(answers) < 5:
                answers.append(storage.search(probe, 8))
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    first_ids, first_lens = batches[0]
    probe = enc.encode_ids(first_ids[:1], first_lens[:1])[0]
    t1, t2 = threading.Thread(target=ingest), threading.Thread(target=search)
    t2.start()
    t1.start()
    t1.join()
    t2.join()
    assert not errors, errors
    assert store.get_nb_point() == n_batches * B
    emb = np.concatenate(all_emb)
    rows = emb.astype(np.float16).astype(np.float32)
    assert (stored_rows(store, len(rows)) == rows).all()
    names = [f"doc{i}-seg{j}" for i in range(n_batches) for j in range(B)]
    # an answer given while n rows were indexed equals the oracle's answer over exactly those rows
    by_prefix = {}
    for n in range(0, n_batches * B + 1, B):
        if n == 0:
            by_prefix[n] = []
            continue
        i_, s_, c_ = cosine.exact_topk(rows[:n], probe[None, :], 8)
        by_prefix[n] = [(names[int(i) - 1], float(s)) for i, s in zip(i_[0, :c_[0]], s_[0, :c_[0]])]
    assert len(answers) >= 5
    for a in answers:
        assert any(a == ref for ref in by_prefix.values()), a
    assert storage.search(probe, 8) == by_prefix[n_batches * B]
    assert storage.search(probe, 8)[0][0] == "doc0-seg0"
    enc.close()
    store.close()


@pytest.mark.parametrize("model", ["AllMiniLmL6V2", "AllDistilrobertaV1"])
def test_sentence_embedder_from_text(model):
    """SentenceEmbedder::{encode, encode_single} (embedding.rs:138-151) from raw text for two of the three models
    segment_text accepts: WordPiece + BERT, and byte-level BPE + RoBERTa (pad id 1, positions from 2).  Tokenizers are
    built from the committed golden vocabularies; the vectors are checked against the HF oracle on the same ids."""
    import dataclasses
    import importlib.util
    import json
    import os

    from memex_b200.embedding import (ARCHITECTURES, EmbeddingsModelType, ModelConfig, SentenceEmbedder, segment_text,
                                      tokenize_batch)
    gold = os.path.join(os.path.dirname(__file__), "golden")

    def load(name):
        spec = importlib.util.spec_from_file_location(name, os.path.join(gold, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    kind = EmbeddingsModelType[model]
    if kind is EmbeddingsModelType.AllDistilrobertaV1:
        mk = load("make_bpe_golden")
        g = json.load(open(os.path.join(gold, "bpe_golden.json")))
        tok = mk.build(g["vocab"], g["merges"])
        base = enc_oracle.DISTILROBERTA
    else:
        mk = load("make_tokenizer_golden")
        vocab, tok = mk.build()
        from tokenizers import processors
        tok.post_processor = processors.TemplateProcessing(single="[CLS] $A [SEP]", special_tokens=[
            ("[CLS]", vocab.index("[CLS]")), ("[SEP]", vocab.index("[SEP]"))])
        g = json.load(open(os.path.join(gold, "tokenizer_golden.json")))
        base = enc_oracle.MINILM_L6
    cfg = dataclasses.replace(base, layers=2, vocab=tok.get_vocab_size())
    arch = dataclasses.replace(ARCHITECTURES[kind], layers=2, vocab=cfg.vocab, max_seq_length=64)
    w = enc_oracle.make_weights(cfg, seed=61)
    enc = B200Encoder(arch, w, precision="bf16", max_tokens=64 * 64)
    mc = ModelConfig(model=kind, max_length=48, stride=16)
    text = max((c["text"] for c in g["cases"]), key=len)
    handle, embedder = SentenceEmbedder.spawn(mc, enc, tok)
    try:
        results = embedder.encode(text)
        single = embedder.encode_single("the quick brown fox")
    finally:
        embedder.shutdown()
        handle.join(timeout=10)
    segments = segment_text(mc, text, tok)
    assert len(segments) > 4 and [r.content for r in results] == segments
    ids, lens = tokenize_batch(tok, segments, arch.max_seq_length, arch.pad_id)
    assert ids.shape[1] <= 64 and (ids[:, 0] == (0 if arch.family == "roberta" else ids[0, 0])).all()
    ref = enc_oracle.hf_encode(cfg, w, ids, lens)
    got = np.array([r.vector for r in results], dtype=np.float32)
    assert got.shape == ref.shape
    assert ((got * ref).sum(1) >= 1 - 2e-4).all()
    ids1, lens1 = tokenize_batch(tok, ["the quick brown fox"], arch.max_seq_length, arch.pad_id)
    ref1 = enc_oracle.hf_encode(cfg, w, ids1, lens1)
    assert (np.array(single.vector, dtype=np.float32) * ref1[0]).sum() >= 1 - 2e-4
    enc.close()
