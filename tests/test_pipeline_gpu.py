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
