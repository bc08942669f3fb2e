"""GPU parity tests of the encoder (through the C ABI) against the CPU oracle.

Tolerances (north_star: cosine within 1e-4 fp32): the f32 path must match the HF/torch golden
outputs to 1e-4 max-abs and 1 - 1e-6 cosine; the 16-bit tensor-core paths are reported against the
same oracle with their own, looser, stated bounds.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from memex_b200 import capi
from memex_b200.embedding import Architecture, B200Encoder, EncodingFailure
from oracle import encoder as enc_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def arch_of(cfg: enc_oracle.EncoderConfig) -> Architecture:
    return Architecture(cfg.layers, cfg.hidden, cfg.heads, cfg.ffn, cfg.vocab, cfg.max_pos, cfg.type_vocab,
                        cfg.ln_eps, cfg.normalize, family=cfg.family, pos_offset=cfg.pos_offset, pad_id=cfg.pad_id,
                        dense_out=cfg.dense_out, dense_act=cfg.dense_act, dense_bias=cfg.dense_bias, ffn_act=cfg.ffn_act,
                        embed_dim=cfg.embed_dim, share_layers=cfg.share_layers, d_kv=cfg.d_kv, rel_buckets=cfg.rel_buckets,
                        rel_max_distance=cfg.rel_max_distance)


def row_cosine(a, b):
    return (a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))


@pytest.mark.parametrize("name,cfg", [("encoder_tiny", enc_oracle.TINY), ("encoder_l6", enc_oracle.MINILM_L6),
                                      ("encoder_tiny_roberta", enc_oracle.TINY_ROBERTA),
                                      ("encoder_tiny_distiluse", enc_oracle.TINY_DISTILUSE),
                                      ("encoder_tiny_albert", enc_oracle.TINY_ALBERT)])
def test_f32_path_matches_golden(name, cfg):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    w = enc_oracle.make_weights(cfg, seed=int(g["weight_seed"]))
    e = B200Encoder(arch_of(cfg), w, precision="f32", max_tokens=4096)
    out = e.encode_ids(g["ids"], g["lens"])
    assert out.shape == g["out"].shape
    assert np.abs(out - g["out"]).max() <= 1e-4
    assert (row_cosine(out, g["out"]) >= 1 - 1e-6).all()
    if cfg.normalize:
        np.testing.assert_allclose(np.linalg.norm(out, axis=1), 1.0, atol=1e-5)


def test_f32_path_vs_numpy_oracle_ragged_and_chunked():
    cfg = enc_oracle.MINILM_L6
    w = enc_oracle.make_weights(cfg, seed=11)
    ids, lens = enc_oracle.make_inputs(cfg, 6, 64, seed=12, ragged=True, min_len=1)
    lens[2] = 1
    ids[2, 1:] = 0
    ref = enc_oracle.np_encode(cfg, w, ids, lens)
    e = B200Encoder(arch_of(cfg), w, precision="f32", max_tokens=3 * 64)   # forces 2 chunks
    out = e.encode_ids(ids, lens)
    assert np.abs(out - ref).max() <= 1e-4
    # padded tail content must not matter
    ids2 = ids.copy()
    ids2[0, lens[0]:] = 999
    np.testing.assert_array_equal(e.encode_ids(ids2, lens), out)


@pytest.mark.parametrize("fmt,tol", [(1, 2e-2), (0, 3e-3)])
@pytest.mark.parametrize("M,N,K,epi", [
    (128, 128, 64, 0), (128, 1152, 384, 0), (300, 1536, 384, 1), (1000, 384, 384, 2), (257, 384, 1536, 2),
    (40000, 1152, 384, 0), (129, 256, 128, 1),
    # large enough for the 2-CTA multicast variant (>= 74 tile pairs), incl. an odd number of row tiles
    (20000, 384, 384, 2), (19100, 384, 1536, 2), (20000, 1536, 384, 1), (19000, 128, 64, 0),
    # N = 768 LayerNorm epilogue (BERT-base): the row is split over a 2-CTA cluster
    (1000, 768, 768, 2), (130, 768, 3072, 2), (20000, 768, 768, 2),
])
def test_tcgen05_gemm_against_torch(M, N, K, epi, fmt, tol):
    import torch
    torch.manual_seed(M + N + K + epi)
    dt = torch.bfloat16 if fmt == 1 else torch.float16
    A = torch.randn(M, K, device="cuda").to(dt)
    W = (torch.randn(N, K, device="cuda") / K ** 0.5).to(dt)
    bias = torch.randn(N, device="cuda") * 0.1
    res = torch.randn(M, N, device="cuda").to(dt)
    gamma = 1 + 0.1 * torch.randn(N, device="cuda")
    beta = 0.1 * torch.randn(N, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda").to(dt)
    rc = capi.lib().mx_debug_gemm(A.data_ptr(), W.data_ptr(), bias.data_ptr(), res.data_ptr(), gamma.data_ptr(),
                                  beta.data_ptr(), out.data_ptr(), M, N, K, fmt, epi, C.c_float(1e-12), 0)
    assert rc == 0, capi.lib().mx_debug_last_error()
    ref = A.float() @ W.float().t() + bias
    if epi == 1:
        ref = torch.nn.functional.gelu(ref)
    if epi == 2:
        ref = torch.nn.functional.layer_norm(ref + res.float(), (N,), gamma, beta, 1e-12)
    err = (out.float() - ref).abs().max().item()
    assert np.isfinite(err) and err <= tol * max(1.0, ref.abs().max().item()), err


capi.lib().mx_debug_gemm.restype = C.c_int32
capi.lib().mx_debug_gemm.argtypes = [C.c_void_p] * 7 + [C.c_uint32] * 5 + [C.c_float, C.c_int32]
capi.lib().mx_debug_last_error.restype = C.c_char_p


capi.lib().mx_debug_attention.restype = C.c_int32
capi.lib().mx_debug_attention.argtypes = [C.c_void_p] * 3 + [C.c_uint32] * 6 + [C.c_int32]


@pytest.mark.parametrize("impl", [0, 1, 2])
@pytest.mark.parametrize("fmt,tol", [(1, 2e-2), (0, 3e-3)])
@pytest.mark.parametrize("B,S,H,heads", [(3, 64, 64, 2), (4, 256, 384, 12), (2, 200, 768, 12), (5, 37, 384, 12),
                                         (2, 130, 256, 2), (1, 512, 384, 12), (40, 256, 384, 12), (7, 128, 384, 12),
                                         (33, 129, 768, 12)])
def test_attention_kernels_against_torch(B, S, H, heads, fmt, tol, impl):
    """K6 in isolation: softmax(q k^T / sqrt(dh) + padding mask) v, rows beyond lens[b] are zero.
    impl 0 = CUDA cores, 1 = mma.sync flash kernel, 2 = tcgen05 kernel (head_dim 32 / 64, S <= 256)"""
    import torch
    if impl == 2 and (H // heads not in (32, 64) or S > 256):
        pytest.skip("shape outside the tcgen05 attention kernel (the encoder falls back to the mma.sync kernel)")
    torch.manual_seed(B * 1000 + S + H)
    dt = torch.bfloat16 if fmt == 1 else torch.float16
    dh = H // heads
    qkv = (torch.randn(B, S, 3 * H, device="cuda") * 1.5).to(dt)
    lens = torch.randint(1, S + 1, (B,), device="cuda", dtype=torch.int32)
    lens[0] = S
    if B > 1:
        lens[1] = 1
    if B > 2:
        lens[2] = 0
    out = torch.full((B, S, H), float("nan"), device="cuda").to(dt)
    rc = capi.lib().mx_debug_attention(qkv.data_ptr(), lens.data_ptr(), out.data_ptr(), B, S, H, heads, fmt, impl, 0)
    assert rc == 0, capi.lib().mx_debug_last_error()
    q, k, v = [x.float().reshape(B, S, heads, dh).transpose(1, 2) for x in qkv.split(H, dim=2)]
    mask = torch.arange(S, device="cuda")[None, :] < lens[:, None]
    sc = q @ k.transpose(-1, -2) / dh ** 0.5
    sc = sc.masked_fill(~mask[:, None, None, :], float("-inf"))
    ref = (torch.softmax(sc, dim=-1).nan_to_num(0.0) @ v).transpose(1, 2).reshape(B, S, H)
    ref = ref * mask[:, :, None]
    err = (out.float() - ref).abs().max().item()
    assert np.isfinite(err) and err <= tol * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("precision,min_cos,max_abs", [("bf16", 1 - 1e-4, 2e-3), ("f16", 1 - 1e-5, 5e-4)])
def test_tensor_core_paths_vs_oracle(precision, min_cos, max_abs):
    cfg = enc_oracle.MINILM_L6
    w = enc_oracle.make_weights(cfg, seed=21)
    ids, lens = enc_oracle.make_inputs(cfg, 8, 128, seed=22, ragged=True, min_len=5)
    ref = enc_oracle.hf_encode(cfg, w, ids, lens)
    e = B200Encoder(arch_of(cfg), w, precision=precision, max_tokens=8 * 128)
    out = e.encode_ids(ids, lens)
    cos = (out * ref).sum(1)
    print(f"{precision}: min cosine to oracle {cos.min():.7f}, max abs diff {np.abs(out - ref).max():.2e}")
    assert (cos >= min_cos).all()
    assert np.abs(out - ref).max() <= max_abs
    # against the library's own f32 path too (same kernels' math in f32)
    e32 = B200Encoder(arch_of(cfg), w, precision="f32", max_tokens=8 * 128)
    out32 = e32.encode_ids(ids, lens)
    assert ((out * out32).sum(1) >= min_cos).all()


@pytest.mark.parametrize("precision", ["bf16", "f16"])
def test_packed_layout_matches_padded_layout_and_oracle(precision, monkeypatch):
    """Ragged batches run without their padding (token i of sequence b in row cu[b] + i, encoder.cuh): same rows, same
    arithmetic per row, so the result must equal the padded layout's (MX_ENCODER_NO_PACKING=1) and the oracle's."""
    cfg = enc_oracle.MINILM_L6
    w = enc_oracle.make_weights(cfg, seed=51)
    ids, lens = enc_oracle.make_inputs(cfg, 9, 160, seed=52, ragged=True, min_len=2)
    lens[0], lens[3], lens[8] = 160, 1, 129          # a full row, a single token, one key past a 128-row tile
    for b in range(9):
        ids[b, lens[b]:] = 0
    ref = enc_oracle.hf_encode(cfg, w, ids, lens)
    e = B200Encoder(arch_of(cfg), w, precision=precision, max_tokens=5 * 160)     # two chunks of 5 and 4 sequences
    out = e.encode_ids(ids, lens)
    monkeypatch.setenv("MX_ENCODER_NO_PACKING", "1")
    e_pad = B200Encoder(arch_of(cfg), w, precision=precision, max_tokens=5 * 160)
    monkeypatch.delenv("MX_ENCODER_NO_PACKING")
    out_pad = e_pad.encode_ids(ids, lens)
    print(f"packed vs padded ({precision}): max abs diff {np.abs(out - out_pad).max():.2e}")
    np.testing.assert_allclose(out, out_pad, atol=1e-6)
    assert ((out * ref).sum(1) >= (1 - 1e-4 if precision == "bf16" else 1 - 1e-5)).all()
    # garbage in the padded tail of the ids must not matter, and an empty sequence gives a zero row in both layouts
    ids2 = ids.copy()
    ids2[1, lens[1]:] = 777
    np.testing.assert_array_equal(e.encode_ids(ids2, lens), out)
    lens0 = lens.copy()
    lens0[4] = 0
    out0, out0_pad = e.encode_ids(ids, lens0), e_pad.encode_ids(ids, lens0)
    assert np.abs(out0[4]).max() == 0.0 and np.abs(out0_pad[4]).max() == 0.0
    np.testing.assert_allclose(np.delete(out0, 4, 0), np.delete(out, 4, 0), atol=1e-6)


def test_encode_errors():
    cfg = enc_oracle.TINY
    w = enc_oracle.make_weights(cfg, seed=1)
    e = B200Encoder(arch_of(cfg), w, precision="f32", max_tokens=256)
    with pytest.raises(EncodingFailure):
        e.encode_ids(np.zeros((1, cfg.max_pos + 1), np.int32), np.ones(1, np.int32))
    del w["embeddings.LayerNorm.bias"]
    from memex_b200.embedding import SetupError
    with pytest.raises(SetupError):
        B200Encoder(arch_of(cfg), w, precision="f32")


def test_bert_base_shape_tensor_core_path():
    """hidden 768 / 12 heads of 64 / ffn 3072 (BertBaseNliMeanTokens, e5-base): split-row LayerNorm GEMMs, dh = 64
    attention.  Two layers keep the CPU oracle quick; every kernel shape is the 12-layer model's."""
    cfg = enc_oracle.EncoderConfig(layers=2, hidden=768, heads=12, ffn=3072)
    w = enc_oracle.make_weights(cfg, seed=31)
    ids, lens = enc_oracle.make_inputs(cfg, 6, 96, seed=32, ragged=True, min_len=3)
    ref = enc_oracle.hf_encode(cfg, w, ids, lens)
    for precision, min_cos in (("bf16", 1 - 1e-4), ("f16", 1 - 1e-5), ("f32", 1 - 1e-6)):
        e = B200Encoder(arch_of(cfg), w, precision=precision, max_tokens=6 * 96)
        out = e.encode_ids(ids, lens)
        cos = (out * ref).sum(1)
        print(f"bert-base shape {precision}: min cosine {cos.min():.7f}")
        assert (cos >= min_cos).all(), (precision, cos.min())
        e.close()


@pytest.mark.parametrize("name,cfg,B,S,sample", [
    ("config 3: MiniLM-L6, B=256, S=256", enc_oracle.MINILM_L6, 256, 256, 16),
    ("memex default: MiniLM-L12, B=256, S=128", enc_oracle.MINILM_L12, 256, 128, 16),
    ("config 5 ingest: BERT-base 12 layers, B=16, S=512", enc_oracle.BERT_BASE, 16, 512, 6),
])
def test_bf16_parity_at_the_benchmark_shapes(name, cfg, B, S, sample):
    """The product path (bf16 activations, tcgen05 GEMMs + attention) at the shapes the numbers are quoted on
    (BASELINE.json configs 3 and 5, and memex's default model embedding.rs:64-72) -- full depth, full batch.  The GPU
    encodes the WHOLE batch; the oracle (HF BertModel fp32 on torch-CPU) re-computes a spread sample of its rows
    (rows are independent, so a sample row's answer does not depend on the rest of the batch).
    Gate: BASELINE.md section 3 / north_star -- cosine >= 1 - 1e-4 per row for the Normalize models, met by the hosts'
    default precision at every depth."""
    w = enc_oracle.make_weights(cfg, seed=61)
    ids, lens = enc_oracle.make_inputs(cfg, B, S, seed=62)
    # a few ragged rows inside the full-length batch (padding mask + packed layout at this size)
    lens[1], lens[B // 2] = max(3, S // 3), S - 1
    ids[1, lens[1]:] = cfg.pad_id
    ids[B // 2, lens[B // 2]:] = cfg.pad_id
    rows = sorted(set([0, 1, B // 2, B - 1] + list(np.linspace(2, B - 2, sample - 4).astype(int))))
    ref = enc_oracle.hf_encode(cfg, w, ids[rows], lens[rows])
    # bf16 (the precision BASELINE.json's config 3 names) meets the 1e-4 gate at 6 layers; at 12 layers its 8-bit
    # mantissa compounds to 1 - 1.05e-4 (measured), which is why the hosts' default ("auto") runs the deeper stacks with
    # f16 activations and weights on the same tcgen05 kernels -- that path must meet the gate at every depth
    gates = {"bf16": 1 - 1e-4 if cfg.layers <= 6 else 1 - 2e-4, "f16": 1 - 1e-5 if cfg.layers <= 6 else 1 - 2e-5}
    for precision in ("bf16", "f16"):
        e = B200Encoder(arch_of(cfg), w, precision=precision, max_tokens=B * S)
        out = e.encode_ids(ids, lens)
        e.close()
        assert out.shape == (B, cfg.hidden) and np.isfinite(out).all()
        np.testing.assert_allclose(np.linalg.norm(out, axis=1), 1.0, atol=1e-5)
        cos = (out[rows] * ref).sum(1)
        print(f"{name}: {precision} min cosine to the oracle over {len(rows)} sampled rows {cos.min():.7f}, "
              f"max abs diff {np.abs(out[rows] - ref).max():.2e}")
        assert (cos >= gates[precision]).all(), (precision, cos.min())
    auto = B200Encoder(arch_of(cfg), w, max_tokens=B * S)
    assert auto.precision == ("bf16" if cfg.layers <= 6 else "f16")
    out = auto.encode_ids(ids[:4], lens[:4])
    auto.close()
    assert (out[0] * ref[rows.index(0)]).sum() >= 1 - 1e-4   # the default meets the gate


@pytest.mark.parametrize("name", ["roberta", "distiluse", "albert"])
def test_other_stacks_of_the_enum_tensor_core_path(name):
    """AllDistilrobertaV1 (RoBERTa: positions from padding_idx + 1, type_vocab 1, eps 1e-5), DistiluseBaseMultilingualCased
    (DistilBERT, no token types, Dense 768 -> 512 + Tanh, no Normalize) and ParaphraseAlbertSmallV2 (ALBERT: 128-wide
    embeddings projected to 768, one shared layer, gelu_new) -- embedding.rs:24-55 -- at the models' true widths (two
    layers keep the CPU oracle quick; smaller vocabularies keep the upload quick), against HF Roberta / DistilBert /
    Albert models on torch CPU."""
    import dataclasses
    base = {"roberta": enc_oracle.DISTILROBERTA, "distiluse": enc_oracle.DISTILUSE, "albert": enc_oracle.ALBERT_SMALL}[name]
    cfg = dataclasses.replace(base, layers=2, vocab=4000)
    w = enc_oracle.make_weights(cfg, seed=41)
    ids, lens = enc_oracle.make_inputs(cfg, 6, 96, seed=42, ragged=True, min_len=3)
    ref = enc_oracle.hf_encode(cfg, w, ids, lens)
    assert ref.shape == (6, cfg.dense_out or cfg.hidden)
    for precision, min_cos in (("bf16", 1 - 2e-4), ("f16", 1 - 1e-5), ("f32", 1 - 1e-6)):
        e = B200Encoder(arch_of(cfg), w, precision=precision, max_tokens=6 * 96)
        out = e.encode_ids(ids, lens)
        cos = row_cosine(out, ref)
        print(f"{name} {precision}: min cosine {cos.min():.7f}, max abs diff {np.abs(out - ref).max():.2e}")
        assert out.shape == ref.shape
        assert (cos >= min_cos).all(), (precision, cos.min())
        if not cfg.normalize:   # the scale matters when the model has no Normalize module
            assert np.abs(out - ref).max() <= (5e-2 if precision == "bf16" else 5e-3) * max(1.0, np.abs(ref).max())
        e.close()


def test_sentence_t5_stack():
    """SentenceT5Base (embedding.rs:32,52): T5 v1.1 encoder (pre-RMSNorm, unscaled attention + bucketed relative position
    bias, gated-GELU feed-forward, no biases) + mean pool + Dense(no bias) + Normalize, against HF T5EncoderModel on
    torch-CPU.  The tiny shape runs the f32 path (every kernel of the stack on CUDA cores); the model's true width runs the
    tensor-core path (the seven GEMMs of a layer on tcgen05) with sequences long enough to reach the logarithmic buckets."""
    import dataclasses
    cfg = enc_oracle.TINY_T5
    w = enc_oracle.make_weights(cfg, seed=71)
    ids, lens = enc_oracle.make_inputs(cfg, 5, 60, seed=72, ragged=True, min_len=2)
    ref = enc_oracle.hf_encode(cfg, w, ids, lens)
    e = B200Encoder(arch_of(cfg), w, precision="f32", max_tokens=5 * 60)
    out = e.encode_ids(ids, lens)
    e.close()
    print(f"tiny T5 f32: max abs diff {np.abs(out - ref).max():.2e}")
    assert np.abs(out - ref).max() <= 1e-4 and (row_cosine(out, ref) >= 1 - 1e-6).all()
    cfg = dataclasses.replace(enc_oracle.SENTENCE_T5_BASE, layers=2, vocab=4000)
    w = enc_oracle.make_weights(cfg, seed=73)
    ids, lens = enc_oracle.make_inputs(cfg, 4, 200, seed=74, ragged=True, min_len=150)
    lens[0] = 200
    ref = enc_oracle.hf_encode(cfg, w, ids, lens)
    # bf16 sits at 1 - 3.5e-4 here (measured): the stack adds sub-layer outputs to an un-normalised stream, and with 8
    # mantissa bits every one of those additions carries the GEMM output's rounding; f16's 11 bits meet the 1e-4 gate and
    # are the hosts' default for this stack ("auto"), with the bf16 twin as the fallback should a checkpoint overflow f16
    for precision, min_cos in (("f32", 1 - 1e-6), ("f16", 1 - 1e-4), ("bf16", 1 - 1e-3)):
        e = B200Encoder(arch_of(cfg), w, precision=precision, max_tokens=4 * 200)
        out = e.encode_ids(ids, lens)
        e.close()
        cos = row_cosine(out, ref)
        print(f"sentence-t5 shape {precision}: min cosine {cos.min():.7f}, max abs diff {np.abs(out - ref).max():.2e}")
        assert out.shape == (4, 768) and (cos >= min_cos).all(), (precision, cos.min())
    auto = B200Encoder(arch_of(cfg), w, max_tokens=4 * 200)
    assert auto.precision == "f16"
    auto.close()


def test_roberta_position_offset_limits_the_sequence_length():
    cfg = enc_oracle.TINY_ROBERTA   # max_pos 66, positions start at 2 -> at most 64 tokens
    w = enc_oracle.make_weights(cfg, seed=1)
    e = B200Encoder(arch_of(cfg), w, precision="f32", max_tokens=256)
    ids, lens = enc_oracle.make_inputs(cfg, 2, 64, seed=2)
    ref = enc_oracle.np_encode(cfg, w, ids, lens)
    assert np.abs(e.encode_ids(ids, lens) - ref).max() <= 1e-4
    with pytest.raises(EncodingFailure):
        e.encode_ids(np.full((1, 65), 5, np.int32), np.full(1, 65, np.int32))


@pytest.mark.parametrize("switch", ["MX_GEMM_MULTICAST", "MX_GEMM_RESIDENT", "MX_GEMM_LN_NO_SPLIT", "MX_GEMM_EPI16", "MX_GEMM_LN_AMC", "MX_GEMM_QKV12",
                                    "MX_GEMM_PAIR"])
def test_gemm_opt_in_variants_in_a_fresh_process(switch):
    """the 2-CTA weight-multicast and resident-weight GEMM variants are selected by an environment switch that
    the library reads once, so they are exercised in a child process: same GEMM parity cases + the end-to-end
    encoder check"""
    import subprocess
    import sys
    switch, _, value = switch.partition("=")   # "NAME" turns an opt-in variant on, "NAME=0" a default one off
    if os.environ.get(switch):
        pytest.skip("already inside the child process")
    env = dict(os.environ, **{switch: value or "1"})
    r = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-x", "-m", "gpu", "-k",
                        "tcgen05_gemm_against_torch or tensor_core_paths_vs_oracle"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
