"""oracle/encoder.py -- TEST INFRASTRUCTURE ONLY.

CPU oracle for the sentence-embedding forward pass memex runs for every ingested segment and
every query: reference ``lib/libmemex/src/llm/embedding.rs:99-109``
(``SentenceEmbeddingsBuilder::remote(model).create_model()`` then ``model.encode(&segments)``).

The arithmetic lives in rust-bert 0.21.0 -> tch 0.13.0 -> libtorch CPU
(reference ``lib/libmemex/Cargo.toml:22``; ``Cargo.lock:3467,4388,4779``), none of which is under
/root/reference or buildable offline, so this file restates the PUBLISHED pipeline those crates
implement for the all-MiniLM / BERT sentence-transformers models:

    BERT encoder (post-LN, erf-GELU, learned absolute positions, token_type 0)
      -> masked mean-pool   sum(h * m) / max(sum(m), 1e-9)
      -> L2 normalise       x / max(||x||_2, 1e-12)

Two independent restatements are kept and checked against each other in tests/:
``np_encode`` (plain numpy, float64 accumulation available) and ``hf_encode`` (HuggingFace
``transformers.BertModel`` on torch CPU fp32 -- the same libtorch kernels ``tch`` dispatches to).
The reference pins NO numerical fixture for this path and no checkpoint is on
the box, so parity is "unpinned against reference output": weights are seeded random tensors at
the true architecture shapes, and the golden vectors under tests/golden/ are outputs of
``hf_encode`` produced by ``tests/golden/make_golden.py``.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict

import numpy as np


@dataclass(frozen=True)
class EncoderConfig:
    layers: int = 6
    hidden: int = 384
    heads: int = 12
    ffn: int = 1536
    vocab: int = 30522
    max_pos: int = 512
    type_vocab: int = 2
    ln_eps: float = 1e-12
    normalize: bool = True

    def to_dict(self):
        return asdict(self)


# embedding.rs:24-55 model enum -> architecture (sentence-transformers model cards)
MINILM_L6 = EncoderConfig(layers=6)
MINILM_L12 = EncoderConfig(layers=12)              # memex default, embedding.rs:64-73
BERT_BASE = EncoderConfig(layers=12, hidden=768, heads=12, ffn=3072)  # BertBaseNliMeanTokens / e5-base
TINY = EncoderConfig(layers=2, hidden=64, heads=2, ffn=128, vocab=200, max_pos=64)


def weight_names(cfg: EncoderConfig):
    """HF ``BertModel(add_pooling_layer=False)`` state_dict names -> shapes."""
    H, F = cfg.hidden, cfg.ffn
    out = {
        "embeddings.word_embeddings.weight": (cfg.vocab, H),
        "embeddings.position_embeddings.weight": (cfg.max_pos, H),
        "embeddings.token_type_embeddings.weight": (cfg.type_vocab, H),
        "embeddings.LayerNorm.weight": (H,),
        "embeddings.LayerNorm.bias": (H,),
    }
    for i in range(cfg.layers):
        p = f"encoder.layer.{i}."
        out.update({
            p + "attention.self.query.weight": (H, H), p + "attention.self.query.bias": (H,),
            p + "attention.self.key.weight": (H, H), p + "attention.self.key.bias": (H,),
            p + "attention.self.value.weight": (H, H), p + "attention.self.value.bias": (H,),
            p + "attention.output.dense.weight": (H, H), p + "attention.output.dense.bias": (H,),
            p + "attention.output.LayerNorm.weight": (H,), p + "attention.output.LayerNorm.bias": (H,),
            p + "intermediate.dense.weight": (F, H), p + "intermediate.dense.bias": (F,),
            p + "output.dense.weight": (H, F), p + "output.dense.bias": (H,),
            p + "output.LayerNorm.weight": (H,), p + "output.LayerNorm.bias": (H,),
        })
    return out


def make_weights(cfg: EncoderConfig, seed: int = 0) -> dict[str, np.ndarray]:
    """Seeded random weights at the true shapes (no checkpoint exists on the box).

    Scales are chosen so that activations stay O(1), attention is not flat, and every fused
    epilogue term is visible: biases and LayerNorm beta are non-zero, gamma is spread around 1.
    numpy's Generator is used (not torch) so the values are identical on every box.
    """
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in weight_names(cfg).items():
        if name.endswith("LayerNorm.weight"):
            a = 1.0 + 0.1 * rng.standard_normal(shape)
        elif name.endswith("LayerNorm.bias") or name.endswith(".bias"):
            a = 0.1 * rng.standard_normal(shape)
        elif "embeddings." in name:
            a = 0.5 * rng.standard_normal(shape)
        else:  # Linear weight [out, in]
            a = rng.standard_normal(shape) * (1.5 / np.sqrt(shape[1]))
        w[name] = a.astype(np.float32)
    return w


def make_inputs(cfg: EncoderConfig, batch: int, seq: int, seed: int = 7, ragged: bool = False,
                min_len: int = 16):
    """ids [B,S] int32 (0 = [PAD] beyond lens), lens [B] int32 -- SURVEY.md section 8d config 3."""
    rng = np.random.default_rng(seed)
    lo = min(1000, cfg.vocab // 4)
    hi = min(30000, cfg.vocab)
    ids = rng.integers(lo, hi, size=(batch, seq), dtype=np.int64).astype(np.int32)
    if ragged:
        lens = rng.integers(min(min_len, seq), seq + 1, size=batch).astype(np.int32)
    else:
        lens = np.full(batch, seq, dtype=np.int32)
    for b in range(batch):
        ids[b, lens[b]:] = 0
    return ids, lens


# ----------------------------------------------------------------------------------------------
# restatement 1: plain numpy
# ----------------------------------------------------------------------------------------------

def _erf(x):
    try:
        from scipy.special import erf
        return erf(x)
    except Exception:  # pragma: no cover
        import math
        return np.vectorize(math.erf)(x)


def _layer_norm(x, g, b, eps):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * g + b


def np_encode(cfg: EncoderConfig, w: dict, ids: np.ndarray, lens: np.ndarray, dtype=np.float64,
              return_hidden: bool = False):
    """BERT forward + masked mean-pool + L2 normalise in numpy (default float64)."""
    W = {k: v.astype(dtype) for k, v in w.items()}
    B, S = ids.shape
    H, nh = cfg.hidden, cfg.heads
    dh = H // nh
    mask = (np.arange(S)[None, :] < lens[:, None])
    x = (W["embeddings.word_embeddings.weight"][ids]
         + W["embeddings.position_embeddings.weight"][np.arange(S)][None]
         + W["embeddings.token_type_embeddings.weight"][0][None, None])
    x = _layer_norm(x, W["embeddings.LayerNorm.weight"], W["embeddings.LayerNorm.bias"], cfg.ln_eps)
    addmask = np.where(mask, 0.0, -1e30)[:, None, None, :]
    for i in range(cfg.layers):
        p = f"encoder.layer.{i}."
        q = x @ W[p + "attention.self.query.weight"].T + W[p + "attention.self.query.bias"]
        k = x @ W[p + "attention.self.key.weight"].T + W[p + "attention.self.key.bias"]
        v = x @ W[p + "attention.self.value.weight"].T + W[p + "attention.self.value.bias"]
        q = q.reshape(B, S, nh, dh).transpose(0, 2, 1, 3)
        k = k.reshape(B, S, nh, dh).transpose(0, 2, 1, 3)
        v = v.reshape(B, S, nh, dh).transpose(0, 2, 1, 3)
        s = q @ k.transpose(0, 1, 3, 2) / np.sqrt(dh) + addmask
        s = s - s.max(-1, keepdims=True)
        pr = np.exp(s)
        pr = pr / pr.sum(-1, keepdims=True)
        ctx = (pr @ v).transpose(0, 2, 1, 3).reshape(B, S, H)
        a = ctx @ W[p + "attention.output.dense.weight"].T + W[p + "attention.output.dense.bias"]
        x = _layer_norm(a + x, W[p + "attention.output.LayerNorm.weight"],
                        W[p + "attention.output.LayerNorm.bias"], cfg.ln_eps)
        h = x @ W[p + "intermediate.dense.weight"].T + W[p + "intermediate.dense.bias"]
        h = 0.5 * h * (1.0 + _erf(h / np.sqrt(2.0)))
        o = h @ W[p + "output.dense.weight"].T + W[p + "output.dense.bias"]
        x = _layer_norm(o + x, W[p + "output.LayerNorm.weight"], W[p + "output.LayerNorm.bias"],
                        cfg.ln_eps)
    m = mask[..., None].astype(dtype)
    pooled = (x * m).sum(1) / np.maximum(m.sum(1), 1e-9)
    if cfg.normalize:
        pooled = pooled / np.maximum(np.linalg.norm(pooled, axis=1, keepdims=True), 1e-12)
    if return_hidden:
        return pooled.astype(np.float32), x
    return pooled.astype(np.float32)


# ----------------------------------------------------------------------------------------------
# restatement 2: HuggingFace BertModel on torch CPU fp32 (what tch/libtorch executes)
# ----------------------------------------------------------------------------------------------

_hf_cache: dict = {}


def hf_model(cfg: EncoderConfig, w: dict):
    import torch
    from transformers import BertConfig, BertModel

    key = (cfg, id(w))
    if key in _hf_cache:
        return _hf_cache[key]
    hc = BertConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers,
                    num_attention_heads=cfg.heads, intermediate_size=cfg.ffn,
                    max_position_embeddings=cfg.max_pos, type_vocab_size=cfg.type_vocab,
                    layer_norm_eps=cfg.ln_eps, hidden_act="gelu", hidden_dropout_prob=0.0,
                    attention_probs_dropout_prob=0.0)
    try:
        hc._attn_implementation = "eager"
    except Exception:
        pass
    model = BertModel(hc, add_pooling_layer=False)
    sd = {k: torch.from_numpy(v.copy()) for k, v in w.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    real_missing = [m for m in missing if "position_ids" not in m and "token_type_ids" not in m]
    if real_missing or unexpected:
        raise RuntimeError(f"weight mismatch: missing={real_missing} unexpected={unexpected}")
    model.eval()
    _hf_cache[key] = model
    return model


def hf_encode(cfg: EncoderConfig, w: dict, ids: np.ndarray, lens: np.ndarray, threads: int | None = None):
    import torch

    if threads:
        torch.set_num_threads(threads)
    model = hf_model(cfg, w)
    B, S = ids.shape
    t_ids = torch.from_numpy(ids.astype(np.int64))
    mask = (torch.arange(S)[None, :] < torch.from_numpy(lens.astype(np.int64))[:, None]).to(torch.int64)
    with torch.no_grad():
        h = model(input_ids=t_ids, attention_mask=mask,
                  token_type_ids=torch.zeros_like(t_ids)).last_hidden_state
        m = mask[..., None].to(h.dtype)
        pooled = (h * m).sum(1) / m.sum(1).clamp(min=1e-9)
        if cfg.normalize:
            pooled = torch.nn.functional.normalize(pooled, p=2, dim=1, eps=1e-12)
    return pooled.numpy().astype(np.float32)
