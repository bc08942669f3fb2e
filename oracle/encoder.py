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

and for the other stacks of the enum that share BERT's layer (embedding.rs:24-55): RoBERTa
(AllDistilrobertaV1: positions from padding_idx + 1), DistilBERT + Dense/Tanh
(DistiluseBaseMultilingualCased: no token types, no Normalize) and ALBERT (ParaphraseAlbertSmallV2:
factorised embeddings, one shared layer, gelu_new).  SentenceT5Base is not restated.

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
    # the other stacks of the enum that share BERT's post-LayerNorm layer (embedding.rs:24-55)
    family: str = "bert"          # "bert" | "roberta" | "distilbert" | "albert"
    pos_offset: int = 0           # RoBERTa: create_position_ids_from_input_ids -> padding_idx + 1 + i
    pad_id: int = 0               # RoBERTa: 1
    dense_out: int = 0            # sentence-transformers Dense module after pooling
    dense_act: str = "identity"   # "identity" | "tanh"
    ffn_act: str = "gelu"         # "gelu" | "gelu_new"
    embed_dim: int = 0            # ALBERT factorised embeddings (0 = hidden)
    share_layers: bool = False    # ALBERT
    # T5 encoder (SentenceT5Base, embedding.rs:32,52): pre-RMSNorm, unscaled attention with a bucketed relative position
    # bias shared by all layers, gated-GELU feed-forward, no biases anywhere; weights under HF T5EncoderModel names
    d_kv: int = 0                 # per-head width (T5: heads * d_kv need not equal hidden)
    rel_buckets: int = 32
    rel_max_distance: int = 128
    dense_bias: bool = True       # sentence-t5's 2_Dense has none

    def to_dict(self):
        return asdict(self)


# embedding.rs:24-55 model enum -> architecture (sentence-transformers model cards)
MINILM_L6 = EncoderConfig(layers=6)
MINILM_L12 = EncoderConfig(layers=12)              # memex default, embedding.rs:64-73
BERT_BASE = EncoderConfig(layers=12, hidden=768, heads=12, ffn=3072)  # BertBaseNliMeanTokens / e5-base
TINY = EncoderConfig(layers=2, hidden=64, heads=2, ffn=128, vocab=200, max_pos=64)
# AllDistilrobertaV1 (one of the three models segment_text accepts, embedding.rs:156-161)
DISTILROBERTA = EncoderConfig(layers=6, hidden=768, heads=12, ffn=3072, vocab=50265, max_pos=514, type_vocab=1,
                              ln_eps=1e-5, family="roberta", pos_offset=2, pad_id=1)
# DistiluseBaseMultilingualCased: DistilBERT + Dense(768 -> 512, Tanh), no Normalize
DISTILUSE = EncoderConfig(layers=6, hidden=768, heads=12, ffn=3072, vocab=119547, max_pos=512, type_vocab=0,
                          normalize=False, family="distilbert", dense_out=512, dense_act="tanh")
# ParaphraseAlbertSmallV2: ALBERT (128-wide embeddings, one shared layer applied 6 times, gelu_new), no Normalize
ALBERT_SMALL = EncoderConfig(layers=6, hidden=768, heads=12, ffn=3072, vocab=30000, max_pos=512, normalize=False,
                             family="albert", ffn_act="gelu_new", embed_dim=128, share_layers=True)
# SentenceT5Base: sentence-transformers/sentence-t5-base = T5 v1.1 base encoder + mean pool + Dense(768 -> 768, no bias) + Normalize
SENTENCE_T5_BASE = EncoderConfig(layers=12, hidden=768, heads=12, ffn=2048, vocab=32128, max_pos=512, type_vocab=0, ln_eps=1e-6,
                                 family="t5", d_kv=64, dense_out=768, dense_bias=False, ffn_act="gated-gelu")
TINY_T5 = EncoderConfig(layers=2, hidden=64, heads=2, ffn=128, vocab=200, max_pos=64, type_vocab=0, ln_eps=1e-6,
                        family="t5", d_kv=32, dense_out=64, dense_bias=False, ffn_act="gated-gelu")
# the same three stacks at a size the numpy restatement runs in milliseconds
TINY_ROBERTA = EncoderConfig(layers=2, hidden=64, heads=2, ffn=128, vocab=200, max_pos=66, type_vocab=1, ln_eps=1e-5,
                             family="roberta", pos_offset=2, pad_id=1)
TINY_DISTILUSE = EncoderConfig(layers=2, hidden=64, heads=2, ffn=128, vocab=200, max_pos=64, type_vocab=0,
                               normalize=False, family="distilbert", dense_out=48, dense_act="tanh")
TINY_ALBERT = EncoderConfig(layers=3, hidden=128, heads=4, ffn=256, vocab=200, max_pos=64, normalize=False,
                            family="albert", ffn_act="gelu_new", embed_dim=64, share_layers=True)


def weight_names(cfg: EncoderConfig):
    """Canonical (HF ``BertModel(add_pooling_layer=False)``) state_dict names -> shapes; the other families use the
    same names (``hf_state_dict`` renames them for the HF model of the family)."""
    H, F = cfg.hidden, cfg.ffn
    if cfg.family == "t5":
        inner = cfg.heads * cfg.d_kv
        out = {"shared.weight": (cfg.vocab, H),
               "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight": (cfg.rel_buckets, cfg.heads)}
        for i in range(cfg.layers):
            p = f"encoder.block.{i}.layer."
            out.update({p + "0.SelfAttention.q.weight": (inner, H), p + "0.SelfAttention.k.weight": (inner, H),
                        p + "0.SelfAttention.v.weight": (inner, H), p + "0.SelfAttention.o.weight": (H, inner),
                        p + "0.layer_norm.weight": (H,),
                        p + "1.DenseReluDense.wi_0.weight": (F, H), p + "1.DenseReluDense.wi_1.weight": (F, H),
                        p + "1.DenseReluDense.wo.weight": (H, F), p + "1.layer_norm.weight": (H,)})
        out["encoder.final_layer_norm.weight"] = (H,)
        if cfg.dense_out:
            out["dense.linear.weight"] = (cfg.dense_out, H)
        return out
    E = cfg.embed_dim or H
    out = {
        "embeddings.word_embeddings.weight": (cfg.vocab, E),
        "embeddings.position_embeddings.weight": (cfg.max_pos, E),
    }
    if cfg.family != "distilbert":
        out["embeddings.token_type_embeddings.weight"] = (cfg.type_vocab, E)
    out["embeddings.LayerNorm.weight"] = (E,)
    out["embeddings.LayerNorm.bias"] = (E,)
    if E != H:
        out["embeddings.projection.weight"] = (H, E)
        out["embeddings.projection.bias"] = (H,)
    for i in range(1 if cfg.share_layers else cfg.layers):
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
    if cfg.dense_out:
        out["dense.linear.weight"] = (cfg.dense_out, H)
        if cfg.dense_bias:
            out["dense.linear.bias"] = (cfg.dense_out,)
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
        if name.endswith("LayerNorm.weight") or name.endswith("layer_norm.weight"):
            a = 1.0 + 0.1 * rng.standard_normal(shape)
        elif name.endswith("relative_attention_bias.weight"):
            a = 1.0 * rng.standard_normal(shape)
        elif name == "shared.weight":
            a = 0.5 * rng.standard_normal(shape)
        elif cfg.family == "t5" and name.endswith("SelfAttention.q.weight"):
            # T5 does not scale q . k by 1 / sqrt(d_kv): the scale lives in the weights (keeps the softmax from saturating)
            a = rng.standard_normal(shape) * (1.5 / np.sqrt(shape[1]) / np.sqrt(cfg.d_kv) ** 0.5)
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
    """ids [B,S] int32 (cfg.pad_id beyond lens), lens [B] int32 -- SURVEY.md section 8d config 3."""
    rng = np.random.default_rng(seed)
    lo = min(1000, cfg.vocab // 4)
    hi = min(30000, cfg.vocab)
    ids = rng.integers(lo, hi, size=(batch, seq), dtype=np.int64).astype(np.int32)
    if ragged:
        lens = rng.integers(min(min_len, seq), seq + 1, size=batch).astype(np.int32)
    else:
        lens = np.full(batch, seq, dtype=np.int32)
    for b in range(batch):
        ids[b, lens[b]:] = cfg.pad_id
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


def t5_relative_buckets(S: int, num_buckets: int = 32, max_distance: int = 128) -> np.ndarray:
    """HF T5Attention._relative_position_bucket, bidirectional: [S (query), S (key)] bucket of (key - query)"""
    rel = np.arange(S)[None, :] - np.arange(S)[:, None]
    nb = num_buckets // 2
    ret = (rel > 0).astype(np.int64) * nb
    n = np.abs(rel)
    max_exact = nb // 2
    is_small = n < max_exact
    with np.errstate(divide="ignore", invalid="ignore"):
        # float32, as torch computes it (the bucket boundaries are decided by this rounding)
        large = max_exact + (np.log(n.astype(np.float32) / max_exact) / np.log(max_distance / max_exact) * (nb - max_exact)).astype(np.int64)
    large = np.minimum(large, nb - 1)
    return ret + np.where(is_small, n, large)


def _np_encode_t5(cfg: EncoderConfig, W: dict, ids, lens, dtype, return_hidden):
    B, S = ids.shape
    H, nh, dk = cfg.hidden, cfg.heads, cfg.d_kv
    mask = (np.arange(S)[None, :] < lens[:, None])

    def rms(x, g):
        return x / np.sqrt((x ** 2).mean(-1, keepdims=True) + cfg.ln_eps) * g

    x = W["shared.weight"][ids]
    bias = W["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"][
        t5_relative_buckets(S, cfg.rel_buckets, cfg.rel_max_distance)]            # [S, S, heads]
    bias = bias.transpose(2, 0, 1)[None] + np.where(mask, 0.0, -1e30)[:, None, None, :]
    for i in range(cfg.layers):
        p = f"encoder.block.{i}.layer."
        n = rms(x, W[p + "0.layer_norm.weight"])
        q = (n @ W[p + "0.SelfAttention.q.weight"].T).reshape(B, S, nh, dk).transpose(0, 2, 1, 3)
        k = (n @ W[p + "0.SelfAttention.k.weight"].T).reshape(B, S, nh, dk).transpose(0, 2, 1, 3)
        v = (n @ W[p + "0.SelfAttention.v.weight"].T).reshape(B, S, nh, dk).transpose(0, 2, 1, 3)
        s = q @ k.transpose(0, 1, 3, 2) + bias                                      # no 1 / sqrt(d_kv)
        s = s - s.max(-1, keepdims=True)
        pr = np.exp(s)
        pr = pr / pr.sum(-1, keepdims=True)
        ctx = (pr @ v).transpose(0, 2, 1, 3).reshape(B, S, nh * dk)
        x = x + ctx @ W[p + "0.SelfAttention.o.weight"].T
        n = rms(x, W[p + "1.layer_norm.weight"])
        g = n @ W[p + "1.DenseReluDense.wi_0.weight"].T
        g = 0.5 * g * (1.0 + np.tanh(np.sqrt(2.0 / np.pi) * (g + 0.044715 * g ** 3)))
        x = x + (g * (n @ W[p + "1.DenseReluDense.wi_1.weight"].T)) @ W[p + "1.DenseReluDense.wo.weight"].T
    x = rms(x, W["encoder.final_layer_norm.weight"])
    m = mask[..., None].astype(dtype)
    pooled = (x * m).sum(1) / np.maximum(m.sum(1), 1e-9)
    if cfg.dense_out:
        pooled = pooled @ W["dense.linear.weight"].T
    if cfg.normalize:
        pooled = pooled / np.maximum(np.linalg.norm(pooled, axis=1, keepdims=True), 1e-12)
    if return_hidden:
        return pooled.astype(np.float32), x
    return pooled.astype(np.float32)


def np_encode(cfg: EncoderConfig, w: dict, ids: np.ndarray, lens: np.ndarray, dtype=np.float64,
              return_hidden: bool = False):
    """BERT forward + masked mean-pool + L2 normalise in numpy (default float64)."""
    W = {k: v.astype(dtype) for k, v in w.items()}
    if cfg.family == "t5":
        return _np_encode_t5(cfg, W, ids, lens, dtype, return_hidden)
    B, S = ids.shape
    H, nh = cfg.hidden, cfg.heads
    dh = H // nh
    mask = (np.arange(S)[None, :] < lens[:, None])
    x = (W["embeddings.word_embeddings.weight"][ids]
         + W["embeddings.position_embeddings.weight"][np.arange(S) + cfg.pos_offset][None])
    if cfg.family != "distilbert":
        x = x + W["embeddings.token_type_embeddings.weight"][0][None, None]
    x = _layer_norm(x, W["embeddings.LayerNorm.weight"], W["embeddings.LayerNorm.bias"], cfg.ln_eps)
    if cfg.embed_dim and cfg.embed_dim != H:
        x = x @ W["embeddings.projection.weight"].T + W["embeddings.projection.bias"]
    addmask = np.where(mask, 0.0, -1e30)[:, None, None, :]
    for i in range(cfg.layers):
        p = "encoder.layer.0." if cfg.share_layers else f"encoder.layer.{i}."
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
        if cfg.ffn_act == "gelu_new":
            h = 0.5 * h * (1.0 + np.tanh(np.sqrt(2.0 / np.pi) * (h + 0.044715 * h ** 3)))
        else:
            h = 0.5 * h * (1.0 + _erf(h / np.sqrt(2.0)))
        o = h @ W[p + "output.dense.weight"].T + W[p + "output.dense.bias"]
        x = _layer_norm(o + x, W[p + "output.LayerNorm.weight"], W[p + "output.LayerNorm.bias"],
                        cfg.ln_eps)
    m = mask[..., None].astype(dtype)
    pooled = (x * m).sum(1) / np.maximum(m.sum(1), 1e-9)
    if cfg.dense_out:   # modules.json order: Transformer, Pooling, Dense, (Normalize)
        pooled = pooled @ W["dense.linear.weight"].T + W["dense.linear.bias"]
        if cfg.dense_act == "tanh":
            pooled = np.tanh(pooled)
    if cfg.normalize:
        pooled = pooled / np.maximum(np.linalg.norm(pooled, axis=1, keepdims=True), 1e-12)
    if return_hidden:
        return pooled.astype(np.float32), x
    return pooled.astype(np.float32)


# ----------------------------------------------------------------------------------------------
# restatement 2: HuggingFace BertModel on torch CPU fp32 (what tch/libtorch executes)
# ----------------------------------------------------------------------------------------------

_hf_cache: dict = {}


def hf_state_dict(cfg: EncoderConfig, w: dict) -> dict:
    """canonical (BERT) names -> the state_dict names of the HF model of cfg.family (Dense module left out)"""
    out = {}
    for name, arr in w.items():
        if name.startswith("dense.linear."):
            continue
        if cfg.family == "distilbert":
            name = name.replace("encoder.layer.", "transformer.layer.")
            for a, b in ((".attention.self.query.", ".attention.q_lin."), (".attention.self.key.", ".attention.k_lin."),
                         (".attention.self.value.", ".attention.v_lin."), (".attention.output.dense.", ".attention.out_lin."),
                         (".attention.output.LayerNorm.", ".sa_layer_norm."), (".intermediate.dense.", ".ffn.lin1."),
                         (".output.dense.", ".ffn.lin2."), (".output.LayerNorm.", ".output_layer_norm.")):
                name = name.replace(a, b)
        elif cfg.family == "albert":
            name = name.replace("embeddings.projection.", "encoder.embedding_hidden_mapping_in.")
            if name.startswith("encoder.layer.0."):
                name = name.replace("encoder.layer.0.", "encoder.albert_layer_groups.0.albert_layers.0.")
                for a, b in ((".attention.self.query.", ".attention.query."), (".attention.self.key.", ".attention.key."),
                             (".attention.self.value.", ".attention.value."), (".attention.output.dense.", ".attention.dense."),
                             (".attention.output.LayerNorm.", ".attention.LayerNorm."), (".intermediate.dense.", ".ffn."),
                             (".output.dense.", ".ffn_output."), (".output.LayerNorm.", ".full_layer_layer_norm.")):
                    name = name.replace(a, b)
        out[name] = arr
    return out


def hf_model(cfg: EncoderConfig, w: dict):
    import torch

    key = (cfg, id(w))
    if key in _hf_cache:
        return _hf_cache[key]
    if cfg.family == "bert":
        from transformers import BertConfig, BertModel
        hc = BertConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers,
                        num_attention_heads=cfg.heads, intermediate_size=cfg.ffn,
                        max_position_embeddings=cfg.max_pos, type_vocab_size=cfg.type_vocab,
                        layer_norm_eps=cfg.ln_eps, hidden_act="gelu", hidden_dropout_prob=0.0,
                        attention_probs_dropout_prob=0.0)
        make = lambda: BertModel(hc, add_pooling_layer=False)
    elif cfg.family == "roberta":
        from transformers import RobertaConfig, RobertaModel
        hc = RobertaConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers,
                           num_attention_heads=cfg.heads, intermediate_size=cfg.ffn,
                           max_position_embeddings=cfg.max_pos, type_vocab_size=cfg.type_vocab,
                           layer_norm_eps=cfg.ln_eps, hidden_act="gelu", hidden_dropout_prob=0.0,
                           attention_probs_dropout_prob=0.0, pad_token_id=cfg.pad_id)
        make = lambda: RobertaModel(hc, add_pooling_layer=False)
    elif cfg.family == "distilbert":
        from transformers import DistilBertConfig, DistilBertModel
        hc = DistilBertConfig(vocab_size=cfg.vocab, dim=cfg.hidden, n_layers=cfg.layers, n_heads=cfg.heads,
                              hidden_dim=cfg.ffn, max_position_embeddings=cfg.max_pos, activation="gelu",
                              dropout=0.0, attention_dropout=0.0, sinusoidal_pos_embds=False, pad_token_id=cfg.pad_id)
        make = lambda: DistilBertModel(hc)
    elif cfg.family == "albert":
        from transformers import AlbertConfig, AlbertModel
        hc = AlbertConfig(vocab_size=cfg.vocab, embedding_size=cfg.embed_dim or cfg.hidden, hidden_size=cfg.hidden,
                          num_hidden_layers=cfg.layers, num_hidden_groups=1, inner_group_num=1,
                          num_attention_heads=cfg.heads, intermediate_size=cfg.ffn,
                          max_position_embeddings=cfg.max_pos, type_vocab_size=cfg.type_vocab,
                          layer_norm_eps=cfg.ln_eps, hidden_act=cfg.ffn_act, hidden_dropout_prob=0.0,
                          attention_probs_dropout_prob=0.0, pad_token_id=cfg.pad_id)
        make = lambda: AlbertModel(hc, add_pooling_layer=False)
    elif cfg.family == "t5":
        from transformers import T5Config, T5EncoderModel
        hc = T5Config(vocab_size=cfg.vocab, d_model=cfg.hidden, d_kv=cfg.d_kv, d_ff=cfg.ffn, num_layers=cfg.layers,
                      num_heads=cfg.heads, relative_attention_num_buckets=cfg.rel_buckets,
                      relative_attention_max_distance=cfg.rel_max_distance, dropout_rate=0.0,
                      layer_norm_epsilon=cfg.ln_eps, feed_forward_proj="gated-gelu", is_encoder_decoder=False,
                      use_cache=False, pad_token_id=cfg.pad_id)
        make = lambda: T5EncoderModel(hc)
    else:
        raise ValueError(cfg.family)
    try:
        hc._attn_implementation = "eager"
    except Exception:
        pass
    model = make()
    sd = {k: torch.from_numpy(v.copy()) for k, v in hf_state_dict(cfg, w).items()}
    if cfg.family == "t5":
        sd["encoder.embed_tokens.weight"] = sd["shared.weight"]
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
        if cfg.family in ("distilbert", "t5"):
            h = model(input_ids=t_ids, attention_mask=mask).last_hidden_state
        else:   # RoBERTa derives its position ids from the pad id inside the model
            h = model(input_ids=t_ids, attention_mask=mask,
                      token_type_ids=torch.zeros_like(t_ids)).last_hidden_state
        m = mask[..., None].to(h.dtype)
        pooled = (h * m).sum(1) / m.sum(1).clamp(min=1e-9)
        if cfg.dense_out:
            pooled = pooled @ torch.from_numpy(w["dense.linear.weight"]).T
            if cfg.dense_bias:
                pooled = pooled + torch.from_numpy(w["dense.linear.bias"])
            if cfg.dense_act == "tanh":
                pooled = torch.tanh(pooled)
        if cfg.normalize:
            pooled = torch.nn.functional.normalize(pooled, p=2, dim=1, eps=1e-12)
    return pooled.numpy().astype(np.float32)
