"""Host-side mirror of memex's sentence embedder over the C ABI (Python face).

Mirrors reference lib/libmemex/src/llm/embedding.rs: EmbeddingError (:10-16), EmbeddingResult
(:18-22), EmbeddingsModelType (:24-55), ModelConfig (:57-73, default AllMiniLmL12V2 / 256 / 86),
SentenceEmbedder::{spawn, encode, encode_single} (:77-152: a dedicated thread owns the model,
requests arrive over a bounded channel of 100) and segment_text (:155-198).

The one line that changes is `model.encode(&segments)` (:109): instead of rust-bert -> libtorch
CPU, the segments are tokenised on the host and the padded ids go to mx_embedder_encode, which
runs the whole BERT forward + mean-pool + L2-normalise on the GPU.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C
import enum
import queue
import threading
from dataclasses import dataclass

import numpy as np

from . import capi


class EmbeddingError(Exception):
    pass


class EncodingFailure(EmbeddingError):       # embedding.rs:12-13
    pass


class SetupError(EmbeddingError):            # embedding.rs:14-15
    pass


@dataclass
class EmbeddingResult:                       # embedding.rs:18-22
    content: str
    vector: list


class EmbeddingsModelType(enum.Enum):        # embedding.rs:24-33
    DistiluseBaseMultilingualCased = "distiluse-base-multilingual-cased"
    BertBaseNliMeanTokens = "bert-base-nli-mean-tokens"
    AllMiniLmL12V2 = "all-MiniLM-L12-v2"
    AllMiniLmL6V2 = "all-MiniLM-L6-v2"
    AllDistilrobertaV1 = "all-distilroberta-v1"
    ParaphraseAlbertSmallV2 = "paraphrase-albert-small-v2"
    SentenceT5Base = "sentence-t5-base"


@dataclass(frozen=True)
class Architecture:
    layers: int
    hidden: int
    heads: int
    ffn: int
    vocab: int = 30522
    max_pos: int = 512
    type_vocab: int = 2
    ln_eps: float = 1e-12
    normalize: bool = True
    max_seq_length: int = 256    # sentence_bert_config.json of the model (what rust-bert truncates to)
    # the stacks that differ from BERT only around the layers (mx_model_ext in include/memex_b200.h)
    family: str = "bert"         # "bert" | "roberta" | "distilbert" | "albert": the checkpoint's naming scheme
    pos_offset: int = 0          # RoBERTa: positions start at padding_idx + 1 = 2
    pad_id: int = 0              # RoBERTa: 1
    dense_out: int = 0           # sentence-transformers Dense module after pooling (0 = none)
    dense_act: str = "identity"  # "identity" | "tanh"
    dense_bias: bool = True
    ffn_act: str = "gelu"        # "gelu" (erf) | "gelu_new" (tanh form, ALBERT)
    embed_dim: int = 0           # ALBERT: factorised embedding width (0 = hidden)
    share_layers: bool = False   # ALBERT: one set of layer weights
    # family "t5" (SentenceT5Base): T5 encoder stack, weights under the HF T5EncoderModel names
    d_kv: int = 0
    rel_buckets: int = 32
    rel_max_distance: int = 128

    @property
    def out_dim(self) -> int:
        return self.dense_out or self.hidden


# Every model of the enum (embedding.rs:24-55; shapes from the sentence-transformers model cards): six stacks around
# BERT's post-LayerNorm block and SentenceT5Base, a T5 v1.1 encoder (pre-RMSNorm, relative position bias, gated-GELU
# feed-forward) + mean pool + Dense(768 -> 768, no bias) + Normalize.  The reference can only SEGMENT L12 / L6 /
# distilroberta (embedding.rs:156-161); the others embed single shots.
ARCHITECTURES = {
    EmbeddingsModelType.SentenceT5Base: Architecture(
        12, 768, 12, 2048, vocab=32128, max_pos=512, type_vocab=0, ln_eps=1e-6, max_seq_length=256,
        family="t5", d_kv=64, dense_out=768, dense_bias=False, ffn_act="gated-gelu"),
    EmbeddingsModelType.AllMiniLmL6V2: Architecture(6, 384, 12, 1536, max_seq_length=256),
    EmbeddingsModelType.AllMiniLmL12V2: Architecture(12, 384, 12, 1536, max_seq_length=128),
    EmbeddingsModelType.BertBaseNliMeanTokens: Architecture(12, 768, 12, 3072, normalize=False, max_seq_length=128),
    EmbeddingsModelType.AllDistilrobertaV1: Architecture(
        6, 768, 12, 3072, vocab=50265, max_pos=514, type_vocab=1, ln_eps=1e-5, max_seq_length=512,
        family="roberta", pos_offset=2, pad_id=1),
    EmbeddingsModelType.DistiluseBaseMultilingualCased: Architecture(
        6, 768, 12, 3072, vocab=119547, max_pos=512, type_vocab=0, normalize=False, max_seq_length=128,
        family="distilbert", dense_out=512, dense_act="tanh"),
    EmbeddingsModelType.ParaphraseAlbertSmallV2: Architecture(
        6, 768, 12, 3072, vocab=30000, max_pos=512, normalize=False, max_seq_length=100,
        family="albert", ffn_act="gelu_new", embed_dim=128, share_layers=True),
}


def canonical_weights(family: str, state: dict) -> dict:
    """Checkpoint tensor names of a RoBERTa / DistilBERT / ALBERT stack -> the BERT names the C ABI takes
    (include/memex_b200.h).  `state` may carry a sentence-transformers Dense module as "linear.weight" /
    "linear.bias" or "dense.linear.*"; a leading "roberta." / "distilbert." / "albert." / "bert." is dropped."""
    out = {}
    for name, arr in state.items():
        for prefix in ("roberta.", "distilbert.", "albert.", "bert."):
            if name.startswith(prefix):
                name = name[len(prefix):]
        if name in ("linear.weight", "linear.bias"):
            name = "dense." + name
        if family == "distilbert":
            name = name.replace("transformer.layer.", "encoder.layer.")
            for a, b in ((".attention.q_lin.", ".attention.self.query."), (".attention.k_lin.", ".attention.self.key."),
                         (".attention.v_lin.", ".attention.self.value."), (".attention.out_lin.", ".attention.output.dense."),
                         (".sa_layer_norm.", ".attention.output.LayerNorm."), (".ffn.lin1.", ".intermediate.dense."),
                         (".ffn.lin2.", ".output.dense."), (".output_layer_norm.", ".output.LayerNorm.")):
                name = name.replace(a, b)
        elif family == "albert":
            name = name.replace("encoder.embedding_hidden_mapping_in.", "embeddings.projection.")
            name = name.replace("encoder.albert_layer_groups.0.albert_layers.0.", "encoder.layer.0.")
            for a, b in ((".attention.query.", ".attention.self.query."), (".attention.key.", ".attention.self.key."),
                         (".attention.value.", ".attention.self.value."), (".attention.dense.", ".attention.output.dense."),
                         (".attention.LayerNorm.", ".attention.output.LayerNorm."), (".ffn_output.", ".output.dense."),
                         (".ffn.", ".intermediate.dense."), (".full_layer_layer_norm.", ".output.LayerNorm.")):
                name = name.replace(a, b)
        if "position_ids" in name or "token_type_ids" in name or name.startswith("pooler."):
            continue
        out[name] = arr
    return out


@dataclass(frozen=True)
class ModelConfig:                           # embedding.rs:57-73
    model: EmbeddingsModelType = EmbeddingsModelType.AllMiniLmL12V2
    max_length: int = 256
    stride: int = 86                         # overlap roughly a third of the previous text


PRECISION = {"bf16": 0, "f32": 1, "f16": 2}


class B200Encoder:
    """Owns an mx_embedder handle: ids [B,S] + lens [B] -> f32 [B, out_dim] (unit-norm when the model has Normalize).
    `weights` use the BERT names (canonical_weights renames DistilBERT / ALBERT checkpoints)."""

    def __init__(self, arch: Architecture, weights: dict, precision: str = "auto", device: int = 0,
                 max_tokens: int = 0):
        """precision: "bf16" | "f16" (tensor-core paths, tcgen05 kind::f16 at the same speed) | "f32" (CUDA-core validation
        path) | "auto" = bf16 up to 6 layers, f16 beyond.  Measured against the fp32 oracle (tests/test_encoder_gpu.py):
        bf16 activations + weights give cosine 1 - 3e-5 at 6 layers but 1 - 1.05e-4 at 12 (the rounding of an 8-bit
        mantissa compounds per layer), f16's 11 bits give 1 - 2e-6 at 12 layers, so the deeper stacks of the enum
        (memex's default all-MiniLM-L12-v2, BERT-base) default to f16.  f16's range ends at 65504: should an exotic
        checkpoint overflow it, the batch comes back non-finite and `encode_ids` re-runs it on a bf16 twin of the model
        (still the GPU path; there is no CPU fallback)."""
        self.arch = arch
        self._auto = precision == "auto"
        if self._auto:
            # (T5: the residual stream is f32 in the kernels, but every sub-layer output is added to it un-normalised, and
            # bf16 measures 1 - 3.5e-4 at two layers; f16 meets the gate.  T5 v1.1 checkpoints are the ones known to
            # overflow f16 in the feed-forward: that is what the bf16 twin below is for.)
            precision = "bf16" if (arch.layers <= 6 and arch.family != "t5") else "f16"
        self._twin = None
        self._args = (arch, weights, device, max_tokens)
        names, tensors, keep = [], [], []
        for name, arr in weights.items():
            a = np.ascontiguousarray(arr, dtype=np.float32)
            keep.append(a)
            names.append(name.encode())
            tensors.append((names[-1], a.ctypes.data_as(C.POINTER(C.c_float)), a.size))
        arr_t = (capi.Tensor * len(tensors))(*[capi.Tensor(n, d, s) for n, d, s in tensors])
        cfg = capi.ModelCfg(layers=arch.layers, hidden=arch.hidden, heads=arch.heads, ffn=arch.ffn,
                            vocab=arch.vocab, max_pos=arch.max_pos, type_vocab=arch.type_vocab,
                            ln_eps=arch.ln_eps, normalize=1 if arch.normalize else 0,
                            precision=PRECISION[precision], max_tokens=max_tokens)
        ext = capi.ModelExt(pos_offset=arch.pos_offset, no_token_type=1 if arch.family == "distilbert" else 0,
                            dense_out=arch.dense_out, dense_act={"identity": 0, "tanh": 1}[arch.dense_act],
                            dense_bias=1 if arch.dense_bias else 0,
                            ffn_act={"gelu": 0, "gelu_new": 1, "gated-gelu": 1}[arch.ffn_act],
                            embed_dim=arch.embed_dim, share_layers=1 if arch.share_layers else 0,
                            family=capi.FAMILY_T5 if arch.family == "t5" else capi.FAMILY_BERT, d_kv=arch.d_kv,
                            rel_buckets=arch.rel_buckets, rel_max_distance=arch.rel_max_distance)
        h = C.c_void_p()
        rc = capi.lib().mx_embedder_create_ex(C.byref(cfg), C.byref(ext), arr_t, len(tensors), device, C.byref(h))
        if rc != capi.OK:
            msg = capi.lib().mx_last_error(None)
            raise SetupError(msg.decode(errors="replace") if msg else f"status {rc}")
        self._h = h
        self.precision = precision

    @property
    def handle(self):
        return self._h

    def encode_ids(self, ids: np.ndarray, lens: np.ndarray) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        B, S = ids.shape
        out = np.zeros((B, self.arch.out_dim), dtype=np.float32)
        rc = capi.lib().mx_embedder_encode(self._h, ids.ctypes.data, lens.ctypes.data, B, S, out.ctypes.data)
        if rc != capi.OK:
            msg = capi.lib().mx_last_error(self._h)
            raise EncodingFailure(msg.decode(errors="replace") if msg else f"status {rc}")
        if self._auto and self.precision == "f16" and not np.isfinite(out).all():
            if self._twin is None:   # the model overflowed f16 somewhere: the wider exponent of bf16 takes this batch
                arch, weights, device, max_tokens = self._args
                self._twin = B200Encoder(arch, weights, precision="bf16", device=device, max_tokens=max_tokens)
            return self._twin.encode_ids(ids, lens)
        return out

    def close(self):
        if getattr(self, "_twin", None) is not None:
            self._twin.close()
            self._twin = None
        if getattr(self, "_h", None) is not None:
            capi.lib().mx_embedder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def segment_text(model_config: ModelConfig, text: str, tokenizer) -> list[str]:
    """embedding.rs:155-198 with the tokenizer passed in (the reference re-loads it per call, :163).

    `tokenizer` is a `tokenizers.Tokenizer`; windows of `max_length` tokens overlapping by `stride`,
    each decoded back to text (the first one also gets `.replace(" ' ", "'")`, :183).
    """
    if model_config.model not in (EmbeddingsModelType.AllMiniLmL12V2, EmbeddingsModelType.AllMiniLmL6V2,
                                  EmbeddingsModelType.AllDistilrobertaV1):
        raise SetupError("Model not supported yet")
    tokenizer.enable_truncation(max_length=model_config.max_length, stride=model_config.stride)
    try:
        encoding = tokenizer.encode(text, add_special_tokens=False)
        segments = [tokenizer.decode(encoding.ids, skip_special_tokens=True).replace(" ' ", "'")]
        for over in encoding.overflowing:
            segments.append(tokenizer.decode(over.ids, skip_special_tokens=True))
    except Exception:
        raise EncodingFailure(text)
    finally:
        tokenizer.no_truncation()
    return segments


def tokenize_batch(tokenizer, segments: list[str], max_seq_length: int, pad_id: int = 0):
    """What rust-bert's SentenceEmbeddingsModel::tokenize does before the forward pass: add
    [CLS]/[SEP], truncate to the model's max_seq_length, pad to the longest of the batch."""
    tokenizer.enable_truncation(max_length=max_seq_length)
    try:
        encs = tokenizer.encode_batch(segments, add_special_tokens=True)
    finally:
        tokenizer.no_truncation()
    S = max(1, max(len(e.ids) for e in encs))
    ids = np.full((len(encs), S), pad_id, dtype=np.int32)
    lens = np.zeros(len(encs), dtype=np.int32)
    for i, e in enumerate(encs):
        ids[i, :len(e.ids)] = e.ids
        lens[i] = len(e.ids)
    return ids, lens


class SentenceEmbedder:
    """embedding.rs:77-152: `spawn` starts the runner thread that owns the model."""

    def __init__(self, sender: "queue.Queue"):
        self.sender = sender

    @classmethod
    def spawn(cls, model_config: ModelConfig, encoder: B200Encoder, tokenizer):
        """-> (thread handle, SentenceEmbedder).  `encoder` / `tokenizer` are the already-loaded
        model and tokenizer (the reference loads both inside the runner, :99-100,163 -- per task;
        SURVEY.md F9 -- here they are process-wide and handed in)."""
        q: queue.Queue = queue.Queue(maxsize=100)                 # mpsc::sync_channel(100), :87
        handle = threading.Thread(target=cls._runner, args=(q, model_config, encoder, tokenizer), daemon=True)
        handle.start()
        return handle, cls(q)

    MAX_BATCH_SEGMENTS = 256      # segments per forward pass (BASELINE.json config 3's batch)
    MAX_BATCH_WAIT_S = 0.0005     # how long the first request of a batch waits for company

    @staticmethod
    def _runner(receiver, model_config, encoder, tokenizer):
        """embedding.rs:94-135 processes ONE document per message: `model.encode(&segments)` sees a batch of ~50 segments
        for a 42 KB text and of 1 for a query, far from the 256-segment batches the kernels are at their best with
        (SURVEY.md 8(f) N4).  Here the runner drains the channel: whatever requests are queued when one arrives (plus what
        arrives within MAX_BATCH_WAIT_S, up to MAX_BATCH_SEGMENTS segments) are segmented and tokenised on the host and go
        through ONE forward pass; every caller gets exactly the rows of its own segments, in order.  Rows of a batch
        are independent (padding is masked, the packed layout drops it), so a request's vectors do not depend on what it
        was batched with."""
        import time as _time
        pending = None
        while True:
            msg = pending if pending is not None else receiver.get()
            pending = None
            if msg is None:
                return
            batch, n_seg, stop = [], 0, False
            deadline = _time.perf_counter() + SentenceEmbedder.MAX_BATCH_WAIT_S
            while True:
                text, segment, reply = msg
                try:
                    segments = segment_text(model_config, text, tokenizer) if segment else [text]
                    batch.append((segments, reply))
                    n_seg += len(segments)
                except Exception as e:  # the reference's runner dies; here the caller gets the error
                    reply.put(e)
                if n_seg >= SentenceEmbedder.MAX_BATCH_SEGMENTS:
                    break
                try:
                    msg = receiver.get(timeout=max(0.0, deadline - _time.perf_counter()))
                except queue.Empty:
                    break
                if msg is None:
                    stop = True
                    break
            if batch:
                try:
                    flat = [s for segs, _ in batch for s in segs]
                    ids, lens = tokenize_batch(tokenizer, flat, encoder.arch.max_seq_length, encoder.arch.pad_id)
                    embeddings = encoder.encode_ids(ids, lens)        # <- model.encode(&segments), :109
                    if len(flat) != len(embeddings):
                        raise EncodingFailure("# of embeddings doesn't match # of segments")
                    at = 0
                    for segs, reply in batch:
                        reply.put([EmbeddingResult(content=s, vector=v.tolist())
                                   for s, v in zip(segs, embeddings[at:at + len(segs)])])
                        at += len(segs)
                except Exception as e:
                    for _, reply in batch:
                        reply.put(e)
            SentenceEmbedder.batches_run += 1
            if stop:
                return

    batches_run = 0               # forward passes issued by runners of this process (tests)

    def _call(self, text: str, segment: bool):
        reply: queue.Queue = queue.Queue(maxsize=1)               # oneshot::channel
        self.sender.put((text, segment, reply))
        res = reply.get()
        if isinstance(res, Exception):
            raise res
        return res

    def encode(self, text: str) -> list[EmbeddingResult]:
        """segment the text into windows and embed each (embedding.rs:138-142)"""
        return self._call(text, True)

    def encode_single(self, text: str):
        """single shot, no segmentation; long text is truncated by the model (embedding.rs:146-151)"""
        res = self._call(text, False)
        return res.pop() if res else None

    def shutdown(self):
        self.sender.put(None)
