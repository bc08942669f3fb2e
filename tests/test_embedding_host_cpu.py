"""CPU tests of the Python host mirror of memex's embedder (memex_b200/embedding.py): the model table, segment_text,
the batch the encoder receives and the SentenceEmbedder actor (reference lib/libmemex/src/llm/embedding.rs:24-198) --
with a fake encoder in place of the GPU one, so nothing here computes an embedding."""
import dataclasses
import importlib.util
import json
import os

import numpy as np
import pytest

from memex_b200.embedding import (ARCHITECTURES, EmbeddingsModelType, EncodingFailure, ModelConfig, SentenceEmbedder,
                                  SetupError, segment_text, tokenize_batch)
from oracle import encoder as enc_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(GOLD, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def bert_tok():
    pytest.importorskip("tokenizers")
    from tokenizers import processors
    vocab, tok = _load("make_tokenizer_golden").build()
    tok.post_processor = processors.TemplateProcessing(single="[CLS] $A [SEP]", special_tokens=[
        ("[CLS]", vocab.index("[CLS]")), ("[SEP]", vocab.index("[SEP]"))])
    return tok, json.load(open(os.path.join(GOLD, "tokenizer_golden.json")))


@pytest.fixture(scope="module")
def bpe_tok():
    pytest.importorskip("tokenizers")
    g = json.load(open(os.path.join(GOLD, "bpe_golden.json")))
    return _load("make_bpe_golden").build(g["vocab"], g["merges"]), g


class FakeEncoder:
    """stands in for B200Encoder: the 'embedding' of a row is (number of real tokens, first id, last real id)"""

    def __init__(self, arch, drop_last=False):
        self.arch = arch
        self.drop_last = drop_last
        self.calls = []

    def encode_ids(self, ids, lens):
        self.calls.append((ids.copy(), lens.copy()))
        out = np.stack([np.array([lens[b], ids[b, 0], ids[b, lens[b] - 1]], dtype=np.float32) for b in range(len(lens))])
        return out[:-1] if self.drop_last else out


def test_model_table_agrees_with_the_oracle_configs():
    """embedding.rs:24-55: EVERY variant of the enum has an architecture, and it is the one the oracle restates"""
    assert set(ARCHITECTURES) == set(EmbeddingsModelType)
    for kind, cfg in ((EmbeddingsModelType.SentenceT5Base, enc_oracle.SENTENCE_T5_BASE),
                      (EmbeddingsModelType.AllMiniLmL6V2, enc_oracle.MINILM_L6),
                      (EmbeddingsModelType.AllMiniLmL12V2, enc_oracle.MINILM_L12),
                      (EmbeddingsModelType.BertBaseNliMeanTokens, dataclasses.replace(enc_oracle.BERT_BASE, normalize=False)),
                      (EmbeddingsModelType.AllDistilrobertaV1, enc_oracle.DISTILROBERTA),
                      (EmbeddingsModelType.DistiluseBaseMultilingualCased, enc_oracle.DISTILUSE),
                      (EmbeddingsModelType.ParaphraseAlbertSmallV2, enc_oracle.ALBERT_SMALL)):
        a = ARCHITECTURES[kind]
        for f in ("layers", "hidden", "heads", "ffn", "vocab", "max_pos", "type_vocab", "ln_eps", "normalize", "family",
                  "pos_offset", "pad_id", "dense_out", "dense_act", "dense_bias", "ffn_act", "embed_dim", "share_layers",
                  "d_kv", "rel_buckets", "rel_max_distance"):
            assert getattr(a, f) == getattr(cfg, f), (kind, f)
    assert ARCHITECTURES[EmbeddingsModelType.DistiluseBaseMultilingualCased].out_dim == 512
    assert ModelConfig().model is EmbeddingsModelType.AllMiniLmL12V2 and (ModelConfig().max_length, ModelConfig().stride) == (256, 86)


@pytest.mark.parametrize("which", ["bert", "bpe"])
def test_segment_text_matches_the_golden_segments(which, bert_tok, bpe_tok):
    tok, g = bert_tok if which == "bert" else bpe_tok
    if which == "bert":   # the fixture was made without a post-processor (one case has stride = max_length - 1)
        tok = _load("make_tokenizer_golden").build()[1]
    model = EmbeddingsModelType.AllMiniLmL6V2 if which == "bert" else EmbeddingsModelType.AllDistilrobertaV1
    for c in g["cases"]:
        mc = ModelConfig(model=model, max_length=c["max_length"], stride=c["stride"])
        assert segment_text(mc, c["text"], tok) == c["segments"]
    for bad in (EmbeddingsModelType.SentenceT5Base, EmbeddingsModelType.BertBaseNliMeanTokens,
                EmbeddingsModelType.ParaphraseAlbertSmallV2):
        with pytest.raises(SetupError):                       # embedding.rs:156-161 "Model not supported yet"
            segment_text(ModelConfig(model=bad), "x", tok)


def test_batches_carry_the_models_special_and_pad_ids(bert_tok, bpe_tok):
    tok, _ = bpe_tok
    ids, lens = tokenize_batch(tok, ["the quick brown fox jumps over the lazy dog", "a"], 8, pad_id=1)
    assert ids.shape == (2, 8) and list(lens) == [8, 3]                      # truncated to max_seq_length, specials kept
    assert ids[0, 0] == 0 and ids[0, 7] == 2 and ids[1, 0] == 0 and ids[1, 2] == 2   # <s> ... </s>
    assert (ids[1, 3:] == 1).all()                                            # <pad>
    tok, _ = bert_tok
    ids, lens = tokenize_batch(tok, ["the quick brown fox", "a"], 16)
    cls, sep = tok.token_to_id("[CLS]"), tok.token_to_id("[SEP]")
    assert ids[0, 0] == cls and ids[0, lens[0] - 1] == sep and (ids[1, lens[1]:] == 0).all()


@pytest.mark.parametrize("which", ["bert", "bpe"])
def test_sentence_embedder_actor(which, bert_tok, bpe_tok):
    """embedding.rs:77-152: encode = segment + embed every window, encode_single = one truncated shot"""
    tok, g = bert_tok if which == "bert" else bpe_tok
    kind = EmbeddingsModelType.AllMiniLmL6V2 if which == "bert" else EmbeddingsModelType.AllDistilrobertaV1
    arch = dataclasses.replace(ARCHITECTURES[kind], max_seq_length=32)
    enc = FakeEncoder(arch)
    mc = ModelConfig(model=kind, max_length=24, stride=8)
    text = max((c["text"] for c in g["cases"]), key=len)
    handle, embedder = SentenceEmbedder.spawn(mc, enc, tok)
    try:
        results = embedder.encode(text)
        segments = segment_text(mc, text, tok)
        assert len(segments) > 10 and [r.content for r in results] == segments
        ids, lens = enc.calls[0]
        assert ids.shape[0] == len(segments) and ids.shape[1] <= 32 and (lens <= 32).all()
        assert (ids[np.arange(len(lens)), lens - 1] == (2 if which == "bpe" else tok.token_to_id("[SEP]"))).all()
        for b in range(len(lens)):
            assert (ids[b, lens[b]:] == arch.pad_id).all()
            assert results[b].vector == [float(lens[b]), float(ids[b, 0]), float(ids[b, lens[b] - 1])]
        single = embedder.encode_single(text)                 # no segmentation: the model truncates (embedding.rs:146-151)
        assert single.content == text and single.vector[0] == 32.0
        assert enc.calls[-1][0].shape == (1, 32)
    finally:
        embedder.shutdown()
        handle.join(timeout=10)
    assert not handle.is_alive()


def test_actor_reports_a_count_mismatch_and_survives(bert_tok):
    """embedding.rs:110-115: '# of embeddings doesn't match # of segments' -- the reference's runner dies with it; here the
    caller gets the error and the actor keeps serving"""
    tok, _ = bert_tok
    arch = dataclasses.replace(ARCHITECTURES[EmbeddingsModelType.AllMiniLmL6V2], max_seq_length=32)
    enc = FakeEncoder(arch, drop_last=True)
    handle, embedder = SentenceEmbedder.spawn(ModelConfig(model=EmbeddingsModelType.AllMiniLmL6V2), enc, tok)
    try:
        with pytest.raises(EncodingFailure):
            embedder.encode("the quick brown fox")
        enc.drop_last = False
        assert len(embedder.encode("the quick brown fox")) == 1
        with pytest.raises(SetupError):
            SentenceEmbedder.spawn(ModelConfig(model=EmbeddingsModelType.SentenceT5Base), enc, tok)[1].encode("x")
    finally:
        embedder.shutdown()
        handle.join(timeout=10)


def test_runner_batches_concurrent_requests_into_one_forward_pass(bert_tok):
    """the reference's runner embeds one document per message (embedding.rs:102-132).  Here requests
    that are queued together share ONE forward pass, and every caller still gets exactly its own rows, in order."""
    import threading
    import time
    tok, g = bert_tok
    arch = dataclasses.replace(ARCHITECTURES[EmbeddingsModelType.AllMiniLmL6V2], max_seq_length=32)

    class SlowEncoder(FakeEncoder):
        def encode_ids(self, ids, lens):
            time.sleep(0.05)                 # while one pass runs, the other callers' requests pile up in the channel
            return super().encode_ids(ids, lens)

    enc = SlowEncoder(arch)
    mc = ModelConfig(model=EmbeddingsModelType.AllMiniLmL6V2, max_length=24, stride=8)
    texts = [c["text"] for c in g["cases"] if c["text"].strip()][:12]
    handle, embedder = SentenceEmbedder.spawn(mc, enc, tok)
    results = [None] * len(texts)

    def call(i):
        results[i] = embedder.encode(texts[i]) if i % 3 else [embedder.encode_single(texts[i])]

    try:
        threads = [threading.Thread(target=call, args=(i,)) for i in range(len(texts))]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=30)
        assert all(r is not None for r in results)
        assert len(enc.calls) < len(texts)                       # fewer forward passes than requests
        assert max(len(lens) for _, lens in enc.calls) > 1
        # every caller got the rows of ITS segments: same content and same vectors as a call made alone
        for i, text in enumerate(texts):
            alone = embedder.encode(text) if i % 3 else [embedder.encode_single(text)]
            assert [r.content for r in results[i]] == [r.content for r in alone]
            assert [r.vector for r in results[i]] == [r.vector for r in alone]
    finally:
        embedder.shutdown()
        handle.join(timeout=10)
