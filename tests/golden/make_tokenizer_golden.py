#!/usr/bin/env python
"""Golden vectors for the C++ BertTokenizer / segment_text (memex_b200/host/tokenizer.cpp), made with the `tokenizers`
package -- the Python build of the crate the reference uses (reference lib/libmemex/src/llm/embedding.rs:163-195;
Cargo.lock pins tokenizers 0.14.0, this image has the version printed into the fixture).

No vocab.txt of a real checkpoint is on this box (no network), so the vocabulary is synthetic: the pipeline
(BertNormalizer -> BertPreTokenizer -> WordPiece, WordPiece decoder with cleanup, truncation with stride) is what is
pinned, on text that exercises accents, case, punctuation, apostrophes, CJK, control characters, unknown and
over-long words, and multi-window documents with memex's default 256 / 86 windowing.

    python tests/golden/make_tokenizer_golden.py      # rewrites tests/golden/tokenizer_golden.json
"""
import json
import os
import random

import tokenizers
from tokenizers import Tokenizer, decoders, models, normalizers, pre_tokenizers

HERE = os.path.dirname(os.path.abspath(__file__))

BASE = """the of and to in a is that for it as was with be by on not he i this are or his from at which but have an had they
you were their one all we can her has there been if more when will would who so no out up said what its about than into them
only other new some could time these two may then do first any my now such like our over man me even most made after also did
many before must through back years where much your way well down should because each just those people mr how too little
state good very make world still own see men work long get here between both life being under never day same another know
while last might us great old year off come since against go came right used take three quick brown fox jumps lazy dog
president congress america american nation union tonight people country economy jobs families energy vaccine ukraine freedom
cafe naive resume uber senor facade zurich francois
run walk talk play jump embed token vector search segment window model sentence transformer attention layer hidden"""
SUFFIX = ["##s", "##ing", "##ed", "##er", "##ly", "##tion", "##ness", "##able", "##es", "##e", "##n", "##t", "##a", "##o", "##i",
          "##r", "##y", "##d", "##m", "##ation", "##al", "##ers", "##est", "##ment"]
PUNCT = list(".,!?;:'\"()-[]{}/\\@#$%^&*_+=<>|~`") + ["—", "’", "“", "”", "。"]
EXTRA = ["中", "文", "日", "本", "α", "β", "γεια", "привет",
         "мир", "straße", "æ", "ø", "1", "2", "3", "2024", "##0", "##1", "##2"]


def build():
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    seen = set(vocab)
    for w in BASE.split() + list("abcdefghijklmnopqrstuvwxyz") + SUFFIX + PUNCT + EXTRA:
        if w not in seen:
            vocab.append(w)
            seen.add(w)
    t = Tokenizer(models.WordPiece({w: i for i, w in enumerate(vocab)}, unk_token="[UNK]", max_input_chars_per_word=100))
    t.normalizer = normalizers.BertNormalizer(clean_text=True, handle_chinese_chars=True, strip_accents=None, lowercase=True)
    t.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
    t.decoder = decoders.WordPiece(prefix="##", cleanup=True)
    t.add_special_tokens(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"])
    return vocab, t


def build_cased():
    """distiluse-base-multilingual-cased's tokenizer.json: the same pipeline with lowercase = false (and with it no accent
    stripping), on a vocabulary that keeps case and accents"""
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    seen = set(vocab)
    words = BASE.split() + [w.capitalize() for w in BASE.split()[:80]] + ["Café", "café", "naïve", "Über", "über", "Zürich",
                                                                            "François", "STRASSE", "Straße", "Привет", "МИР", "мир"]
    for w in words + list("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZéèïüßñç") + SUFFIX + PUNCT + EXTRA:
        if w not in seen:
            vocab.append(w)
            seen.add(w)
    t = Tokenizer(models.WordPiece({w: i for i, w in enumerate(vocab)}, unk_token="[UNK]", max_input_chars_per_word=100))
    t.normalizer = normalizers.BertNormalizer(clean_text=True, handle_chinese_chars=True, strip_accents=None, lowercase=False)
    t.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
    t.decoder = decoders.WordPiece(prefix="##", cleanup=True)
    t.add_special_tokens(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"])
    return vocab, t


def cases_for(vocab, t, texts):
    cases = []
    for text, max_length, stride in texts:
        t.no_truncation()
        enc = t.encode(text, add_special_tokens=False)
        ids = enc.ids
        ids_special = [vocab.index("[CLS]")] + ids + [vocab.index("[SEP]")]
        decoded = t.decode(ids, skip_special_tokens=True)
        t.enable_truncation(max_length=max_length, stride=stride)
        e2 = t.encode(text, add_special_tokens=False)
        windows = [e2.ids] + [o.ids for o in e2.overflowing]
        segments = [t.decode(e2.ids, skip_special_tokens=True).replace(" ' ", "'")]
        segments += [t.decode(o.ids, skip_special_tokens=True) for o in e2.overflowing]
        t.no_truncation()
        cases.append(dict(text=text, ids=ids, ids_special=ids_special, decoded=decoded, max_length=max_length, stride=stride,
                          windows=windows, segments=segments))
    return cases


def cased_texts():
    rng = random.Random(20260103)
    doc = long_document(rng, 300)
    return [
        ("The Quick brown Fox jumps over the lazy Dog.", 256, 86),
        ("Café café CAFÉ naïve Naïve Über über Zürich François STRASSE Straße straße", 256, 86),
        ("Привет МИР мир α β Γεια 中文 mixed中Text", 256, 86),
        ("don't It's We're I'M   tabs\tand\nnewlines \u00e9 vs e\u0301 combining", 256, 86),
        (" ".join(w.capitalize() if i % 3 == 0 else w for i, w in enumerate(doc.split())), 32, 8),
    ]


def long_document(rng, n_words):
    words = BASE.split()
    out = []
    for i in range(n_words):
        w = rng.choice(words)
        r = rng.random()
        if r < 0.08:
            w = w.capitalize()
        elif r < 0.12:
            w += rng.choice(["s", "ing", "ed", "ly", "ers"])
        elif r < 0.14:
            w = "zzqx" + w          # unknown
        out.append(w)
        if rng.random() < 0.07:
            out[-1] += rng.choice([".", ",", "!", "?", ";"])
        if rng.random() < 0.02:
            out.append(rng.choice(["don't", "it's", "we're", "I'm", "they've", "do not", "isn't"]))
    return " ".join(out)


def main():
    vocab, t = build()
    rng = random.Random(20260101)
    texts = [
        ("this is a test string", 256, 128),          # the reference's own test_tokenizer text (embedding.rs:206)
        ("", 256, 86),
        ("The quick brown fox jumps over the lazy dog.", 256, 86),
        ("Café naïve résumé Über Señor façade Zürich François STRASSE straße Æ Ø", 256, 86),
        ("don't it's we're I'm they've   do not  isn't , \"quoted\" (paren) [x] {y} a/b a\\b #1 $2 50% x^2 a&b *c* _d_ e+f=g <h> i|j ~k `l`", 256, 86),
        ("中文日本 mixed中text 。 αβ Γεια ΣΟΦΙΑ Привет МИР", 256, 86),
        ("tabs\tand\nnewlines\r\nand nbsp emspace ctl\x00\x01\x7f​﻿zero � repl", 256, 86),
        ("x" * 101 + " " + "y" * 100 + " running jumps talked walker quickly embedding tokens vectors searched", 256, 86),
        ("emoji \U0001F600 and é combining and ẛ̣ and 가나 hangul Å angstrom ﬁ ligature", 256, 86),
        ("America—the nation’s “Union” tonight: jobs, energy, freedom; Ukraine & vaccine!", 256, 86),
        (long_document(rng, 40), 8, 3),
        (long_document(rng, 90), 16, 5),
        (long_document(rng, 700), 256, 86),
        (long_document(rng, 1100), 256, 86),
        (long_document(rng, 256), 256, 86),
        (long_document(rng, 60), 16, 15),
    ]
    cases = []
    for text, max_length, stride in texts:
        t.no_truncation()
        enc = t.encode(text, add_special_tokens=False)
        ids = enc.ids
        # HF tokenizer.json adds "[CLS] $A [SEP]" through a TemplateProcessing post-processor; without one, the same thing:
        ids_special = [vocab.index("[CLS]")] + ids + [vocab.index("[SEP]")]
        decoded = t.decode(ids, skip_special_tokens=True)
        # segment_text, embedding.rs:170-195
        t.enable_truncation(max_length=max_length, stride=stride)
        e2 = t.encode(text, add_special_tokens=False)
        windows = [e2.ids] + [o.ids for o in e2.overflowing]
        segments = [t.decode(e2.ids, skip_special_tokens=True).replace(" ' ", "'")]
        segments += [t.decode(o.ids, skip_special_tokens=True) for o in e2.overflowing]
        cases.append(dict(text=text, ids=ids, ids_special=ids_special, decoded=decoded, max_length=max_length, stride=stride,
                          windows=windows, segments=segments))
    out = dict(tokenizers_version=tokenizers.__version__, vocab=vocab, cases=cases)
    with open(os.path.join(HERE, "tokenizer_golden.json"), "w") as f:
        json.dump(out, f, ensure_ascii=True)
    print(f"{len(cases)} cases, {len(vocab)} vocab entries, windows per case: {[len(c['windows']) for c in cases]}")
    vocab_c, t_c = build_cased()
    cased = cases_for(vocab_c, t_c, cased_texts())
    with open(os.path.join(HERE, "tokenizer_cased_golden.json"), "w") as f:
        json.dump(dict(tokenizers_version=tokenizers.__version__, lowercase=False, vocab=vocab_c, cases=cased), f, ensure_ascii=True)
    print(f"cased: {len(cased)} cases, {len(vocab_c)} vocab entries, windows per case: {[len(c['windows']) for c in cased]}")


if __name__ == "__main__":
    main()
