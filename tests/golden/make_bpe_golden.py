#!/usr/bin/env python
"""Golden vectors for the C++ ByteLevelBpeTokenizer / segment_text (memex_b200/host/tokenizer.cpp), made with the
`tokenizers` package -- the Python build of the crate the reference uses (reference
lib/libmemex/src/llm/embedding.rs:163-195) -- for the third model segment_text accepts, all-distilroberta-v1
(embedding.rs:159): no normalizer, ByteLevel pre-tokenizer (add_prefix_space = false), BPE, ByteLevel decoder,
RobertaProcessing.

No vocab.json / merges.txt of the real checkpoint is on this box (no network), so a small byte-level BPE vocabulary is
trained here on the synthetic text below (full 256-byte alphabet, RoBERTa's special tokens at ids 0-4); what is pinned
is the pipeline: the GPT-2 split regex (contractions, optional leading space, \\p{L} / \\p{N} over all planes, Unicode
white space and its look-ahead rule), the byte <-> code point map, merge order, lossy UTF-8 decoding of windows that cut
a character, and memex's 256 / 86 windowing.  The trained vocabulary and merges are stored in the fixture, so the check
does not depend on the trainer being reproducible.

    python tests/golden/make_bpe_golden.py      # rewrites tests/golden/bpe_golden.json
"""
import json
import os
import random

import tokenizers
from tokenizers import Tokenizer, decoders, models, pre_tokenizers, processors, trainers

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "bpe_golden.json")

BASE = """the of and to in a is that for it as was with be by on not he i this are or his from at which but have an had they
you were their one all we can her has there been if more when will would who so no out up said what its about than into them
only other new some could time these two may then do first any my now such like our over man me even most made after also did
president congress america american nation union tonight people country economy jobs families energy vaccine ukraine freedom
quick brown fox jumps lazy dog embed token vector search segment window model sentence transformer attention layer hidden
café naïve résumé über señor straße привет мир 中文 日本 2024 1999 42 3.14"""
SPECIALS = ["<s>", "<pad>", "</s>", "<unk>", "<mask>"]


def long_document(rng, n_words):
    words = BASE.split()
    out = []
    for _ in range(n_words):
        w = rng.choice(words)
        r = rng.random()
        if r < 0.08:
            w = w.capitalize()
        elif r < 0.12:
            w += rng.choice(["s", "ing", "ed", "ly", "ers"])
        elif r < 0.14:
            w = "zzqx" + w
        out.append(w)
        if rng.random() < 0.07:
            out[-1] += rng.choice([".", ",", "!", "?", ";"])
        if rng.random() < 0.03:
            out.append(rng.choice(["don't", "it's", "we're", "I'm", "they've", "she'll", "he'd", "été", "中文日本"]))
        if rng.random() < 0.02:
            out[-1] += rng.choice(["\n", "\n\n", "  ", "\t"])
    return " ".join(out)


def train():
    rng = random.Random(7)
    corpus = [long_document(rng, 400) for _ in range(30)] + [BASE]
    t = Tokenizer(models.BPE())
    t.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    t.decoder = decoders.ByteLevel()
    trainer = trainers.BpeTrainer(vocab_size=700, special_tokens=SPECIALS, show_progress=False,
                                  initial_alphabet=pre_tokenizers.ByteLevel.alphabet())
    t.train_from_iterator(corpus, trainer)
    model = json.loads(t.to_str())["model"]
    vocab = [None] * len(model["vocab"])
    for tok, i in model["vocab"].items():
        vocab[i] = tok
    merges = [m.split(" ") if isinstance(m, str) else list(m) for m in model["merges"]]
    return vocab, merges


def build(vocab, merges):
    """the tokenizer a RoBERTa tokenizer.json describes, from a stored vocabulary"""
    t = Tokenizer(models.BPE({tok: i for i, tok in enumerate(vocab)}, [tuple(m) for m in merges]))
    t.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    t.decoder = decoders.ByteLevel()
    t.post_processor = processors.RobertaProcessing(("</s>", vocab.index("</s>")), ("<s>", vocab.index("<s>")))
    t.add_special_tokens(SPECIALS)
    return t


def texts():
    rng = random.Random(20260102)
    return [
        ("this is a test string", 256, 128),          # the reference's own test_tokenizer text (embedding.rs:206)
        ("", 256, 86),
        (" ", 256, 86),
        ("The quick brown fox jumps over the lazy dog.", 256, 86),
        ("don't DON'T it'll we're I'm they've he'd 's 't x'sy ''s 'llama '", 256, 86),
        ("  two  spaces   three\n\nnewlines \n mixed\t\ttabs  trailing  ", 256, 86),
        ("123abc ab12 3.14 1,000 ٣٤ Ⅷ ² x² \U0001d7d8\U0001d7d9", 256, 86),
        ("café CAFÉ naïve é कि _x_ a_b snake_case __init__", 256, 86),
        ("a\u00a0\u00a0b\u001c\u001cc\u3000\u3000d\u2028\u2028e\u0085\u0085f\u200b\u200bg\u180e\u180eh\u1680\u1680i"
         "\u2003\u2003j\u202f\u202fk\u205f\u205fl\u000b\u000bm\u000c\u000cn\ufeff\ufeffo\u0000\u0000p \u00a0x\u00a0 y", 256, 86),
        ("a\U0001d49cb \U00020000c \U0001F600d emoji \U0001F600\U0001F601 end", 256, 86),
        ("中文日本 mixed中text 。 αβ Γεια Привет МИР", 256, 86),
        ("America—the nation’s “Union” tonight: jobs, energy, freedom; Ukraine & vaccine!", 256, 86),
        ("中文日本語のテキストを窓で切る \U0001F600\U0001F601\U0001F602 éèêë", 5, 2),   # windows cut inside characters
        (long_document(rng, 40), 8, 3),
        (long_document(rng, 90), 16, 5),
        (long_document(rng, 700), 256, 86),
        (long_document(rng, 1100), 256, 86),
        (long_document(rng, 60), 16, 13),     # stride must stay below max_length - 2 (the post-processor's specials)
    ]


def cases_for(t, vocab):
    cases = []
    for text, max_length, stride in texts():
        t.no_truncation()
        ids = t.encode(text, add_special_tokens=False).ids
        ids_special = t.encode(text, add_special_tokens=True).ids
        decoded = t.decode(ids, skip_special_tokens=True)
        # segment_text, embedding.rs:170-195
        t.enable_truncation(max_length=max_length, stride=stride)
        e2 = t.encode(text, add_special_tokens=False)
        windows = [e2.ids] + [o.ids for o in e2.overflowing]
        segments = [t.decode(e2.ids, skip_special_tokens=True).replace(" ' ", "'")]
        segments += [t.decode(o.ids, skip_special_tokens=True) for o in e2.overflowing]
        t.no_truncation()
        cases.append(dict(text=text, ids=ids, ids_special=ids_special, decoded=decoded, max_length=max_length, stride=stride,
                          windows=windows, segments=segments))
    return cases


def main():
    vocab, merges = train()
    t = build(vocab, merges)
    cases = cases_for(t, vocab)
    out = dict(tokenizers_version=tokenizers.__version__, vocab=vocab, merges=merges, cases=cases)
    with open(OUT, "w") as f:
        json.dump(out, f, ensure_ascii=True)
    print(f"{len(cases)} cases, {len(vocab)} vocab entries, {len(merges)} merges, windows per case: {[len(c['windows']) for c in cases]}")


if __name__ == "__main__":
    main()
