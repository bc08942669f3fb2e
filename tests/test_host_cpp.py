"""The C++ host layer (memex_b200/host/): the reference's Rust surfaces restated in C++ above the C ABI.

CPU: tokenizer / segmenter against golden vectors made with the `tokenizers` package, the SentenceEmbedder actor,
the micro-batcher, factory errors, JSON -- and that a store cannot be created without a CUDA device.
GPU: the reference's own store tests (storage/local.rs:175-242) through B200Store, the registry, the batcher, and
B200Encoder (weights read from a .safetensors file) against the Python host path on the same weights.
"""
import os
import struct
import subprocess
import json

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "tokenizer_golden.json")
BPE_GOLDEN = os.path.join(ROOT, "tests", "golden", "bpe_golden.json")


@pytest.fixture(scope="module")
def test_host():
    from memex_b200 import build as mx_build
    mx_build.build_host()
    return mx_build.HOST_TEST


def run(binary, *args):
    r = subprocess.run([binary, *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, f"{args}: rc {r.returncode}\n{r.stdout}\n{r.stderr}"
    return r.stdout


def test_tokenizer_matches_tokenizers_package(test_host):
    out = run(test_host, "tokenizer", GOLDEN)
    assert "tokenizer ok" in out


def test_cased_tokenizer_matches_tokenizers_package(test_host):
    """lowercase = false (no accent stripping): the tokenizer of DistiluseBaseMultilingualCased"""
    out = run(test_host, "tokenizer", os.path.join(ROOT, "tests", "golden", "tokenizer_cased_golden.json"))
    assert "tokenizer ok: 5 cases" in out


def test_bpe_tokenizer_matches_tokenizers_package(test_host):
    """ByteLevelBpeTokenizer (all-distilroberta-v1's pipeline, the third model segment_text accepts, embedding.rs:159):
    ids, <s> .. </s>, decode, windows and segments identical to the `tokenizers` package on tests/golden/bpe_golden.json"""
    out = run(test_host, "bpe", BPE_GOLDEN)
    assert "bpe ok: 18 cases" in out


def test_bpe_golden_is_reproducible_here():
    """the committed BPE fixture is what the installed `tokenizers` gives for the stored vocabulary and merges"""
    pytest.importorskip("tokenizers")
    import importlib.util
    spec = importlib.util.spec_from_file_location("mkbpe", os.path.join(ROOT, "tests", "golden", "make_bpe_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    g = json.load(open(BPE_GOLDEN))
    t = mk.build(g["vocab"], g["merges"])
    assert [c["text"] for c in g["cases"]] == [x[0] for x in mk.texts()]
    assert mk.cases_for(t, g["vocab"]) == g["cases"]


def test_golden_is_reproducible_here():
    """the committed fixture is what tests/golden/make_tokenizer_golden.py produces with the installed `tokenizers`"""
    tokenizers = pytest.importorskip("tokenizers")
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tests", "golden", "make_tokenizer_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    vocab, t = mk.build()
    g = json.load(open(GOLDEN))
    assert g["vocab"] == vocab
    for c in g["cases"][:10]:
        t.no_truncation()
        assert t.encode(c["text"], add_special_tokens=False).ids == c["ids"]


def test_reference_windowing_defaults():
    """memex's default windowing (embedding.rs:64-72): 256 tokens, stride 86 -> windows start every 170 tokens"""
    g = json.load(open(GOLDEN))
    multi = [c for c in g["cases"] if c["max_length"] == 256 and len(c["windows"]) > 2]
    assert multi
    for c in multi:
        n = len(c["ids"])
        starts = list(range(0, n, 170))
        assert c["windows"][0] == c["ids"][:256]
        for i, w in enumerate(c["windows"]):
            assert w == c["ids"][starts[i]:starts[i] + 256]
        assert starts[len(c["windows"]) - 1] + 256 >= n


def test_host_logic_cpu(test_host):
    out = run(test_host, "cpu")
    assert "cpu ok" in out


def test_host_library_exports():
    from memex_b200 import build as mx_build
    mx_build.build_host()
    syms = subprocess.run(["nm", "-D", "--defined-only", "-C", mx_build.HOST_LIB], capture_output=True, text=True).stdout
    for name in ("memex::get_vector_storage", "memex::B200Store::load", "memex::B200Store::search", "memex::segment_text",
                 "memex::SentenceEmbedder::spawn", "memex::SentenceEmbedder::encode_single", "memex::SearchBatcher::submit",
                 "memex::Tokenizer::encode_windows", "memex::ByteLevelBpeTokenizer::encode", "memex::Weights::from_safetensors"):
        assert name in syms, name


def write_safetensors(path, tensors):
    header, offset, blobs = {}, 0, []
    for name, arr in tensors.items():
        a = np.ascontiguousarray(arr, dtype=np.float32)
        header[name] = {"dtype": "F32", "shape": list(a.shape), "data_offsets": [offset, offset + a.nbytes]}
        blobs.append(a.tobytes())
        offset += a.nbytes
    h = json.dumps(header, separators=(",", ":")).encode()
    h += b" " * ((8 - len(h) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(h)))
        f.write(h)
        for b in blobs:
            f.write(b)


@pytest.mark.gpu
def test_store_registry_batcher_gpu(test_host, tmp_path):
    out = run(test_host, "gpu", str(tmp_path))
    assert "gpu ok" in out


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [0, 1])
def test_cpp_encoder_equals_python_host(test_host, tmp_path, precision):
    """B200Encoder (C++, weights from .safetensors) and memex_b200.embedding.B200Encoder (Python) drive the same
    library: identical bits on the same weights and ids"""
    from oracle import encoder as enc_oracle
    from memex_b200.embedding import Architecture, B200Encoder
    cfg = enc_oracle.MINILM_L6
    w = enc_oracle.make_weights(cfg, seed=5)
    ids, lens = enc_oracle.make_inputs(cfg, 6, 48, seed=6, ragged=True, min_len=3)
    st = str(tmp_path / "model.safetensors")
    write_safetensors(st, {("bert." + k if i % 2 else k): v for i, (k, v) in enumerate(w.items())})
    with open(tmp_path / "ids.bin", "wb") as f:
        f.write(struct.pack("<II", ids.shape[0], ids.shape[1]))
        f.write(np.ascontiguousarray(ids, dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(lens, dtype=np.int32).tobytes())
    run(test_host, "encode", st, str(tmp_path / "ids.bin"), str(tmp_path / "out.bin"), str(cfg.layers), str(cfg.hidden),
        str(cfg.heads), str(cfg.ffn), str(cfg.vocab), str(cfg.max_pos), str(precision))
    got = np.fromfile(tmp_path / "out.bin", dtype=np.float32).reshape(ids.shape[0], cfg.hidden)
    arch = Architecture(cfg.layers, cfg.hidden, cfg.heads, cfg.ffn, cfg.vocab, cfg.max_pos, cfg.type_vocab, cfg.ln_eps, cfg.normalize)
    e = B200Encoder(arch, w, precision={0: "bf16", 1: "f32"}[precision], max_tokens=ids.size)
    want = e.encode_ids(ids, lens)
    assert (got.view(np.uint32) == want.view(np.uint32)).all()
    ref = enc_oracle.np_encode(cfg, w, ids, lens)
    assert ((got * ref).sum(1) >= (1 - 2e-4 if precision == 0 else 1 - 1e-6)).all()
