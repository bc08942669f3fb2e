"""memex_b200 -- B200-native embedding + vector-search hot path of memex.

csrc/        hand-written CUDA for sm_100a + the C ABI (include/memex_b200.h)
capi.py      ctypes binding of that ABI (what the Rust shim binds)
storage.py   host-side mirror of memex's VectorStore surface (storage/mod.rs, storage/local.rs)
embedding.py host-side mirror of memex's SentenceEmbedder (llm/embedding.rs)
sharded.py   row-sharded multi-GPU search: one process per GPU, one all-gather of top-k candidates
"""
from . import capi  # noqa: F401

__all__ = ["capi", "storage", "embedding", "sharded", "build"]
