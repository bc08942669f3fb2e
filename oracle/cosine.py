"""oracle/cosine.py -- TEST INFRASTRUCTURE ONLY.

Python face of the CPU oracle for the vector-store search path:

* ``exact_topk`` / ``scores_of`` / ``hnsw_*``  -> ctypes into ``oracle/_build/libmemex_oracle.so``
  (cosine_oracle.c, hnsw_oracle.cpp; built by ``oracle/Makefile``).
* ``dist_cosine_np`` / ``exact_topk_np``        -> an independent numpy restatement of the same
  arithmetic, used to cross-check the C one on small cases.

Both follow reference ``lib/libmemex/src/storage/local.rs:62-91`` (1-based ids, ascending
distance, ``score = 1 - 1/(1/d)``) and hnsw_rs 0.1.20's ``DistCosine`` (f32 products folded
left-to-right in f64).  Only tests/, bench.py's CPU-baseline legs and ``__graft_entry__.smoke``
may import this module; memex_b200/ never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmemex_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C/C++ oracle (gcc only; no reference sources are involved)."""
    if force or not os.path.exists(_SO):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        f32p = ctypes.POINTER(ctypes.c_float)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        L.mxo_dist_cosine.restype = ctypes.c_float
        L.mxo_dist_cosine.argtypes = [f32p, f32p, ctypes.c_size_t]
        L.mxo_dot.restype = ctypes.c_float
        L.mxo_dot.argtypes = [f32p, f32p, ctypes.c_size_t]
        L.mxo_similarity.restype = ctypes.c_float
        L.mxo_similarity.argtypes = [ctypes.c_float]
        L.mxo_exact_topk.restype = ctypes.c_int
        L.mxo_exact_topk.argtypes = [f32p, ctypes.c_uint64, ctypes.c_uint32, f32p, ctypes.c_uint32,
                                     ctypes.c_uint32, ctypes.c_uint32, u64p, f32p, u32p]
        L.mxo_scores_of.restype = None
        L.mxo_scores_of.argtypes = [f32p, ctypes.c_uint32, f32p, u64p, ctypes.c_uint32,
                                    ctypes.c_uint32, f32p]
        L.mxo_num_threads.restype = ctypes.c_int
        L.mxo_hnsw_new.restype = ctypes.c_void_p
        L.mxo_hnsw_new.argtypes = [ctypes.c_uint32, ctypes.c_uint64]
        L.mxo_hnsw_free.argtypes = [ctypes.c_void_p]
        L.mxo_hnsw_insert.argtypes = [ctypes.c_void_p, f32p, ctypes.c_uint64]
        L.mxo_hnsw_insert_parallel.argtypes = [ctypes.c_void_p, f32p, ctypes.c_uint64, ctypes.c_int]
        L.mxo_hnsw_len.restype = ctypes.c_uint64
        L.mxo_hnsw_len.argtypes = [ctypes.c_void_p]
        L.mxo_hnsw_search.argtypes = [ctypes.c_void_p, f32p, ctypes.c_uint32, ctypes.c_uint32,
                                      ctypes.c_uint32, ctypes.c_int, u64p, f32p, u32p]
        _lib = L
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


METRIC = {"cosine": 0, "dot": 1}


def dist_cosine(a, b) -> float:
    a, pa = _f32(a)
    b, pb = _f32(b)
    return float(lib().mxo_dist_cosine(pa, pb, a.size))


def similarity(d: float) -> float:
    return float(lib().mxo_similarity(ctypes.c_float(d)))


def exact_topk(corpus, queries, k: int, metric: str = "cosine"):
    """-> (ids [nq,k] uint64 1-based, scores [nq,k] f32, counts [nq] uint32)"""
    corpus, pc = _f32(corpus)
    queries, pq = _f32(np.atleast_2d(queries))
    n, d = corpus.shape if corpus.ndim == 2 else (0, queries.shape[1])
    nq = queries.shape[0]
    ids = np.zeros((nq, k), dtype=np.uint64)
    scores = np.zeros((nq, k), dtype=np.float32)
    counts = np.zeros(nq, dtype=np.uint32)
    rc = lib().mxo_exact_topk(pc, n, d, pq, nq, k, METRIC[metric],
                              ids.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                              scores.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                              counts.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    if rc != 0:
        raise RuntimeError(f"mxo_exact_topk failed: {rc}")
    return ids, scores, counts


def scores_of(corpus, query, ids, metric: str = "cosine"):
    corpus, pc = _f32(corpus)
    query, pq = _f32(query)
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    out = np.zeros(ids.size, dtype=np.float32)
    lib().mxo_scores_of(pc, corpus.shape[1], pq, ids.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                        ids.size, METRIC[metric], out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def num_threads() -> int:
    return int(lib().mxo_num_threads())


class HnswOracle:
    """HNSW restatement with the reference's parameters (timed baseline / recall only)."""

    def __init__(self, dim: int, seed: int = 0):
        self._h = lib().mxo_hnsw_new(dim, seed)
        self.dim = dim

    def insert(self, vecs, threads: int = 1):
        """threads = 1: one point at a time as local.rs:62-69; > 1: the crate's parallel_insert scheme
        (bench samples only -- the build is never the timed part)"""
        vecs, pv = _f32(np.atleast_2d(vecs))
        assert vecs.shape[1] == self.dim
        if threads > 1:
            lib().mxo_hnsw_insert_parallel(self._h, pv, vecs.shape[0], threads)
        else:
            lib().mxo_hnsw_insert(self._h, pv, vecs.shape[0])

    def __len__(self):
        return int(lib().mxo_hnsw_len(self._h))

    def search(self, queries, k: int, ef: int = 32, threads: int = 1):
        """-> (ids, scores, counts); scores follow local.rs:86"""
        queries, pq = _f32(np.atleast_2d(queries))
        nq = queries.shape[0]
        ids = np.zeros((nq, k), dtype=np.uint64)
        dists = np.zeros((nq, k), dtype=np.float32)
        counts = np.zeros(nq, dtype=np.uint32)
        lib().mxo_hnsw_search(self._h, pq, nq, k, ef, threads,
                              ids.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                              dists.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                              counts.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
        with np.errstate(divide="ignore"):
            scores = (np.float32(1.0) - np.float32(1.0) / (np.float32(1.0) / dists)).astype(np.float32)
        return ids, scores, counts

    def __del__(self):
        try:
            lib().mxo_hnsw_free(self._h)
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------
# independent numpy restatement (small cases; cross-checks the C oracle)
# ----------------------------------------------------------------------------------------------

def dist_cosine_np(a, b) -> np.float32:
    """hnsw_rs DistCosine::eval: f32 products, sequential f64 fold (np.cumsum is sequential)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    ab = np.cumsum((a * b).astype(np.float32).astype(np.float64))[-1]
    aa = np.cumsum((a * a).astype(np.float32).astype(np.float64))[-1]
    bb = np.cumsum((b * b).astype(np.float32).astype(np.float64))[-1]
    if aa > 0.0 and bb > 0.0:
        return np.float32(max(0.0, 1.0 - ab / np.sqrt(aa * bb)))
    return np.float32(0.0)


def similarity_np(d) -> np.float32:
    """local.rs:86 in f32"""
    d = np.float32(d)
    with np.errstate(divide="ignore"):
        return np.float32(1.0) - np.float32(1.0) / (np.float32(1.0) / d)


def exact_topk_np(corpus, query, k: int):
    """-> list[(id 1-based, score f32)] best first; ties -> lower id"""
    corpus = np.asarray(corpus, dtype=np.float32)
    d = [dist_cosine_np(query, row) for row in corpus]
    order = sorted(range(len(d)), key=lambda i: (d[i], i))[:k]
    return [(i + 1, similarity_np(d[i])) for i in order]
