"""Builds libmemex_b200.so (hand-written CUDA for sm_100a + the C ABI of include/memex_b200.h).

In-tree build: objects and the shared library go to ``memex_b200/_lib/`` (git-ignored, but it
travels to the GPU box with the gpurun snapshot).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_lib")
OBJ = os.path.join(OUT, "obj")
LIB = os.path.join(OUT, "libmemex_b200.so")
HOST_LIB = os.path.join(OUT, "libmemex_host.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unknown-pragmas", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint(src: str) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for f in sorted(os.listdir(CSRC)):
        if f == src or f.endswith((".cuh", ".h", ".hpp")):
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(HERE, "..", "include", "memex_b200.h"), "rb").read())
    return h.hexdigest()


def _compile_one(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ, src[:-3] + ".o")
    stamp = obj + ".sha"
    fp = _fingerprint(src)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == fp:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    open(stamp, "w").write(fp)
    return obj


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, verbose), srcs))
    newest = max(os.path.getmtime(o) for o in objs)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


HOST = os.path.join(HERE, "host")
HOST_TEST = os.path.join(OUT, "test_host")


def build_host() -> str:
    """C++ host layer (memex_b200/host/) -> _lib/libmemex_host.so + _lib/test_host, linked against libmemex_b200.so"""
    if not os.path.exists(LIB):
        build()
    gxx = shutil.which("g++") or "/usr/bin/g++"
    srcs = [os.path.join(HOST, f) for f in ("store.cpp", "tokenizer.cpp", "embedder.cpp")]
    deps = srcs + [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith((".hpp", ".cpp"))] + [
        os.path.join(HERE, "..", "include", "memex_b200.h"), LIB]
    newest = max(os.path.getmtime(d) for d in deps)
    common = ["-std=c++17", "-O2", "-fPIC", "-Wall", "-Wextra", "-pthread"]
    link = ["-L" + OUT, "-lmemex_b200", "-Wl,-rpath,$ORIGIN"]
    if not os.path.exists(HOST_LIB) or os.path.getmtime(HOST_LIB) < newest:
        r = subprocess.run([gxx, *common, "-shared", "-o", HOST_LIB, *srcs, *link], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"host library build failed:\n{r.stdout}\n{r.stderr}")
    if not os.path.exists(HOST_TEST) or os.path.getmtime(HOST_TEST) < max(newest, os.path.getmtime(HOST_LIB)):
        r = subprocess.run([gxx, *common, "-o", HOST_TEST, os.path.join(HOST, "test_host.cpp"), "-L" + OUT, "-lmemex_host",
                            "-lmemex_b200", "-Wl,-rpath,$ORIGIN"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"host test build failed:\n{r.stdout}\n{r.stderr}")
    return HOST_LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
    print(build_host())
