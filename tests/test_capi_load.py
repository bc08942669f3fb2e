"""CPU: the C-ABI library builds, loads and exports every symbol include/*.h declares; without a
GPU every computing entry point fails loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from memex_b200 import capi, storage, embedding


def test_exports_every_declared_symbol():
    L = capi.lib()
    names = capi.declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} is declared in memex_b200.h but not exported"
    dbg = open(os.path.join(os.path.dirname(capi.HEADER), "memex_b200_debug.h")).read()
    dbg = re.sub(r"/\*.*?\*/", "", dbg, flags=re.S)
    for n in set(re.findall(r"\b(mx_debug_[a-z0-9_]+)\s*\(", dbg)):
        assert hasattr(L, n)


def test_abi_version_and_launch_counter():
    L = capi.lib()
    assert L.mx_abi_version() == 2   # 2: mx_model_ext gained the T5 fields (family, d_kv, rel_buckets, rel_max_distance)
    assert L.mx_launch_count() >= 0


def test_sass_is_sm100a_only_and_has_tcgen05_tma():
    """cuobjdump evidence that the tensor-core kernels are tcgen05 + TMA (B200_PROFILING.md table)."""
    import shutil
    import subprocess
    from memex_b200 import build
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
    sass = subprocess.run(["cuobjdump", "-sass", build.LIB], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass          # tcgen05.mma kind::f16
    assert "UTMALDG" in sass          # cp.async.bulk.tensor
    assert "LDTM" in sass             # tcgen05.ld
    # warp-level mma.sync appears in exactly one place: the softmax-bound flash attention kernel
    # (attention_mma.cu explains why); every GEMM-shaped kernel is tcgen05
    fn = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
        elif "HMMA." in line and "UTCHMMA" not in line:
            assert fn is not None and "attention_mma_kernel" in fn, fn


@pytest.mark.skipif(capi.lib().mx_device_count() > 0, reason="box has a GPU")
def test_no_cpu_fallback_store():
    cfg = capi.StoreCfg(dim=8, dtype=0, metric=0, device=0, capacity=0, id_offset=0, id_stride=1)
    h = C.c_void_p()
    rc = capi.lib().mx_store_create(C.byref(cfg), C.byref(h))
    assert rc == capi.ERR_CONNECTION and not h.value
    assert b"no CPU path" in capi.lib().mx_last_error(None)
    with pytest.raises(storage.VectorStoreError) as ei:
        storage.B200Store.new("/tmp/mx-nogpu", dim=8)
    assert ei.value.variant == "ConnectionError"


@pytest.mark.skipif(capi.lib().mx_device_count() > 0, reason="box has a GPU")
def test_no_cpu_fallback_embedder():
    import numpy as np
    arch = embedding.Architecture(1, 64, 2, 128, vocab=50, max_pos=16)
    with pytest.raises(embedding.SetupError) as ei:
        embedding.B200Encoder(arch, {"x": np.zeros(1, np.float32)}, precision="f32")
    assert "no CPU path" in str(ei.value)


def test_argument_validation_needs_no_device():
    L = capi.lib()
    h = C.c_void_p()
    assert L.mx_store_create(None, C.byref(h)) == capi.ERR_INVALID
    bad = capi.StoreCfg(dim=0, dtype=0, metric=0, device=0, capacity=0, id_offset=0, id_stride=1)
    assert L.mx_store_create(C.byref(bad), C.byref(h)) == capi.ERR_INVALID
    bad = capi.StoreCfg(dim=4, dtype=7, metric=0, device=0, capacity=0, id_offset=0, id_stride=1)
    assert L.mx_store_create(C.byref(bad), C.byref(h)) == capi.ERR_UNSUPPORTED
    assert L.mx_store_delete(None, 1) == capi.ERR_UNSUPPORTED      # local.rs:29-32, as a status
    assert L.mx_store_has_file(b"/nonexistent-dir") == 0


def test_factory_rejects_unknown_schemes():
    # storage/mod.rs:104-136: unknown scheme -> Unsupported(uri)
    for uri in ("qdrant://x", "hnsw://tmp", "nonsense"):
        with pytest.raises(storage.VectorStoreError) as ei:
            storage.get_vector_storage(uri, "c")
        assert ei.value.variant == "Unsupported"


def test_staging_copy_names_the_first_non_finite_row():
    """Host logic of the host-buffer search calls (no device): one pass copies the query block into the staging buffer and
    finds the first row with a NaN / +-inf at ANY position, for widths that do and do not fill whole SIMD vectors."""
    import numpy as np
    L = capi.lib()
    fn = L.mx_debug_copy_checking_finite
    fn.restype = C.c_int64
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]
    rng = np.random.default_rng(5)
    for rows, dim in ((1, 1), (3, 7), (64, 384), (5, 33), (2, 1000)):
        src = rng.standard_normal((rows, dim)).astype(np.float32)
        src[0, 0] = np.float32(3.4e38)          # the largest finite magnitudes and denormals are fine
        src[-1, -1] = np.float32(1e-45)
        dst = np.full_like(src, -7.0)
        assert fn(dst.ctypes.data, src.ctypes.data, rows, dim) == -1
        np.testing.assert_array_equal(dst, src)
        for bad in (np.nan, np.inf, -np.inf):
            for r, c in {(0, 0), (rows - 1, dim - 1), (rows // 2, dim // 2), (rows - 1, 0)}:
                x = src.copy()
                x[r, c] = bad
                if rows > 1 and r + 1 < rows:
                    x[rows - 1, 0] = np.nan    # a later bad row does not change the answer
                assert fn(dst.ctypes.data, x.ctypes.data, rows, dim) == r, (rows, dim, r, c, bad)
    assert fn(None, None, 0, 384) == -1
