"""ctypes binding of the C ABI declared in include/memex_b200.h.

This is the same surface the Rust shim binds; the parity tests and bench.py go
through it.  There is no fallback: if the shared library is missing it is built with nvcc, and
if that fails the import raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.normpath(os.path.join(_HERE, "..", "include", "memex_b200.h"))

OK = 0
ERR_CONNECTION, ERR_DELETE, ERR_FILE_IO, ERR_INSERTION, ERR_SEARCH = -1, -2, -3, -4, -5
ERR_SERDE, ERR_SAVE, ERR_UNSUPPORTED, ERR_INVALID, ERR_ENCODE, ERR_SETUP = -6, -7, -8, -9, -10, -11
DTYPE_F32, DTYPE_F16 = 0, 1
METRIC_COSINE, METRIC_DOT = 0, 1
MAX_K = 256
IPC_HANDLE_BYTES = 64

ERROR_NAMES = {
    ERR_CONNECTION: "ConnectionError", ERR_DELETE: "DeleteError", ERR_FILE_IO: "FileIOError",
    ERR_INSERTION: "InsertionError", ERR_SEARCH: "SearchError", ERR_SERDE: "SerdeError",
    ERR_SAVE: "SaveError", ERR_UNSUPPORTED: "Unsupported", ERR_INVALID: "InvalidArgument",
    ERR_ENCODE: "EncodingFailure", ERR_SETUP: "SetupError",
}


class StoreCfg(C.Structure):
    _fields_ = [("dim", C.c_uint32), ("dtype", C.c_uint32), ("metric", C.c_uint32),
                ("device", C.c_int32), ("capacity", C.c_uint64), ("id_offset", C.c_uint64),
                ("id_stride", C.c_uint64)]


class ModelCfg(C.Structure):
    _fields_ = [("layers", C.c_uint32), ("hidden", C.c_uint32), ("heads", C.c_uint32),
                ("ffn", C.c_uint32), ("vocab", C.c_uint32), ("max_pos", C.c_uint32),
                ("type_vocab", C.c_uint32), ("ln_eps", C.c_float), ("normalize", C.c_uint32),
                ("precision", C.c_uint32), ("max_tokens", C.c_uint32)]


class ModelExt(C.Structure):
    _fields_ = [("pos_offset", C.c_uint32), ("no_token_type", C.c_uint32), ("dense_out", C.c_uint32),
                ("dense_act", C.c_uint32), ("dense_bias", C.c_uint32), ("ffn_act", C.c_uint32),
                ("embed_dim", C.c_uint32), ("share_layers", C.c_uint32),
                ("family", C.c_uint32), ("d_kv", C.c_uint32), ("rel_buckets", C.c_uint32), ("rel_max_distance", C.c_uint32)]


FAMILY_BERT, FAMILY_T5 = 0, 1


class Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.POINTER(C.c_float)), ("numel", C.c_uint64)]


def declared_symbols() -> list[str]:
    """Every function name include/memex_b200.h declares."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mx_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path):
        _build.build()
    # measurement aid: MX_B200_LIB=<another build of the library> (same-box A/B of two builds, scripts/ab_builds.sh);
    # symbols that build lacks are skipped, every other use loads the in-tree build and nothing else
    override = os.environ.get("MX_B200_LIB")
    if override:
        path = override
    L = C.CDLL(path)
    vp, f32p, u64p, u32p, i32p = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint64), \
        C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    f64p = C.POINTER(C.c_double)

    def sig(name, res, *args):
        if override and not hasattr(L, name):
            return
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("mx_store_create", C.c_int32, C.POINTER(StoreCfg), C.POINTER(vp))
    sig("mx_store_destroy", None, vp)
    sig("mx_store_add", C.c_int32, vp, vp, C.c_uint64, u64p)
    sig("mx_store_add_device", C.c_int32, vp, vp, C.c_uint64, u64p)
    sig("mx_store_add_device_stream", C.c_int32, vp, vp, C.c_uint64, u64p, vp)
    sig("mx_store_set_sm_limit", C.c_int32, vp, C.c_uint32)
    sig("mx_embedder_set_sm_limit", C.c_int32, vp, C.c_uint32)
    sig("mx_sm_partition_create", C.c_int32, C.c_int32, C.c_uint32, C.POINTER(vp))
    sig("mx_sm_partition_destroy", None, vp)
    sig("mx_sm_partition_stream", vp, vp, C.c_uint32)
    sig("mx_sm_partition_sms", C.c_uint32, vp, C.c_uint32)
    sig("mx_store_search", C.c_int32, vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp)
    sig("mx_store_search_submit", C.c_int32, vp, vp, C.c_uint32, C.c_uint32, u64p)
    sig("mx_store_search_collect", C.c_int32, vp, C.c_uint64, vp, vp, vp)
    sig("mx_store_search_device", C.c_int32, vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp)
    sig("mx_merge_topk_device", C.c_int32, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32,
        C.c_uint32, vp, vp, vp, C.c_int32, vp)
    sig("mx_topk_blob_bytes", C.c_uint64, C.c_uint32, C.c_uint32)
    sig("mx_store_search_blob_device", C.c_int32, vp, vp, C.c_uint32, C.c_uint32, vp, vp)
    sig("mx_merge_topk_blobs_device", C.c_int32, vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
        C.c_uint32, vp, vp, vp, C.c_int32, vp)
    sig("mx_exchange_push_device", C.c_int32, vp, C.c_uint64, u64p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32,
        C.c_int32, vp)
    sig("mx_merge_topk_blobs_wait_device", C.c_int32, vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, vp,
        vp, C.c_uint32, C.c_int32, vp)
    sig("mx_shard_group_create", C.c_int32, C.c_int32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp))
    sig("mx_shard_group_destroy", None, vp)
    sig("mx_shard_group_export", C.c_int32, vp, vp)
    sig("mx_shard_group_connect", C.c_int32, vp, vp)
    sig("mx_shard_group_connect_local", C.c_int32, C.POINTER(vp), C.c_uint32)
    sig("mx_shard_group_search_device", C.c_int32, vp, vp, vp, C.c_int32, C.c_uint32, C.c_uint32, vp, vp, vp, vp)
    sig("mx_shard_group_search", C.c_int32, vp, vp, vp, C.c_int32, C.c_uint32, C.c_uint32, vp, vp, vp)
    sig("mx_shard_group_search_submit", C.c_int32, vp, vp, vp, C.c_int32, C.c_uint32, C.c_uint32, u64p)
    sig("mx_shard_group_search_collect", C.c_int32, vp, C.c_uint64, vp, vp, vp)
    sig("mx_shard_group_search_local", C.c_int32, C.POINTER(vp), C.POINTER(vp), C.c_uint32, vp, C.c_uint32, C.c_uint32, vp, vp, vp)
    sig("mx_shard_group_info", C.c_int32, vp, u32p, u32p, u32p, i32p)
    sig("mx_store_len", C.c_int32, vp, u64p)
    sig("mx_store_clear", C.c_int32, vp)
    sig("mx_store_delete", C.c_int32, vp, C.c_uint64)
    sig("mx_store_save", C.c_int32, vp, C.c_char_p)
    sig("mx_store_load", C.c_int32, C.c_char_p, C.c_int32, C.POINTER(vp))
    sig("mx_store_has_file", C.c_int32, C.c_char_p)
    sig("mx_store_remove_file", C.c_int32, C.c_char_p)
    sig("mx_store_get_rows", C.c_int32, vp, C.c_uint64, C.c_uint64, vp)
    sig("mx_store_sync", C.c_int32, vp)
    sig("mx_store_info", C.c_int32, vp, u32p, u32p, u32p, u64p)
    sig("mx_store_scan_path", C.c_int32, vp, C.c_uint32, C.c_uint32, C.c_int32)
    sig("mx_store_verify_stats", C.c_int32, vp, u64p, u64p)
    sig("mx_store_set_verify", C.c_int32, vp, C.c_int32)
    sig("mx_store_set_timing", C.c_int32, vp, C.c_int32)
    sig("mx_store_get_timing", C.c_int32, vp, f64p, u64p, f64p, u64p)
    sig("mx_embedder_create", C.c_int32, C.POINTER(ModelCfg), C.POINTER(Tensor), C.c_uint32,
        C.c_int32, C.POINTER(vp))
    sig("mx_embedder_create_ex", C.c_int32, C.POINTER(ModelCfg), C.POINTER(ModelExt), C.POINTER(Tensor), C.c_uint32,
        C.c_int32, C.POINTER(vp))
    sig("mx_embedder_out_dim", C.c_int32, vp, u32p)
    sig("mx_embedder_destroy", None, vp)
    sig("mx_embedder_encode", C.c_int32, vp, vp, vp, C.c_uint32, C.c_uint32, vp)
    sig("mx_embedder_encode_device", C.c_int32, vp, vp, vp, C.c_uint32, C.c_uint32, vp, vp)
    sig("mx_embedder_sync", C.c_int32, vp)
    sig("mx_embedder_set_timing", C.c_int32, vp, C.c_int32)
    sig("mx_embedder_get_timing", C.c_int32, vp, f64p, u64p, f64p, u64p)
    sig("mx_last_error", C.c_char_p, vp)
    sig("mx_launch_count", C.c_uint64)
    sig("mx_device_count", C.c_int32)
    sig("mx_abi_version", C.c_int32)
    _ = i32p
    _lib = L
    return L


class MxError(RuntimeError):
    """A non-zero status from the C ABI; `.kind` is the reference's error variant name."""

    def __init__(self, code: int, message: str):
        self.code = code
        self.kind = ERROR_NAMES.get(code, f"status {code}")
        super().__init__(f"{self.kind}: {message}")


def check(code: int, handle=None):
    if code != OK:
        msg = lib().mx_last_error(handle)
        raise MxError(code, msg.decode(errors="replace") if msg else "")
