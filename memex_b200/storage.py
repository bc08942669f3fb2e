"""Host-side mirror of memex's vector-store surface over the C ABI (Python face; the C++ face is
memex_b200/host/).

Mirrors reference lib/libmemex/src/storage/mod.rs:16-139 (VectorData, VectorStoreError,
trait VectorStore, VectorStorage, get_vector_storage) and storage/local.rs:21-166 (HnswStore ->
B200Store): same method names, argument meaning and error behaviour, so the tests in tests/
read like local.rs:175-242.  All arithmetic happens in libmemex_b200.so on the GPU; this file
only keeps the `_id_map` (usize -> uuid string) and the meta file, as the Rust side does.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import threading
from dataclasses import dataclass, field
from urllib.parse import urlparse

import numpy as np

from . import capi

META_FILE = "vectors.meta.json"      # local.rs:19 -- byte-compatible: {"<usize>": "<uuid>", ...}
DATA_FILE = "vectors.b200.bin"       # replaces vectors.hnsw.graph / vectors.hnsw.data (local.rs:17-18)


@dataclass
class VectorData:
    """storage/mod.rs:16-28"""
    _id: str
    document_id: str = ""
    text: str = ""
    vector: list = field(default_factory=list)
    segment_id: int = 0


class VectorStoreError(Exception):
    """storage/mod.rs:30-48; `.variant` carries the enum variant name."""
    variant = "VectorStoreError"

    def __init__(self, message: str = ""):
        super().__init__(f"{self.variant}: {message}")
        self.message = message


def _variant(name):
    return type(name, (VectorStoreError,), {"variant": name})


ConnectionError_ = _variant("ConnectionError")
DeleteError = _variant("DeleteError")
FileIOError = _variant("FileIOError")
InsertionError = _variant("InsertionError")
SearchError = _variant("SearchError")
SerdeError = _variant("SerdeError")
SaveError = _variant("SaveError")
Unsupported = _variant("Unsupported")

_BY_CODE = {
    capi.ERR_CONNECTION: ConnectionError_, capi.ERR_DELETE: DeleteError, capi.ERR_FILE_IO: FileIOError,
    capi.ERR_INSERTION: InsertionError, capi.ERR_SEARCH: SearchError, capi.ERR_SERDE: SerdeError,
    capi.ERR_SAVE: SaveError, capi.ERR_UNSUPPORTED: Unsupported, capi.ERR_INVALID: SearchError,
}


def _raise(code: int, handle=None, default=VectorStoreError):
    msg = capi.lib().mx_last_error(handle)
    raise _BY_CODE.get(code, default)(msg.decode(errors="replace") if msg else f"status {code}")


class VectorStore:
    """trait VectorStore, storage/mod.rs:54-66"""

    def delete(self, id: str) -> None: raise NotImplementedError
    def delete_all(self) -> None: raise NotImplementedError
    def bulk_insert(self, data) -> None: raise NotImplementedError
    def insert(self, data: VectorData) -> None: raise NotImplementedError
    def search(self, vec, limit: int): raise NotImplementedError


class B200Store(VectorStore):
    """Drop-in for HnswStore (local.rs:21-166): flat [N, d] matrix in HBM, exact cosine top-k.

    Differences a caller can observe, all deliberate:
      * `search` returns the exact (distance asc, id asc) ranking the HNSW walk approximates;
      * `delete` raises Unsupported instead of panicking (local.rs:29-32);
      * `bulk_insert` is one batched append and `insert` does not re-dump the whole index
        (local.rs:66-67) -- call `save()` (the factory-made VectorStorage does after each batch).
    """

    def __init__(self, storage_path, dim: int | None = None, dtype: str = "f32", metric: str = "cosine",
                 device: int = 0, capacity: int = 0, id_offset: int = 0, id_stride: int = 1, _handle=None):
        self.storage_path = str(storage_path)
        self._id_map: dict[int, str] = {}
        self._h = C.c_void_p(_handle) if _handle else None
        self._dim = dim
        self._dtype = dtype
        self._metric = metric
        self._device = device
        self._capacity = capacity
        self._id_offset, self._id_stride = id_offset, id_stride
        if self._h is None and dim is not None:
            self._create(dim)

    # --- HnswStore::new / has_store / load / save (local.rs:94-165) -------------------------
    @classmethod
    def new(cls, storage_path, **kw) -> "B200Store":
        return cls(storage_path, **kw)

    @staticmethod
    def has_store(store_path) -> bool:
        return os.path.exists(os.path.join(str(store_path), META_FILE))

    @classmethod
    def load(cls, store_path, device: int = 0) -> "B200Store":
        store_path = str(store_path)
        meta_path = os.path.join(store_path, META_FILE)
        L = capi.lib()
        h = C.c_void_p()
        if not L.mx_store_has_file(store_path.encode()):
            # a meta file without its matrix: only an EMPTY map is a store (one that never received a row, e.g.
            # left by an older build after add_vectors([])); anything else is the reference's load failure
            try:
                with open(meta_path) as f:
                    if json.load(f) == {}:
                        return cls(store_path, device=device)
            except (OSError, ValueError):
                pass
            raise FileIOError(f"{os.path.join(store_path, DATA_FILE)}: No such file or directory")
        rc = L.mx_store_load(store_path.encode(), device, C.byref(h))
        if rc != capi.OK:
            _raise(rc)
        try:
            with open(meta_path) as f:
                raw = json.load(f)
            id_map = {int(k): str(v) for k, v in raw.items()}
        except OSError as e:
            L.mx_store_destroy(h)
            raise FileIOError(str(e))
        except (ValueError, AttributeError) as e:
            L.mx_store_destroy(h)
            raise SerdeError(str(e))
        self = cls(store_path, _handle=h.value, device=device)
        dim, dt, me = C.c_uint32(), C.c_uint32(), C.c_uint32()
        L.mx_store_info(self._h, C.byref(dim), C.byref(dt), C.byref(me), None)
        self._dim, self._dtype = dim.value, ("f16" if dt.value == capi.DTYPE_F16 else "f32")
        self._metric = "dot" if me.value == capi.METRIC_DOT else "cosine"
        self._id_map = id_map
        return self

    def save(self, store_path=None) -> None:
        store_path = str(store_path) if store_path is not None else self.storage_path
        if self._h is None:
            return   # nothing was ever inserted: no files, so has_store() stays false (a fresh HnswStore directory)
        os.makedirs(store_path, exist_ok=True)
        rc = capi.lib().mx_store_save(self._h, store_path.encode())
        if rc != capi.OK:
            _raise(rc, self._h, SaveError)
        try:
            tmp = os.path.join(store_path, META_FILE + ".tmp")
            with open(tmp, "w") as f:
                json.dump({str(k): v for k, v in self._id_map.items()}, f, separators=(",", ":"))
            os.replace(tmp, os.path.join(store_path, META_FILE))
        except OSError as e:
            raise FileIOError(str(e))

    # --- trait VectorStore ------------------------------------------------------------------
    def delete(self, id: str) -> None:
        if self._h is None:
            raise Unsupported("removing a single point is not supported by the file store")
        _raise(capi.lib().mx_store_delete(self._h, 0), self._h)

    def delete_all(self) -> None:
        for name in (DATA_FILE, META_FILE):
            p = os.path.join(self.storage_path, name)
            if os.path.exists(p):
                try:
                    os.remove(p)
                except OSError:
                    pass
        if self._h is not None:
            rc = capi.lib().mx_store_clear(self._h)
            if rc != capi.OK:
                _raise(rc, self._h, DeleteError)
        self._id_map.clear()

    def bulk_insert(self, data) -> None:
        data = list(data)
        if not data:
            return
        vecs = np.asarray([d.vector for d in data], dtype=np.float32)
        if vecs.ndim != 2:
            raise InsertionError("vectors of one batch must share a dimension")
        first = self.add_matrix(vecs)
        # next_id = len + 1 (local.rs:63); the store's ids and the map stay in lockstep
        for i, d in enumerate(data):
            self._id_map[(first - self._id_offset - 1) // self._id_stride + 1 + i] = str(d._id)

    def insert(self, data: VectorData) -> None:
        self.bulk_insert([data])

    def search(self, vec, limit: int):
        """-> [(doc_id, score)] best first, at most `limit` (local.rs:71-91)"""
        if limit <= 0 or self._h is None or len(self._id_map) == 0:
            return []
        ids, scores, counts = self.search_matrix(np.asarray(vec, dtype=np.float32)[None, :], limit)
        out = []
        for j in range(int(counts[0])):
            local = (int(ids[0, j]) - self._id_offset - 1) // self._id_stride + 1
            doc_id = self._id_map.get(local)
            if doc_id is None:  # local.rs:80-83 panics here; the ABI surface reports it
                raise SearchError("Internal inconsistency. Id from vector store not mapped.")
            out.append((doc_id, float(scores[0, j])))
        return out

    # --- matrix-level entry points (what VectorStorage / the bench use) ----------------------
    def _create(self, dim: int):
        cfg = capi.StoreCfg(dim=dim, dtype=capi.DTYPE_F16 if self._dtype == "f16" else capi.DTYPE_F32,
                            metric=capi.METRIC_DOT if self._metric == "dot" else capi.METRIC_COSINE,
                            device=self._device, capacity=self._capacity, id_offset=self._id_offset,
                            id_stride=self._id_stride)
        h = C.c_void_p()
        rc = capi.lib().mx_store_create(C.byref(cfg), C.byref(h))
        if rc != capi.OK:
            _raise(rc, None, ConnectionError_)
        self._h, self._dim = h, dim

    def add_matrix(self, vecs: np.ndarray) -> int:
        """append [n, d] f32 rows; returns the id of the first one"""
        vecs = np.ascontiguousarray(vecs, dtype=np.float32)
        if self._h is None:
            self._create(vecs.shape[1])
        if vecs.shape[1] != self._dim:
            raise InsertionError(f"vector has dimension {vecs.shape[1]}, store has {self._dim}")
        first = C.c_uint64()
        rc = capi.lib().mx_store_add(self._h, vecs.ctypes.data, vecs.shape[0], C.byref(first))
        if rc != capi.OK:
            _raise(rc, self._h, InsertionError)
        return first.value

    def search_matrix(self, queries: np.ndarray, k: int):
        """[nq, d] f32 -> (ids [nq,k] u64, scores [nq,k] f32, counts [nq] u32)"""
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        if queries.ndim != 2 or queries.shape[1] != self._dim:
            raise SearchError(f"query has dimension {queries.shape[-1]}, store has {self._dim}")
        if k > capi.MAX_K:   # one behaviour in every host (C++, Python, Rust): an error, never a silent cut
            raise SearchError(f"limit {k} exceeds the store's maximum of {capi.MAX_K} neighbours per query")
        nq = queries.shape[0]
        ids = np.zeros((nq, k), dtype=np.uint64)
        scores = np.zeros((nq, k), dtype=np.float32)
        counts = np.zeros(nq, dtype=np.uint32)
        rc = capi.lib().mx_store_search(self._h, queries.ctypes.data, nq, k, ids.ctypes.data,
                                        scores.ctypes.data, counts.ctypes.data)
        if rc != capi.OK:
            _raise(rc, self._h, SearchError)
        return ids, scores, counts

    def search_submit(self, queries: np.ndarray, k: int):
        """First half of search_matrix (mx_store_search_submit): stages and enqueues the search, returns a ticket at once.
        Two searches may be in flight; collect them in the order they were submitted."""
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        if queries.ndim != 2 or queries.shape[1] != self._dim:
            raise SearchError(f"query has dimension {queries.shape[-1]}, store has {self._dim}")
        if k > capi.MAX_K:
            raise SearchError(f"limit {k} exceeds the store's maximum of {capi.MAX_K} neighbours per query")
        ticket = C.c_uint64()
        rc = capi.lib().mx_store_search_submit(self._h, queries.ctypes.data, queries.shape[0], k, C.byref(ticket))
        if rc != capi.OK:
            _raise(rc, self._h, SearchError)
        return (ticket.value, queries.shape[0], k)

    def search_collect(self, ticket):
        """Second half: waits for the search behind `ticket` -> (ids [nq,k] u64, scores [nq,k] f32, counts [nq] u32)"""
        t, nq, k = ticket
        ids = np.zeros((nq, k), dtype=np.uint64)
        scores = np.zeros((nq, k), dtype=np.float32)
        counts = np.zeros(nq, dtype=np.uint32)
        rc = capi.lib().mx_store_search_collect(self._h, t, ids.ctypes.data, scores.ctypes.data, counts.ctypes.data)
        if rc != capi.OK:
            _raise(rc, self._h, SearchError)
        return ids, scores, counts

    def get_nb_point(self) -> int:
        """hnsw.get_nb_point() (local.rs:238)"""
        if self._h is None:
            return 0
        n = C.c_uint64()
        capi.lib().mx_store_len(self._h, C.byref(n))
        return n.value

    def __len__(self):
        return self.get_nb_point()

    @property
    def handle(self):
        return self._h

    @property
    def dim(self):
        return self._dim

    def close(self):
        if getattr(self, "_h", None) is not None:
            capi.lib().mx_store_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class VectorStorage:
    """storage/mod.rs:68-93: every call takes the one lock (tokio Mutex there, a thread lock here)."""

    def __init__(self, client: VectorStore, autosave: bool = True):
        self.client = client
        self._lock = threading.Lock()
        self._autosave = autosave

    def add_vectors(self, points) -> None:
        with self._lock:
            self.client.bulk_insert(points)
            if self._autosave and hasattr(self.client, "save"):
                try:
                    self.client.save()   # once per batch, not once per vector (local.rs:66-67)
                except VectorStoreError:
                    pass                 # the reference ignores save errors too (`let _ =`)

    def delete_collection(self) -> None:
        with self._lock:
            self.client.delete_all()

    def search(self, query, limit: int):
        with self._lock:
            return self.client.search(query, limit)


def get_vector_storage(uri: str, collection: str, device: int = 0) -> VectorStorage:
    """storage/mod.rs:95-139 with the new `b200://<dir>` scheme (`b200+f16://` keeps rows in fp16).

    Collections are folders; an existing vectors.meta.json means load, else new (mod.rs:107-121).
    """
    try:
        scheme = urlparse(uri).scheme
    except ValueError:
        raise Unsupported(uri)
    if scheme not in ("b200", "b200+f16", "b200+f32"):
        raise Unsupported(uri)
    storage = os.path.join(uri.split("://", 1)[1], collection)
    try:
        os.makedirs(storage, exist_ok=True)
    except OSError as e:
        raise FileIOError(str(e))
    if B200Store.has_store(storage):
        store = B200Store.load(storage, device=device)
    else:
        store = B200Store.new(storage, dtype="f16" if scheme == "b200+f16" else "f32", device=device)
    return VectorStorage(store)
