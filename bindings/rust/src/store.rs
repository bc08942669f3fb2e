//! lib/libmemex/src/storage/b200.rs -- `B200Store: VectorStore`, the GPU twin of `HnswStore`
//! (storage/local.rs:21-166).  NEVER COMPILED in the build image (no rustc); the C++ `memex::B200Store`
//! (memex_b200/host/store.cpp) is its compiled twin and the ctypes `B200Store` (memex_b200/storage.py) its tested one.
//!
//! What stays exactly as in `HnswStore`: the 1-based row ids (`next_id = len + 1`, local.rs:63), the
//! `_id_map: HashMap<usize, String>` and its `vectors.meta.json` rendering (local.rs:155-161), `score = 1 - d`
//! (local.rs:86, computed on the device with the same f32 operations), `delete` unsupported (local.rs:29-32).
//! What changes: the rows live in HBM (`vectors.b200.bin` on disk instead of the hnsw_rs dumps), search is exact,
//! a batch is saved once instead of once per vector (local.rs:66-67), and nothing panics across the boundary.
use async_trait::async_trait;
use std::{
    collections::HashMap,
    ffi::CString,
    fs::File,
    io::{BufReader, Write},
    os::raw::c_void,
    path::{Path, PathBuf},
};

use super::{StoreResult, VectorData, VectorSearchResult, VectorStore, VectorStoreError};
use crate::b200::ffi;

const META_FILE: &str = "vectors.meta.json";
const DEFAULT_DIM: u32 = 384; // all-MiniLM; the first insert fixes the real width (see `ensure`)

pub struct B200Store {
    pub storage_path: PathBuf,
    handle: *mut ffi::mx_store, // created on the first insert, when the dimension is known
    half: bool,                 // fp16 rows (b200+f16://): the tensor-core scan serves batched queries
    device: i32,
    dim: usize, // width of the device store (0 until the first insert / load); every FFI call is checked against it
    pub _id_map: HashMap<usize, String>,
}

// The handle is not re-entrant; VectorStorage's tokio Mutex (storage/mod.rs:70-92) already serialises every call.
unsafe impl Send for B200Store {}
unsafe impl Sync for B200Store {}

fn cstr(p: &Path) -> CString {
    CString::new(p.display().to_string()).unwrap_or_default()
}

/// status code -> the reference's error variant (storage/mod.rs:30-48)
fn check(rc: i32, handle: *const c_void) -> StoreResult<()> {
    if rc == ffi::MX_OK {
        return Ok(());
    }
    let msg = ffi::last_error(handle);
    Err(match rc {
        ffi::MX_ERR_CONNECTION => VectorStoreError::ConnectionError(msg),
        ffi::MX_ERR_DELETE => VectorStoreError::DeleteError(msg),
        ffi::MX_ERR_FILE_IO => VectorStoreError::FileIOError(std::io::Error::new(std::io::ErrorKind::Other, msg)),
        ffi::MX_ERR_INSERTION => VectorStoreError::InsertionError(msg),
        ffi::MX_ERR_SAVE => VectorStoreError::SaveError(msg),
        ffi::MX_ERR_UNSUPPORTED => VectorStoreError::Unsupported(msg),
        // SerdeError wraps serde_json::Error, which cannot be built from a message: a bad row file is a load failure
        ffi::MX_ERR_SERDE => VectorStoreError::ConnectionError(msg),
        _ => VectorStoreError::SearchError(msg),
    })
}

impl B200Store {
    pub fn new(storage_path: &Path, half: bool) -> Self {
        log::info!("Initializing B200 vector storage @ \"{}\"", storage_path.display());
        Self { storage_path: storage_path.to_path_buf(), handle: std::ptr::null_mut(), half, device: 0, dim: 0, _id_map: HashMap::new() }
    }

    pub fn has_store(store_path: &Path) -> bool {
        store_path.join(META_FILE).exists() // as local.rs:115-118
    }

    pub fn load(store_path: &Path) -> Result<Self, VectorStoreError> {
        log::info!("Loading B200 vector storage @ \"{}\"", store_path.display());
        let dir = cstr(store_path);
        let mut handle: *mut ffi::mx_store = std::ptr::null_mut();
        if unsafe { ffi::mx_store_has_file(dir.as_ptr()) } == 1 {
            check(unsafe { ffi::mx_store_load(dir.as_ptr(), 0, &mut handle) }, std::ptr::null())?;
        }
        let meta_reader = BufReader::new(File::open(store_path.join(META_FILE))?);
        let _id_map: HashMap<usize, String> = serde_json::from_reader(meta_reader)?;
        // the C ABI's add / search take no width: read it back so that every slice can be checked before the call
        let (mut dim, mut dtype, mut metric, mut cap) = (0u32, 0u32, 0u32, 0u64);
        if !handle.is_null() {
            check(unsafe { ffi::mx_store_info(handle, &mut dim, &mut dtype, &mut metric, &mut cap) }, handle as *const c_void)?;
        }
        Ok(Self {
            storage_path: store_path.to_path_buf(),
            handle,
            half: dtype == ffi::MX_DTYPE_F16,
            device: 0,
            dim: dim as usize,
            _id_map,
        })
    }

    pub fn save(&self, store_path: PathBuf) -> Result<(), VectorStoreError> {
        if !store_path.exists() {
            let _ = std::fs::create_dir_all(store_path.clone());
        }
        if self.handle.is_null() {
            return Ok(()); // nothing was ever inserted: no files, so has_store() stays false
        }
        let dir = cstr(&store_path);
        check(unsafe { ffi::mx_store_save(self.handle, dir.as_ptr()) }, self.handle as *const c_void)?;
        // the id map, byte-compatible with local.rs:155-161
        let result = serde_json::to_string(&self._id_map).map_err(|err| VectorStoreError::SaveError(err.to_string()))?;
        let mut f = File::create(store_path.join(META_FILE))?;
        let _ = f.write(result.as_bytes())?;
        f.flush()?;
        Ok(())
    }

    fn ensure(&mut self, dim: usize) -> StoreResult<()> {
        if !self.handle.is_null() {
            // an existing (or loaded) store has ONE width: a narrower slice would make the device read past its end
            return if dim == self.dim {
                Ok(())
            } else {
                Err(VectorStoreError::InsertionError(format!("vector has dimension {dim}, store has {}", self.dim)))
            };
        }
        let cfg = ffi::mx_store_cfg {
            dim: if dim == 0 { DEFAULT_DIM } else { dim as u32 },
            dtype: if self.half { ffi::MX_DTYPE_F16 } else { ffi::MX_DTYPE_F32 },
            metric: ffi::MX_METRIC_COSINE,
            device: self.device,
            capacity: 0,
            id_offset: 0,
            id_stride: 1,
        };
        check(unsafe { ffi::mx_store_create(&cfg, &mut self.handle) }, std::ptr::null())?;
        self.dim = cfg.dim as usize;
        Ok(())
    }
}

impl Drop for B200Store {
    fn drop(&mut self) {
        if !self.handle.is_null() {
            unsafe { ffi::mx_store_destroy(self.handle) };
        }
    }
}

#[async_trait]
impl VectorStore for B200Store {
    async fn delete(&mut self, _: &str) -> StoreResult<()> {
        // local.rs:29-32 is `unimplemented!()`; here the same fact is an error value
        Err(VectorStoreError::Unsupported("single-point delete".into()))
    }

    async fn delete_all(&mut self) -> StoreResult<()> {
        // local.rs:34-53: remove the files, start from an empty index, clear the map
        let dir = cstr(&self.storage_path);
        unsafe { ffi::mx_store_remove_file(dir.as_ptr()) };
        let _ = std::fs::remove_file(self.storage_path.join(META_FILE));
        if !self.handle.is_null() {
            check(unsafe { ffi::mx_store_clear(self.handle) }, self.handle as *const c_void)?;
        }
        self._id_map.clear();
        Ok(())
    }

    async fn bulk_insert(&mut self, data: &[VectorData]) -> StoreResult<()> {
        // local.rs:55-60 loops over `insert` (and saves per vector); one batch, one save here
        if data.is_empty() {
            return Ok(());
        }
        let dim = data[0].vector.len();
        if data.iter().any(|d| d.vector.len() != dim) {
            return Err(VectorStoreError::InsertionError("vectors of different dimensions in one batch".into()));
        }
        self.ensure(dim)?;
        let mut flat = Vec::with_capacity(data.len() * dim);
        for d in data {
            flat.extend_from_slice(&d.vector);
        }
        let mut first: u64 = 0;
        check(
            unsafe { ffi::mx_store_add(self.handle, flat.as_ptr(), data.len() as u64, &mut first) },
            self.handle as *const c_void,
        )?;
        for (i, d) in data.iter().enumerate() {
            self._id_map.insert(first as usize + i, d._id.to_string()); // next_id = len + 1, local.rs:63-64
        }
        let _ = self.save(self.storage_path.clone());
        Ok(())
    }

    async fn insert(&mut self, data: &VectorData) -> StoreResult<()> {
        self.bulk_insert(std::slice::from_ref(data)).await
    }

    async fn search(&self, vec: &[f32], limit: usize) -> StoreResult<Vec<VectorSearchResult>> {
        // local.rs:71-91
        if limit == 0 || self.handle.is_null() {
            return Ok(Vec::new());
        }
        // one behaviour in every host (C++, Python, Rust): more than MX_MAX_K neighbours is an error, never a silent cut
        if limit > ffi::MX_MAX_K as usize {
            return Err(VectorStoreError::SearchError(format!("limit {limit} exceeds the store's maximum of {}", ffi::MX_MAX_K)));
        }
        if vec.len() != self.dim {
            return Err(VectorStoreError::SearchError(format!("query has dimension {}, store has {}", vec.len(), self.dim)));
        }
        let k = limit as u32;
        let (mut ids, mut scores, mut count) = (vec![0u64; k as usize], vec![0f32; k as usize], 0u32);
        check(
            unsafe {
                ffi::mx_store_search(self.handle, vec.as_ptr(), 1, k, ids.as_mut_ptr(), scores.as_mut_ptr(), &mut count)
            },
            self.handle as *const c_void,
        )?;
        (0..count as usize)
            .map(|j| {
                // local.rs:80-83 panics on an unmapped id; here it is a SearchError
                let doc = self
                    ._id_map
                    .get(&(ids[j] as usize))
                    .ok_or_else(|| VectorStoreError::SearchError("Id from vector store not mapped".into()))?;
                Ok((doc.to_string(), scores[j])) // score = 1 - 1/(1/d), computed on the device (local.rs:86)
            })
            .collect()
    }
}

// ---- storage/mod.rs:104-136: the branch `get_vector_storage` gains ------------------------------------------
//
//     } else if scheme == "b200" || scheme == "b200+f16" {
//         let storage: PathBuf = uri.split_once("://").map(|x| x.1).unwrap_or_default().into();
//         let storage = storage.join(collection);                      // collections are folders, as hnsw://
//         if !storage.exists() {
//             std::fs::create_dir_all(storage.clone())?;
//         }
//         let store = if B200Store::has_store(&storage) {
//             B200Store::load(&storage)?
//         } else {
//             B200Store::new(&storage, scheme == "b200+f16")
//         };
//         Arc::new(Mutex::new(store))
//     }
//
// and VECTOR_CONNECTION=b200:///var/lib/memex/vectors.  SURVEY F9: the reference calls get_vector_storage per task and
// per request (worker/lib.rs:190, handlers.rs:35,63), i.e. it would re-upload the whole matrix each time -- keep the
// VectorStorage values in a process-wide map keyed by (uri, collection), as memex::get_vector_storage does in the
// C++ host layer (memex_b200/host/store.cpp).
