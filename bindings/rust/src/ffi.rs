//! `extern "C"` mirror of include/memex_b200.h (ABI version 1).  Plain pointers and sizes, opaque handles, i32 status.
#![allow(non_camel_case_types)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct mx_store {
    _p: [u8; 0],
}
#[repr(C)]
pub struct mx_embedder {
    _p: [u8; 0],
}
#[repr(C)]
pub struct mx_shard_group {
    _p: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct mx_store_cfg {
    pub dim: u32,
    pub dtype: u32,
    pub metric: u32,
    pub device: i32,
    pub capacity: u64,
    pub id_offset: u64,
    pub id_stride: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct mx_model_cfg {
    pub layers: u32,
    pub hidden: u32,
    pub heads: u32,
    pub ffn: u32,
    pub vocab: u32,
    pub max_pos: u32,
    pub type_vocab: u32,
    pub ln_eps: f32,
    pub normalize: u32,
    pub precision: u32,
    pub max_tokens: u32,
}

/// The non-BERT stacks of `EmbeddingsModelType` (llm/embedding.rs:24-55); all zero = BERT.
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct mx_model_ext {
    pub pos_offset: u32,
    pub no_token_type: u32,
    pub dense_out: u32,
    pub dense_act: u32,
    pub dense_bias: u32,
    pub ffn_act: u32,
    pub embed_dim: u32,
    pub share_layers: u32,
    pub family: u32,           // MX_FAMILY_*
    pub d_kv: u32,             // T5: per-head width
    pub rel_buckets: u32,      // T5: relative_attention_num_buckets
    pub rel_max_distance: u32, // T5: relative_attention_max_distance
}

#[repr(C)]
pub struct mx_tensor {
    pub name: *const c_char,
    pub data: *const f32,
    pub numel: u64,
}

pub const MX_OK: i32 = 0;
pub const MX_ERR_CONNECTION: i32 = -1;
pub const MX_ERR_DELETE: i32 = -2;
pub const MX_ERR_FILE_IO: i32 = -3;
pub const MX_ERR_INSERTION: i32 = -4;
pub const MX_ERR_SEARCH: i32 = -5;
pub const MX_ERR_SERDE: i32 = -6;
pub const MX_ERR_SAVE: i32 = -7;
pub const MX_ERR_UNSUPPORTED: i32 = -8;
pub const MX_ERR_INVALID: i32 = -9;
pub const MX_ERR_ENCODE: i32 = -10;
pub const MX_ERR_SETUP: i32 = -11;

pub const MX_DTYPE_F32: u32 = 0;
pub const MX_DTYPE_F16: u32 = 1;
pub const MX_METRIC_COSINE: u32 = 0;
pub const MX_METRIC_DOT: u32 = 1;
pub const MX_MAX_K: u32 = 256;
pub const MX_IPC_HANDLE_BYTES: usize = 64;
pub const MX_ACT_IDENTITY: u32 = 0;
pub const MX_ACT_TANH: u32 = 1;
pub const MX_FFN_GELU_ERF: u32 = 0;
pub const MX_FFN_GELU_TANH: u32 = 1;
pub const MX_FAMILY_BERT: u32 = 0;
pub const MX_FAMILY_T5: u32 = 1;

extern "C" {
    pub fn mx_store_create(cfg: *const mx_store_cfg, out: *mut *mut mx_store) -> i32;
    pub fn mx_store_destroy(s: *mut mx_store);
    pub fn mx_store_add(s: *mut mx_store, vecs: *const f32, n: u64, first_id_out: *mut u64) -> i32;
    pub fn mx_store_search(
        s: *mut mx_store, queries: *const f32, nq: u32, k: u32, ids_out: *mut u64, scores_out: *mut f32,
        counts_out: *mut u32,
    ) -> i32;
    // the same call in two halves: two searches may be in flight on a store (tickets are collected in issue order)
    pub fn mx_store_search_submit(s: *mut mx_store, queries: *const f32, nq: u32, k: u32, ticket_out: *mut u64) -> i32;
    pub fn mx_store_search_collect(
        s: *mut mx_store, ticket: u64, ids_out: *mut u64, scores_out: *mut f32, counts_out: *mut u32,
    ) -> i32;
    pub fn mx_store_len(s: *mut mx_store, n_out: *mut u64) -> i32;
    pub fn mx_store_info(s: *mut mx_store, dim: *mut u32, dtype: *mut u32, metric: *mut u32, capacity: *mut u64) -> i32;
    pub fn mx_store_clear(s: *mut mx_store) -> i32;
    pub fn mx_store_delete(s: *mut mx_store, id: u64) -> i32; // always MX_ERR_UNSUPPORTED (local.rs:29-32)
    pub fn mx_store_save(s: *mut mx_store, dir: *const c_char) -> i32;
    pub fn mx_store_load(dir: *const c_char, device: i32, out: *mut *mut mx_store) -> i32;
    pub fn mx_store_has_file(dir: *const c_char) -> i32;
    pub fn mx_store_remove_file(dir: *const c_char) -> i32;

    // mx_shard_group: the multi-GPU search below the C ABI (one member per GPU; rendezvous by CUDA IPC handles between
    // processes, or mx_shard_group_connect_local inside one process) -- what a sharded `impl VectorStore` calls
    pub fn mx_shard_group_create(
        device: i32, world: u32, rank: u32, dim: u32, max_nq: u32, max_k: u32, out: *mut *mut mx_shard_group,
    ) -> i32;
    pub fn mx_shard_group_destroy(g: *mut mx_shard_group);
    pub fn mx_shard_group_export(g: *mut mx_shard_group, handle_out: *mut u8) -> i32; // MX_IPC_HANDLE_BYTES
    pub fn mx_shard_group_connect(g: *mut mx_shard_group, handles: *const u8) -> i32; // [world][MX_IPC_HANDLE_BYTES]
    pub fn mx_shard_group_connect_local(groups: *const *mut mx_shard_group, world: u32) -> i32;
    pub fn mx_shard_group_search(
        g: *mut mx_shard_group, s: *mut mx_store, queries: *const f32, query_root: i32, nq: u32, k: u32,
        ids_out: *mut u64, scores_out: *mut f32, counts_out: *mut u32,
    ) -> i32;
    pub fn mx_shard_group_search_submit(
        g: *mut mx_shard_group, s: *mut mx_store, queries: *const f32, query_root: i32, nq: u32, k: u32,
        ticket_out: *mut u64,
    ) -> i32;
    pub fn mx_shard_group_search_collect(
        g: *mut mx_shard_group, ticket: u64, ids_out: *mut u64, scores_out: *mut f32, counts_out: *mut u32,
    ) -> i32;
    pub fn mx_shard_group_search_local(
        groups: *const *mut mx_shard_group, stores: *const *mut mx_store, world: u32, queries: *const f32, nq: u32,
        k: u32, ids_out: *mut u64, scores_out: *mut f32, counts_out: *mut u32,
    ) -> i32;

    pub fn mx_embedder_create(
        cfg: *const mx_model_cfg, w: *const mx_tensor, n: u32, device: i32, out: *mut *mut mx_embedder,
    ) -> i32;
    pub fn mx_embedder_create_ex(
        cfg: *const mx_model_cfg, ext: *const mx_model_ext, w: *const mx_tensor, n: u32, device: i32,
        out: *mut *mut mx_embedder,
    ) -> i32;
    pub fn mx_embedder_out_dim(e: *mut mx_embedder, dim: *mut u32) -> i32;
    pub fn mx_embedder_destroy(e: *mut mx_embedder);
    pub fn mx_embedder_encode(
        e: *mut mx_embedder, ids: *const i32, lens: *const i32, b: u32, s: u32, out: *mut f32,
    ) -> i32;

    pub fn mx_last_error(handle: *const c_void) -> *const c_char;
    pub fn mx_abi_version() -> i32;
}

/// Message of the last failing call on `handle` (null = last create / load failure on this thread).
pub fn last_error(handle: *const c_void) -> String {
    let p = unsafe { mx_last_error(handle) };
    if p.is_null() {
        String::new()
    } else {
        unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned()
    }
}
