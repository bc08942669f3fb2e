/*
 * memex_b200.h -- C ABI of the B200-native embedding + vector-search hot path of memex.
 *
 * This is the drop-in boundary: a Rust (cxx / extern "C") shim binds exactly these symbols
 * (INTEGRATION.md shows it) and keeps memex's own surfaces unchanged above them:
 *
 *   trait VectorStore           reference lib/libmemex/src/storage/mod.rs:54-66
 *   HnswStore (file store)      reference lib/libmemex/src/storage/local.rs:21-166
 *   SentenceEmbedder runner     reference lib/libmemex/src/llm/embedding.rs:94-135
 *                               (the one line `model.encode(&segments)`, :109)
 *
 * Conventions
 *   - plain pointers and sizes only; handles are opaque; all buffers are caller-owned.
 *   - every call returns an int32 status: MX_OK or a negative code that maps 1:1 onto the
 *     reference's VectorStoreError / EmbeddingError variants (storage/mod.rs:30-48,
 *     llm/embedding.rs:10-16).  Nothing aborts or unwinds across the boundary (the reference
 *     panics at local.rs:31 and :80-83; this ABI returns MX_ERR_UNSUPPORTED / MX_ERR_SEARCH).
 *   - mx_last_error(handle) gives the message for the last failing call on that handle
 *     (valid until the next call on it); mx_last_error(NULL) the last create/load failure
 *     on the calling thread.
 *   - a handle is not re-entrant (memex already serialises a store behind a tokio Mutex,
 *     storage/mod.rs:70-92, and owns the embedder on one thread, embedding.rs:84-91);
 *     distinct handles may be driven from different threads -- each owns a CUDA stream.
 *   - row ids are 1-based and contiguous in insertion order, as `next_id = len + 1`
 *     (local.rs:63); the uuid strings stay on the host side of the binding.
 *   - there is NO CPU fallback: every entry point that computes needs a CUDA device
 *     (sm_100a cubins only) and fails with MX_ERR_CONNECTION without one.
 */
#ifndef MEMEX_B200_H
#define MEMEX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MX_ABI_VERSION 2

/* status codes <-> reference error variants */
#define MX_OK 0
#define MX_ERR_CONNECTION (-1)  /* VectorStoreError::ConnectionError  (no device / CUDA failure)   */
#define MX_ERR_DELETE (-2)      /* VectorStoreError::DeleteError                                  */
#define MX_ERR_FILE_IO (-3)     /* VectorStoreError::FileIOError                                  */
#define MX_ERR_INSERTION (-4)   /* VectorStoreError::InsertionError (non-finite input, OOM, ...)   */
#define MX_ERR_SEARCH (-5)      /* VectorStoreError::SearchError                                  */
#define MX_ERR_SERDE (-6)       /* VectorStoreError::SerdeError    (bad file header)              */
#define MX_ERR_SAVE (-7)        /* VectorStoreError::SaveError                                    */
#define MX_ERR_UNSUPPORTED (-8) /* VectorStoreError::Unsupported   (also single-row delete)       */
#define MX_ERR_INVALID (-9)     /* bad argument: null pointer, dim mismatch, k == 0                */
#define MX_ERR_ENCODE (-10)     /* EmbeddingError::EncodingFailure                                */
#define MX_ERR_SETUP (-11)      /* EmbeddingError::SetupError      (bad config / missing weight)   */

#define MX_DTYPE_F32 0u
#define MX_DTYPE_F16 1u
#define MX_METRIC_COSINE 0u /* score = 1 - DistCosine, as local.rs:86 */
#define MX_METRIC_DOT 1u    /* score = dot product (north_star's dot-product scan) */

#define MX_MAX_K 256u

typedef struct mx_store mx_store;
typedef struct mx_embedder mx_embedder;

/* ------------------------------------------------------------------------------------------
 * vector store -- replaces HnswStore's use of hnsw_rs (local.rs:48,65,76,101,127-129,150-153)
 * ---------------------------------------------------------------------------------------- */
typedef struct mx_store_cfg {
    uint32_t dim;       /* vector dimension (384 for the MiniLM models, storage/mod.rs:126)  */
    uint32_t dtype;     /* MX_DTYPE_*: how rows are kept in HBM                               */
    uint32_t metric;    /* MX_METRIC_*                                                        */
    int32_t device;     /* CUDA ordinal                                                       */
    uint64_t capacity;  /* rows to preallocate; the store grows geometrically beyond it        */
    uint64_t id_offset; /* ids reported = id_offset + local_row * id_stride + 1 (row sharding) */
    uint64_t id_stride; /* 0 is read as 1                                                     */
} mx_store_cfg;

/* HnswStore::new (local.rs:95-108) */
int32_t mx_store_create(const mx_store_cfg *cfg, mx_store **out);
void mx_store_destroy(mx_store *s);

/* HnswStore::insert / bulk_insert (local.rs:55-69): appends n rows [n, dim] f32 row-major from
 * HOST memory; *first_id_out (may be NULL) receives the 1-based id of the first new row. */
int32_t mx_store_add(mx_store *s, const float *vecs, uint64_t n, uint64_t *first_id_out);
/* same, rows already in DEVICE memory (ingest straight from the embedder's output) */
int32_t mx_store_add_device(mx_store *s, const float *vecs_dev, uint64_t n, uint64_t *first_id_out);
/* streamed ingest WHILE another host thread searches the same store (worker/src/tasks.rs:15-59 next to
 * api handlers.rs:61-81): the rows are produced on `cuda_stream` (the embedder's), appended on it, and join the
 * searchable range (the committed-rows watermark) when the append has completed.  The store never grows on this path:
 * create it with the capacity it will reach. */
int32_t mx_store_add_device_stream(mx_store *s, const float *vecs_dev, uint64_t n, uint64_t *first_id_out,
                                   void *cuda_stream);
/* SM partitioning between a store's scan and an embedder's forward pass running concurrently on one GPU: both are
 * persistent kernels that otherwise fill every SM and could only take turns.  `sms` = the CTA budget of the scan grid
 * (the scan stays HBM-bound far below the full chip) / of the embedder's grids; 0 = the whole device. */
int32_t mx_store_set_sm_limit(mx_store *s, uint32_t sms);
/* ... and the partition itself: two disjoint sets of SMs of one GPU (CUDA green contexts) with one stream each.  EVERY
 * kernel launched into a partition's stream runs on its SMs only, so a scan on part 0 and a forward pass on part 1
 * overlap instead of taking turns.  sms_first is rounded up to the hardware's granularity (8 SMs on sm_100); part 1 gets
 * the rest.  MX_ERR_UNSUPPORTED where the driver has no green contexts (the caller then shares the GPU by grid budgets
 * alone).  Pass mx_sm_partition_stream(p, i) as the cuda_stream argument of the *_device calls. */
typedef struct mx_sm_partition mx_sm_partition;
int32_t mx_sm_partition_create(int32_t device, uint32_t sms_first, mx_sm_partition **out);
void mx_sm_partition_destroy(mx_sm_partition *p);
void *mx_sm_partition_stream(mx_sm_partition *p, uint32_t which);
uint32_t mx_sm_partition_sms(mx_sm_partition *p, uint32_t which);

/* HnswStore::search (local.rs:71-91) for nq queries at once.  HOST buffers.
 *   ids_out [nq, k], scores_out [nq, k] best first, counts_out [nq] = min(k, len).
 * Ranking is the EXACT (distance asc, id asc) order the reference's HNSW walk approximates;
 * scores are bit-identical to DistCosine + local.rs:86 evaluated on the stored rows. */
int32_t mx_store_search(mx_store *s, const float *queries, uint32_t nq, uint32_t k,
                        uint64_t *ids_out, float *scores_out, uint32_t *counts_out);
/* The same call split in two, so that TWO searches may be in flight on a store: submit stages the queries, enqueues the
 * copy and the kernels on the store's stream and returns a ticket at once; collect waits for that search and copies its
 * answer out.  While the device works on search i the host prepares search i + 1 -- a server with several requests
 * outstanding (memex's handlers are async, lib/api/src/endpoints/collections/handlers.rs:61-81) keeps the GPU busy instead
 * of alternating with it.  Tickets are collected in the order they were issued; a third submit before a collect is
 * MX_ERR_INVALID.  Same results, bit for bit, as mx_store_search.  Not thread-safe (as every other call on a handle): the
 * host's VectorStorage mutex (storage/mod.rs:68) covers each half. */
int32_t mx_store_search_submit(mx_store *s, const float *queries, uint32_t nq, uint32_t k, uint64_t *ticket_out);
int32_t mx_store_search_collect(mx_store *s, uint64_t ticket, uint64_t *ids_out, float *scores_out, uint32_t *counts_out);
/* same with DEVICE buffers, asynchronous on `cuda_stream` (a cudaStream_t; NULL = the store's
 * own stream).  dists_out (may be NULL) receives the raw sort keys (cosine distance, or -dot)
 * that mx_merge_topk_device consumes. */
int32_t mx_store_search_device(mx_store *s, const float *queries_dev, uint32_t nq, uint32_t k,
                               uint64_t *ids_dev, float *scores_dev, float *dists_dev,
                               uint32_t *counts_dev, void *cuda_stream);

/* merge of G per-shard results (after the all-gather): inputs [G, nq, k] / [G, nq], DEVICE
 * buffers, ordered by (key asc, id asc) -> [nq, k].  metric picks the key -> score map. */
int32_t mx_merge_topk_device(const uint64_t *ids_dev, const float *dists_dev,
                             const uint32_t *counts_dev, uint32_t n_shards, uint32_t nq, uint32_t k,
                             uint32_t metric, uint64_t *ids_out_dev, float *scores_out_dev,
                             uint32_t *counts_out_dev, int32_t device, void *cuda_stream);

/* The same two steps with ONE buffer per shard, so that a single all-gather moves a shard's whole
 * answer: blob = { ids u64 [nq,k] | keys f32 [nq,k] | counts u32 [nq] }, mx_topk_blob_bytes(nq,k)
 * bytes (16-byte multiple).  blobs_dev holds n_shards blobs blob_stride_bytes apart. */
uint64_t mx_topk_blob_bytes(uint32_t nq, uint32_t k);
int32_t mx_store_search_blob_device(mx_store *s, const float *queries_dev, uint32_t nq, uint32_t k,
                                    void *blob_dev, void *cuda_stream);
int32_t mx_merge_topk_blobs_device(const void *blobs_dev, uint64_t blob_stride_bytes, uint32_t n_shards,
                                   uint32_t nq, uint32_t k, uint32_t metric, uint64_t *ids_out_dev,
                                   float *scores_out_dev, uint32_t *counts_out_dev, int32_t device,
                                   void *cuda_stream);

/* Peer-memory form of the exchange step (instead of the all-gather): every rank owns an exchange buffer that all
 * peers can address (CUDA IPC / symmetric memory, set up by the host side): [2][world] blob slots of
 * blob_stride bytes, then `world` u32 flags (zero-initialised).  mx_exchange_push_device stores this rank's blob into
 * slot_offset_bytes of EVERY peer's buffer (peer_bases: HOST array of `world` device-visible base addresses, this
 * rank's own included) over NVLink and then release-stores `epoch` into that peer's flag [rank];
 * mx_merge_topk_blobs_wait_device waits (bounded) until its `world` flags have reached `epoch`, then merges.  The
 * host alternates the two slot sets and increments the epoch per batch. */
int32_t mx_exchange_push_device(const void *blob_dev, uint64_t blob_bytes, const uint64_t *peer_bases, uint32_t world,
                                uint32_t rank, uint64_t slot_offset_bytes, uint64_t flag_offset_bytes, uint32_t epoch,
                                int32_t device, void *cuda_stream);
int32_t mx_merge_topk_blobs_wait_device(const void *blobs_dev, uint64_t blob_stride_bytes, uint32_t n_shards,
                                        uint32_t nq, uint32_t k, uint32_t metric, uint64_t *ids_out_dev,
                                        float *scores_out_dev, uint32_t *counts_out_dev, const uint32_t *flags_dev,
                                        uint32_t epoch, int32_t device, void *cuda_stream);

/* ------------------------------------------------------------------------------------------
 * mx_shard_group -- the multi-GPU search of a row-sharded corpus, rendezvous included, with no
 * torch and no collective library: what lets `VectorStorage::search` (storage/mod.rs:85-92)
 * reach the 8 GPUs of a box from a C++ or Rust host.
 *
 * Every member owns one exchange buffer in its GPU's memory ([2][world] blob slots, [2] query
 * slots, world + 1 epoch flags) that all peers address over NVLink.  Rendezvous, either
 *   - one PROCESS per GPU: mx_shard_group_export -> 64-byte CUDA IPC handle; the host moves the
 *     `world` handles over its own transport (file, socket, MPI, torch.distributed);
 *     mx_shard_group_connect(handles [world][64], rank order) opens them; or
 *   - one process driving SEVERAL GPUs (memex's server is one process): create one member per
 *     device and call mx_shard_group_connect_local (peer access + raw pointers).
 * A search = the shard's own scan + rerank (mx_store_search_blob_device), one push of the blob into
 * every peer's slot (P2P stores + release-stored epoch flag), one merge that waits for its `world`
 * flags inside the kernel.  All members make the same sequence of search calls (SPMD).
 *   query_root < 0 : every member passes the same queries.
 *   query_root = r : member r passes them and pushes the block into every peer's query slot; the
 *                    others pass NULL and their first kernel waits for the root's flag -- the
 *                    host-buffer call needs no broadcast.
 * Stores of a group report GLOBAL ids (mx_store_cfg.id_offset / id_stride). */
typedef struct mx_shard_group mx_shard_group;
#define MX_IPC_HANDLE_BYTES 64
int32_t mx_shard_group_create(int32_t device, uint32_t world, uint32_t rank, uint32_t dim, uint32_t max_nq,
                              uint32_t max_k, mx_shard_group **out);
void mx_shard_group_destroy(mx_shard_group *g);
int32_t mx_shard_group_export(mx_shard_group *g, void *handle_out /* MX_IPC_HANDLE_BYTES */);
int32_t mx_shard_group_connect(mx_shard_group *g, const void *handles /* [world][MX_IPC_HANDLE_BYTES] */);
int32_t mx_shard_group_connect_local(mx_shard_group *const *groups, uint32_t world);
/* DEVICE buffers, asynchronous on cuda_stream (NULL = the member's own stream) */
int32_t mx_shard_group_search_device(mx_shard_group *g, mx_store *s, const float *queries_dev, int32_t query_root,
                                     uint32_t nq, uint32_t k, uint64_t *ids_dev, float *scores_dev,
                                     uint32_t *counts_dev, void *cuda_stream);
/* HOST buffers: H2D of the queries (where this member has them), the search, ONE D2H of the answer */
int32_t mx_shard_group_search(mx_shard_group *g, mx_store *s, const float *queries, int32_t query_root, uint32_t nq,
                              uint32_t k, uint64_t *ids_out, float *scores_out, uint32_t *counts_out);
/* the host-buffer call split in two (see mx_store_search_submit): two searches in flight per member, every member submits
 * and collects in the same order */
int32_t mx_shard_group_search_submit(mx_shard_group *g, mx_store *s, const float *queries, int32_t query_root, uint32_t nq,
                                     uint32_t k, uint64_t *ticket_out);
int32_t mx_shard_group_search_collect(mx_shard_group *g, uint64_t ticket, uint64_t *ids_out, float *scores_out,
                                      uint32_t *counts_out);
/* single-process form: `world` members / stores in rank order; HOST buffers; the answer is member 0's */
int32_t mx_shard_group_search_local(mx_shard_group *const *groups, mx_store *const *stores, uint32_t world,
                                    const float *queries, uint32_t nq, uint32_t k, uint64_t *ids_out,
                                    float *scores_out, uint32_t *counts_out);
int32_t mx_shard_group_info(const mx_shard_group *g, uint32_t *world, uint32_t *rank, uint32_t *epoch,
                            int32_t *connected);

int32_t mx_store_len(mx_store *s, uint64_t *n_out); /* hnsw.get_nb_point(), local.rs:238 */
int32_t mx_store_clear(mx_store *s);                /* delete_all's index reset, local.rs:48-50 */
int32_t mx_store_delete(mx_store *s, uint64_t id);  /* local.rs:29-32: always MX_ERR_UNSUPPORTED */
/* HnswStore::save / load (local.rs:115-165): one flat file `vectors.b200.bin` in `dir`
 * (the id map `vectors.meta.json` is written by the host side of the binding). */
int32_t mx_store_save(mx_store *s, const char *dir);
int32_t mx_store_load(const char *dir, int32_t device, mx_store **out);
int32_t mx_store_has_file(const char *dir); /* 1 / 0 */
int32_t mx_store_remove_file(const char *dir);
/* stored rows [first_row, first_row + n) widened to f32 into HOST memory (tests, migration) */
int32_t mx_store_get_rows(mx_store *s, uint64_t first_row, uint64_t n, float *out);
int32_t mx_store_sync(mx_store *s);
int32_t mx_store_info(mx_store *s, uint32_t *dim, uint32_t *dtype, uint32_t *metric,
                      uint64_t *capacity);
/* which scan kernel mx_store_search* will use for (nq, k): 0 = fp32 CUDA-core stream,
 * 1 = fp16 CUDA-core stream, 2 = fp16 tcgen05.  force >= 0 pins a path (tests / bench). */
int32_t mx_store_scan_path(mx_store *s, uint32_t nq, uint32_t k, int32_t force);

/* Superset certificate.  The scan kernels rank rows by an approximate score and the answer is
 * re-scored exactly; after every search the library checks, per query, that no row it turned away could have reached
 * the k-th exact score (approximation radius from the fp16 query rounding and the f32 accumulation), and answers the
 * queries that fail the check again with an exact f64 scan -- so results equal the reference's (distance, id) order on
 * ANY data (near-tie crowds, mass duplicates).  mx_store_verify_stats: lifetime counts of queries answered / sent to the
 * exact scan (synchronises the device).  mx_store_set_verify(0) turns the check off (A/B measurements only). */
int32_t mx_store_verify_stats(mx_store *s, uint64_t *queries_out, uint64_t *flagged_out);
int32_t mx_store_set_verify(mx_store *s, int32_t on);

/* device-side timing of the store's own kernels (CUDA events on the launching stream) */
int32_t mx_store_set_timing(mx_store *s, int32_t on);
int32_t mx_store_get_timing(mx_store *s, double *scan_ms_total, uint64_t *scan_launches,
                            double *other_ms_total, uint64_t *other_launches);

/* ------------------------------------------------------------------------------------------
 * sentence embedder -- replaces rust-bert's SentenceEmbeddingsModel::encode (embedding.rs:109)
 * ---------------------------------------------------------------------------------------- */
typedef struct mx_model_cfg {
    uint32_t layers, hidden, heads, ffn, vocab, max_pos, type_vocab;
    float ln_eps;
    uint32_t normalize;  /* sentence-transformers Normalize module present (all-MiniLM: 1) */
    uint32_t precision;  /* 0 = bf16 activations on tcgen05 (product path), 1 = fp32 CUDA-core
                            validation path (same kernels' math in fp32) */
    uint32_t max_tokens; /* workspace bound: max B*S per encode call */
} mx_model_cfg;

/* weights by HF BertModel name ("embeddings.word_embeddings.weight",
 * "encoder.layer.0.attention.self.query.weight", ...), f32, Linear weights [out, in] */
typedef struct mx_tensor {
    const char *name;
    const float *data;
    uint64_t numel;
} mx_tensor;

int32_t mx_embedder_create(const mx_model_cfg *cfg, const mx_tensor *weights, uint32_t n_weights,
                           int32_t device, mx_embedder **out);

/* The other encoder stacks of memex's EmbeddingsModelType (embedding.rs:24-55) that share BERT's
 * post-LayerNorm layer: RoBERTa (AllDistilrobertaV1 -- one of the three models segment_text
 * accepts, embedding.rs:156-161), DistilBERT + Dense (DistiluseBaseMultilingualCased) and ALBERT
 * (ParaphraseAlbertSmallV2).  Weights keep the BERT names above; the host side renames
 * DistilBERT / ALBERT checkpoints (memex_b200/embedding.py::canonical_weights).  Extra tensors:
 *   "dense.linear.weight" [dense_out, hidden], "dense.linear.bias" [dense_out]   (2_Dense module)
 *   "embeddings.projection.weight" [hidden, embed_dim], "embeddings.projection.bias" [hidden]
 * All-zero ext == mx_embedder_create. */
#define MX_ACT_IDENTITY 0u
#define MX_ACT_TANH 1u
#define MX_FFN_GELU_ERF 0u  /* BERT / RoBERTa / DistilBERT "gelu" */
#define MX_FFN_GELU_TANH 1u /* ALBERT "gelu_new" */
typedef struct mx_model_ext {
    uint32_t pos_offset;    /* row of the position table used by token 0 (RoBERTa: padding_idx + 1 = 2) */
    uint32_t no_token_type; /* 1: the stack has no token-type table (DistilBERT) */
    uint32_t dense_out;     /* > 0: sentence-transformers Dense module after pooling, this many outputs */
    uint32_t dense_act;     /* MX_ACT_* */
    uint32_t dense_bias;    /* 1: "dense.linear.bias" is present */
    uint32_t ffn_act;       /* MX_FFN_* */
    uint32_t embed_dim;     /* > 0 and != hidden: factorised embeddings (ALBERT), projected to hidden */
    uint32_t share_layers;  /* 1: every layer uses the weights of "encoder.layer.0." (ALBERT) */
    /* SentenceT5Base (embedding.rs:32,52): family = MX_FAMILY_T5 selects the T5 encoder stack -- pre-RMSNorm layers around an
     * f32 residual stream, unscaled attention with a bucketed relative position bias shared by all layers, gated-GELU
     * feed-forward, no biases.  Weights under the HF T5EncoderModel names: "shared.weight" [vocab, hidden],
     * "encoder.block.{i}.layer.0.SelfAttention.{q,k,v,o}.weight", "encoder.block.0.layer.0.SelfAttention.
     * relative_attention_bias.weight" [rel_buckets, heads], "encoder.block.{i}.layer.{0,1}.layer_norm.weight",
     * "encoder.block.{i}.layer.1.DenseReluDense.{wi_0,wi_1,wo}.weight", "encoder.final_layer_norm.weight" (+ the Dense
     * module's "dense.linear.weight").  heads * d_kv must equal hidden (T5-base: 12 x 64 = 768); cfg.max_pos bounds the
     * sequence length (the bias table is built for it), ffn_act is read as gated gelu_new.  The tensor-core path runs the
     * seven GEMMs of a layer on tcgen05; attention (score bias) and RMSNorm run on CUDA cores. */
    uint32_t family;           /* MX_FAMILY_* */
    uint32_t d_kv;             /* T5: per-head width */
    uint32_t rel_buckets;      /* T5: relative_attention_num_buckets (32) */
    uint32_t rel_max_distance; /* T5: relative_attention_max_distance (128) */
} mx_model_ext;
#define MX_FAMILY_BERT 0u /* BERT / RoBERTa / DistilBERT / ALBERT: post-LayerNorm layers */
#define MX_FAMILY_T5 1u
int32_t mx_embedder_create_ex(const mx_model_cfg *cfg, const mx_model_ext *ext, const mx_tensor *weights,
                              uint32_t n_weights, int32_t device, mx_embedder **out);
/* width of one output row: dense_out when a Dense module is present, else hidden */
int32_t mx_embedder_out_dim(mx_embedder *e, uint32_t *dim);
void mx_embedder_destroy(mx_embedder *e);
/* ids [B, S] int32 padded, lens [B] (tokens beyond lens[b] are ignored), out [B, out_dim] f32;
 * HOST buffers.  Padding costs nothing: the activations are stored without it (row offsets from
 * lens), so the GEMMs run on sum(lens) rows; the result is bit-identical to the padded layout. */
int32_t mx_embedder_encode(mx_embedder *e, const int32_t *ids, const int32_t *lens, uint32_t B,
                           uint32_t S, float *out);
/* ids / out in DEVICE memory, lens on the HOST; asynchronous on cuda_stream (NULL = own) */
int32_t mx_embedder_encode_device(mx_embedder *e, const int32_t *ids_dev, const int32_t *lens,
                                  uint32_t B, uint32_t S, float *out_dev, void *cuda_stream);
int32_t mx_embedder_sync(mx_embedder *e);
int32_t mx_embedder_set_sm_limit(mx_embedder *e, uint32_t sms);
int32_t mx_embedder_set_timing(mx_embedder *e, int32_t on);
int32_t mx_embedder_get_timing(mx_embedder *e, double *gemm_ms_total, uint64_t *gemm_launches,
                               double *other_ms_total, uint64_t *other_launches);

/* ------------------------------------------------------------------------------------------ */
const char *mx_last_error(const void *handle);
uint64_t mx_launch_count(void);     /* kernels this library has launched in this process */
int32_t mx_device_count(void);
int32_t mx_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MEMEX_B200_H */
