// scan.cuh -- interfaces between the store and its scan / rerank kernels.
#pragma once
#include "common.cuh"

namespace mx {

// One approximate-scan launch: every query in [0, nq) against rows [0, n_rows).
// Output: cand_s / cand_r [nq][n_lists][lcap] -- per (query, CTA) unordered top lists, padded
// with (-inf, kNoRow).
struct ScanParams {
    const void *rows;        // [capacity, ld] in the stored dtype
    const float *inv_norm;   // [capacity] 1 / |row|  (cosine), 0 for zero rows
    const float *queries;    // [nq, ldq] f32, zero padded to ldq
    float *cand_s;
    uint32_t *cand_r;
    uint32_t n_rows;
    uint32_t ld;             // elements per stored row (multiple of 16 bytes)
    uint32_t ldq;            // floats per staged query (>= ld, multiple of 8)
    uint32_t nq;
    uint32_t n_lists;        // lists per query (= CTAs of the scan grid)
    uint32_t use_inv;        // 1 = cosine (scale by inv_norm), 0 = dot
};

uint32_t scan_stream_lcap(uint32_t k);
uint32_t scan_stream_nq_per_pass(uint32_t nq, uint32_t k);
cudaError_t launch_scan_stream(const ScanParams &p, uint32_t dtype, uint32_t k, uint32_t n_ctas,
                               cudaStream_t st);

// merge of the candidate lists + exact re-scoring with the reference's arithmetic
struct RerankParams {
    const void *rows;
    const float *queries;     // [nq, ldq] f32 (the caller's values, unrounded)
    const float *cand_s;
    const uint32_t *cand_r;
    const uint32_t *zero_rows;  // first <= MX_MAX_K rows whose norm is zero (cosine only)
    const uint32_t *n_zero;     // device scalar
    uint64_t *ids_out;          // [nq, k]
    float *scores_out;          // [nq, k]
    float *dists_out;           // [nq, k] or nullptr
    uint32_t *counts_out;       // [nq]
    uint64_t id_offset, id_stride;
    uint32_t n_rows, ld, ldq, dim, nq, k;
    uint32_t n_lists, lcap;
    uint32_t dtype, metric;
    // ---- superset certificate ----
    // The scan ranks rows by an APPROXIMATE score; the answer is provably the exact one when every row the scan or the
    // rerank's own top-(32 E) cut rejected scored, approximately, more than the approximation radius below the k-th
    // exact score.  Queries that fail the test are appended to q_map / fb_thr and answered again by the exact scan.
    const float *qerr = nullptr;        // [nq] fp16 rounding radius of the prepared query (tcgen05 scan) or nullptr
    const float *scan_floor = nullptr;  // [nq] tau0 of the tcgen05 scan's threshold seeding or nullptr
    const float *max_norm = nullptr;    // device scalar: largest stored row norm (dot metric radius)
    uint32_t unit_queries = 0;          // 1: approximate scores are those of the unit-norm query (tcgen05 scan)
    uint32_t *n_flagged = nullptr;      // device scalar, zero on entry; nullptr = no certificate
    uint32_t *n_flagged_next = nullptr; // the counter of the NEXT search: zeroed here (nobody reads it during this one)
    unsigned long long *prof = nullptr; // test-only (MX_RERANK_PROF=1): %globaltimer of CTA 0 at the phase boundaries, [8]
    uint32_t *q_map = nullptr;          // [nq] flagged query indices
    float *fb_thr = nullptr;            // [nq] fp32-scan threshold below which a row cannot be in the flagged query's top k
    unsigned long long *stats = nullptr;  // [2] lifetime counters: queries answered, queries flagged
    // ---- second pass over the flagged queries only: CTA b answers query active_map[b] from candidate lists b ----
    const uint32_t *active_n = nullptr;
    const uint32_t *active_map = nullptr;
};
cudaError_t launch_rerank(const RerankParams &p, cudaStream_t st);

// exact fallback scan (scan_stream.cu): for each flagged query, every row whose fp32 score reaches fb_thr is re-scored
// with the reference's f64 arithmetic IN the scan, and the lists are kept on the exact key -- no approximation left.
// Candidate lists [flagged index][cta][lcap] with score field = -key.
struct ExactScanParams {
    ScanParams scan;              // queries = all queries of the batch; cand_* = the lists
    const uint32_t *active_n;     // device scalar: flagged queries
    const uint32_t *active_map;   // [nq]
    const float *fb_thr;          // [nq]
    uint32_t dim, metric;
};
cudaError_t launch_scan_exact(const ExactScanParams &p, uint32_t dtype, uint32_t k, uint32_t n_ctas, cudaStream_t st);

struct MergeParams {
    const uint64_t *ids;     // shard g: [nq, k] at (char *)ids + g * stride_ids
    const float *dists;      // shard g: [nq, k] at (char *)dists + g * stride_dists
    const uint32_t *counts;  // shard g: [nq]    at (char *)counts + g * stride_counts
    uint64_t stride_ids, stride_dists, stride_counts;  // bytes between consecutive shards
    uint64_t *ids_out;
    float *scores_out;
    uint32_t *counts_out;
    uint32_t n_shards, nq, k, metric;
    // peer-memory exchange: when set, shard g's blob is only read once wait_flags[g] has reached wait_epoch
    // (release-stored by the peer that pushed it, launch_exchange_push)
    const uint32_t *wait_flags = nullptr;
    uint32_t wait_epoch = 0;
};
cudaError_t launch_merge(const MergeParams &p, cudaStream_t st);

// this rank's blob -> slot of every peer's gathered buffer (P2P stores over NVLink), then flag[rank] = epoch on that peer
constexpr int kMaxPeers = 16;
struct PushParams {
    const void *blob;
    uint64_t blob_bytes;             // multiple of 16
    uint64_t peer_base[kMaxPeers];   // device-visible base address of every peer's exchange buffer (this rank's included)
    uint64_t slot_offset, flag_offset;
    uint32_t world, rank, epoch;
};
cudaError_t launch_exchange_push(const PushParams &p, cudaStream_t st);

// ingest: f32 rows -> stored dtype (+ inv_norm, zero-row list, non-finite flag)
struct IngestParams {
    const float *src;     // [n, dim]
    void *rows;           // stored matrix base
    float *inv_norm;
    uint32_t *zero_rows;
    uint32_t *n_zero;
    uint32_t *bad_flag;
    float *max_norm;      // device scalar: running maximum of the stored rows' norms
    uint64_t first_row, n;
    uint32_t dim, ld, dtype, metric;
};
cudaError_t launch_ingest(const IngestParams &p, cudaStream_t st);
cudaError_t launch_export_rows(const void *rows, uint32_t dtype, uint32_t ld, uint32_t dim,
                               uint64_t first_row, uint64_t n, float *out, cudaStream_t st);
cudaError_t launch_stage_queries(const float *q, uint32_t nq, uint32_t dim, uint32_t ldq, float *out,
                                 cudaStream_t st);

// K2: fp16 batched scan on tcgen05 (scan_tc.cu).  Same candidate-list contract as the stream scan.
struct TcScanState;
TcScanState *tc_scan_create(int sm_count, uint32_t ld, uint32_t dim);  // nullptr if shape unsupported
void tc_scan_destroy(TcScanState *t);
void tc_scan_invalidate(TcScanState *t);  // the row matrix moved: tensor maps must be rebuilt
void tc_scan_set_sms(TcScanState *t, int sm_count);   // the scan grid's CTA budget (SM partitioning with another kernel)
bool tc_scan_supports(const TcScanState *t, uint32_t k);
uint32_t tc_scan_max_k();
uint32_t tc_scan_lists(const TcScanState *t, uint64_t n_rows);
uint32_t tc_scan_lcap(uint32_t k);
const float *tc_scan_qerr(const TcScanState *t);    // valid after tc_scan_launch, [nq]
const float *tc_scan_floor(const TcScanState *t);
const unsigned long long *tc_scan_prof(const TcScanState *t);   // MX_SCAN_TC_PROF=1: [8] phase timestamps of CTA 0, else null
cudaError_t tc_scan_launch(TcScanState *t, const ScanParams &p, uint64_t capacity, uint32_t k, KernelTimer *timer,
                           cudaStream_t st, const char **why);

}  // namespace mx
