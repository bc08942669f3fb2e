// store.cu -- mx_store: the device-resident flat vector store behind memex's VectorStore surface.
//
// Host-side bookkeeping mirrors HnswStore (reference lib/libmemex/src/storage/local.rs:21-166):
// rows are appended in insertion order and identified by 1-based ids (local.rs:63); search returns
// (id, score) best first (local.rs:71-91); delete of one row is unsupported (local.rs:29-32);
// clear resets the index (local.rs:48-50); save/load persist one flat file next to the host
// side's vectors.meta.json (local.rs:115-165).
//
// HBM layout: rows [capacity, ld] in the stored dtype (ld = dim rounded up to 16 bytes, zero
// padded), inv_norm [capacity] f32.  Search = approximate scan (scan_stream.cu or scan_tc.cu)
// -> candidate merge + exact re-score (rerank.cu).  There is no CPU path.
#include <cerrno>
#include <cmath>
#include <cstring>
#include <mutex>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/memex_b200_debug.h"
#include "common.cuh"
#include "scan.cuh"

namespace mx {

std::atomic<uint64_t> g_launches{0};
thread_local std::string g_last_error;

cudaError_t set_max_smem_impl(const void *func, int bytes)
{
    struct Entry {
        const void *func;
        int device, bytes;
    };
    static std::mutex mu;
    static std::vector<Entry> seen;   // a few dozen kernels x devices: a linear scan beats a map
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    for (Entry &en : seen)
        if (en.func == func && en.device == device) {
            if (en.bytes >= bytes) return cudaSuccess;
            e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            if (e == cudaSuccess) en.bytes = bytes;
            return e;
        }
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) seen.push_back({func, device, bytes});
    return e;
}

cudaError_t launch_collect_zero_rows(const float *inv_norm, uint64_t n, uint32_t *zero_rows, uint32_t *n_zero,
                                     cudaStream_t st);

static const char kFileName[] = "vectors.b200.bin";
static const char kFileMagic[8] = {'M', 'X', 'B', '2', '0', '0', 'V', '1'};

struct FileHeader {
    char magic[8];
    uint32_t dim, dtype, metric, ld;
    uint64_t n;
    uint64_t reserved[4];
};

}  // namespace mx

using namespace mx;

struct mx_store : HandleBase {
    mx_store_cfg cfg{};
    cudaStream_t stream = nullptr;
    int sm_count = kNumSMsDefault;
    uint32_t ld = 0;      // elements per stored row
    uint32_t ldq = 0;     // floats per staged query
    size_t elem = 4;
    uint64_t n = 0, capacity = 0;
    uint64_t n_view = 0;   // the committed-rows watermark ONE search works with (another host thread may append meanwhile)
    void *rows = nullptr;
    float *inv_norm = nullptr;
    uint32_t *zero_rows = nullptr;  // [MX_MAX_K]
    uint32_t *n_zero = nullptr;     // device scalar
    uint32_t *flags = nullptr;      // device scalar: bit0 non-finite, bit1 zero row seen
    uint32_t *flags_host = nullptr; // PINNED copy target: a device-to-host copy into pageable memory holds a runtime lock until
                                    // the stream has drained, which stalls the kernel launches of every other host thread
    bool zero_dirty = false;
    int32_t force_path = -1;
    // scratch
    float *q_stage = nullptr;
    size_t q_stage_cap = 0;  // floats
    float *cand_s = nullptr;
    uint32_t *cand_r = nullptr;
    size_t cand_cap = 0;  // entries
    void *pinned = nullptr;
    size_t pinned_cap = 0;
    void *dev_io = nullptr;
    size_t dev_io_cap = 0;
    TcScanState *tc = nullptr;
    float *blob_scores = nullptr;  // scratch of mx_store_search_blob_device
    size_t blob_scores_cap = 0;
    // superset certificate + exact fallback
    bool verify = true;
    uint32_t *n_flagged_buf = nullptr;   // device [2]: search i counts into [i & 1]; its rerank zeroes the other one for search i + 1
    uint32_t flag_cur = 0;
    uint32_t *n_flagged = nullptr;       // = n_flagged_buf + flag_cur during a search
    uint32_t *q_map = nullptr;       // [verify_cap]
    float *fb_thr = nullptr;         // [verify_cap]
    size_t verify_cap = 0;
    unsigned long long *stats = nullptr;   // device [2]: queries answered, queries flagged
    float *max_norm = nullptr;       // device scalar
    unsigned long long *prof = nullptr;   // test-only: rerank phase timestamps (MX_RERANK_PROF=1)
    KernelTimer timer;
    IoSlot slots[2];                 // mx_store_search_submit / _collect: two host-buffer searches in flight
    uint64_t submit_seq = 0;
};

namespace {

int32_t set_device(mx_store *s) { MX_CUDA(s, MX_ERR_CONNECTION, cudaSetDevice(s->cfg.device)); return MX_OK; }

// test-only (MX_HOST_PROF=1): wall-clock phases of the host-buffer search call, read by mx_debug_host_prof
struct HostProf {
    static constexpr int kSlots = 12;
    double acc[kSlots] = {};
    uint64_t calls = 0;
    timespec last{};
    bool on = getenv("MX_HOST_PROF") != nullptr;
    void start()
    {
        if (on) clock_gettime(CLOCK_MONOTONIC, &last);
    }
    void mark(int slot)
    {
        if (!on) return;
        timespec t;
        clock_gettime(CLOCK_MONOTONIC, &t);
        acc[slot] += (double)(t.tv_sec - last.tv_sec) * 1e6 + (double)(t.tv_nsec - last.tv_nsec) * 1e-3;
        last = t;
    }
};
HostProf g_host_prof;

int32_t ensure_pinned(mx_store *s, size_t bytes)
{
    if (bytes <= s->pinned_cap) return MX_OK;
    if (s->pinned) cudaFreeHost(s->pinned);
    s->pinned = nullptr;
    s->pinned_cap = 0;
    MX_CUDA(s, MX_ERR_CONNECTION, cudaMallocHost(&s->pinned, bytes));
    s->pinned_cap = bytes;
    return MX_OK;
}

int32_t ensure_dev_io(mx_store *s, size_t bytes)
{
    if (bytes <= s->dev_io_cap) return MX_OK;
    if (s->dev_io) cudaFree(s->dev_io);
    s->dev_io = nullptr;
    s->dev_io_cap = 0;
    MX_CUDA(s, MX_ERR_CONNECTION, cudaMalloc(&s->dev_io, bytes));
    s->dev_io_cap = bytes;
    return MX_OK;
}

int32_t reserve(mx_store *s, uint64_t want, int32_t err_code)
{
    if (want <= s->capacity) return MX_OK;
    uint64_t cap = s->capacity ? s->capacity : 1024;
    while (cap < want) cap = cap + cap / 2 + 1024;
    if (cap > 0xfffffff0ull) {
        if (want > 0xfffffff0ull) return fail(s, err_code, "store is limited to 2^32 - 16 rows per device");
        cap = 0xfffffff0ull;
    }
    void *nrows = nullptr;
    float *ninv = nullptr;
    MX_CUDA(s, err_code, cudaMalloc(&nrows, cap * s->ld * s->elem));
    cudaError_t e = cudaMalloc(&ninv, (cap + 128) * sizeof(float));  // scan_tc reads whole 128-row tiles
    if (e != cudaSuccess) {
        cudaFree(nrows);
        return fail(s, err_code, "cudaMalloc(inv_norm) failed: %s", cudaGetErrorString(e));
    }
    if (s->n) {
        cudaMemcpyAsync(nrows, s->rows, s->n * s->ld * s->elem, cudaMemcpyDeviceToDevice, s->stream);
        cudaMemcpyAsync(ninv, s->inv_norm, s->n * sizeof(float), cudaMemcpyDeviceToDevice, s->stream);
    }
    MX_CUDA(s, err_code, cudaStreamSynchronize(s->stream));
    cudaFree(s->rows);
    cudaFree(s->inv_norm);
    s->rows = nrows;
    s->inv_norm = ninv;
    s->capacity = cap;
    if (s->tc) tc_scan_invalidate(s->tc);
    return MX_OK;
}

// rows already on the device as f32 [n, dim]
int32_t ingest_device(mx_store *s, const float *src_dev, uint64_t n, cudaStream_t st)
{
    IngestParams ip{};
    ip.src = src_dev;
    ip.rows = s->rows;
    ip.inv_norm = s->inv_norm;
    ip.zero_rows = s->zero_rows;
    ip.n_zero = s->n_zero;
    ip.bad_flag = s->flags;
    ip.max_norm = s->max_norm;
    ip.first_row = s->n;
    ip.n = n;
    ip.dim = s->cfg.dim;
    ip.ld = s->ld;
    ip.dtype = s->cfg.dtype;
    ip.metric = s->cfg.metric;
    s->timer.begin(st, 1);
    MX_CUDA(s, MX_ERR_INSERTION, launch_ingest(ip, st));
    s->timer.end(st);
    return MX_OK;
}

int32_t finish_add(mx_store *s, uint64_t n, cudaStream_t st, uint64_t *first_id_out)
{
    MX_CUDA(s, MX_ERR_INSERTION, cudaMemcpyAsync(s->flags_host, s->flags, 4, cudaMemcpyDeviceToHost, st));
    MX_CUDA(s, MX_ERR_INSERTION, cudaStreamSynchronize(st));
    const uint32_t flags = *s->flags_host;
    if (flags) {
        cudaMemsetAsync(s->flags, 0, 4, st);
        cudaStreamSynchronize(st);
    }
    if (flags & 1u) return fail(s, MX_ERR_INSERTION, "non-finite value in inserted vectors (batch of %llu rejected)",
                                (unsigned long long)n);
    if (flags & 2u) s->zero_dirty = true;
    if (first_id_out) *first_id_out = s->cfg.id_offset + s->n * s->cfg.id_stride + 1;
    s->n += n;
    return MX_OK;
}

uint32_t pick_path(mx_store *s, uint32_t nq, uint32_t k)
{
    if (s->force_path >= 0) return (uint32_t)s->force_path;
    if (s->cfg.dtype == MX_DTYPE_F16 && s->tc && tc_scan_supports(s->tc, k) && nq >= 8) return 2;
    return s->cfg.dtype == MX_DTYPE_F16 ? 1 : 0;
}

uint32_t stream_lists(const mx_store *s)
{
    const uint64_t rows_per_cta = 16 * 16;  // a CTA iteration covers at least this many rows
    return (uint32_t)std::min<uint64_t>((uint64_t)s->sm_count, ceil_div<uint64_t>(s->n_view, rows_per_cta));
}

// Second pass over the queries the certificate flagged (normally none: both kernels return at once).  The exact scan
// re-scores every row that can still matter with the reference's f64 fold and keeps its lists on the exact key; the
// rerank kernel then answers the flagged queries from those lists, overwriting their slots in the outputs.
int32_t search_fallback(mx_store *s, const float *q_use, uint32_t nq, uint32_t k, uint64_t *ids_dev, float *scores_dev,
                        float *dists_dev, uint32_t *counts_dev, cudaStream_t st)
{
    const uint32_t n_lists = stream_lists(s), lcap = scan_stream_lcap(k);
    ExactScanParams ep{};
    ep.scan.rows = s->rows;
    ep.scan.inv_norm = s->inv_norm;
    ep.scan.queries = q_use;
    ep.scan.cand_s = s->cand_s;
    ep.scan.cand_r = s->cand_r;
    ep.scan.n_rows = (uint32_t)s->n_view;
    ep.scan.ld = s->ld;
    ep.scan.ldq = s->ldq;
    ep.scan.nq = nq;
    ep.scan.n_lists = n_lists;
    ep.scan.use_inv = s->cfg.metric == MX_METRIC_COSINE ? 1u : 0u;
    ep.active_n = s->n_flagged;
    ep.active_map = s->q_map;
    ep.fb_thr = s->fb_thr;
    ep.dim = s->cfg.dim;
    ep.metric = s->cfg.metric;
    s->timer.begin(st, 1);
    MX_CUDA(s, MX_ERR_SEARCH, launch_scan_exact(ep, s->cfg.dtype, k, n_lists, st));
    s->timer.end(st);
    RerankParams rp{};
    rp.rows = s->rows;
    rp.queries = q_use;
    rp.cand_s = s->cand_s;
    rp.cand_r = s->cand_r;
    rp.zero_rows = s->zero_rows;
    rp.n_zero = s->n_zero;
    rp.ids_out = ids_dev;
    rp.scores_out = scores_dev;
    rp.dists_out = dists_dev;
    rp.counts_out = counts_dev;
    rp.id_offset = s->cfg.id_offset;
    rp.id_stride = s->cfg.id_stride;
    rp.n_rows = (uint32_t)s->n_view;
    rp.ld = s->ld;
    rp.ldq = s->ldq;
    rp.dim = s->cfg.dim;
    rp.nq = nq;
    rp.k = k;
    rp.n_lists = n_lists;
    rp.lcap = lcap;
    rp.dtype = s->cfg.dtype;
    rp.metric = s->cfg.metric;
    rp.active_n = s->n_flagged;
    rp.active_map = s->q_map;
    s->timer.begin(st, 1);
    MX_CUDA(s, MX_ERR_SEARCH, launch_rerank(rp, st));
    s->timer.end(st);
    return MX_OK;
}

// defer_fallback: the caller reads *n_flagged back itself and runs search_fallback only when it is non-zero (the
// host-buffer call synchronises anyway); otherwise the two fallback kernels are always enqueued and exit at once
// when nothing was flagged (the asynchronous device-buffer call cannot look)
int32_t search_device_impl(mx_store *s, const float *q_dev, uint32_t nq, uint32_t k, uint64_t *ids_dev,
                           float *scores_dev, float *dists_dev, uint32_t *counts_dev, cudaStream_t st,
                           bool defer_fallback = false, const float **q_used = nullptr)
{
    if (!q_dev || !ids_dev || !scores_dev || !counts_dev) return fail(s, MX_ERR_INVALID, "null buffer");
    if (k == 0 || k > MX_MAX_K) return fail(s, MX_ERR_INVALID, "k must be in [1, %u]", MX_MAX_K);
    if (nq == 0) return MX_OK;
    s->n_view = s->n;
    if (s->n_view == 0) {
        // empty index: HnswStore::search returns no neighbours (local.rs:76-90)
        MX_CUDA(s, MX_ERR_SEARCH, cudaMemsetAsync(counts_dev, 0, sizeof(uint32_t) * nq, st));
        MX_CUDA(s, MX_ERR_SEARCH, cudaMemsetAsync(ids_dev, 0, sizeof(uint64_t) * nq * k, st));
        MX_CUDA(s, MX_ERR_SEARCH, cudaMemsetAsync(scores_dev, 0, sizeof(float) * nq * k, st));
        if (dists_dev) MX_CUDA(s, MX_ERR_SEARCH, cudaMemsetAsync(dists_dev, 0, sizeof(float) * nq * k, st));
        return MX_OK;
    }
    if (s->zero_dirty) {
        MX_CUDA(s, MX_ERR_SEARCH, launch_collect_zero_rows(s->inv_norm, s->n_view, s->zero_rows, s->n_zero, st));
        s->zero_dirty = false;
    }
    const uint32_t path = pick_path(s, nq, k);
    if (path == 2 && !(s->cfg.dtype == MX_DTYPE_F16 && s->tc && tc_scan_supports(s->tc, k)))
        return fail(s, MX_ERR_UNSUPPORTED, "tcgen05 scan needs an fp16 store and k <= %u", tc_scan_max_k());
    if ((path == 0) != (s->cfg.dtype == MX_DTYPE_F32))
        return fail(s, MX_ERR_UNSUPPORTED, "scan path %u does not match the store dtype", path);

    // queries as [nq, ldq] f32, zero padded to a multiple of 8 floats: when dim already is one (384, 768, ...)
    // the caller's buffer is used in place, otherwise it is staged once
    const float *q_use = q_dev;
    if (s->ldq != s->cfg.dim) {
        const size_t qfloats = (size_t)nq * s->ldq;
        if (qfloats > s->q_stage_cap) {
            if (s->q_stage) cudaFree(s->q_stage);
            s->q_stage = nullptr;
            s->q_stage_cap = 0;
            MX_CUDA(s, MX_ERR_SEARCH, cudaMalloc(&s->q_stage, qfloats * sizeof(float)));
            s->q_stage_cap = qfloats;
        }
        s->timer.begin(st, 1);
        MX_CUDA(s, MX_ERR_SEARCH, launch_stage_queries(q_dev, nq, s->cfg.dim, s->ldq, s->q_stage, st));
        s->timer.end(st);
        q_use = s->q_stage;
    }
    if (q_used) *q_used = q_use;

    uint32_t n_lists, lcap;
    if (path == 2) {
        n_lists = tc_scan_lists(s->tc, s->n_view);
        lcap = tc_scan_lcap(k);
    } else {
        n_lists = stream_lists(s);
        lcap = scan_stream_lcap(k);
    }
    // the exact fallback writes [flagged][stream lists][stream lcap] into the same candidate buffer
    const size_t cand = std::max((size_t)nq * n_lists * lcap, (size_t)nq * stream_lists(s) * scan_stream_lcap(k));
    if (cand > s->cand_cap) {
        if (s->cand_s) cudaFree(s->cand_s);
        if (s->cand_r) cudaFree(s->cand_r);
        s->cand_s = nullptr;
        s->cand_r = nullptr;
        s->cand_cap = 0;
        MX_CUDA(s, MX_ERR_SEARCH, cudaMalloc(&s->cand_s, cand * sizeof(float)));
        MX_CUDA(s, MX_ERR_SEARCH, cudaMalloc(&s->cand_r, cand * sizeof(uint32_t)));
        s->cand_cap = cand;
    }
    if (s->verify && nq > s->verify_cap) {
        cudaFree(s->q_map);
        cudaFree(s->fb_thr);
        s->q_map = nullptr;
        s->fb_thr = nullptr;
        s->verify_cap = 0;
        MX_CUDA(s, MX_ERR_SEARCH, cudaMalloc(&s->q_map, (size_t)nq * sizeof(uint32_t)));
        MX_CUDA(s, MX_ERR_SEARCH, cudaMalloc(&s->fb_thr, (size_t)nq * sizeof(float)));
        s->verify_cap = nq;
    }
    // the flagged-query counter alternates between two cells: this search's rerank zeroes the other cell for the next
    // search (no memset node in the stream)
    s->flag_cur ^= 1u;
    s->n_flagged = s->n_flagged_buf + s->flag_cur;

    ScanParams sp{};
    sp.rows = s->rows;
    sp.inv_norm = s->inv_norm;
    sp.queries = q_use;
    sp.cand_s = s->cand_s;
    sp.cand_r = s->cand_r;
    sp.n_rows = (uint32_t)s->n_view;
    sp.ld = s->ld;
    sp.ldq = s->ldq;
    sp.nq = nq;
    sp.n_lists = n_lists;
    sp.use_inv = s->cfg.metric == MX_METRIC_COSINE ? 1u : 0u;
    if (path == 2) {
        const char *why = nullptr;
        cudaError_t e = tc_scan_launch(s->tc, sp, s->capacity, k, &s->timer, st, &why);
        if (e != cudaSuccess)
            return fail(s, MX_ERR_SEARCH, "tcgen05 scan failed: %s (%s)", cudaGetErrorString(e), why ? why : "");
    } else {
        s->timer.begin(st, 0);
        MX_CUDA(s, MX_ERR_SEARCH, launch_scan_stream(sp, s->cfg.dtype, k, n_lists, st));
        s->timer.end(st);
    }

    RerankParams rp{};
    rp.rows = s->rows;
    rp.queries = q_use;
    rp.cand_s = s->cand_s;
    rp.cand_r = s->cand_r;
    rp.zero_rows = s->zero_rows;
    rp.n_zero = s->n_zero;
    rp.ids_out = ids_dev;
    rp.scores_out = scores_dev;
    rp.dists_out = dists_dev;
    rp.counts_out = counts_dev;
    rp.id_offset = s->cfg.id_offset;
    rp.id_stride = s->cfg.id_stride;
    rp.n_rows = (uint32_t)s->n_view;
    rp.ld = s->ld;
    rp.ldq = s->ldq;
    rp.dim = s->cfg.dim;
    rp.nq = nq;
    rp.k = k;
    rp.n_lists = n_lists;
    rp.lcap = lcap;
    rp.dtype = s->cfg.dtype;
    rp.metric = s->cfg.metric;
    if (s->verify) {
        rp.n_flagged = s->n_flagged;
        rp.n_flagged_next = s->n_flagged_buf + (s->flag_cur ^ 1u);
        rp.prof = s->prof;
        rp.q_map = s->q_map;
        rp.fb_thr = s->fb_thr;
        rp.stats = s->stats;
        rp.max_norm = s->max_norm;
        if (path == 2) {
            rp.qerr = tc_scan_qerr(s->tc);
            rp.scan_floor = tc_scan_floor(s->tc);
            rp.unit_queries = 1;
        }
    }
    s->timer.begin(st, 1);
    MX_CUDA(s, MX_ERR_SEARCH, launch_rerank(rp, st));
    s->timer.end(st);
    if (s->verify && !defer_fallback) return search_fallback(s, q_use, nq, k, ids_dev, scores_dev, dists_dev, counts_dev, st);
    return MX_OK;
}

bool dir_join(const char *dir, std::string &out)
{
    if (!dir) return false;
    out = dir;
    if (!out.empty() && out.back() != '/') out += '/';
    out += kFileName;
    return true;
}

}  // namespace

extern "C" {

int32_t mx_abi_version(void) { return MX_ABI_VERSION; }
uint64_t mx_launch_count(void) { return g_launches.load(); }
int32_t mx_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char *mx_last_error(const void *handle)
{
    if (!handle) return g_last_error.c_str();
    return static_cast<const HandleBase *>(handle)->last_error.c_str();
}

int32_t mx_store_create(const mx_store_cfg *cfg, mx_store **out)
{
    if (!cfg || !out) return fail(nullptr, MX_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->dim == 0 || cfg->dim > 65536) return fail(nullptr, MX_ERR_INVALID, "dim must be in [1, 65536]");
    if (cfg->dtype > MX_DTYPE_F16) return fail(nullptr, MX_ERR_UNSUPPORTED, "unknown dtype %u", cfg->dtype);
    if (cfg->metric > MX_METRIC_DOT) return fail(nullptr, MX_ERR_UNSUPPORTED, "unknown metric %u", cfg->metric);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, MX_ERR_CONNECTION, "no CUDA device: %s (this library has no CPU path)",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (cfg->device < 0 || cfg->device >= ndev)
        return fail(nullptr, MX_ERR_CONNECTION, "device %d out of range [0, %d)", cfg->device, ndev);
    mx_store *s = new mx_store();
    s->magic = kStoreMagic;
    s->cfg = *cfg;
    if (s->cfg.id_stride == 0) s->cfg.id_stride = 1;
    s->elem = cfg->dtype == MX_DTYPE_F32 ? 4 : 2;
    const uint32_t per16 = (uint32_t)(16 / s->elem);
    s->ld = ceil_div<uint32_t>(cfg->dim, per16) * per16;
    s->ldq = ceil_div<uint32_t>(cfg->dim, 8) * 8;
    auto bail = [&](int32_t code, const char *what, cudaError_t ce) {
        int32_t r = fail(nullptr, code, "%s: %s", what, cudaGetErrorString(ce));
        mx_store_destroy(s);
        return r;
    };
    if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return bail(MX_ERR_CONNECTION, "cudaSetDevice", e);
    cudaDeviceProp prop{};
    if ((e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess)
        return bail(MX_ERR_CONNECTION, "cudaGetDeviceProperties", e);
    if (prop.major != 10)
        {
            int32_t r = fail(nullptr, MX_ERR_CONNECTION, "device %d is sm_%d%d; this library ships sm_100a code only",
                             cfg->device, prop.major, prop.minor);
            mx_store_destroy(s);
            return r;
        }
    s->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(MX_ERR_CONNECTION, "cudaStreamCreate", e);
    if ((e = cudaMalloc(&s->zero_rows, sizeof(uint32_t) * MX_MAX_K)) != cudaSuccess ||
        (e = cudaMalloc(&s->n_zero, 4)) != cudaSuccess || (e = cudaMalloc(&s->flags, 4)) != cudaSuccess)
        return bail(MX_ERR_CONNECTION, "cudaMalloc", e);
    if ((e = cudaMallocHost(&s->flags_host, 64)) != cudaSuccess) return bail(MX_ERR_CONNECTION, "cudaMallocHost", e);
    *s->flags_host = 0;
    if ((e = cudaMalloc(&s->n_flagged_buf, 8)) != cudaSuccess || (e = cudaMalloc(&s->stats, 16)) != cudaSuccess ||
        (e = cudaMalloc(&s->max_norm, 4)) != cudaSuccess)
        return bail(MX_ERR_CONNECTION, "cudaMalloc", e);
    cudaMemsetAsync(s->n_zero, 0, 4, s->stream);
    cudaMemsetAsync(s->flags, 0, 4, s->stream);
    cudaMemsetAsync(s->n_flagged_buf, 0, 8, s->stream);
    s->n_flagged = s->n_flagged_buf;
    cudaMemsetAsync(s->stats, 0, 16, s->stream);
    cudaMemsetAsync(s->max_norm, 0, 4, s->stream);
    if (const char *v = getenv("MX_SEARCH_VERIFY")) s->verify = atoi(v) != 0;   // A/B measurements only
    if (getenv("MX_RERANK_PROF")) {
        if (cudaMalloc(&s->prof, 64) == cudaSuccess) cudaMemsetAsync(s->prof, 0, 64, s->stream);
    }
    if (cfg->dtype == MX_DTYPE_F16) s->tc = tc_scan_create(s->sm_count, s->ld, cfg->dim);
    int32_t rc = reserve(s, cfg->capacity ? cfg->capacity : 1024, MX_ERR_CONNECTION);
    if (rc != MX_OK) {
        g_last_error = s->last_error;
        mx_store_destroy(s);
        return rc;
    }
    *out = s;
    return MX_OK;
}

void mx_store_destroy(mx_store *s)
{
    if (!s) return;
    cudaSetDevice(s->cfg.device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->tc) tc_scan_destroy(s->tc);
    cudaFree(s->rows);
    cudaFree(s->inv_norm);
    cudaFree(s->zero_rows);
    cudaFree(s->n_zero);
    cudaFree(s->flags);
    cudaFree(s->q_stage);
    cudaFree(s->cand_s);
    cudaFree(s->cand_r);
    cudaFree(s->dev_io);
    cudaFree(s->blob_scores);
    cudaFree(s->n_flagged_buf);
    cudaFree(s->q_map);
    cudaFree(s->fb_thr);
    cudaFree(s->stats);
    cudaFree(s->max_norm);
    cudaFree(s->prof);
    if (s->pinned) cudaFreeHost(s->pinned);
    if (s->flags_host) cudaFreeHost(s->flags_host);
    for (IoSlot &sl : s->slots) sl.release();
    if (s->stream) cudaStreamDestroy(s->stream);
    s->magic = 0;
    delete s;
}

int32_t mx_store_add(mx_store *s, const float *vecs, uint64_t n, uint64_t *first_id_out)
{
    if (!s) return MX_ERR_INVALID;
    if (n == 0) {
        if (first_id_out) *first_id_out = s->cfg.id_offset + s->n * s->cfg.id_stride + 1;
        return MX_OK;
    }
    if (!vecs) return fail(s, MX_ERR_INVALID, "null vectors");
    int32_t rc;
    if ((rc = set_device(s)) != MX_OK) return rc;
    if ((rc = reserve(s, s->n + n, MX_ERR_INSERTION)) != MX_OK) return rc;
    // stream the batch through a pinned + device staging buffer in chunks
    const size_t row_bytes = (size_t)s->cfg.dim * sizeof(float);
    const uint64_t chunk_rows = std::max<uint64_t>(1, std::min<uint64_t>(n, (64ull << 20) / row_bytes));
    if ((rc = ensure_pinned(s, chunk_rows * row_bytes)) != MX_OK) return rc;
    if ((rc = ensure_dev_io(s, chunk_rows * row_bytes)) != MX_OK) return rc;
    const uint64_t n_before = s->n;
    // ingest appends at s->n, so the count moves with the chunks; EVERY exit below puts it back, and only finish_add
    // commits the batch (a failed copy must not leave rows the host's id map knows nothing about)
    struct Rollback {
        mx_store *s;
        uint64_t n;
        ~Rollback() { s->n = n; }
    } rollback{s, n_before};
    uint64_t done = 0;
    while (done < n) {
        const uint64_t c = std::min(chunk_rows, n - done);
        memcpy(s->pinned, vecs + done * s->cfg.dim, c * row_bytes);
        MX_CUDA(s, MX_ERR_INSERTION,
                cudaMemcpyAsync(s->dev_io, s->pinned, c * row_bytes, cudaMemcpyHostToDevice, s->stream));
        s->n = n_before + done;
        rc = ingest_device(s, (const float *)s->dev_io, c, s->stream);
        if (rc != MX_OK) return rc;
        MX_CUDA(s, MX_ERR_INSERTION, cudaStreamSynchronize(s->stream));  // pinned buffer is reused
        done += c;
    }
    s->n = n_before;
    rc = finish_add(s, n, s->stream, first_id_out);
    rollback.n = s->n;   // committed (or unchanged when the batch was rejected)
    return rc;
}

int32_t mx_store_add_device(mx_store *s, const float *vecs_dev, uint64_t n, uint64_t *first_id_out)
{
    if (!s) return MX_ERR_INVALID;
    if (n == 0) {
        if (first_id_out) *first_id_out = s->cfg.id_offset + s->n * s->cfg.id_stride + 1;
        return MX_OK;
    }
    if (!vecs_dev) return fail(s, MX_ERR_INVALID, "null vectors");
    int32_t rc;
    if ((rc = set_device(s)) != MX_OK) return rc;
    if ((rc = reserve(s, s->n + n, MX_ERR_INSERTION)) != MX_OK) return rc;
    // the caller's producer may run on another stream: order after everything issued so far
    MX_CUDA(s, MX_ERR_INSERTION, cudaDeviceSynchronize());
    if ((rc = ingest_device(s, vecs_dev, n, s->stream)) != MX_OK) return rc;
    return finish_add(s, n, s->stream, first_id_out);
}

int32_t mx_store_add_device_stream(mx_store *s, const float *vecs_dev, uint64_t n, uint64_t *first_id_out, void *cuda_stream)
{
    if (!s) return MX_ERR_INVALID;
    if (n == 0) {
        if (first_id_out) *first_id_out = s->cfg.id_offset + s->n * s->cfg.id_stride + 1;
        return MX_OK;
    }
    if (!vecs_dev) return fail(s, MX_ERR_INVALID, "null vectors");
    int32_t rc;
    if ((rc = set_device(s)) != MX_OK) return rc;
    // concurrent with searches of another host thread: the row matrix must not move (no growth on this path)
    if (s->n + n > s->capacity)
        return fail(s, MX_ERR_INSERTION, "capacity %llu exhausted (streamed ingest does not grow the store: create it with the "
                                         "capacity it will reach)", (unsigned long long)s->capacity);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : s->stream;
    IngestParams ip{};
    ip.src = vecs_dev;
    ip.rows = s->rows;
    ip.inv_norm = s->inv_norm;
    ip.zero_rows = s->zero_rows;
    ip.n_zero = s->n_zero;
    ip.bad_flag = s->flags;
    ip.max_norm = s->max_norm;
    ip.first_row = s->n;
    ip.n = n;
    ip.dim = s->cfg.dim;
    ip.ld = s->ld;
    ip.dtype = s->cfg.dtype;
    ip.metric = s->cfg.metric;
    MX_CUDA(s, MX_ERR_INSERTION, launch_ingest(ip, st));
    // the rows become visible to searches (the committed-rows watermark s->n) only once they are in place
    return finish_add(s, n, st, first_id_out);
}

int32_t mx_store_set_sm_limit(mx_store *s, uint32_t sms)
{
    if (!s) return MX_ERR_INVALID;
    int32_t rc;
    if ((rc = set_device(s)) != MX_OK) return rc;
    cudaDeviceProp prop{};
    MX_CUDA(s, MX_ERR_CONNECTION, cudaGetDeviceProperties(&prop, s->cfg.device));
    const uint32_t all = (uint32_t)prop.multiProcessorCount;
    s->sm_count = (int)(sms == 0 || sms >= all ? all : std::max<uint32_t>(1u, sms));
    MX_CUDA(s, MX_ERR_CONNECTION, cudaStreamSynchronize(s->stream));
    if (s->tc) tc_scan_set_sms(s->tc, s->sm_count);
    return MX_OK;
}

int32_t mx_store_search_device(mx_store *s, const float *queries_dev, uint32_t nq, uint32_t k, uint64_t *ids_dev,
                               float *scores_dev, float *dists_dev, uint32_t *counts_dev, void *cuda_stream)
{
    if (!s) return MX_ERR_INVALID;
    int32_t rc;
    if ((rc = set_device(s)) != MX_OK) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : s->stream;
    return search_device_impl(s, queries_dev, nq, k, ids_dev, scores_dev, dists_dev, counts_dev, st);
}

int32_t mx_store_search(mx_store *s, const float *queries, uint32_t nq, uint32_t k, uint64_t *ids_out,
                        float *scores_out, uint32_t *counts_out)
{
    if (!s) return MX_ERR_INVALID;
    if (!queries || !ids_out || !scores_out || !counts_out) return fail(s, MX_ERR_INVALID, "null buffer");
    if (k == 0 || k > MX_MAX_K) return fail(s, MX_ERR_INVALID, "k must be in [1, %u]", MX_MAX_K);
    if (nq == 0) return MX_OK;
    if (s->n == 0) {
        // empty index: HnswStore::search returns no neighbours (local.rs:76-90)
        memset(ids_out, 0, (size_t)nq * k * sizeof(uint64_t));
        memset(scores_out, 0, (size_t)nq * k * sizeof(float));
        memset(counts_out, 0, (size_t)nq * sizeof(uint32_t));
        return MX_OK;
    }
    g_host_prof.start();
    int32_t rc;
    if ((rc = set_device(s)) != MX_OK) return rc;
    const size_t qb = (size_t)nq * s->cfg.dim * sizeof(float);
    const size_t ib = (size_t)nq * k * sizeof(uint64_t), sb = (size_t)nq * k * sizeof(float), cb = (size_t)nq * 4;
    const size_t off_i = (qb + 255) & ~(size_t)255, off_s = off_i + ((ib + 255) & ~(size_t)255),
                 off_c = off_s + ((sb + 255) & ~(size_t)255), total = off_c + ((cb + 255) & ~(size_t)255);
    if ((rc = ensure_pinned(s, total + 16)) != MX_OK) return rc;
    if ((rc = ensure_dev_io(s, total)) != MX_OK) return rc;
    char *hp = (char *)s->pinned, *dp = (char *)s->dev_io;
    g_host_prof.mark(1);
    const int64_t bad_row = copy_checking_finite(reinterpret_cast<float *>(hp), queries, nq, s->cfg.dim);
    if (bad_row >= 0) return fail(s, MX_ERR_SEARCH, "non-finite value in query %lld", (long long)bad_row);
    g_host_prof.mark(2);
    MX_CUDA(s, MX_ERR_SEARCH, cudaMemcpyAsync(dp, hp, qb, cudaMemcpyHostToDevice, s->stream));
    g_host_prof.mark(3);
    // the answer goes STRAIGHT to the pinned host buffer: the rerank kernel's ~8 KB of stores travel over PCIe as posted
    // writes (the kernel never reads them back), which takes a device-to-host copy operation and its ~5 us of latency off
    // the end of every call; only the 4-byte flagged-query counter is still copied
    const float *q_used = nullptr;
    rc = search_device_impl(s, (const float *)dp, nq, k, (uint64_t *)(hp + off_i), (float *)(hp + off_s), nullptr,
                            (uint32_t *)(hp + off_c), s->stream, true, &q_used);
    if (rc != MX_OK) return rc;
    g_host_prof.mark(4);
    uint32_t *flagged_h = reinterpret_cast<uint32_t *>(hp + total);
    *flagged_h = 0;
    if (s->verify && s->n_view > 0)
        MX_CUDA(s, MX_ERR_SEARCH, cudaMemcpyAsync(flagged_h, s->n_flagged, 4, cudaMemcpyDeviceToHost, s->stream));
    g_host_prof.mark(5);
    MX_CUDA(s, MX_ERR_SEARCH, cudaStreamSynchronize(s->stream));
    g_host_prof.mark(6);
    if (*flagged_h > 0) {
        // the certificate could not vouch for some queries: answer those again with the exact scan
        rc = search_fallback(s, q_used, nq, k, (uint64_t *)(hp + off_i), (float *)(hp + off_s), nullptr,
                             (uint32_t *)(hp + off_c), s->stream);
        if (rc != MX_OK) return rc;
        MX_CUDA(s, MX_ERR_SEARCH, cudaStreamSynchronize(s->stream));
    }
    memcpy(ids_out, hp + off_i, ib);
    memcpy(scores_out, hp + off_s, sb);
    memcpy(counts_out, hp + off_c, cb);
    g_host_prof.mark(7);
    g_host_prof.calls++;
    return MX_OK;
}

int32_t mx_store_search_submit(mx_store *s, const float *queries, uint32_t nq, uint32_t k, uint64_t *ticket_out)
{
    if (!s) return MX_ERR_INVALID;
    if (!queries || !ticket_out) return fail(s, MX_ERR_INVALID, "null buffer");
    if (k == 0 || k > MX_MAX_K) return fail(s, MX_ERR_INVALID, "k must be in [1, %u]", MX_MAX_K);
    IoSlot &sl = s->slots[s->submit_seq & 1];
    if (sl.busy)
        return fail(s, MX_ERR_INVALID, "two searches are already in flight: collect ticket %llu first", (unsigned long long)sl.ticket);
    const size_t qb = (size_t)nq * s->cfg.dim * sizeof(float);
    sl.layout(qb, nq, k);
    sl.empty = nq == 0 || s->n == 0;     // HnswStore::search on an empty index returns no neighbours (local.rs:76-90)
    if (!sl.empty) {
        int32_t rc;
        if ((rc = set_device(s)) != MX_OK) return rc;
        MX_CUDA(s, MX_ERR_CONNECTION, sl.reserve(sl.total()));
        char *hp = static_cast<char *>(sl.pinned), *dp = static_cast<char *>(sl.dev);
        const int64_t bad_row = copy_checking_finite(reinterpret_cast<float *>(hp), queries, nq, s->cfg.dim);
        if (bad_row >= 0) return fail(s, MX_ERR_SEARCH, "non-finite value in query %lld", (long long)bad_row);
        MX_CUDA(s, MX_ERR_SEARCH, cudaMemcpyAsync(dp, hp, qb, cudaMemcpyHostToDevice, s->stream));
        // nobody looks at the flagged-query counter between the rerank and the fallback here (the call returns at once): the
        // exact-scan pair is always enqueued and exits immediately when no query was flagged
        rc = search_device_impl(s, reinterpret_cast<const float *>(dp), nq, k, reinterpret_cast<uint64_t *>(hp + sl.off_i),
                                reinterpret_cast<float *>(hp + sl.off_s), nullptr, reinterpret_cast<uint32_t *>(hp + sl.off_c),
                                s->stream);
        if (rc != MX_OK) return rc;
        MX_CUDA(s, MX_ERR_SEARCH, cudaEventRecord(sl.done, s->stream));
    }
    sl.busy = true;
    sl.ticket = s->submit_seq;
    *ticket_out = s->submit_seq++;
    return MX_OK;
}

int32_t mx_store_search_collect(mx_store *s, uint64_t ticket, uint64_t *ids_out, float *scores_out, uint32_t *counts_out)
{
    if (!s) return MX_ERR_INVALID;
    if (!ids_out || !scores_out || !counts_out) return fail(s, MX_ERR_INVALID, "null buffer");
    IoSlot &sl = s->slots[ticket & 1];
    if (!sl.busy || sl.ticket != ticket) return fail(s, MX_ERR_INVALID, "no search with ticket %llu is in flight", (unsigned long long)ticket);
    sl.busy = false;
    if (sl.empty) {
        memset(ids_out, 0, (size_t)sl.nq * sl.k * sizeof(uint64_t));
        memset(scores_out, 0, (size_t)sl.nq * sl.k * sizeof(float));
        memset(counts_out, 0, (size_t)sl.nq * sizeof(uint32_t));
        return MX_OK;
    }
    int32_t rc;
    if ((rc = set_device(s)) != MX_OK) return rc;
    MX_CUDA(s, MX_ERR_SEARCH, cudaEventSynchronize(sl.done));
    sl.copy_out(ids_out, scores_out, counts_out);
    return MX_OK;
}

static int32_t merge_impl(const MergeParams &mp, int32_t device, void *cuda_stream)
{
    if (!mp.ids || !mp.dists || !mp.counts || !mp.ids_out || !mp.scores_out || !mp.counts_out)
        return fail(nullptr, MX_ERR_INVALID, "null buffer");
    if (mp.k == 0 || mp.k > MX_MAX_K || mp.n_shards == 0 || mp.metric > MX_METRIC_DOT)
        return fail(nullptr, MX_ERR_INVALID, "bad merge shape");
    if (mp.nq == 0) return MX_OK;
    MX_CUDA(nullptr, MX_ERR_CONNECTION, cudaSetDevice(device));
    MX_CUDA(nullptr, MX_ERR_SEARCH, launch_merge(mp, (cudaStream_t)cuda_stream));
    return MX_OK;
}

int32_t mx_merge_topk_device(const uint64_t *ids_dev, const float *dists_dev, const uint32_t *counts_dev,
                             uint32_t n_shards, uint32_t nq, uint32_t k, uint32_t metric, uint64_t *ids_out_dev,
                             float *scores_out_dev, uint32_t *counts_out_dev, int32_t device, void *cuda_stream)
{
    MergeParams mp{ids_dev, dists_dev, counts_dev, (uint64_t)nq * k * 8, (uint64_t)nq * k * 4, (uint64_t)nq * 4,
                   ids_out_dev, scores_out_dev, counts_out_dev, n_shards, nq, k, metric};
    return merge_impl(mp, device, cuda_stream);
}

uint64_t mx_topk_blob_bytes(uint32_t nq, uint32_t k) { return (((uint64_t)nq * k * 12 + (uint64_t)nq * 4) + 15) & ~15ull; }

int32_t mx_store_search_blob_device(mx_store *s, const float *queries_dev, uint32_t nq, uint32_t k, void *blob_dev,
                                    void *cuda_stream)
{
    if (!blob_dev) return fail(s, MX_ERR_INVALID, "null buffer");
    char *b = static_cast<char *>(blob_dev);
    // scores are not part of the blob (the merge recomputes them from the keys): park them behind the keys' slot
    // of a scratch area the store owns
    if (!s) return MX_ERR_INVALID;
    int32_t rc;
    if ((rc = set_device(s)) != MX_OK) return rc;
    const size_t sb = (size_t)nq * k * sizeof(float);
    if (sb > s->blob_scores_cap) {
        if (s->blob_scores) cudaFree(s->blob_scores);
        s->blob_scores = nullptr;
        s->blob_scores_cap = 0;
        MX_CUDA(s, MX_ERR_SEARCH, cudaMalloc(&s->blob_scores, sb));
        s->blob_scores_cap = sb;
    }
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : s->stream;
    return search_device_impl(s, queries_dev, nq, k, reinterpret_cast<uint64_t *>(b), s->blob_scores,
                              reinterpret_cast<float *>(b + (size_t)nq * k * 8),
                              reinterpret_cast<uint32_t *>(b + (size_t)nq * k * 12), st);
}

int32_t mx_merge_topk_blobs_device(const void *blobs_dev, uint64_t blob_stride_bytes, uint32_t n_shards, uint32_t nq,
                                   uint32_t k, uint32_t metric, uint64_t *ids_out_dev, float *scores_out_dev,
                                   uint32_t *counts_out_dev, int32_t device, void *cuda_stream)
{
    if (!blobs_dev) return fail(nullptr, MX_ERR_INVALID, "null buffer");
    if (blob_stride_bytes < mx_topk_blob_bytes(nq, k) || blob_stride_bytes % 8 != 0)
        return fail(nullptr, MX_ERR_INVALID, "blob stride too small or not 8-byte aligned");
    const char *b = static_cast<const char *>(blobs_dev);
    MergeParams mp{reinterpret_cast<const uint64_t *>(b), reinterpret_cast<const float *>(b + (size_t)nq * k * 8),
                   reinterpret_cast<const uint32_t *>(b + (size_t)nq * k * 12), blob_stride_bytes, blob_stride_bytes,
                   blob_stride_bytes, ids_out_dev, scores_out_dev, counts_out_dev, n_shards, nq, k, metric};
    return merge_impl(mp, device, cuda_stream);
}

int32_t mx_merge_topk_blobs_wait_device(const void *blobs_dev, uint64_t blob_stride_bytes, uint32_t n_shards, uint32_t nq,
                                        uint32_t k, uint32_t metric, uint64_t *ids_out_dev, float *scores_out_dev,
                                        uint32_t *counts_out_dev, const uint32_t *flags_dev, uint32_t epoch, int32_t device,
                                        void *cuda_stream)
{
    if (!blobs_dev || !flags_dev) return fail(nullptr, MX_ERR_INVALID, "null buffer");
    if (blob_stride_bytes < mx_topk_blob_bytes(nq, k) || blob_stride_bytes % 8 != 0)
        return fail(nullptr, MX_ERR_INVALID, "blob stride too small or not 8-byte aligned");
    if (n_shards > (uint32_t)kMaxPeers) return fail(nullptr, MX_ERR_INVALID, "more than %d shards", kMaxPeers);
    const char *b = static_cast<const char *>(blobs_dev);
    MergeParams mp{reinterpret_cast<const uint64_t *>(b), reinterpret_cast<const float *>(b + (size_t)nq * k * 8),
                   reinterpret_cast<const uint32_t *>(b + (size_t)nq * k * 12), blob_stride_bytes, blob_stride_bytes,
                   blob_stride_bytes, ids_out_dev, scores_out_dev, counts_out_dev, n_shards, nq, k, metric};
    mp.wait_flags = flags_dev;
    mp.wait_epoch = epoch;
    return merge_impl(mp, device, cuda_stream);
}

int32_t mx_exchange_push_device(const void *blob_dev, uint64_t blob_bytes, const uint64_t *peer_bases, uint32_t world,
                                uint32_t rank, uint64_t slot_offset_bytes, uint64_t flag_offset_bytes, uint32_t epoch,
                                int32_t device, void *cuda_stream)
{
    if (!blob_dev || !peer_bases) return fail(nullptr, MX_ERR_INVALID, "null buffer");
    if (world == 0 || world > (uint32_t)kMaxPeers || rank >= world || blob_bytes % 16 != 0 || slot_offset_bytes % 16 != 0 ||
        flag_offset_bytes % 4 != 0)
        return fail(nullptr, MX_ERR_INVALID, "bad exchange shape");
    PushParams pp{};
    pp.blob = blob_dev;
    pp.blob_bytes = blob_bytes;
    for (uint32_t i = 0; i < world; ++i) {
        if (!peer_bases[i]) return fail(nullptr, MX_ERR_INVALID, "null peer buffer");
        pp.peer_base[i] = peer_bases[i];
    }
    pp.slot_offset = slot_offset_bytes;
    pp.flag_offset = flag_offset_bytes;
    pp.world = world;
    pp.rank = rank;
    pp.epoch = epoch;
    MX_CUDA(nullptr, MX_ERR_CONNECTION, cudaSetDevice(device));
    MX_CUDA(nullptr, MX_ERR_SEARCH, launch_exchange_push(pp, (cudaStream_t)cuda_stream));
    return MX_OK;
}

int32_t mx_debug_scan_tc_prof(void *store, uint64_t *out32)
{
    mx_store *s = static_cast<mx_store *>(store);
    if (!s || !out32 || !s->tc || !tc_scan_prof(s->tc)) return MX_ERR_INVALID;
    cudaSetDevice(s->cfg.device);
    return cudaMemcpy(out32, tc_scan_prof(s->tc), 256, cudaMemcpyDeviceToHost) == cudaSuccess ? MX_OK : MX_ERR_CONNECTION;
}

int32_t mx_debug_host_prof(double *out12, uint64_t *calls, int32_t reset)
{
    if (!out12 || !calls) return MX_ERR_INVALID;
    for (int i = 0; i < HostProf::kSlots; ++i) out12[i] = g_host_prof.acc[i];
    *calls = g_host_prof.calls;
    if (reset) {
        for (int i = 0; i < HostProf::kSlots; ++i) g_host_prof.acc[i] = 0;
        g_host_prof.calls = 0;
    }
    return MX_OK;
}

int32_t mx_debug_rerank_prof(void *store, uint64_t *out8)
{
    mx_store *s = static_cast<mx_store *>(store);
    if (!s || !out8 || !s->prof) return MX_ERR_INVALID;
    cudaSetDevice(s->cfg.device);
    return cudaMemcpy(out8, s->prof, 64, cudaMemcpyDeviceToHost) == cudaSuccess ? MX_OK : MX_ERR_CONNECTION;
}

int32_t mx_store_len(mx_store *s, uint64_t *n_out)
{
    if (!s || !n_out) return MX_ERR_INVALID;
    *n_out = s->n;
    return MX_OK;
}

int32_t mx_store_clear(mx_store *s)
{
    if (!s) return MX_ERR_INVALID;
    int32_t rc;
    if ((rc = set_device(s)) != MX_OK) return rc;
    MX_CUDA(s, MX_ERR_DELETE, cudaStreamSynchronize(s->stream));
    s->n = 0;
    s->zero_dirty = false;
    MX_CUDA(s, MX_ERR_DELETE, cudaMemsetAsync(s->n_zero, 0, 4, s->stream));
    MX_CUDA(s, MX_ERR_DELETE, cudaMemsetAsync(s->max_norm, 0, 4, s->stream));
    return MX_OK;
}

int32_t mx_store_delete(mx_store *s, uint64_t id)
{
    // reference local.rs:29-32 is `unimplemented!()` (a panic); across the C ABI it is a status
    return fail(s, MX_ERR_UNSUPPORTED, "removing a single point (id %llu) is not supported by the file store",
                (unsigned long long)id);
}

int32_t mx_store_sync(mx_store *s)
{
    if (!s) return MX_ERR_INVALID;
    int32_t rc;
    if ((rc = set_device(s)) != MX_OK) return rc;
    MX_CUDA(s, MX_ERR_CONNECTION, cudaStreamSynchronize(s->stream));
    return MX_OK;
}

int32_t mx_store_info(mx_store *s, uint32_t *dim, uint32_t *dtype, uint32_t *metric, uint64_t *capacity)
{
    if (!s) return MX_ERR_INVALID;
    if (dim) *dim = s->cfg.dim;
    if (dtype) *dtype = s->cfg.dtype;
    if (metric) *metric = s->cfg.metric;
    if (capacity) *capacity = s->capacity;
    return MX_OK;
}

int32_t mx_store_scan_path(mx_store *s, uint32_t nq, uint32_t k, int32_t force)
{
    if (!s) return MX_ERR_INVALID;
    if (force >= -1 && force <= 2) s->force_path = force;
    return (int32_t)pick_path(s, nq, k);
}

int32_t mx_store_set_timing(mx_store *s, int32_t on)
{
    if (!s) return MX_ERR_INVALID;
    cudaSetDevice(s->cfg.device);
    s->timer.reset();
    s->timer.on = on != 0;
    return MX_OK;
}

int32_t mx_store_get_timing(mx_store *s, double *scan_ms_total, uint64_t *scan_launches, double *other_ms_total,
                            uint64_t *other_launches)
{
    if (!s) return MX_ERR_INVALID;
    cudaSetDevice(s->cfg.device);
    s->timer.collect();
    if (scan_ms_total) *scan_ms_total = s->timer.total_ms[0];
    if (scan_launches) *scan_launches = s->timer.launches[0];
    if (other_ms_total) *other_ms_total = s->timer.total_ms[1];
    if (other_launches) *other_launches = s->timer.launches[1];
    return MX_OK;
}

int32_t mx_store_verify_stats(mx_store *s, uint64_t *queries_out, uint64_t *flagged_out)
{
    if (!s) return MX_ERR_INVALID;
    int32_t rc;
    if ((rc = set_device(s)) != MX_OK) return rc;
    unsigned long long h[2] = {0, 0};
    MX_CUDA(s, MX_ERR_CONNECTION, cudaDeviceSynchronize());
    MX_CUDA(s, MX_ERR_CONNECTION, cudaMemcpy(h, s->stats, sizeof h, cudaMemcpyDeviceToHost));
    if (queries_out) *queries_out = h[0];
    if (flagged_out) *flagged_out = h[1];
    return MX_OK;
}

int32_t mx_store_set_verify(mx_store *s, int32_t on)
{
    if (!s) return MX_ERR_INVALID;
    s->verify = on != 0;
    return MX_OK;
}

int32_t mx_store_get_rows(mx_store *s, uint64_t first_row, uint64_t n, float *out)
{
    if (!s || !out) return MX_ERR_INVALID;
    if (first_row + n > s->n) return fail(s, MX_ERR_INVALID, "row range [%llu, %llu) exceeds len %llu",
                                          (unsigned long long)first_row, (unsigned long long)(first_row + n),
                                          (unsigned long long)s->n);
    int32_t rc;
    if ((rc = set_device(s)) != MX_OK) return rc;
    const size_t row_bytes = (size_t)s->cfg.dim * sizeof(float);
    const uint64_t chunk_rows = std::max<uint64_t>(1, std::min<uint64_t>(n ? n : 1, (64ull << 20) / row_bytes));
    if ((rc = ensure_pinned(s, chunk_rows * row_bytes)) != MX_OK) return rc;
    if ((rc = ensure_dev_io(s, chunk_rows * row_bytes)) != MX_OK) return rc;
    for (uint64_t done = 0; done < n;) {
        const uint64_t c = std::min(chunk_rows, n - done);
        MX_CUDA(s, MX_ERR_FILE_IO, launch_export_rows(s->rows, s->cfg.dtype, s->ld, s->cfg.dim, first_row + done, c,
                                                      (float *)s->dev_io, s->stream));
        MX_CUDA(s, MX_ERR_FILE_IO,
                cudaMemcpyAsync(s->pinned, s->dev_io, c * row_bytes, cudaMemcpyDeviceToHost, s->stream));
        MX_CUDA(s, MX_ERR_FILE_IO, cudaStreamSynchronize(s->stream));
        memcpy(out + done * s->cfg.dim, s->pinned, c * row_bytes);
        done += c;
    }
    return MX_OK;
}

int32_t mx_store_has_file(const char *dir)
{
    std::string path;
    if (!dir_join(dir, path)) return 0;
    struct stat st;
    return stat(path.c_str(), &st) == 0 ? 1 : 0;
}

int32_t mx_store_remove_file(const char *dir)
{
    std::string path;
    if (!dir_join(dir, path)) return MX_ERR_INVALID;
    if (unlink(path.c_str()) != 0 && errno != ENOENT)
        return fail(nullptr, MX_ERR_FILE_IO, "unlink %s: %s", path.c_str(), strerror(errno));
    return MX_OK;
}

int32_t mx_store_save(mx_store *s, const char *dir)
{
    if (!s) return MX_ERR_INVALID;
    std::string path;
    if (!dir_join(dir, path)) return fail(s, MX_ERR_INVALID, "null dir");
    int32_t rc;
    if ((rc = set_device(s)) != MX_OK) return rc;
    std::string tmp = path + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return fail(s, MX_ERR_SAVE, "open %s: %s", tmp.c_str(), strerror(errno));
    FileHeader h{};
    memcpy(h.magic, kFileMagic, 8);
    h.dim = s->cfg.dim;
    h.dtype = s->cfg.dtype;
    h.metric = s->cfg.metric;
    h.ld = s->ld;
    h.n = s->n;
    bool ok = fwrite(&h, sizeof h, 1, f) == 1;
    const size_t row_bytes = (size_t)s->ld * s->elem;
    const uint64_t chunk_rows = std::max<uint64_t>(1, (64ull << 20) / row_bytes);
    if (ok && (rc = ensure_pinned(s, chunk_rows * row_bytes)) != MX_OK) {
        fclose(f);
        return rc;
    }
    for (uint64_t done = 0; ok && done < s->n;) {
        const uint64_t c = std::min(chunk_rows, s->n - done);
        cudaError_t e = cudaMemcpyAsync(s->pinned, (const char *)s->rows + done * row_bytes, c * row_bytes,
                                        cudaMemcpyDeviceToHost, s->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        if (e != cudaSuccess) {
            fclose(f);
            unlink(tmp.c_str());
            return fail(s, MX_ERR_SAVE, "device read failed: %s", cudaGetErrorString(e));
        }
        ok = fwrite(s->pinned, row_bytes, c, f) == c;
        done += c;
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) {
        unlink(tmp.c_str());
        return fail(s, MX_ERR_SAVE, "write %s: %s", path.c_str(), strerror(errno));
    }
    return MX_OK;
}

int32_t mx_store_load(const char *dir, int32_t device, mx_store **out)
{
    if (!out) return fail(nullptr, MX_ERR_INVALID, "null argument");
    *out = nullptr;
    std::string path;
    if (!dir_join(dir, path)) return fail(nullptr, MX_ERR_INVALID, "null dir");
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return fail(nullptr, MX_ERR_FILE_IO, "open %s: %s", path.c_str(), strerror(errno));
    FileHeader h{};
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, kFileMagic, 8) != 0) {
        fclose(f);
        return fail(nullptr, MX_ERR_SERDE, "%s: bad header", path.c_str());
    }
    mx_store_cfg cfg{};
    cfg.dim = h.dim;
    cfg.dtype = h.dtype;
    cfg.metric = h.metric;
    cfg.device = device;
    cfg.capacity = h.n + h.n / 8 + 1024;
    mx_store *s = nullptr;
    int32_t rc = mx_store_create(&cfg, &s);
    if (rc != MX_OK) {
        fclose(f);
        return rc;
    }
    auto bail = [&](int32_t code, const char *msg) {
        int32_t r = fail(nullptr, code, "%s: %s", path.c_str(), msg);
        fclose(f);
        mx_store_destroy(s);
        return r;
    };
    if (h.ld != s->ld) return bail(MX_ERR_SERDE, "row stride mismatch");
    // raw stored rows -> widen to f32 on the device -> normal ingest (rounding is idempotent)
    const size_t row_bytes = (size_t)s->ld * s->elem;
    const uint64_t chunk_rows = std::max<uint64_t>(1, (32ull << 20) / row_bytes);
    void *raw_dev = nullptr;
    float *f32_dev = nullptr;
    if (cudaMalloc(&raw_dev, chunk_rows * row_bytes) != cudaSuccess ||
        cudaMalloc(&f32_dev, chunk_rows * (size_t)h.dim * 4) != cudaSuccess) {
        cudaFree(raw_dev);
        return bail(MX_ERR_CONNECTION, "cudaMalloc failed");
    }
    rc = ensure_pinned(s, chunk_rows * row_bytes);
    if (rc != MX_OK) g_last_error = s->last_error;   // the handle is destroyed below: keep its message
    for (uint64_t done = 0; rc == MX_OK && done < h.n;) {
        const uint64_t c = std::min(chunk_rows, h.n - done);
        if (fread(s->pinned, row_bytes, c, f) != c) {
            rc = fail(nullptr, MX_ERR_SERDE, "%s: truncated file", path.c_str());
            break;
        }
        cudaMemcpyAsync(raw_dev, s->pinned, c * row_bytes, cudaMemcpyHostToDevice, s->stream);
        launch_export_rows(raw_dev, s->cfg.dtype, s->ld, s->cfg.dim, 0, c, f32_dev, s->stream);
        rc = ingest_device(s, f32_dev, c, s->stream);
        if (rc == MX_OK) rc = finish_add(s, c, s->stream, nullptr);
        if (rc != MX_OK) g_last_error = s->last_error;
        done += c;
    }
    cudaFree(raw_dev);
    cudaFree(f32_dev);
    fclose(f);
    if (rc != MX_OK) {
        mx_store_destroy(s);
        return rc;
    }
    *out = s;
    return MX_OK;
}

}  // extern "C"
