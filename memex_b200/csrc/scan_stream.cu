// scan_stream.cu -- K1: HBM-streaming dot-product scan with in-kernel top-k (CUDA cores).
//
// Replaces the per-candidate `DistCosine::eval` calls of hnsw_rs' beam search that memex
// reaches through `self.hnsw.search(vec, limit, 16 * 2)` (reference storage/local.rs:76) with an
// exhaustive scan of the flat [N, d] row matrix.  This stage only has to produce a candidate
// SUPERSET ranked by an fp32 approximation of the cosine; rerank.cu re-scores the survivors with
// the reference's exact arithmetic, so ordering and scores are decided there.
//
// Roofline: HBM.  Algorithmic bytes per launch = N * ld * sizeof(T) (+ 4 N for inv_norm).
// Layout: each group of LPR lanes owns one row per slot and reads it as 16-byte chunks,
// lane-interleaved (a warp-wide load touches contiguous 512 B); R slots are unrolled so that
// R * CPL 128-bit loads are in flight per lane before the first FMA.  Partial sums are reduced
// with a transposing butterfly (R values over LPR lanes in ~R + log2(LPR) shuffles), then one
// lane per row offers (score, row) to a warp-resident top-k list; lists are merged per CTA and
// written as [query][cta][LCAP] candidates.
#include "common.cuh"
#include "exact.cuh"
#include "scan.cuh"

namespace mx {

template <typename T>
struct Chunk;
template <>
struct Chunk<float> {
    static constexpr int kElems = 4;
    __device__ static __forceinline__ float dot(const float4 &v, const float *q)
    {
        const float4 qv = *reinterpret_cast<const float4 *>(q);
        return fmaf(v.x, qv.x, fmaf(v.y, qv.y, fmaf(v.z, qv.z, v.w * qv.w)));
    }
};
template <>
struct Chunk<__half> {
    static constexpr int kElems = 8;
    __device__ static __forceinline__ float dot(const float4 &v, const float *q)
    {
        const float4 q0 = *reinterpret_cast<const float4 *>(q);
        const float4 q1 = *reinterpret_cast<const float4 *>(q + 4);
        const __half2 *h = reinterpret_cast<const __half2 *>(&v);
        const float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
        const float2 c = __half22float2(h[2]), d = __half22float2(h[3]);
        float s = a.x * q0.x;
        s = fmaf(a.y, q0.y, s);
        s = fmaf(b.x, q0.z, s);
        s = fmaf(b.y, q0.w, s);
        s = fmaf(c.x, q1.x, s);
        s = fmaf(c.y, q1.y, s);
        s = fmaf(d.x, q1.z, s);
        s = fmaf(d.y, q1.w, s);
        return s;
    }
};

constexpr int kScanThreads = 512;
constexpr int kScanWarps = kScanThreads / 32;

// T: stored element type; LPR: lanes per row; CPL: 16-byte chunks per lane (0 = runtime loop);
// R: row slots in flight per lane group; NQ: queries per pass; E: list entries per lane.
// EXACT (the fallback of the superset certificate, NQ = 1): the query is flagged query number `slot`
// (x.active_map[slot]); a row whose f32 score reaches x.fb_thr[slot] is re-scored with the reference's f64 fold right
// here, by the lane that speaks for it, and offered to the list with score = -key.  The lists then rank by the EXACT key
// (ties -> lower row), so their union holds the exact top 32 E whatever the data looks like.
struct ExactExtra {
    const uint32_t *active_map;
    const float *fb_thr;
    uint32_t dim, metric;
};

template <typename T, int LPR, int CPL, int R, int NQ, int E, bool EXACT>
__device__ __forceinline__ void scan_stream_body(const ScanParams &p, const ExactExtra &x, uint32_t slot)
{
    static_assert(LPR >= R && (LPR & (LPR - 1)) == 0 && (R & (R - 1)) == 0, "bad LPR / R");
    static_assert(!EXACT || NQ == 1, "the exact scan takes one query per pass");
    constexpr int kGroups = 32 / LPR;           // rows per slot per warp
    constexpr int kRowsPerIter = R * kGroups;   // rows per warp iteration
    constexpr int kCE = Chunk<T>::kElems;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *qs = reinterpret_cast<float *>(smem_raw);  // [NQ][ldq]

    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t q0 = EXACT ? x.active_map[slot] : blockIdx.y * NQ;
    const uint32_t nq_here = EXACT ? 1u : min((uint32_t)NQ, p.nq - q0);
    const float exact_thr = EXACT ? x.fb_thr[slot] : 0.f;

    for (uint32_t i = threadIdx.x; i < NQ * p.ldq; i += blockDim.x) {
        const uint32_t qi = i / p.ldq;
        qs[i] = qi < nq_here ? p.queries[(size_t)(q0 + qi) * p.ldq + (i - qi * p.ldq)] : 0.f;
    }
    __syncthreads();

    WarpTopK<E> top[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) top[qi].init();

    const uint32_t group = lane / LPR;       // which row of the slot this lane works on
    const uint32_t gl = lane % LPR;          // lane within the group
    const uint32_t V = p.ld / kCE;           // 16-byte chunks per row
    const float4 *rows = reinterpret_cast<const float4 *>(p.rows);
    const uint32_t total_warps = gridDim.x * kScanWarps;
    const uint32_t gw = blockIdx.x * kScanWarps + warp;
    const uint32_t last_row = p.n_rows - 1;

    // row index (within the R slots) this lane ends up holding after the transposing reduce
    uint32_t my_slot = 0;
    {
        int cnt = R;
#pragma unroll
        for (int m = LPR / 2; m >= 1; m >>= 1) {
            if (cnt > 1) {
                cnt >>= 1;
                if (gl & m) my_slot += cnt;
            }
        }
    }

    for (uint64_t base = (uint64_t)gw * kRowsPerIter; base < p.n_rows;
         base += (uint64_t)total_warps * kRowsPerIter) {
        float acc[NQ][R];
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi)
#pragma unroll
            for (int r = 0; r < R; ++r) acc[qi][r] = 0.f;

        if constexpr (CPL > 0) {
            float4 v[R][CPL];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t row = min((uint32_t)(base + r * kGroups + group), last_row);
                const float4 *rp = rows + (size_t)row * V + gl;
#pragma unroll
                for (int c = 0; c < CPL; ++c) v[r][c] = ldg_stream_f4(rp + c * LPR);
            }
#pragma unroll
            for (int c = 0; c < CPL; ++c)
#pragma unroll
                for (int qi = 0; qi < NQ; ++qi) {
                    const float *qp = qs + qi * p.ldq + (gl + c * LPR) * kCE;
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[qi][r] += Chunk<T>::dot(v[r][c], qp);
                }
        } else {
            for (uint32_t c = gl; c < V; c += LPR) {
                float4 v[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t row = min((uint32_t)(base + r * kGroups + group), last_row);
                    v[r] = ldg_stream_f4(rows + (size_t)row * V + c);
                }
#pragma unroll
                for (int qi = 0; qi < NQ; ++qi) {
                    const float *qp = qs + qi * p.ldq + c * kCE;
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[qi][r] += Chunk<T>::dot(v[r], qp);
                }
            }
        }

        // transposing butterfly inside each LPR-lane group: R partials -> 1 complete sum per lane
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
            int cnt = R;
#pragma unroll
            for (int m = LPR / 2; m >= 1; m >>= 1) {
                if (cnt > 1) {
                    const int half = cnt >> 1;
                    const bool upper = (gl & m) != 0;
#pragma unroll
                    for (int i = 0; i < half; ++i) {
                        const float send = upper ? acc[qi][i] : acc[qi][i + half];
                        const float keep = upper ? acc[qi][i + half] : acc[qi][i];
                        acc[qi][i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
                    }
                    cnt = half;
                } else {
                    acc[qi][0] += __shfl_xor_sync(0xffffffffu, acc[qi][0], m);
                }
            }
        }
        // lanes sharing a slot hold the same sum; the first of them speaks for the row.
        // Which lanes share a slot: those equal on the halving masks, i.e. differing only in the
        // low log2(LPR / R) bits of gl.
        constexpr int kDup = (LPR >= R) ? (LPR / R) : 1;
        const bool speaker = (gl % kDup) == 0;
        const uint64_t row64 = base + (uint64_t)my_slot * kGroups + group;
        const bool valid = speaker && row64 < p.n_rows;
        const uint32_t row = (uint32_t)row64;
        const float inv = (valid && p.use_inv) ? __ldg(p.inv_norm + row) : 1.f;
        if constexpr (EXACT) {
            const bool pass = valid && acc[0][0] * inv >= exact_thr;
            float v = kNegInf;
            if (pass)
                v = -exact_key(qs, p.rows, sizeof(T) == 4 ? MX_DTYPE_F32 : MX_DTYPE_F16, x.metric, (size_t)row * p.ld, x.dim,
                               nullptr);
            top[0].offer(pass, v, row);
        } else {
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi) top[qi].offer(valid, acc[qi][0] * inv, row);
        }
    }

    // ---- per-CTA merge of the warp lists, then one list per (query, cta) to global ----
    __syncthreads();
    float *ls = reinterpret_cast<float *>(smem_raw);                    // [warps][32*E]
    uint32_t *lr = reinterpret_cast<uint32_t *>(ls + kScanWarps * 32 * E);
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            ls[(warp * E + e) * 32 + lane] = top[qi].s[e];
            lr[(warp * E + e) * 32 + lane] = top[qi].r[e];
        }
        __syncthreads();
        if (warp == 0) {
            WarpTopK<E> fin = top[qi];
            for (int w = 1; w < kScanWarps; ++w)
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const float v = ls[(w * E + e) * 32 + lane];
                    const uint32_t r = lr[(w * E + e) * 32 + lane];
                    fin.offer(r != kNoRow, v, r);
                }
            if ((uint32_t)qi < nq_here) {
                const size_t o = ((size_t)(EXACT ? slot : q0 + qi) * p.n_lists + blockIdx.x) * (32 * E);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    p.cand_s[o + e * 32 + lane] = fin.s[e];
                    p.cand_r[o + e * 32 + lane] = fin.r[e];
                }
            }
        }
        __syncthreads();
    }
}

template <typename T, int LPR, int CPL, int R, int NQ, int E>
__global__ void __launch_bounds__(kScanThreads, 1) scan_stream_kernel(ScanParams p)
{
    scan_stream_body<T, LPR, CPL, R, NQ, E, false>(p, ExactExtra{}, 0);
}

// one pass over the corpus per flagged query; normally there is none and the kernel returns at once
template <typename T, int LPR, int CPL, int R, int E>
__global__ void __launch_bounds__(kScanThreads, 1) scan_exact_kernel(ScanParams p, ExactExtra x, const uint32_t *active_n)
{
    pdl_trigger();
    pdl_wait();   // the rerank's certificate counts the flagged queries
    const uint32_t n = *active_n;
    for (uint32_t slot = 0; slot < n; ++slot) scan_stream_body<T, LPR, CPL, R, 1, E, true>(p, x, slot);
}

template <typename T, int LPR, int CPL, int R, int NQ, int E>
static cudaError_t launch_one(const ScanParams &p, uint32_t n_ctas, cudaStream_t st)
{
    auto kern = scan_stream_kernel<T, LPR, CPL, R, NQ, E>;
    size_t smem = std::max((size_t)NQ * p.ldq * sizeof(float), (size_t)kScanWarps * 32 * E * 8);
    if (smem > 48 * 1024) {
        cudaError_t e = set_max_smem(kern, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(n_ctas, ceil_div<uint32_t>(p.nq, NQ));
    kern<<<grid, kScanThreads, smem, st>>>(p);
    count_launch();
    return cudaGetLastError();
}

template <typename T, int NQ, int E>
static cudaError_t launch_shape(const ScanParams &p, uint32_t n_ctas, cudaStream_t st)
{
    const uint32_t V = p.ld / Chunk<T>::kElems;
    if (V == 96) return launch_one<T, 32, 3, (NQ == 1 ? 8 : 4), NQ, E>(p, n_ctas, st);   // f32 d=384 / f16 d=768
    if (V == 48) return launch_one<T, 16, 3, (NQ == 1 ? 8 : 4), NQ, E>(p, n_ctas, st);   // f16 d=384
    if (V == 192) return launch_one<T, 32, 6, 4, NQ, E>(p, n_ctas, st);                  // f32 d=768
    if (V >= 32) return launch_one<T, 32, 0, 4, NQ, E>(p, n_ctas, st);
    if (V >= 8) return launch_one<T, 8, 0, 4, NQ, E>(p, n_ctas, st);
    return launch_one<T, 1, 0, 1, NQ, E>(p, n_ctas, st);
}

// entries-per-lane for the lists this path produces (LCAP = 32 * E)
uint32_t scan_stream_lcap(uint32_t k) { return 32u * (k + 8 <= 32 ? 1 : k + 8 <= 64 ? 2 : k + 8 <= 128 ? 4 : 8); }

cudaError_t launch_scan_stream(const ScanParams &p, uint32_t dtype, uint32_t k, uint32_t n_ctas, cudaStream_t st)
{
    const uint32_t E = scan_stream_lcap(k) / 32;
    const bool multi = p.nq > 1;
#define MX_DISPATCH(T)                                                                      \
    switch (E) {                                                                            \
        case 1: return multi ? launch_shape<T, 4, 1>(p, n_ctas, st) : launch_shape<T, 1, 1>(p, n_ctas, st); \
        case 2: return multi ? launch_shape<T, 4, 2>(p, n_ctas, st) : launch_shape<T, 1, 2>(p, n_ctas, st); \
        case 4: return launch_shape<T, 1, 4>(p, n_ctas, st);                                \
        default: return launch_shape<T, 1, 8>(p, n_ctas, st);                               \
    }
    if (dtype == MX_DTYPE_F32) {
        MX_DISPATCH(float)
    } else {
        MX_DISPATCH(__half)
    }
#undef MX_DISPATCH
}

template <typename T, int LPR, int CPL, int R, int E>
static cudaError_t launch_exact_one(const ExactScanParams &p, uint32_t n_ctas, cudaStream_t st)
{
    auto kern = scan_exact_kernel<T, LPR, CPL, R, E>;
    size_t smem = std::max((size_t)p.scan.ldq * sizeof(float), (size_t)kScanWarps * 32 * E * 8);
    if (smem > 48 * 1024) {
        cudaError_t e = set_max_smem(kern, (int)smem);
        if (e != cudaSuccess) return e;
    }
    LaunchAttrs attrs;
    attrs.pdl();
    cudaError_t e = launch_ex(kern, dim3(n_ctas), dim3(kScanThreads), smem, st, attrs, p.scan,
                              ExactExtra{p.active_map, p.fb_thr, p.dim, p.metric}, p.active_n);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

template <typename T, int E>
static cudaError_t launch_exact_shape(const ExactScanParams &p, uint32_t n_ctas, cudaStream_t st)
{
    const uint32_t V = p.scan.ld / Chunk<T>::kElems;
    if (V == 96) return launch_exact_one<T, 32, 3, 8, E>(p, n_ctas, st);
    if (V == 48) return launch_exact_one<T, 16, 3, 8, E>(p, n_ctas, st);
    if (V >= 32) return launch_exact_one<T, 32, 0, 4, E>(p, n_ctas, st);
    if (V >= 8) return launch_exact_one<T, 8, 0, 4, E>(p, n_ctas, st);
    return launch_exact_one<T, 1, 0, 1, E>(p, n_ctas, st);
}

cudaError_t launch_scan_exact(const ExactScanParams &p, uint32_t dtype, uint32_t k, uint32_t n_ctas, cudaStream_t st)
{
    const uint32_t E = scan_stream_lcap(k) / 32;
#define MX_DISPATCH_EXACT(T)                                      \
    switch (E) {                                                  \
        case 1: return launch_exact_shape<T, 1>(p, n_ctas, st);   \
        case 2: return launch_exact_shape<T, 2>(p, n_ctas, st);   \
        case 4: return launch_exact_shape<T, 4>(p, n_ctas, st);   \
        default: return launch_exact_shape<T, 8>(p, n_ctas, st);  \
    }
    if (dtype == MX_DTYPE_F32) {
        MX_DISPATCH_EXACT(float)
    } else {
        MX_DISPATCH_EXACT(__half)
    }
#undef MX_DISPATCH_EXACT
}

// queries-per-pass the dispatch above uses for (nq, k): lets the caller size grid.y / candidates
uint32_t scan_stream_nq_per_pass(uint32_t nq, uint32_t k)
{
    const uint32_t E = scan_stream_lcap(k) / 32;
    return (nq > 1 && E <= 2) ? 4u : 1u;
}

}  // namespace mx
