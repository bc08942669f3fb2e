// rerank.cu -- candidate merge + exact re-scoring, shard merge, ingest and small data movers.
//
// The scan kernels rank rows by an fp32 / tensor-core approximation of the score.  This file
// turns their per-CTA candidate lists into the final answer with the REFERENCE'S arithmetic:
//
//   hnsw_rs 0.1.20 DistCosine::eval  (called under reference storage/local.rs:76):
//       products a*b, a*a, b*b in f32, widened to f64, folded left to right in f64;
//       d = max(0, 1 - ab / sqrt(aa * bb)) as f32; either norm zero -> d = 0
//   reference storage/local.rs:86:   similarity = 1.0 - (1.0 / (1.0 / d))   in f32
//   reference storage/local.rs:63:   ids are 1-based in insertion order
//
// Every operation is the IEEE round-to-nearest one (explicit __fmul_rn / __dadd_rn / __fdiv_rn,
// double sqrt and division are correctly rounded on the device), evaluated in the same order,
// so scores are bit-identical to oracle/cosine_oracle.c and ordering is (d asc, id asc).
#include "common.cuh"
#include "exact.cuh"
#include "scan.cuh"

namespace mx {

__device__ __forceinline__ float key_to_score(float key, uint32_t metric)
{
    if (metric == MX_METRIC_DOT) return -key;
    return __fsub_rn(1.0f, __fdiv_rn(1.0f, __fdiv_rn(1.0f, key)));  // local.rs:86
}

// ---- warp-wide sorted lists of 32 (score, row) pairs, best first (cand_before order) ------------------
// bitonic sort across the lanes of a warp: lane 0 ends up with the best entry
__device__ __forceinline__ void warp_sort32(float &v, uint32_t &r)
{
    const uint32_t lane = lane_id();
#pragma unroll
    for (uint32_t k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, j);
            const uint32_t orow = __shfl_xor_sync(0xffffffffu, r, j);
            const bool take_better = ((lane & k) == 0) == ((lane & j) == 0);
            const bool mine_better = cand_before(v, r, ov, orow);
            if (take_better != mine_better && !(v == ov && r == orow)) {
                v = ov;
                r = orow;
            }
        }
    }
}
// (v, r) sorted across lanes and (bv, br) sorted across lanes -> the 32 best of the union, sorted
__device__ __forceinline__ void warp_merge32(float &v, uint32_t &r, float bv, uint32_t br)
{
    const uint32_t lane = lane_id();
    // best(A[i], B[31 - i]) is the top half of the union as a bitonic sequence
    const float rv = __shfl_sync(0xffffffffu, bv, 31 - lane);
    const uint32_t rr = __shfl_sync(0xffffffffu, br, 31 - lane);
    if (cand_before(rv, rr, v, r)) {
        v = rv;
        r = rr;
    }
#pragma unroll
    for (uint32_t j = 16; j > 0; j >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, j);
        const uint32_t orow = __shfl_xor_sync(0xffffffffu, r, j);
        const bool take_better = (lane & j) == 0;
        const bool mine_better = cand_before(v, r, ov, orow);
        if (take_better != mine_better && !(v == ov && r == orow)) {
            v = ov;
            r = orow;
        }
    }
}

__device__ __forceinline__ void prof_mark(const RerankParams &p, int slot)
{
    if (p.prof && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        p.prof[slot] = t;
    }
}

constexpr int kRerankThreads = 512;
constexpr int kRerankWarps = kRerankThreads / 32;
constexpr int kFoldChunk = 384;           // elements of a row staged per pass
constexpr int kFoldPitch = kFoldChunk + 1;  // odd pitch: lane j reading row j is bank-conflict free
constexpr int kFoldBatches = 2;           // 32-entry batches folded concurrently (one warp each)

// One CTA per query:
//   1. every warp folds a slice of the scan's candidate lists into a warp-resident top-(32 E) list;
//   2. the 16 warp lists are merged by counting ranks (one thread per entry) -> the 32 E best rows;
//   3. the lowest-id zero-norm rows are appended (cosine: d = 0 for them whatever the query);
//   4. exact keys: candidate rows are staged in shared memory with coalesced loads, then ONE THREAD
//      per candidate folds its row left to right in f64 -- the reference's order of operations;
//   5. duplicates dropped, entries rank-sorted by (key asc, id asc), best k written.
template <int E>
__global__ void __launch_bounds__(kRerankThreads) rerank_kernel(RerankParams p)
{
    constexpr int kList = 32 * E;
    constexpr int kMaxEntries = kList + (int)MX_MAX_K;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *qs = reinterpret_cast<float *>(smem_raw);                 // [kFoldChunk] query chunk
    float *ws = qs + kFoldChunk;                                     // [warps][kList] scores
    uint32_t *wr = reinterpret_cast<uint32_t *>(ws + kRerankWarps * kList);
    float *ekey = reinterpret_cast<float *>(wr + kRerankWarps * kList);  // [kMaxEntries]
    uint32_t *erow = reinterpret_cast<uint32_t *>(ekey + kMaxEntries);
    float *rowbuf = reinterpret_cast<float *>(erow + kMaxEntries);   // [kFoldBatches * 32][kFoldPitch]
    __shared__ uint32_t n_entries_s;
    __shared__ int zero_query_s;
    __shared__ float sel_v_s, kth_key_s, red_s[kRerankWarps];
    __shared__ uint32_t sel_r_s;

    pdl_trigger();
    pdl_wait();   // the candidate lists (and, second pass, the flagged-query count) are the previous kernel's output
    // second pass (exact fallback): CTA b answers flagged query active_map[b] from candidate lists b
    const bool second_pass = p.active_n != nullptr;
    if (second_pass && blockIdx.x >= *p.active_n) return;
    const uint32_t q = second_pass ? p.active_map[blockIdx.x] : blockIdx.x;
    const bool certify = !second_pass && p.n_flagged != nullptr;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        sel_r_s = kNoRow;
        sel_v_s = kNegInf;
        kth_key_s = 0.f;
    }
    const float *qg = p.queries + (size_t)q * p.ldq;
    prof_mark(p, 0);

    for (uint32_t j = threadIdx.x; j < kMaxEntries; j += blockDim.x) {
        ekey[j] = 0.f;
        erow[j] = kNoRow;
    }
    const uint32_t total = p.n_lists * p.lcap;   // the (query, list) candidate lists are contiguous
    const size_t cbase = (size_t)blockIdx.x * total;
    // certificate, part 1 (see below): the best APPROXIMATE score among everything that was rejected before the exact
    // re-scoring -- by a scan CTA (a full list's minimum bounds whatever that CTA turned away, and the global threshold tau
    // is the maximum of such minima), by the seeding floor, or by the top-(32 E) cut
    float a_rej = kNegInf;
    if (certify && p.n_flagged_next && blockIdx.x == 0 && threadIdx.x == 0) *p.n_flagged_next = 0;
    constexpr int kRegSlots = 10;   // candidates per thread held in registers by the E == 1 fast path
    bool reg_path = false;
    if constexpr (E == 1) reg_path = total <= (uint32_t)kRegSlots * kRerankThreads && (p.lcap == 16 || p.lcap == 32);
    if (E == 1 && reg_path) {
        // 1+2 (k <= 24), ONE pass over global memory: every candidate of the query goes to a register first (all loads
        // in flight together), the certificate's per-list facts come from segmented warp reductions of those registers
        // (a list is 16 or 32 consecutive entries = half a warp-load or a whole one), then sorted-list algebra: each
        // warp sorts 32 candidates at a time and merges them into its running sorted top-32, a tree merges the 16 warps.
        float cv[kRegSlots];
        uint32_t cr[kRegSlots];
#pragma unroll
        for (int i = 0; i < kRegSlots; ++i) {
            const uint32_t idx = (uint32_t)i * kRerankThreads + threadIdx.x;
            const bool in = idx < total;
            cv[i] = in ? p.cand_s[cbase + idx] : kNegInf;
            cr[i] = in ? p.cand_r[cbase + idx] : kNoRow;
        }
        prof_mark(p, 1);   // (issue only: the loads complete at their first use)
        if (certify) {
            const uint32_t segmask = p.lcap == 32 ? 0xffffffffu : (0xffffu << (lane & 16));
#pragma unroll
            for (int i = 0; i < kRegSlots; ++i) {
                if ((uint32_t)i * kRerankThreads + warp * 32 < total) {            // warp-uniform
                    const uint32_t have = __ballot_sync(0xffffffffu, cr[i] != kNoRow);
                    float mn = cv[i];
                    if (p.lcap == 32) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, 16));
#pragma unroll
                    for (int o = 8; o >= 1; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    if ((have & segmask) == segmask) a_rej = fmaxf(a_rej, mn);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < kRegSlots; ++i)
            if (cr[i] == kNoRow) cv[i] = kNegInf;
        float bv = kNegInf;
        uint32_t br = kNoRow;
#pragma unroll
        for (int i = 0; i < kRegSlots; ++i) {
            if ((uint32_t)i * kRerankThreads + warp * 32 < total) {               // warp-uniform
                float v = cv[i];
                uint32_t r = cr[i];
                warp_sort32(v, r);
                if (i == 0) {
                    bv = v;
                    br = r;
                } else {
                    warp_merge32(bv, br, v, r);
                }
            }
        }
        prof_mark(p, 2);
        ws[warp * 32 + lane] = bv;
        wr[warp * 32 + lane] = br;
        // tree merge of the 16 warp lists (the top 32 of a union does not depend on the merge order)
        for (int half = kRerankWarps / 2; half >= 1; half >>= 1) {
            __syncthreads();
            if ((int)warp < half) {
                warp_merge32(bv, br, ws[(warp + half) * 32 + lane], wr[(warp + half) * 32 + lane]);
                if (half > 1) {
                    ws[warp * 32 + lane] = bv;
                    wr[warp * 32 + lane] = br;
                }
            }
        }
        if (warp == 0) {
            erow[lane] = br;
            if (lane == 31) {   // the worst entry that made the cut (kNoRow: nothing was cut)
                sel_v_s = bv;
                sel_r_s = br;
            }
        }
        if (certify) {
            __syncthreads();
            const float sv = sel_v_s;
            const uint32_t sr = sel_r_s;
            if (sr != kNoRow) {
#pragma unroll
                for (int i = 0; i < kRegSlots; ++i)
                    if (cr[i] != kNoRow && cand_before(sv, sr, cv[i], cr[i])) a_rej = fmaxf(a_rej, cv[i]);
            }
        }
    } else if constexpr (E == 1) {
        // generic form of the same (more candidates than the register path holds): each warp sorts 32 candidates at a
        // time and merges them into its running sorted top-32; warp 0 then merges the 16 warp lists.
        float bv = kNegInf;
        uint32_t br = kNoRow;
        for (uint32_t i = warp * 32; i < total; i += kRerankWarps * 32) {
            const uint32_t idx = i + lane;
            float v = idx < total ? p.cand_s[cbase + idx] : kNegInf;
            uint32_t r = idx < total ? p.cand_r[cbase + idx] : kNoRow;
            if (r == kNoRow) v = kNegInf;
            warp_sort32(v, r);
            if (i == warp * 32) {
                bv = v;
                br = r;
            } else {
                warp_merge32(bv, br, v, r);
            }
        }
        ws[warp * 32 + lane] = bv;
        wr[warp * 32 + lane] = br;
        for (int half = kRerankWarps / 2; half >= 1; half >>= 1) {
            __syncthreads();
            if ((int)warp < half) {
                warp_merge32(bv, br, ws[(warp + half) * 32 + lane], wr[(warp + half) * 32 + lane]);
                if (half > 1) {
                    ws[warp * 32 + lane] = bv;
                    wr[warp * 32 + lane] = br;
                }
            }
        }
        if (warp == 0) {
            erow[lane] = br;
            if (lane == 31) {
                sel_v_s = bv;
                sel_r_s = br;
            }
        }
    } else {
        // 1. every warp folds a slice of the candidate lists into its own list
        WarpTopK<E> top;
        top.init();
        for (uint32_t i = warp * 32; i < total; i += kRerankWarps * 32) {
            const uint32_t idx = i + lane;
            const bool in = idx < total;
            const float v = in ? p.cand_s[cbase + idx] : kNegInf;
            const uint32_t r = in ? p.cand_r[cbase + idx] : kNoRow;
            top.offer(r != kNoRow, v, r);
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            ws[warp * kList + e * 32 + lane] = top.s[e];
            wr[warp * kList + e * 32 + lane] = top.r[e];
        }
        __syncthreads();
        // 2. rank of every entry among all warp lists; rows are unique across lists (a row is scanned by
        //    one CTA), so ranks of real entries are unique and the kList best land in erow[0 .. kList)
        for (uint32_t j = threadIdx.x; j < kRerankWarps * kList; j += blockDim.x) {
            const float v = ws[j];
            const uint32_t r = wr[j];
            if (r == kNoRow) continue;
            uint32_t rank = 0;
            for (uint32_t i = 0; i < kRerankWarps * kList; ++i) rank += cand_before(ws[i], wr[i], v, r) ? 1u : 0u;
            if (rank < kList) erow[rank] = r;
            if (rank == kList - 1) {
                sel_v_s = v;
                sel_r_s = r;
            }
        }
    }
    if (certify) {
        __syncthreads();
        if (!(E == 1 && reg_path)) {
            for (uint32_t c = threadIdx.x; c < p.n_lists; c += blockDim.x) {
                float mn = __int_as_float(0x7f800000);
                bool full = true;
                for (uint32_t e = 0; e < p.lcap; ++e) {
                    full &= p.cand_r[cbase + (size_t)c * p.lcap + e] != kNoRow;
                    mn = fminf(mn, p.cand_s[cbase + (size_t)c * p.lcap + e]);
                }
                if (full) a_rej = fmaxf(a_rej, mn);
            }
            const float sv = sel_v_s;
            const uint32_t sr = sel_r_s;
            if (sr != kNoRow)
                for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
                    const float v = p.cand_s[cbase + i];
                    const uint32_t r = p.cand_r[cbase + i];
                    if (r != kNoRow && cand_before(sv, sr, v, r)) a_rej = fmaxf(a_rej, v);
                }
        }
        if (threadIdx.x == 0 && p.scan_floor) a_rej = fmaxf(a_rej, p.scan_floor[q]);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) a_rej = fmaxf(a_rej, __shfl_xor_sync(0xffffffffu, a_rej, o));
        if (lane == 0) red_s[warp] = a_rej;
        __syncthreads();
        a_rej = red_s[0];
        for (int w = 1; w < kRerankWarps; ++w) a_rej = fmaxf(a_rej, red_s[w]);
    }
    prof_mark(p, 3);
    // 3. the lowest-id zero-norm rows (cosine: d = 0)
    if (threadIdx.x == 0) {
        uint32_t n = kList;
        if (p.metric == MX_METRIC_COSINE) n += min(min(*p.n_zero, p.k), (uint32_t)MX_MAX_K);
        n_entries_s = n;
    }
    __syncthreads();
    const uint32_t n_entries = n_entries_s;
    for (uint32_t i = threadIdx.x; i + kList < n_entries; i += blockDim.x) erow[kList + i] = p.zero_rows[i];
    __syncthreads();

    // 4. exact key per entry.  Batches of 32 entries; the rows of kFoldBatches batches are staged per pass
    //    (chunks of kFoldChunk elements), then lane j folds entry 32 * (b0 + b) + j left to right in f64.  The three
    //    sums of DistCosine are independent chains, so they go to different warps (a f32 -> f64 conversion is an XU op
    //    and a dependent DADD is ~10 cycles: one thread running all three took 12 of the kernel's 28 us):
    //    warps [0, F) fold a.b, warps [F, 2F) the row's b.b, warp 2F the query's a.a (once).  Same order of
    //    operations per sum as the reference, hence the same bits.
    __shared__ double bb_s[kFoldBatches * 32];
    __shared__ double aa_s;
    prof_mark(p, 4);
    const uint32_t n_batches = (n_entries + 31) / 32;
    const uint32_t frole = warp / kFoldBatches, fb = warp % kFoldBatches;   // role 0: a.b, 1: b.b, warp 2F: a.a
    // entry e of a pass -> rowbuf[e][0 .. cn) as f32.  16-byte loads, all of a row's loads issued before the first use
    // (a warp owns entries e = warp, warp + 16, ...); the query chunk goes to qs
    auto stage = [&](uint32_t first, uint32_t n_here, uint32_t c0, uint32_t cn) {
        __syncthreads();   // previous chunk fully consumed
        for (uint32_t i = threadIdx.x; i < cn; i += blockDim.x) qs[i] = qg[c0 + i];
        const uint32_t per16 = p.dtype == MX_DTYPE_F32 ? 4u : 8u;        // elements per 16-byte chunk
        const uint32_t n16 = (cn + per16 - 1) / per16;                   // c0 and ld are multiples of per16
        for (uint32_t e = warp; e < n_here; e += kRerankWarps) {
            const uint32_t row = erow[first + e];
            if (row == kNoRow || row >= p.n_rows) continue;
            const uint4 *src = reinterpret_cast<const uint4 *>(
                reinterpret_cast<const char *>(p.rows) + ((size_t)row * p.ld + c0) * (p.dtype == MX_DTYPE_F32 ? 4 : 2));
            uint4 buf[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const uint32_t ch = lane + 32 * u;
                buf[u] = ch < n16 ? src[ch] : make_uint4(0, 0, 0, 0);
            }
            float *dst = rowbuf + e * kFoldPitch;
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const uint32_t ch = lane + 32 * u;
                if (ch >= n16) continue;
                if (p.dtype == MX_DTYPE_F32) {
                    const float *f = reinterpret_cast<const float *>(&buf[u]);
#pragma unroll
                    for (int x = 0; x < 4; ++x)
                        if (ch * 4 + x < cn) dst[ch * 4 + x] = f[x];
                } else {
                    const __half *hh = reinterpret_cast<const __half *>(&buf[u]);
#pragma unroll
                    for (int x = 0; x < 8; ++x)
                        if (ch * 8 + x < cn) dst[ch * 8 + x] = __half2float(hh[x]);
                }
            }
        }
        __syncthreads();
    };
    // (r2, measured: a "fast fold" that summed an entry's products with a warp-parallel tree whenever the sum is provably
    // exact -- all non-zero products within 29 - log2(dim) binades, so that no f64 addition can round in ANY order -- was
    // bit-identical but SLOWER, 13.4 us against 7.0: the fold is bound by the FP64 pipe's THROUGHPUT, ~16 cycles per
    // warp-wide DADD and SM sub-partition, not by the chain's latency, and the tree does the same number of additions plus
    // the bookkeeping.  What helps is spreading the three chains over three sub-partitions: a.b on warp 0, b.b on warp
    // 2, a.a on warp 5 -- warp 4 would share warp 0's sub-partition and double the critical path.)
    __shared__ int aa_known_s;
    if (threadIdx.x == 0) aa_known_s = 0;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < n_batches; b0 += kFoldBatches) {
        const uint32_t first = b0 * 32;
        const uint32_t n_here = min((uint32_t)kFoldBatches * 32, n_entries - first);
        double acc = 0.0;
        const bool folder = frole < 2 && fb * 32 + lane < n_here;
        const bool q_folder = warp == 2 * kFoldBatches + 1 && !aa_known_s;
        const uint32_t my_row = folder ? erow[first + fb * 32 + lane] : kNoRow;
        for (uint32_t c0 = 0; c0 < p.dim; c0 += kFoldChunk) {
            const uint32_t cn = min((uint32_t)kFoldChunk, p.dim - c0);
            stage(first, n_here, c0, cn);
            if (folder && my_row != kNoRow && my_row < p.n_rows) {
                const float *rb = rowbuf + (fb * 32 + lane) * kFoldPitch;
                // 16 products and their f32 -> f64 conversions are computed ahead of the chain, so the chain itself is 16
                // dependent DADDs
                const float *xa = frole == 0 ? qs : rb;
                uint32_t i = 0;
                for (; i + 16 <= cn; i += 16) {
                    double t[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) t[j] = (double)__fmul_rn(xa[i + j], rb[i + j]);
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc = __dadd_rn(acc, t[j]);
                }
                for (; i < cn; ++i) acc = __dadd_rn(acc, (double)__fmul_rn(xa[i], rb[i]));
            } else if (q_folder) {
                uint32_t i = 0;
                for (; i + 16 <= cn; i += 16) {
                    double t[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) t[j] = (double)__fmul_rn(qs[i + j], qs[i + j]);
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc = __dadd_rn(acc, t[j]);
                }
                for (; i < cn; ++i) acc = __dadd_rn(acc, (double)__fmul_rn(qs[i], qs[i]));
            }
        }
        if (frole == 1 && folder) bb_s[fb * 32 + lane] = acc;
        if (q_folder && lane == 0) aa_s = acc;
        __syncthreads();
        if (q_folder && lane == 0) aa_known_s = 1;
        if (frole == 0 && folder) {
            const uint32_t j = first + fb * 32 + lane;
            if (my_row != kNoRow && my_row < p.n_rows) {
                const double ab = acc, aa = aa_s, bb = bb_s[fb * 32 + lane];
                float key;
                if (p.metric == MX_METRIC_DOT)
                    key = -(float)ab;
                else if (aa > 0.0 && bb > 0.0)
                    key = (float)fmax(__dsub_rn(1.0, __ddiv_rn(ab, __dsqrt_rn(__dmul_rn(aa, bb)))), 0.0);
                else
                    key = 0.f;
                ekey[j] = key;
            } else {
                erow[j] = kNoRow;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // the query's own norm decides the degenerate case (DistCosine: aa == 0 -> d = 0 for all rows).
        // thread 0 folded entry 0 when that entry was real; otherwise (no candidates at all) fold here.
        double a2 = n_batches > 0 ? aa_s : 0.0;   // folded by warp 2F during the first pass
        if (n_batches == 0) {
            for (uint32_t i = 0; i < p.dim; ++i) a2 = __dadd_rn(a2, (double)__fmul_rn(qg[i], qg[i]));
        }
        zero_query_s = (p.metric == MX_METRIC_COSINE && !(a2 > 0.0)) ? 1 : 0;
    }
    __syncthreads();

    prof_mark(p, 5);
    const uint32_t count = min(p.k, p.n_rows);
    if (zero_query_s) {
        // all distances are 0 -> (d asc, id asc) is simply the first rows
        for (uint32_t j = threadIdx.x; j < p.k; j += blockDim.x) {
            const bool ok = j < count;
            p.ids_out[(size_t)q * p.k + j] = ok ? p.id_offset + (uint64_t)j * p.id_stride + 1 : 0;
            p.scores_out[(size_t)q * p.k + j] = ok ? key_to_score(0.f, p.metric) : 0.f;
            if (p.dists_out) p.dists_out[(size_t)q * p.k + j] = 0.f;
        }
        if (threadIdx.x == 0) {
            p.counts_out[q] = count;
            if (certify && p.stats) atomicAdd(p.stats, 1ull);
        }
        return;
    }

    // 5. drop duplicates (a zero row can also arrive through the scan), then rank-sort
    for (uint32_t j = threadIdx.x; j < n_entries; j += blockDim.x) {
        const uint32_t row = erow[j];
        bool dup = false;
        if (row != kNoRow)
            for (uint32_t i = 0; i < j; ++i) dup |= (erow[i] == row);
        if (dup) ekey[j] = __int_as_float(0x7fc00000);  // mark; row cleared after the barrier
    }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < n_entries; j += blockDim.x)
        if (ekey[j] != ekey[j]) erow[j] = kNoRow;
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < p.k; j += blockDim.x) {
        if (j >= count) {
            p.ids_out[(size_t)q * p.k + j] = 0;
            p.scores_out[(size_t)q * p.k + j] = 0.f;
            if (p.dists_out) p.dists_out[(size_t)q * p.k + j] = 0.f;
        }
    }
    for (uint32_t j = threadIdx.x; j < n_entries; j += blockDim.x) {
        const uint32_t row = erow[j];
        if (row == kNoRow) continue;
        const float key = ekey[j];
        uint32_t rank = 0;
        for (uint32_t i = 0; i < n_entries; ++i) {
            const uint32_t ri = erow[i];
            if (ri == kNoRow) continue;
            const float ki = ekey[i];
            rank += (ki < key || (ki == key && ri < row)) ? 1u : 0u;
        }
        if (rank < count) {
            p.ids_out[(size_t)q * p.k + rank] = p.id_offset + (uint64_t)row * p.id_stride + 1;
            p.scores_out[(size_t)q * p.k + rank] = key_to_score(key, p.metric);
            if (p.dists_out) p.dists_out[(size_t)q * p.k + rank] = key;
            if (rank + 1 == count) kth_key_s = key;
        }
    }
    if (threadIdx.x == 0) p.counts_out[q] = count;
    prof_mark(p, 6);
    if (!certify) return;
    // certificate, part 2.  In the exact score's units (cosine: cos; dot: q . c) the approximate score of a row is within
    //   eps = [ |q16 - q/|q||_2  (tcgen05 scan: fp16 query, Cauchy-Schwarz) + (dim + 64) 2^-23  (f32 / tensor-core
    //           accumulation, inverse norm, the f32 products of the exact fold) + 4e-6 ]  x  (dot: |q| max|row|)
    // of its exact score, so no rejected row can reach the k-th exact score X when  a_rej + eps < X.
    __syncthreads();
    if (threadIdx.x == 0) {
        if (p.stats) atomicAdd(p.stats, 1ull);
        const double qn = sqrt(aa_s);
        if (count > 0 && a_rej > kNegInf && qn > 0.0) {
            const float rel1 = (float)(p.dim + 64) * 1.1920929e-7f + 4e-6f;
            const float rel = rel1 + (p.qerr ? p.qerr[q] : 0.f);
            const double mn = p.metric == MX_METRIC_DOT ? (double)*p.max_norm : 1.0;
            double a_eu, eps, eps1, X;
            if (p.metric == MX_METRIC_COSINE) {
                a_eu = p.unit_queries ? (double)a_rej : (double)a_rej / qn;
                eps = rel;
                eps1 = rel1;
                X = 1.0 - (double)kth_key_s;
            } else {
                a_eu = p.unit_queries ? (double)a_rej * qn : (double)a_rej;
                eps = (double)rel * qn * mn;
                eps1 = (double)rel1 * qn * mn;
                X = -(double)kth_key_s;
            }
            if (!(a_eu + eps < X)) {
                // threshold of the exact scan's f32 pre-filter, in ITS units (cosine: q . c / |c|; dot: q . c)
                double t1 = p.metric == MX_METRIC_COSINE ? (X - eps1) * qn : X - eps1;
                t1 -= fabs(t1) * 1e-6 + 1e-30;
                const uint32_t pos = atomicAdd(p.n_flagged, 1u);
                p.q_map[pos] = q;
                p.fb_thr[pos] = __double2float_rd(t1);
                if (p.stats) atomicAdd(p.stats + 1, 1ull);
            }
        }
    }
}

cudaError_t launch_rerank(const RerankParams &p, cudaStream_t st)
{
    // the rerank's own list holds the 32 * E best approximate candidates, E chosen from k as the stream scan does
    const uint32_t lk = scan_stream_lcap(p.k);
    const uint32_t E = lk / 32;
    const size_t smem = (size_t)kFoldChunk * 4 + (size_t)kRerankWarps * lk * 8 + (size_t)(lk + MX_MAX_K) * 8 +
                        (size_t)kFoldBatches * 32 * kFoldPitch * 4;
#define MX_RR(EE)                                                                                  \
    {                                                                                              \
        auto kern = rerank_kernel<EE>;                                                             \
        if (smem > 48 * 1024) {                                                                    \
            cudaError_t e = set_max_smem(kern, (int)smem); \
            if (e != cudaSuccess) return e;                                                        \
        }                                                                                          \
        LaunchAttrs attrs;                                                                         \
        attrs.pdl();                                                                               \
        cudaError_t le = launch_ex(kern, dim3(p.nq), dim3(kRerankThreads), smem, st, attrs, p);    \
        if (le != cudaSuccess) return le;                                                          \
    }
    switch (E) {
        case 1: MX_RR(1) break;
        case 2: MX_RR(2) break;
        case 4: MX_RR(4) break;
        case 8: MX_RR(8) break;
        default: return cudaErrorInvalidValue;
    }
#undef MX_RR
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K3 tail: merge of G shard answers after the all-gather.  One CTA per query, rank-sort of
// G*k entries by (key asc, id asc).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) merge_kernel(MergeParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t T = p.n_shards * p.k;
    float *key = reinterpret_cast<float *>(smem_raw);
    uint64_t *id = reinterpret_cast<uint64_t *>(smem_raw + (((size_t)T * 4 + 7) & ~(size_t)7));
    __shared__ uint32_t total_s;
    const uint32_t q = blockIdx.x;
    if (threadIdx.x == 0) total_s = 0;
    pdl_trigger();
    pdl_wait();
    if (p.wait_flags != nullptr && threadIdx.x < p.n_shards) {
        // peer-memory exchange: shard g's answer was stored into this GPU's memory by rank g, followed by a release store of
        // the epoch into flag g.  Bounded wait: a lost peer must surface as a launch failure, not as a hung GPU.
        uint32_t seen, spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.wait_flags + threadIdx.x) : "memory");
            if (++spins > (1u << 27)) {
                printf("memex_b200: merge wait for shard %u timed out (flag %u, epoch %u)\n", threadIdx.x, seen, p.wait_epoch);
                __trap();
            }
        } while ((int32_t)(seen - p.wait_epoch) < 0);
    }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < T; j += blockDim.x) {
        const uint32_t g = j / p.k, e = j - g * p.k;
        const uint32_t *cg = reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(p.counts) + g * p.stride_counts);
        const float *dg = reinterpret_cast<const float *>(reinterpret_cast<const char *>(p.dists) + g * p.stride_dists);
        const uint64_t *ig = reinterpret_cast<const uint64_t *>(reinterpret_cast<const char *>(p.ids) + g * p.stride_ids);
        // .cg loads: the blobs may have been written by another GPU since this SM last read these addresses
        const bool ok = e < __ldcg(cg + q);
        key[j] = ok ? __ldcg(dg + (size_t)q * p.k + e) : 0.f;
        id[j] = ok ? __ldcg(reinterpret_cast<const unsigned long long *>(ig) + (size_t)q * p.k + e) : 0;
        if (ok) atomicAdd(&total_s, 1u);
    }
    __syncthreads();
    const uint32_t count = min(p.k, total_s);
    for (uint32_t j = threadIdx.x; j < p.k; j += blockDim.x)
        if (j >= count) {
            p.ids_out[(size_t)q * p.k + j] = 0;
            p.scores_out[(size_t)q * p.k + j] = 0.f;
        }
    for (uint32_t j = threadIdx.x; j < T; j += blockDim.x) {
        const uint64_t mine = id[j];
        if (mine == 0) continue;
        const float kj = key[j];
        uint32_t rank = 0;
        for (uint32_t i = 0; i < T; ++i) {
            const uint64_t oi = id[i];
            if (oi == 0) continue;
            const float ki = key[i];
            rank += (ki < kj || (ki == kj && oi < mine)) ? 1u : 0u;
        }
        if (rank < count) {
            p.ids_out[(size_t)q * p.k + rank] = mine;
            p.scores_out[(size_t)q * p.k + rank] = key_to_score(kj, p.metric);
        }
    }
    if (threadIdx.x == 0) p.counts_out[q] = count;
}

// One CTA per peer: 16-byte stores of this rank's blob into the peer's slot, a system-scope fence, then the flag.
__global__ void __launch_bounds__(256) exchange_push_kernel(PushParams p)
{
    const uint32_t peer = blockIdx.x;
    pdl_trigger();
    pdl_wait();
    const uint4 *src = reinterpret_cast<const uint4 *>(p.blob);
    uint4 *dst = reinterpret_cast<uint4 *>(p.peer_base[peer] + p.slot_offset);
    for (uint64_t i = threadIdx.x; i < p.blob_bytes / 16; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t *flag = reinterpret_cast<uint32_t *>(p.peer_base[peer] + p.flag_offset) + p.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(p.epoch) : "memory");
    }
}

cudaError_t launch_exchange_push(const PushParams &p, cudaStream_t st)
{
    if (p.world == 0 || p.world > (uint32_t)kMaxPeers || p.blob_bytes % 16 != 0) return cudaErrorInvalidValue;
    LaunchAttrs attrs;
    attrs.pdl();
    cudaError_t e = launch_ex(exchange_push_kernel, dim3(p.world), dim3(256), (size_t)0, st, attrs, p);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_merge(const MergeParams &p, cudaStream_t st)
{
    const size_t T = (size_t)p.n_shards * p.k;
    const size_t smem = ((T * 4 + 7) & ~(size_t)7) + T * 8;
    if (smem > 48 * 1024) {
        cudaError_t e = set_max_smem(merge_kernel, (int)smem);
        if (e != cudaSuccess) return e;
    }
    LaunchAttrs attrs;
    attrs.pdl();
    cudaError_t e = launch_ex(merge_kernel, dim3(p.nq), dim3(256), smem, st, attrs, p);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// ingest: one warp per row.  f32 source -> stored dtype (round to nearest even, as numpy's
// astype(float16)), inv_norm from the STORED values, flags for zero-norm / non-finite rows.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ingest_kernel(IngestParams p)
{
    const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (row >= p.n) return;
    const uint32_t lane = lane_id();
    const float *src = p.src + row * p.dim;
    const uint64_t dst_row = p.first_row + row;
    float ss = 0.f;
    bool bad = false;
    for (uint32_t c = lane; c < p.ld; c += 32) {
        float v = c < p.dim ? src[c] : 0.f;
        if (p.dtype == MX_DTYPE_F16) {
            const __half h = __float2half_rn(v);
            reinterpret_cast<__half *>(p.rows)[dst_row * p.ld + c] = h;
            v = __half2float(h);
        } else {
            reinterpret_cast<float *>(p.rows)[dst_row * p.ld + c] = v;
        }
        bad |= !isfinite(v);
        ss = fmaf(v, v, ss);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    bad = __any_sync(0xffffffffu, bad) || !isfinite(ss);
    if (lane == 0) {
        if (bad) atomicOr(p.bad_flag, 1u);
        const bool zero = !(ss > 0.f);
        p.inv_norm[dst_row] = zero ? 0.f : 1.0f / sqrtf(ss);
        if (p.max_norm && !bad) {
            // running maximum of the row norms (rounded up): the dot metric's approximation radius scales with it
            const float nrm = sqrtf(ss) * 1.0001f;
            if (nrm > *reinterpret_cast<volatile float *>(p.max_norm))
                atomicMax(reinterpret_cast<unsigned int *>(p.max_norm), __float_as_uint(nrm));
        }
        if (zero && p.metric == MX_METRIC_COSINE) atomicOr(p.bad_flag, 2u);
    }
}

cudaError_t launch_ingest(const IngestParams &p, cudaStream_t st)
{
    if (p.n == 0) return cudaSuccess;
    const uint32_t rows_per_cta = 256 / 32;
    ingest_kernel<<<(unsigned)ceil_div<uint64_t>(p.n, rows_per_cta), 256, 0, st>>>(p);
    count_launch();
    return cudaGetLastError();
}

// the k lowest zero-norm rows matter for cosine (d = 0 -> score 1.0): ordered compaction of the
// first MX_MAX_K rows with inv_norm == 0, one CTA, early exit.
__global__ void __launch_bounds__(1024) collect_zero_rows_kernel(const float *inv_norm, uint64_t n,
                                                                 uint32_t *zero_rows, uint32_t *n_zero)
{
    __shared__ uint32_t warp_cnt[32];
    __shared__ uint32_t found_s;
    if (threadIdx.x == 0) found_s = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n; base += blockDim.x) {
        const uint64_t i = base + threadIdx.x;
        const bool z = i < n && inv_norm[i] == 0.f;
        const uint32_t m = __ballot_sync(0xffffffffu, z);
        const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        uint32_t before = found_s;
        for (uint32_t w = 0; w < warp; ++w) before += warp_cnt[w];
        const uint32_t pos = before + __popc(m & ((1u << lane) - 1));
        if (z && pos < MX_MAX_K) zero_rows[pos] = (uint32_t)i;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = found_s;
            for (uint32_t w = 0; w < blockDim.x / 32; ++w) t += warp_cnt[w];
            found_s = t;
        }
        __syncthreads();
        if (found_s >= MX_MAX_K) break;
    }
    if (threadIdx.x == 0) *n_zero = min(found_s, (uint32_t)MX_MAX_K);
}

cudaError_t launch_collect_zero_rows(const float *inv_norm, uint64_t n, uint32_t *zero_rows, uint32_t *n_zero,
                                     cudaStream_t st)
{
    collect_zero_rows_kernel<<<1, 1024, 0, st>>>(inv_norm, n, zero_rows, n_zero);
    count_launch();
    return cudaGetLastError();
}

__global__ void export_rows_kernel(const void *rows, uint32_t dtype, uint32_t ld, uint32_t dim, uint64_t first_row,
                                   uint64_t n, float *out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * dim) return;
    const uint64_t r = i / dim;
    const uint32_t c = (uint32_t)(i - r * dim);
    out[i] = load_elem(rows, dtype, (size_t)(first_row + r) * ld + c);
}

cudaError_t launch_export_rows(const void *rows, uint32_t dtype, uint32_t ld, uint32_t dim, uint64_t first_row,
                               uint64_t n, float *out, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    export_rows_kernel<<<(unsigned)ceil_div<uint64_t>(n * dim, 256), 256, 0, st>>>(rows, dtype, ld, dim, first_row, n,
                                                                                   out);
    count_launch();
    return cudaGetLastError();
}

__global__ void stage_queries_kernel(const float *q, uint32_t nq, uint32_t dim, uint32_t ldq, float *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * ldq) return;
    const uint32_t r = i / ldq, c = i - r * ldq;
    out[i] = c < dim ? q[(size_t)r * dim + c] : 0.f;
}

cudaError_t launch_stage_queries(const float *q, uint32_t nq, uint32_t dim, uint32_t ldq, float *out, cudaStream_t st)
{
    stage_queries_kernel<<<ceil_div<uint32_t>(nq * ldq, 256), 256, 0, st>>>(q, nq, dim, ldq, out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mx
