// scan_tc.cu -- K2: batched cosine / dot-product scan of an fp16 corpus on tcgen05 tensor cores.
//
// Replaces, for a BATCH of queries, the per-candidate `DistCosine::eval` calls hnsw_rs makes under
// `self.hnsw.search(vec, limit, 16 * 2)` (reference lib/libmemex/src/storage/local.rs:76) with one
// exhaustive pass over the flat [N, d] fp16 row matrix:
//
//     S[q, n] = Q16[q, :] . C[n, :]  (* inv_norm[n] for cosine)        q < 128 per pass
//
// The scan is HBM-bound only if every corpus byte is read ONCE for all queries of the batch, which
// at 64-128 queries needs ~400-800 TFLOP/s -- tensor cores -- and the [q, N] score matrix must never
// leave the SM.  So: the queries sit in shared memory as the M = 128 operand for the whole kernel,
// 128-row corpus tiles stream through a TMA ring as the N operand, scores accumulate in TMEM
// (4 accumulator stages of 128 columns), and the epilogue warps (one THREAD per query = one TMEM
// lane) filter the 128 scores of each tile against a running threshold and keep an L-entry
// candidate list per (query, CTA) in shared memory.  rerank.cu merges the lists and re-scores the
// survivors with the reference's exact f64 arithmetic, so this stage only has to deliver a superset.
//
// Threshold = max(local list minimum [strict >], global per-query lower bound tau[q] [>=]).
// tau[q] is published with an atomic max by any CTA whose list for q is full: L rows with
// score >= tau exist somewhere, so a row scoring below tau cannot be in the top L <= needed k.
// It keeps the number of list insertions per thread ~L*ln(N/L)/CTAs instead of ~L*ln(N/CTAs/L).
//
// Threshold seeding.  An insertion is the slow path (the warp diverges and rescans an L-entry list), and a cold list
// takes ~L ln(n / L) of them per thread: ~100 us per launch at L = 16 whatever the shard size, 330 us at L = 32
// (measured, scripts/diag_scan_fixed.py).  So every CTA first scans its first P tiles keeping only the best score per
// query (no list), folds it into the maximum of its group with one atomic, and after ONE grid-wide barrier (cooperative launch) every thread takes tau0 = the
// smallest of L group maxima (CTA c in group c % L): L different CTAs hold a row scoring >= tau0, so tau0 is a valid
// lower bound for the top L, and it sits at the ~0.07 % quantile instead of -inf.  The P tiles are then scanned
// again, normally, RIGHT AFTER the barrier (they are still in L2: the sampling pass loads them with the normal eviction
// priority, everything after it evict-first), so a CTA visits its rows in increasing order and the strict `>` admission
// of the list keeps the LOWER row of two equal scores -- the (distance, id) order of the reference's result.
//
// Roofline: HBM.  Algorithmic bytes per launch = N * ld * 2 (+ 4 N inv_norm).
// Warp roles (192 threads): warp 0 TMA producer, warp 1 tcgen05.mma issuer + TMEM owner,
// warps 2..5 epilogue (warp % 4 = TMEM lane quarter).
#include <cstdlib>

#include "common.cuh"
#include "scan.cuh"
#include "tc.cuh"

namespace mx {

using namespace tc;

namespace {

constexpr int kTcThreads = 192;
constexpr int kQM = 128;            // list stride; queries per pass = UMMA M = 128 (dim <= 384) or 64 (dim <= 768)
constexpr int kTileN = 128;         // corpus rows per tile = UMMA N
constexpr int kBK = 64;             // fp16 elements per k-block (one 128-byte swizzle atom)
constexpr int kMaxKB = 12;          // dim <= 768 (with M = 64: the query block must stay within 96 KB of smem)
constexpr int kAccStages = 4;       // 4 x 128 TMEM columns
constexpr int kInvSlots = 24;        // inverse-norm tiles in flight: the producer runs up to kMaxStages + kAccStages tiles ahead (k_blocks = 1)
constexpr int kKBBytes = kTileN * kBK * 2;  // 16 KB: one k-block of a corpus tile (a k-block of Q is QM x 128 bytes)

constexpr int kSampLd = 32;         // group maxima per query of the threshold seeding (>= the longest list)
constexpr int kMaxStages = 12;      // the corpus ring takes whatever shared memory the query block and the lists leave

template <int L>
struct TcCfg {
    static constexpr int kListBytes = L * kQM * 8;
    static constexpr int kInvBytes = kInvSlots * kTileN * 4;
    static constexpr int kBarBytes = 512;
    static constexpr int fixed_bytes(int kb, int qm) { return kb * qm * kBK * 2 + kListBytes + kInvBytes + kBarBytes + 1024; }
    // bytes in flight are what hides the HBM latency: 6 stages (96 KB) next to a 128-query block, 9 (144 KB) next to 64 queries
    static constexpr int stages(int kb, int qm)
    {
        const int s = (227 * 1024 - fixed_bytes(kb, qm)) / kKBBytes;
        return s > kMaxStages ? kMaxStages : s;
    }
    static constexpr int smem_bytes(int kb, int qm, int n_stages) { return fixed_bytes(kb, qm) + n_stages * kKBBytes; }
};
static_assert((2 * kMaxStages + 2 * kAccStages + kInvSlots + 1) * 8 + 8 <= 512, "barrier block");

__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ float ld_relaxed(const float *p)
{
    float v;
    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// monotone max on a float cell initialised to -inf
__device__ __forceinline__ void atomic_max_float(float *addr, float v)
{
    if (v >= 0.f)
        atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
    else
        atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

__device__ __forceinline__ void tc_prof_mark(unsigned long long *prof, int slot)
{
    if (prof && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 64) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        prof[slot] = t;
    }
}

struct TcParams {
    const float *inv_norm;  // [>= n_tiles * 128] (cosine) or nullptr (dot)
    float *tau;             // [nq_pad] global lower bounds, -inf on entry
    float *cand_s;
    uint32_t *cand_r;
    uint32_t n_rows, nq, n_lists, k_blocks;
    uint32_t stages;        // depth of the corpus ring (<= kMaxStages)
    // threshold seeding (see the kernel): the first `sample_tiles` tiles of every CTA are scanned for their best score
    // only; samp [nq_pad][kSampLd] collects the maxima of the CTA groups (-inf on entry, like tau), sync is the grid-wide
    // arrival counter
    uint32_t sample_tiles, nq_pad;
    float *samp;
    uint32_t *sync;
    float *floor_out;       // [nq_pad] tau0 of each query (-inf without seeding), read by the rerank's certificate
    // query preparation in the prologue (no separate launch, no fp16 staging buffer): queries f32 [nq, ldq], zero padded
    const float *queries;
    uint32_t ldq, dim;
    float *qerr;            // [nq_pad] fp16 rounding radius of each prepared query (written by CTA x = 0)
    uint32_t *done;         // exit ticket: the LAST CTA out resets tau / sync / done for the next launch
    uint32_t plain_barrier; // host-side only: launch the seeded form without the cooperative attribute (exclusive SM partition)
    unsigned long long *prof;   // test-only (MX_SCAN_TC_PROF=1): %globaltimer of CTA 0's epilogue thread 64 at phase boundaries, [32]
    uint32_t diag;          // TIMING / POWER DIAGNOSTIC ONLY (MX_SCAN_TC_DIAG, wrong results): bit 0 = skip the query preparation,
                            // and, in a -DMX_TC_DIAG build only: bit 1 = no tcgen05.mma (barriers only), bit 2 = no epilogue work
};

// Query preparation by the four epilogue warps of a CTA (warp `quarter` takes rows quarter, quarter + 4, ...): R rows in
// flight per warp (all their loads are issued before the first reduction), U 32-lane rounds of 16-byte chunk pairs per row.
// ~7 us per launch at 64 x 384, and not for want of loads in flight: 148 CTAs pull the SAME 98 KB out of L2 at the same
// moment.  Measured r2 (profiles/r02b_query_staging_ab.txt): ONE cp.async.bulk of the raw block per CTA into the ring's upper
// stages + conversion from shared memory took 7.9 us against 7.3 -- the broadcast is what costs, whoever issues it; only
// fewer bytes per SM (a pre-converted fp16 block, cluster multicast) would shorten it.
template <int QM, int R, int U>
__device__ __forceinline__ void prepare_queries(const TcParams &p, unsigned char *sq, uint32_t q0, uint32_t quarter, uint32_t lane)
{
    constexpr uint32_t kQKB = QM * kBK * 2;
    const uint32_t n_chunks = p.k_blocks * 8;            // 8-element (16-byte) chunks per prepared row
    // every CTA of the grid reads the same block at the same moment: each starts at a different group of rows so that the
    // requests spread over the L2 slices instead of queueing on the same lines
    constexpr uint32_t kGroups = QM / (4 * R);
    for (uint32_t it = 0; it < kGroups; ++it) {
        const uint32_t rb = quarter + 4 * R * ((it + blockIdx.x) % kGroups);
        float4 va[R][U], vb[R][U];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const uint32_t gq = q0 + rb + 4 * j;
            const bool live = gq < p.nq;
            const float *src = p.queries + (size_t)gq * p.ldq;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t c = lane + 32 * u;
                va[j][u] = vb[j][u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live && c < n_chunks && c * 8 < p.ldq) {     // ldq is a multiple of 8, zero beyond dim
                    va[j][u] = __ldg(reinterpret_cast<const float4 *>(src + c * 8));
                    vb[j][u] = __ldg(reinterpret_cast<const float4 *>(src + c * 8 + 4));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const uint32_t r = rb + 4 * j;
            float ss = 0.f;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                ss = fmaf(va[j][u].x, va[j][u].x, fmaf(va[j][u].y, va[j][u].y, fmaf(va[j][u].z, va[j][u].z, fmaf(va[j][u].w, va[j][u].w, ss))));
                ss = fmaf(vb[j][u].x, vb[j][u].x, fmaf(vb[j][u].y, vb[j][u].y, fmaf(vb[j][u].z, vb[j][u].z, fmaf(vb[j][u].w, vb[j][u].w, ss))));
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            const float inv = ss > 0.f ? rsqrtf(ss) : 0.f;
            float ee = 0.f;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t c = lane + 32 * u;
                if (c >= n_chunks) continue;
                const float f[8] = {va[j][u].x * inv, va[j][u].y * inv, va[j][u].z * inv, va[j][u].w * inv,
                                    vb[j][u].x * inv, vb[j][u].y * inv, vb[j][u].z * inv, vb[j][u].w * inv};
                uint32_t w[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const __half2 h2 = __floats2half2_rn(f[2 * x], f[2 * x + 1]);
                    const float2 back = __half22float2(h2);
                    const float e0 = back.x - f[2 * x], e1 = back.y - f[2 * x + 1];
                    ee = fmaf(e0, e0, fmaf(e1, e1, ee));
                    w[x] = *reinterpret_cast<const uint32_t *>(&h2);
                }
                const uint32_t kb = c >> 3, cj = c & 7;
                *reinterpret_cast<uint4 *>(sq + kb * kQKB + r * 128 + ((cj ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            if (blockIdx.x == 0) {
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) ee += __shfl_xor_sync(0xffffffffu, ee, o);
                if (lane == 0) p.qerr[q0 + r] = sqrtf(ee) * 1.001f;
            }
        }
    }
}

// QM = 128: TMEM lane = query.  QM = 64 (cta_group::1, M = 64): accumulator row r sits in TMEM lane
// 32 * (r / 16) + r % 16, i.e. the first 16 lanes of each 32-lane quarter -- epilogue lanes 16..31 idle.
template <int L, bool USE_INV, int QM>
__global__ void __launch_bounds__(kTcThreads, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap tmC, TcParams p)
{
    using Cfg = TcCfg<L>;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment by OFFSET: the pointer stays in the shared address space (LDS/STS, not generic LD/ST)
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *sq = smem;                                          // [k_blocks][128 x 64] fp16, swizzled
    constexpr uint32_t kQKB = QM * kBK * 2;                            // bytes of one k-block of the query block
    unsigned char *ring = sq + p.k_blocks * kQKB;                      // [kStages][128 x 64]
    const uint32_t n_stages = p.stages;
    float *list_s = reinterpret_cast<float *>(ring + n_stages * kKBBytes);  // [L][128]
    uint32_t *list_r = reinterpret_cast<uint32_t *>(list_s + L * kQM);
    float *sinv = reinterpret_cast<float *>(list_r + L * kQM);         // [kInvSlots][128]
    uint64_t *full = reinterpret_cast<uint64_t *>(sinv + kInvSlots * kTileN);
    uint64_t *empty = full + kMaxStages;
    uint64_t *tmem_full = empty + kMaxStages;
    uint64_t *tmem_empty = tmem_full + kAccStages;
    uint64_t *inv_full = tmem_empty + kAccStages;
    uint64_t *q_full = inv_full + kInvSlots;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(q_full + 1);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n_tiles = (p.n_rows + kTileN - 1) / kTileN;
    const uint32_t q0 = blockIdx.y * QM;
    // this CTA's tiles: blockIdx.x + i * gridDim.x, i < n_local; with seeding the first P of them come twice:
    // positions [0, P) of the sequence are the sampling pass over local tiles 0 .. P-1, positions [P, P + n_local) the
    // real scan over local tiles 0 .. n_local-1 -- rows in increasing order
    const uint32_t n_local = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t n_sample = p.sample_tiles;
    const uint32_t n_seq = n_local + n_sample;
    auto tile_of = [&](uint32_t i) { return blockIdx.x + (i < n_sample ? i : i - n_sample) * gridDim.x; };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmC);
        for (uint32_t i = 0; i < n_stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < kAccStages; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);
        }
        for (int i = 0; i < kInvSlots; ++i) mbar_init(&inv_full[i], 1);
        mbar_init(q_full, 4);   // one arrive per epilogue warp once its share of the query block is in shared memory
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();   // launched as a programmatic dependent (plain-barrier form): the queries are the previous kernel's output

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;   // the corpus ring fills while the epilogue warps prepare the queries
            for (uint32_t local = 0; local < n_seq; ++local) {
                const uint32_t tile = tile_of(local);
                if (USE_INV) {
                    const uint32_t slot = local % kInvSlots;
                    mbar_arrive_expect_tx(&inv_full[slot], kTileN * 4);
                    bulk_load_1d(sinv + slot * kTileN, p.inv_norm + (size_t)tile * kTileN, kTileN * 4, &inv_full[slot]);
                }
                for (uint32_t kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], kKBBytes);
                    tma_load_2d(ring + stage * kKBBytes, &tmC, &full[stage], kb * kBK, tile * kTileN,
                                local < n_sample ? kEvictNormal : kEvictFirst);
                    if (++stage == n_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(QM, kTileN, 0 /* f16 */);
            mbar_wait(q_full, 0);
            tc_fence_after();
            const uint32_t sq_addr = smem_u32(sq);
            uint32_t stage = 0, phase = 0;
            // diagnostic switches 2 and 4 exist only in a -DMX_TC_DIAG build (scripts/power_probe.py): testing the kernel
            // parameter here put a constant-bank load + branch between the barrier wait and the first MMA of every k-block,
            // +6 % on the main loop of a 1.25 M-row shard (measured on one box with scripts/ab_builds.sh), and hoisting
            // the test does not help -- the compiler rematerialises it inside the loop
#ifdef MX_TC_DIAG
            const bool diag_no_mma = (p.diag & 2u) != 0;
#else
            constexpr bool diag_no_mma = false;
#endif
            for (uint32_t local = 0; local < n_seq; ++local) {
                const uint32_t as = local % kAccStages, aphase = (local / kAccStages) & 1;
                mbar_wait(&tmem_empty[as], aphase ^ 1);
                tc_fence_after();
                for (uint32_t kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = sq_addr + kb * kQKB;
                    const uint32_t sb = smem_u32(ring + stage * kKBBytes);
                    if (!diag_no_mma) {
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k)
                            umma(tmem_base + as * kTileN, make_smem_desc(sa + k * 32), make_smem_desc(sb + k * 32), idesc,
                                 (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == n_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tmem_full[as]);
            }
        }
    } else {
        // ================= epilogue: thread = query (TMEM lane), columns = corpus rows =================
        const uint32_t quarter = warp & 3;
        tc_prof_mark(p.prof, 0);
        // ---- query preparation (was a kernel of its own): rows q0 .. q0 + QM of the f32 query block -> unit norm ->
        // fp16 -> the K-major SWIZZLE_128B layout the UMMA descriptor expects ([k-block][QM rows][64 halfs], 16-byte chunk
        // index ^= row & 7), zero rows beyond nq.  Scaling a query by a positive constant changes neither ranking, and keeps
        // every fp16 component in [-1, 1].  qerr[row] = |q16 - q / |q||_2: by Cauchy-Schwarz the tensor-core score of ANY
        // unit row differs from the exact cosine by at most this -- the rerank's certificate uses it as the radius.
        if (!(p.diag & 1u)) {
            // 8 rows in flight when a row is at most 64 chunks (dim <= 512: two loads of two float4 per lane and row), else 4
#ifndef MX_TC_PREP4
            if (p.k_blocks <= 8)
                prepare_queries<QM, 8, 2>(p, sq, q0, quarter, lane);
            else
#endif
                prepare_queries<QM, 4, 3>(p, sq, q0, quarter, lane);
        }
        fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(q_full);
        tc_prof_mark(p.prof, 1);
        const uint32_t t = QM == 128 ? quarter * 32 + lane : quarter * 16 + (lane & 15);   // query row of this thread
        const bool q_ok = (QM == 128 || lane < 16) && q0 + t < p.nq;
        float *ls = list_s + quarter * 32 + lane;    // one list column per thread (idle lanes own an unused one)
        uint32_t *lr = list_r + quarter * 32 + lane;
#pragma unroll
        for (int e = 0; e < L; ++e) {
            ls[e * kQM] = kNegInf;
            lr[e * kQM] = kNoRow;
        }
        float lthr = kNegInf;     // minimum of the local list (-inf until it is full)
        uint32_t lpos = 0;        // where that minimum sits
        uint32_t filled = 0;
        const float *tau = p.tau + q0 + t;
        const float kPosInf = __int_as_float(0x7f800000);
        float g_floor = kNegInf;          // tau0 from the sampling pass
        if (n_sample == 0 && q_ok && blockIdx.x == 0) p.floor_out[q0 + t] = kNegInf;
        float b1 = kNegInf;               // best score of the sampling pass
#ifdef MX_TC_DIAG
        const int n_chunks_real = (p.diag & 4u) ? 0 : kTileN / 32;   // diagnostic switch (see the MMA issuer)
#else
        constexpr int n_chunks_real = kTileN / 32;
#endif
        for (uint32_t local = 0; local < n_seq; ++local) {
            const uint32_t tile = tile_of(local);
            const uint32_t as = local % kAccStages, aphase = (local / kAccStages) & 1;
            const uint32_t slot = local % kInvSlots, sphase = (local / kInvSlots) & 1;
            if (local < n_sample) {
                // ---- sampling pass: the best score of this query over the tile, nothing else ----
                if (USE_INV) mbar_wait(&inv_full[slot], sphase);
                mbar_wait(&tmem_full[as], aphase);
                tc_fence_after();
                const float *inv = sinv + slot * kTileN;
                const uint32_t row0 = tile * kTileN;
                const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + as * kTileN;
#pragma unroll 1
                for (int c = 0; c < kTileN / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float s = __uint_as_float(v[j]);
                        if (USE_INV) s *= inv[c * 32 + j];
                        if (row0 + c * 32 + j >= p.n_rows) s = kNegInf;
                        b1 = fmaxf(b1, s);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[as]);
                if (local + 1 == n_sample) {
                    tc_prof_mark(p.prof, 2);
                    // ---- publish, ONE grid-wide barrier (all CTAs are co-resident: cooperative launch), take tau0 ----
                    if (q_ok) atomic_max_float(p.samp + (size_t)(q0 + t) * kSampLd + blockIdx.x % L, b1);
                    __threadfence();
                    bar_sync(1, 128);
                    if (threadIdx.x == 64) {
                        atomicAdd(p.sync, 1u);
                        uint32_t spins = 0, seen = 0;
                        do {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.sync) : "memory");
                            if (++spins > (1u << 26)) {
                                printf("memex_b200: scan_tc grid barrier timed out (block %d, %u of %u)\n", blockIdx.x, seen, gridDim.x);
                                __trap();
                            }
                        } while (seen < gridDim.x);
                    }
                    bar_sync(1, 128);
                    tc_prof_mark(p.prof, 6);
                    if (q_ok && gridDim.x >= (uint32_t)L) {
                        // CTA c belongs to group c % L and has folded its best score into samp[query][group] with an atomic
                        // max; tau0 = the smallest of the L group maxima: L different CTAs each hold a row scoring >= tau0.
                        // One round trip of L / 4 float4 loads per thread.  (r2 history, scripts/scan_tc_prof.py: the exact
                        // (L/2)-th largest of the CTAs' second-best scores, kept as a sorted list per thread over a
                        // [query][CTA] matrix, bounded the same 0.07 % quantile (simulated) but took 8.9 us of dependent
                        // compares; reading the matrix branch-free still took 5 us -- every SM reads all of it.)
                        const float4 *sp = reinterpret_cast<const float4 *>(p.samp + (size_t)(q0 + t) * kSampLd);
                        float4 vv[L / 4];
#pragma unroll
                        for (int j = 0; j < L / 4; ++j) vv[j] = __ldcg(sp + j);
                        float lo = vv[0].x;
#pragma unroll
                        for (int j = 0; j < L / 4; ++j) lo = fminf(fminf(lo, vv[j].x), fminf(fminf(vv[j].y, vv[j].z), vv[j].w));
                        g_floor = lo;
                    }
                    if (q_ok && blockIdx.x == 0) p.floor_out[q0 + t] = g_floor;   // rerank's certificate needs every threshold
                    tc_prof_mark(p.prof, 3);
                }
                continue;
            }
            if (local == 2 * n_sample) tc_prof_mark(p.prof, 4);
#ifdef MX_TC_DIAG
            const bool fine = p.prof != nullptr && local - n_sample < 2u;   // diagnostic build: the first two real tiles in detail
#else
            constexpr bool fine = false;
#endif
            if (fine) tc_prof_mark(p.prof, 8 + 4 * (local - n_sample));
            float g = q_ok ? fmaxf(ld_relaxed(tau), g_floor) : kPosInf;
            if (fine) {
                if (g == 12345.678f) p.prof[31] = 1;   // the mark below must wait for the load
                tc_prof_mark(p.prof, 9 + 4 * (local - n_sample));
            }
            if (USE_INV) mbar_wait(&inv_full[slot], sphase);
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            if (fine) tc_prof_mark(p.prof, 10 + 4 * (local - n_sample));
            // s >= g  <=>  s > gm
            const float gm = (g == kNegInf) ? kNegInf : nextafterf(g, kNegInf);
            float thr = fmaxf(lthr, gm);
            const float *inv = sinv + slot * kTileN;
            const uint32_t row0 = tile * kTileN;
            const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + as * kTileN;
#pragma unroll 1
            for (int c = 0; c < n_chunks_real; ++c) {
                uint32_t v[32];
                tmem_ld32(taddr + c * 32, v);
                tmem_ld_wait();
                bool any = false;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    float4 iv = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (USE_INV) iv = *reinterpret_cast<const float4 *>(inv + c * 32 + j4 * 4);
                    const float s0 = __uint_as_float(v[4 * j4 + 0]) * iv.x;
                    const float s1 = __uint_as_float(v[4 * j4 + 1]) * iv.y;
                    const float s2 = __uint_as_float(v[4 * j4 + 2]) * iv.z;
                    const float s3 = __uint_as_float(v[4 * j4 + 3]) * iv.w;
                    v[4 * j4 + 0] = __float_as_uint(s0);
                    v[4 * j4 + 1] = __float_as_uint(s1);
                    v[4 * j4 + 2] = __float_as_uint(s2);
                    v[4 * j4 + 3] = __float_as_uint(s3);
                    any |= (s0 > thr) | (s1 > thr) | (s2 > thr) | (s3 > thr);
                }
                if (any && q_ok) {   // (idle lanes of the M = 64 layout read unowned TMEM lanes: never act on them)
                    // slow path (rare after warm-up): the scores go to a dynamically indexed local array so
                    // that the insertion code exists once instead of 32 times
                    uint32_t m = 0;
                    float sv[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        sv[j] = __uint_as_float(v[j]);
                        m |= (sv[j] > thr ? 1u : 0u) << j;
                    }
                    while (m) {
                        const int j = __ffs(m) - 1;
                        m &= m - 1;
                        const float s = sv[j];
                        const uint32_t row = row0 + c * 32 + j;
                        if (s > thr && row < p.n_rows) {
                            ls[lpos * kQM] = s;
                            lr[lpos * kQM] = row;
                            if (filled < L) ++filled;
#ifndef MX_TC_NO_APPEND
                            if (filled < L) {
                                // the list still has empty slots: append, nothing to evict, the threshold does not move.
                                // With a seeded threshold a shard of ~1 M rows leaves ~6 rows per (CTA, query) above it:
                                // its lists never fill, and EVERY insertion used to pay for the 16-entry scan below
                                lpos = filled;
                                continue;
                            }
#endif
                            // new eviction entry: minimal score, ties -> maximal row
                            float ms = ls[0];
                            uint32_t mr = lr[0], mp = 0;
#pragma unroll
                            for (int e = 1; e < L; ++e) {
                                const float se = ls[e * kQM];
                                const uint32_t re = lr[e * kQM];
                                if (se < ms || (se == ms && re > mr)) {
                                    ms = se;
                                    mr = re;
                                    mp = e;
                                }
                            }
                            lpos = mp;
                            lthr = ms;
                            if (lthr > g) {
                                atomic_max_float(p.tau + q0 + t, lthr);
                                g = lthr;
                            }
                            thr = fmaxf(lthr, gm);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
            if (fine) tc_prof_mark(p.prof, 11 + 4 * (local - n_sample));
        }
        tc_prof_mark(p.prof, 5);
        // the rerank is launched as a programmatic dependent: its CTAs may be scheduled (and run their prologue) as soon as
        // every CTA of this grid is past its last tile; its pdl_wait() still waits for this grid to complete
        if (threadIdx.x == 64) pdl_trigger();
        if (q_ok) {
            const size_t o = ((size_t)(q0 + t) * p.n_lists + blockIdx.x) * L;
#pragma unroll
            for (int e = 0; e < L; ++e) {
                p.cand_s[o + e] = ls[e * kQM];
                p.cand_r[o + e] = lr[e * kQM];
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    if (warp == 0) {
        // last CTA out leaves the cross-launch state as the next launch expects it: tau = -inf, counters zero (every other
        // CTA made its last tau / sync access before taking its ticket)
        uint32_t ticket = 0;
        if (lane == 0) {
            __threadfence();
            ticket = atomicAdd(p.done, 1u);
        }
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket == gridDim.x * gridDim.y - 1) {
            for (uint32_t i = lane; i < p.nq_pad; i += 32) p.tau[i] = kNegInf;
            if (n_sample > 0)
                for (uint32_t i = lane; i < p.nq_pad * kSampLd; i += 32) p.samp[i] = kNegInf;
            if (lane == 0) {
                *p.sync = 0;
                __threadfence();
                *p.done = 0;
            }
        }
    }
}

// cross-launch state of a freshly (re)allocated query capacity: tau = -inf, counters zero.  Every launch's last CTA puts
// it back, so this runs once per allocation, not once per search.
__global__ void tc_init_state_kernel(float *tau, float *floor_out, float *qerr, uint32_t n, uint32_t *sync, uint32_t *done,
                                     float *samp)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        for (uint32_t c = 0; c < (uint32_t)kSampLd; ++c) samp[(size_t)i * kSampLd + c] = kNegInf;
    if (i < n) {
        tau[i] = kNegInf;
        floor_out[i] = kNegInf;
        qerr[i] = 0.f;
    }
    if (i == 0) {
        *sync = 0;
        *done = 0;
    }
}

}  // namespace

struct TcScanState {
    int sm_count;
    int device_sms = 0;         // SMs of the whole device (a smaller sm_count = an SM budget, mx_store_set_sm_limit)
    uint32_t ld, dim, k_blocks;
    float *tau = nullptr;
    float *samp = nullptr;      // [q_cap][kSampLd] group maxima of the threshold seeding
    uint32_t *sync = nullptr;   // [0] grid barrier arrivals, [32] exit tickets (separate 128-byte lines)
    unsigned long long *prof = nullptr;   // MX_SCAN_TC_PROF=1: phase timestamps of CTA 0
    float *qerr = nullptr;      // [q_cap] fp16 rounding radius of each prepared query
    float *floor_out = nullptr; // [q_cap] tau0 of the last launch
    uint32_t q_cap = 0;  // rows
};

TcScanState *tc_scan_create(int sm_count, uint32_t ld, uint32_t dim)
{
    if (dim == 0 || dim > kMaxKB * kBK || ld % 8 != 0) return nullptr;   // > 768: the stream scan takes over
    if (!encode_tiled_fn()) return nullptr;
    TcScanState *t = new TcScanState();
    t->sm_count = sm_count;
    t->device_sms = sm_count;
    t->ld = ld;
    t->dim = dim;
    t->k_blocks = ceil_div<uint32_t>(dim, kBK);
    if (getenv("MX_SCAN_TC_PROF") && cudaMalloc(&t->prof, 256) == cudaSuccess) cudaMemset(t->prof, 0, 256);
    return t;
}

void tc_scan_destroy(TcScanState *t)
{
    if (!t) return;
    cudaFree(t->prof);
    cudaFree(t->tau);
    cudaFree(t->samp);
    cudaFree(t->sync);
    cudaFree(t->qerr);
    cudaFree(t->floor_out);
    delete t;
}

void tc_scan_invalidate(TcScanState *) {}  // tensor maps are rebuilt on every launch
void tc_scan_set_sms(TcScanState *t, int sm_count)
{
    if (t && sm_count > 0 && sm_count != t->sm_count) {
        t->sm_count = sm_count;
        t->q_cap = 0;   // the cross-launch state is re-made (and re-initialised) on the next launch
    }
}

uint32_t tc_scan_max_k() { return 26; }
bool tc_scan_supports(const TcScanState *t, uint32_t k) { return t != nullptr && k <= tc_scan_max_k(); }
uint32_t tc_scan_lcap(uint32_t k) { return k <= 10 ? 16u : 32u; }
const float *tc_scan_qerr(const TcScanState *t) { return t->qerr; }
const float *tc_scan_floor(const TcScanState *t) { return t->floor_out; }
const unsigned long long *tc_scan_prof(const TcScanState *t) { return t ? t->prof : nullptr; }
uint32_t tc_scan_lists(const TcScanState *t, uint64_t n_rows)
{
    return (uint32_t)std::min<uint64_t>((uint64_t)t->sm_count, ceil_div<uint64_t>(n_rows, kTileN));
}

template <int L, bool USE_INV, int QM>
static cudaError_t launch_tc_one(const CUtensorMap &tmC, TcParams tp, dim3 grid, cudaStream_t st)
{
    int n_stages = TcCfg<L>::stages((int)tp.k_blocks, QM);
    static const int cap = getenv("MX_SCAN_TC_STAGES") ? atoi(getenv("MX_SCAN_TC_STAGES")) : kMaxStages;   // measurement aid
    if (cap >= 2 && cap < n_stages) n_stages = cap;
    if (n_stages < 2) return cudaErrorInvalidValue;
    tp.stages = (uint32_t)n_stages;
    const int smem = TcCfg<L>::smem_bytes((int)tp.k_blocks, QM, n_stages);
    auto kern = scan_tc_kernel<L, USE_INV, QM>;
    cudaError_t e = set_max_smem(kern, smem);
    if (e != cudaSuccess) return e;
    // A cooperative launch does not overlap with kernels of other streams (measured r2, config 5: scan and forward pass ran
    // in lock-step, 5.8 ms per pair against 1.8 + 2.6 ms alone, with or without an SM partition).  When the store has been
    // given an SM budget of its own (mx_store_set_sm_limit: one CTA per SM of an exclusive partition, nothing else ever
    // resident there), co-residency holds by construction and the barrier runs under a plain launch; its bounded spin
    // still turns a violated assumption into a launch failure rather than a hang.
    if (tp.sample_tiles > 0 && tp.plain_barrier) {
        if (tp.plain_barrier == 2) {   // measurement aid (MX_SCAN_TC_PLAIN=2): plain launch as a programmatic dependent
            LaunchAttrs attrs;
            attrs.pdl();
            e = launch_ex(kern, grid, dim3(kTcThreads), (size_t)smem, st, attrs, tmC, tp);
            count_launch();
            return e != cudaSuccess ? e : cudaGetLastError();
        }
        kern<<<grid, kTcThreads, smem, st>>>(tmC, tp);
        count_launch();
        return cudaGetLastError();
    }
    if (tp.sample_tiles > 0) {
        // the grid barrier needs every CTA resident at once: cooperative launch (the driver refuses rather than deadlocks)
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(kTcThreads);
        cfg.dynamicSmemBytes = (size_t)smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kern, tmC, tp);
        if (e == cudaSuccess) {
            count_launch();
            return cudaGetLastError();
        }
        cudaGetLastError();   // not launchable cooperatively here: plain launch without seeding
        tp.sample_tiles = 0;
    }
    kern<<<grid, kTcThreads, smem, st>>>(tmC, tp);
    count_launch();
    return cudaGetLastError();
}

template <int L>
static cudaError_t launch_tc(const CUtensorMap &tmC, const TcParams &tp, bool use_inv, uint32_t qm, dim3 grid, cudaStream_t st)
{
    if (qm == 128)
        return use_inv ? launch_tc_one<L, true, 128>(tmC, tp, grid, st) : launch_tc_one<L, false, 128>(tmC, tp, grid, st);
    return use_inv ? launch_tc_one<L, true, 64>(tmC, tp, grid, st) : launch_tc_one<L, false, 64>(tmC, tp, grid, st);
}

cudaError_t tc_scan_launch(TcScanState *t, const ScanParams &p, uint64_t capacity, uint32_t k, KernelTimer *timer,
                           cudaStream_t st, const char **why)
{
    (void)capacity;
    // queries per pass = UMMA M.  128 needs the query block to fit 96 KB (dim <= 384); batches of up to 64 queries use
    // M = 64 anyway: half the query block buys three more ring stages (MX_SCAN_TC_QM=128 forces the wide form)
    static const int qm_force = getenv("MX_SCAN_TC_QM") ? atoi(getenv("MX_SCAN_TC_QM")) : 0;
    uint32_t qm = (t->k_blocks <= 6 && p.nq > 64) ? 128u : 64u;
    if (qm_force == 128 && t->k_blocks <= 6) qm = 128u;
    if (qm_force == 64) qm = 64u;
    const uint32_t nq_pad = ceil_div<uint32_t>(p.nq, qm) * qm;
    if (nq_pad > t->q_cap) {
        cudaFree(t->tau);
        cudaFree(t->samp);
        cudaFree(t->sync);
        cudaFree(t->qerr);
        cudaFree(t->floor_out);
        t->tau = nullptr;
        t->samp = nullptr;
        t->sync = nullptr;
        t->qerr = nullptr;
        t->floor_out = nullptr;
        t->q_cap = 0;
        cudaError_t e = cudaMalloc(&t->tau, (size_t)nq_pad * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&t->samp, (size_t)kSampLd * nq_pad * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&t->sync, 256);
        if (e == cudaSuccess) e = cudaMalloc(&t->qerr, (size_t)nq_pad * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&t->floor_out, (size_t)nq_pad * sizeof(float));
        if (e != cudaSuccess) {
            if (why) *why = "query staging allocation failed";
            return e;
        }
        tc_init_state_kernel<<<ceil_div<uint32_t>(nq_pad, 128), 128, 0, st>>>(t->tau, t->floor_out, t->qerr, nq_pad, t->sync, t->sync + 32,
                                                                              t->samp);
        count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        t->q_cap = nq_pad;
    }

    CUtensorMap tmC;
    if (!make_tmap_k_major_16bit(&tmC, p.rows, p.n_rows, t->dim, p.ld, kTileN, false)) {
        if (why) *why = "cuTensorMapEncodeTiled failed";
        return cudaErrorInvalidValue;
    }
    TcParams tp{};
    tp.inv_norm = p.use_inv ? p.inv_norm : nullptr;
    tp.tau = t->tau;
    tp.cand_s = p.cand_s;
    tp.cand_r = p.cand_r;
    tp.n_rows = p.n_rows;
    tp.nq = p.nq;
    tp.n_lists = p.n_lists;
    tp.k_blocks = t->k_blocks;
    tp.nq_pad = nq_pad;
    tp.samp = t->samp;
    tp.sync = t->sync;
    tp.floor_out = t->floor_out;
    tp.queries = p.queries;
    tp.ldq = p.ldq;
    tp.dim = t->dim;
    tp.qerr = t->qerr;
    tp.done = t->sync + 32;
    const uint32_t diag = getenv("MX_SCAN_TC_DIAG") ? (uint32_t)atoi(getenv("MX_SCAN_TC_DIAG")) : 0u;   // read per launch: scripts flip it
    tp.diag = diag;
    tp.prof = t->prof;
    tp.plain_barrier = t->sm_count < t->device_sms ? 1u : 0u;
    if (const char *pl = getenv("MX_SCAN_TC_PLAIN")) tp.plain_barrier = (uint32_t)atoi(pl);   // measurement aid, read per launch
    dim3 grid(p.n_lists, nq_pad / qm);
    // threshold seeding needs every CTA at the barrier: one query pass (grid.y == 1), a full grid, enough tiles per CTA
    // that scanning P of them twice is cheap; MX_SCAN_TC_SAMPLE=0 turns it off (A/B measurements)
    const int sample_cap = getenv("MX_SCAN_TC_SAMPLE") ? atoi(getenv("MX_SCAN_TC_SAMPLE")) : 4;   // read per launch: sweeps flip it
    const uint32_t tiles_per_cta = ceil_div<uint32_t>(p.n_rows, kTileN) / std::max<uint32_t>(1u, p.n_lists);
    // sampled tiles = min(4, tiles per CTA / 4): measured r2 (scripts/diag_scan_fixed.py, k = 10): 4 tiles beat 1 / 2 / 3 / 6 / 8
    // on shards of 1.25 M rows and more (216 / 198 / 181 / 180 / 182 / 186 us at 1.25 M), and sampling 4 instead of 1-2 on
    // smaller shards is worth 10-14 % (625 k rows: 123 -> 111 us, 312 k: 91 -> 79 us)
    static const uint32_t sample_div = getenv("MX_SCAN_TC_SAMPLE_DIV") ? (uint32_t)std::max(1, atoi(getenv("MX_SCAN_TC_SAMPLE_DIV"))) : 4u;
    tp.sample_tiles = (grid.y == 1 && p.n_lists == (uint32_t)t->sm_count && sample_cap > 0)
                          ? std::min<uint32_t>((uint32_t)sample_cap, tiles_per_cta / sample_div)
                          : 0u;
    if (timer) timer->begin(st, 0);
    cudaError_t e = tc_scan_lcap(k) == 16 ? launch_tc<16>(tmC, tp, p.use_inv != 0, qm, grid, st)
                                          : launch_tc<32>(tmC, tp, p.use_inv != 0, qm, grid, st);
    if (timer) timer->end(st);
    if (e != cudaSuccess && why) *why = "scan_tc_kernel launch failed";
    return e;
}

}  // namespace mx
