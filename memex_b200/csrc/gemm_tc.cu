// gemm_tc.cu -- K5/K7/K8/K9: the encoder's dense layers on tcgen05 tensor cores.
//
//   out[M, N] = epilogue( A[M, K] . W[N, K]^T )          A, W bf16 K-major; f32 accumulate in TMEM
//
// replaces the nn.Linear GEMMs libtorch executes for rust-bert's BertEncoder under
// `model.encode(&segments)` (reference lib/libmemex/src/llm/embedding.rs:109), with the
// element-wise work that follows each of them fused into the epilogue:
//   EPI_BIAS       + bias                                   (fused Q|K|V projection)
//   EPI_BIAS_GELU  + bias, erf-GELU                         (intermediate.dense)
//   EPI_BIAS_RES_LN+ bias, + residual, LayerNorm(gamma,beta)(attention.output / output; BN == N)
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one elected
// lane; also owns the TMEM allocation), warps 2.. = epilogue (8 or 16 warps: TMEM -> registers -> global).
// Three pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue, two
// accumulator stages when 2 * BN <= 512 columns), and the static tile schedule
// tile = blockIdx.x + i * gridDim.x with the N index fastest (concurrent CTAs share A rows in L2).
#include <cstdlib>

#include "common.cuh"
#include "gemm.cuh"
#include "tc.cuh"

namespace mx {

using namespace tc;

constexpr int kBM = 128;
constexpr int kBK = 64;  // 128 bytes of bf16: one swizzle atom

constexpr int kResMaxKB = 6;   // resident-weights variant: K <= 384

// RES = the CTA keeps its [BN, K] weight slice resident in shared memory and streams only activations
// (K <= 384): weights are read from L2 once per CTA instead of once per tile, which takes the K = 384
// GEMMs off the L2 -> SM bandwidth limit (a 128 x 192 x 384 tile would otherwise pull 240 KB for 2304 MMA cycles).
// CG != 0 overrides the number of epilogue column groups.  CG = 2 gives 8 epilogue warps instead of 16 / 12, which
// halves the output staging and, with a smaller bias staging, leaves room for one more ring stage: four 48 KB stages at
// BN = 256 (three before: 1.4 us of MMA work in flight against a ~1.9 us TMA round trip; measured r1: the
// intermediate GEMM 93 -> 81 us), five 40 KB stages at BN = 192.
// PAIR = cta_group::2: the two CTAs of a cluster run ONE MMA of M = 256 on vertically adjacent row tiles; each CTA streams
// its own 128 activation rows and only HALF of every weight k-block (BN / 2 rows), and the tensor cores of both SMs read
// both halves.  Shared memory bandwidth is what paces these GEMMs (every operand byte is written once by TMA and read once
// by the MMA: 188 B / clk at 128 x 256 against ~128 B / clk per SM, plus 42 B / clk of output staging -- which is where the
// measured ~50 % tensor-pipe activity of the single-CTA kernels comes from); the pair form removes a quarter of it.
template <int BN, bool RES, bool LN, int CG = 0, bool PAIR = false>
struct GemmCfg {
    static constexpr int kChunks = BN > 256 ? 2 : 1;          // one tcgen05.mma covers N <= 256
    static constexpr int kChunkN = BN / kChunks;
    static constexpr int kAccStages = (2 * BN <= 512) ? 2 : 1;
    static constexpr int kTmemCols = (kAccStages * BN <= 128) ? 128 : (kAccStages * BN <= 256 ? 256 : 512);
    static constexpr int kABytes = kBM * kBK * 2;             // one k-block of activations
    static constexpr int kBBytes = (PAIR ? BN / 2 : BN) * kBK * 2;   // one k-block of weights (PAIR: this CTA's half)
    static constexpr int kStageBytes = RES ? kABytes : kABytes + kBBytes;
    static constexpr int kResBytes = RES ? kResMaxKB * kBBytes : 0;
    // epilogue: 4 TMEM lane quarters x kColGroups column groups, one warp each -- several warps per SM
    // sub-partition so that the bias / GELU / LayerNorm arithmetic is issue-bound, not latency-bound
    static constexpr int kColGroups = CG ? CG : ((BN % 128 == 0) ? 4 : 3);   // 192 -> 3 x 64 columns
    static constexpr int kColsPerWarp = BN / kColGroups;
    static constexpr int kEpiWarps = 4 * kColGroups;
    static constexpr int kThreads = 64 + 32 * kEpiWarps;
    // LayerNorm partial (sum, sq) per column group, double buffered, + the peer CTA's row totals (split-N variant)
    static constexpr int kStatBytes = LN ? 2 * kColGroups * kBM * 8 + 2 * kBM * 8 : 0;
    static constexpr int kMaxBiasN = RES ? 2048 : (CG ? (BN == 192 ? 2304 : 3072) : 4096);
    static constexpr int kVecBytes = LN ? 3 * BN * 4 : kMaxBiasN * 4;      // bias | gamma | beta, or the whole bias vector
    // output staging for TMA stores: one [32 rows x 32 columns] 16-bit tile (2 KB, 64-byte swizzle) per epilogue warp
    // (LayerNorm with 64 columns per warp: a [32 rows x 128 bytes] tile per warp, used to transpose the residual
    //  in and the result out between the coalesced global pattern and the row-per-thread accumulator pattern)
    static constexpr bool kLnStaged = LN && kColsPerWarp == 64;
    static constexpr int kOutWarpBytes = LN ? (kLnStaged ? 4096 : 0) : 2048;
    static constexpr int kOutBytes = kEpiWarps * kOutWarpBytes;
    static constexpr int kTailBytes = ((256 + kStatBytes + kVecBytes + 1023) / 1024) * 1024 + kOutBytes;
    // the activation (+ weight) ring takes what is left of the 227 KB
    static constexpr int kRingBudget = 227 * 1024 - 1024 - kTailBytes - kResBytes;
    static constexpr int kStages = kRingBudget / kStageBytes < 8 ? kRingBudget / kStageBytes : 8;
    static constexpr int kRingOff = kResBytes;
    static constexpr int kBarOff = kRingOff + kStages * kStageBytes;
    static constexpr int kOutOff = kBarOff + kTailBytes - kOutBytes;    // 1024-byte aligned (all parts before it are)
    static constexpr int kSmemBytes = kBarOff + 1024 /*align*/ + kTailBytes;
    static_assert(kChunkN % 16 == 0 && kChunkN <= 256, "invalid UMMA N");
    static_assert((BN * kBK * 2 / kChunks) % 1024 == 0, "B chunks must stay 1024-byte aligned");
    static_assert(!PAIR || (kChunks == 1 && !RES && !LN && (BN / 2 * kBK * 2) % 1024 == 0), "pair variant: streaming, BN <= 256");
    static_assert(kColsPerWarp % 32 == 0, "an epilogue warp works in 32-column chunks");
    static_assert(kStages >= 2, "ring too small");
    static_assert(!RES || kChunks == 1, "resident variant: BN <= 256");
};

// erf-GELU for the 16-bit paths.  erf(t) = 1 - 2^(-q(t)) on t in [0, 4] with q a degree-5 polynomial
// without constant term (least-squares/minimax fit of -log2(erfc(t)); max |error| of erf 6.7e-7 when
// evaluated in f32, far below the bf16/f16 output resolution); beyond t = 4 the polynomial keeps growing.
// One SFU op (ex2) instead of erff's two polynomial branches.  The fp32 validation path keeps erff.
__device__ __forceinline__ float gelu_erf(float x)
{
    const float t = fabsf(x) * 0.70710678118654752f;   // q is increasing beyond t = 4, so 2^-q just underflows to 0
    float q = fmaf(t, 0.00294416f, -0.02959005f);
    q = fmaf(t, q, 0.14866564f);
    q = fmaf(t, q, 0.91850936f);
    q = fmaf(t, q, 1.627889f);
    q *= t;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-q));
    const float erf_x = copysignf(1.0f - e, x);
    const float h = 0.5f * x;
    return fmaf(h, erf_x, h);
}

// ALBERT's "gelu_new": 0.5 x (1 + tanh(u)), u = sqrt(2 / pi) (x + 0.044715 x^3); tanh(u) = 1 - 2 / (1 + e^(2u)) through one
// ex2 and one reciprocal (e = inf gives 1, e = 0 gives -1: no clamping needed)
__device__ __forceinline__ float gelu_tanh(float x)
{
    const float u = 0.7978845608028654f * fmaf(0.044715f * x, x * x, x);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(u * 2.8853900817779268f));   // e^(2u) = 2^(2 u log2 e)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    const float t = fmaf(-2.0f, r, 1.0f);
    const float h = 0.5f * x;
    return fmaf(h, t, h);
}

// The same polynomial for two values at once with the packed f32x2 instructions of sm_100 (FFMA2 / FMUL2: one issue
// slot for two FMAs -- the epilogue of the intermediate GEMM is bound by instruction issue, not by the FMA pipe).
__device__ __forceinline__ uint64_t pk_f32x2(float lo, float hi)
{
    return (uint64_t)__float_as_uint(lo) | ((uint64_t)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b)
{
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// x0, x1 <- gelu_erf(x0 + b0), gelu_erf(x1 + b1)
__device__ __forceinline__ void bias_gelu_erf2(float &x0, float &x1, float b0, float b1)
{
    constexpr float kC5 = 0.00294416f, kC4 = -0.02959005f, kC3 = 0.14866564f, kC2 = 0.91850936f, kC1 = 1.627889f;
    uint64_t x;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(x) : "l"(pk_f32x2(x0, x1)), "l"(pk_f32x2(b0, b1)));
    const uint64_t t = mul2(x & 0x7fffffff7fffffffull, pk_f32x2(0.70710678118654752f, 0.70710678118654752f));   // |x| / sqrt 2
    uint64_t q = fma2(t, pk_f32x2(kC5, kC5), pk_f32x2(kC4, kC4));
    q = fma2(t, q, pk_f32x2(kC3, kC3));
    q = fma2(t, q, pk_f32x2(kC2, kC2));
    q = fma2(t, q, pk_f32x2(kC1, kC1));
    q = mul2(q, t ^ 0x8000000080000000ull);   // -q(t): 2^(-q) needs no separate negation
    float e0, e1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(__uint_as_float((uint32_t)q)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(__uint_as_float((uint32_t)(q >> 32))));
    // erf(|x| / sqrt 2) = 1 - e, with the sign of x; gelu = h + h * erf, h = x / 2
    const uint64_t one = pk_f32x2(1.0f, 1.0f), m_one = pk_f32x2(-1.0f, -1.0f);
    uint64_t erf_abs = fma2(pk_f32x2(e0, e1), m_one, one);
    const uint64_t erf_x = erf_abs | (x & 0x8000000080000000ull);
    const uint64_t h = mul2(x, pk_f32x2(0.5f, 0.5f));
    const uint64_t g = fma2(h, erf_x, h);
    x0 = __uint_as_float((uint32_t)g);
    x1 = __uint_as_float((uint32_t)(g >> 32));
}

template <int FMT>
__device__ __forceinline__ uint32_t pack16(float a, float b)
{
    if constexpr (FMT == 1) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&h);
    } else {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&h);
    }
}
template <int FMT>
__device__ __forceinline__ float2 unpack16(uint32_t u)
{
    if constexpr (FMT == 1)
        return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162 *>(&u));
    else
        return __half22float2(*reinterpret_cast<__half2 *>(&u));
}

// MC = 2-CTA cluster: the two CTAs work on vertically adjacent tiles (same n_blk, m_blk = 2 i + rank), each loads
// HALF of every weight k-block and multicasts it into both CTAs' rings, which halves the L2 -> SM weight traffic
// (the streaming GEMMs are otherwise L2-bandwidth-bound: a 128 x 256 x 384 tile pulls 288 KB for 3072 MMA cycles).
//
// SPLIT (LayerNorm epilogue only) = 2-CTA cluster that splits the ROW: CTA rank r owns columns [r BN, (r + 1) BN) of
// a 128-row tile (N = 2 BN), the two CTAs exchange their per-row (sum, sum of squares) through distributed shared
// memory and each normalises its half.  With N = 384 that makes BN = 192: two TMEM accumulator stages fit, so the
// two-pass LayerNorm epilogue overlaps the next tile's MMAs; with N = 768 (BERT-base) it is what makes the fused
// epilogue possible at all (768 f32 columns do not fit the 512 TMEM columns of one SM).
// AMC (split-row LayerNorm variant only): the two CTAs of a cluster work on the SAME row tile, so each loads half of
// every activation k-block and multicasts it to both -- the activations cross the L2 -> SM path once per cluster
template <int BN, int EPI, int FMT, bool RES, bool MC, bool SPLIT, int CG = 0, bool AMC = false, bool PAIR = false>
__global__ void __launch_bounds__(GemmCfg<BN, RES, EPI == EPI_BIAS_RES_LN, CG, PAIR>::kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, GemmParams p)
{
    using Cfg = GemmCfg<BN, RES, EPI == EPI_BIAS_RES_LN, CG, PAIR>;
    static_assert(!PAIR || (!MC && !SPLIT && !AMC && !RES), "the pair variant is a form of its own");
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment by OFFSET (keeps the pointer in the shared address space: LDS/STS, not generic LD/ST)
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *ring = smem + Cfg::kRingOff;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Cfg::kBarOff);
    uint64_t *empty = full + Cfg::kStages;
    uint64_t *tmem_full = empty + Cfg::kStages;
    uint64_t *tmem_empty = tmem_full + Cfg::kAccStages;
    uint64_t *b_full = tmem_empty + Cfg::kAccStages;
    uint64_t *b_empty = b_full + 1;
    // SPLIT: the peer's row totals have arrived.  Two barriers used by alternate tiles: the peer can only send for
    // tile i + 2 after every warp here has passed its wait for tile i (it needs this CTA's tile i + 1 totals first,
    // which are sent after the CTA-wide epilogue barrier of tile i + 1), so a barrier never runs a phase ahead.
    uint64_t *x_full = b_empty + 1;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(x_full + 2);
    float2 *stat = reinterpret_cast<float2 *>(smem + Cfg::kBarOff + 256);  // [2][kColGroups][kBM]
    float *svec = reinterpret_cast<float *>(smem + Cfg::kBarOff + 256 + Cfg::kStatBytes);
    float2 *xstat = stat + 2 * Cfg::kColGroups * kBM;   // [2][kBM], written by the peer CTA

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tiles_n = p.N / BN;   // SPLIT: 2
    const uint32_t tiles_m = (p.M + kBM - 1) / kBM;
    const uint32_t n_tiles = tiles_m * tiles_n;
    const uint32_t k_blocks = (p.K + kBK - 1) / kBK;
    // tile schedule.  streaming: tile = blockIdx.x + i * gridDim.x, n fastest (concurrent CTAs share A rows in L2).
    // resident: a contiguous range of the n-major list, so a CTA changes its weight slice at most a few times.
    uint32_t t_begin, t_end, t_step;
    uint32_t cta_rank = 0;
    if constexpr (MC || SPLIT || PAIR) cta_rank = cluster_ctarank();
    if constexpr (SPLIT) {
        // a cluster strides over the row tiles; the rank picks the column half
        t_begin = blockIdx.x >> 1;
        t_end = tiles_m;
        t_step = gridDim.x >> 1;
    } else if constexpr (MC || PAIR) {
        // units = (pair of vertically adjacent tiles); a cluster strides over them, n fastest
        t_begin = blockIdx.x >> 1;
        t_end = ((tiles_m + 1) / 2) * tiles_n;
        t_step = gridDim.x >> 1;
    } else if constexpr (RES) {
        const uint32_t per = (n_tiles + gridDim.x - 1) / gridDim.x;
        t_begin = min(n_tiles, blockIdx.x * per);
        t_end = min(n_tiles, t_begin + per);
        t_step = 1;
    } else {
        t_begin = blockIdx.x;
        t_end = n_tiles;
        t_step = gridDim.x;
    }
    auto tile_m = [&](uint32_t t) {
        return SPLIT ? t : ((MC || PAIR) ? 2 * (t / tiles_n) + cta_rank : (RES ? t % tiles_m : t / tiles_n));
    };
    auto tile_n = [&](uint32_t t) { return SPLIT ? cta_rank : (RES ? t / tiles_m : t % tiles_n); };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < Cfg::kStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], (MC || AMC) ? 2 : 1);   // multicast: a stage is free once BOTH CTAs' MMAs have read it
        }
        for (int i = 0; i < Cfg::kAccStages; ++i) {
            mbar_init(&tmem_full[i], 1);
            // pair: the leader's MMA warp waits for the epilogue warps of BOTH CTAs (the peer's arrive remotely)
            mbar_init(&tmem_empty[i], PAIR ? 2 * Cfg::kEpiWarps : Cfg::kEpiWarps);
        }
        mbar_init(b_full, 1);
        mbar_init(b_empty, 1);
        mbar_init(&x_full[0], kBM);   // one remote arrive per row
        mbar_init(&x_full[1], kBM);
        fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (PAIR)
            tmem_alloc_pair(tmem_ptr, Cfg::kTmemCols);
        else
            tmem_alloc(tmem_ptr, Cfg::kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (MC || SPLIT || PAIR) cluster_sync_all();   // the peer's barriers exist before anything is sent at them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // everything above is this kernel's own set-up; from here on it reads what the previous kernel of the stream wrote
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, seg = 0, cur_n = 0xffffffffu;
            for (uint32_t tile = t_begin; tile < t_end; tile += t_step) {
                const uint32_t m_blk = tile_m(tile), n_blk = tile_n(tile);
                if constexpr (RES) {
                    if (n_blk != cur_n) {
                        // new weight slice: wait until every MMA that reads the old one has completed
                        if (seg > 0) mbar_wait(b_empty, (seg - 1) & 1);
                        mbar_arrive_expect_tx(b_full, k_blocks * Cfg::kBBytes);
                        for (uint32_t kb = 0; kb < k_blocks; ++kb)
                            tma_load_2d(smem + kb * Cfg::kBBytes, &tmB, b_full, kb * kBK, n_blk * BN, kEvictLast);
                        cur_n = n_blk;
                        ++seg;
                    }
                }
                for (uint32_t kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char *sa = ring + stage * Cfg::kStageBytes;
                    if constexpr (PAIR) {
                        // both CTAs load into their own ring; the bytes of both are counted on the LEADER's barrier, which
                        // the leader arms for the two stages' worth (the phase cannot complete before that arrive)
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
                        tma_load_2d_pair(sa, &tmA, &full[stage], kb * kBK, m_blk * kBM, kEvictFirst);
                        tma_load_2d_pair(sa + Cfg::kABytes, &tmB, &full[stage], kb * kBK, n_blk * BN + cta_rank * (BN / 2),
                                         kEvictLast);
                        if (++stage == Cfg::kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                        continue;
                    }
                    mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
                    if constexpr (AMC) {
                        // this CTA's 64 rows of the activation k-block, delivered to both CTAs (and both `full` barriers)
                        tma_load_2d_multicast(sa + cta_rank * (Cfg::kABytes / 2), &tmA, &full[stage], kb * kBK,
                                              m_blk * kBM + cta_rank * (kBM / 2), (uint16_t)3, kEvictFirst);
                    } else {
                        tma_load_2d(sa, &tmA, &full[stage], kb * kBK, m_blk * kBM, RES ? kEvictLast : kEvictFirst);
                    }
                    if constexpr (!RES) {
                        unsigned char *sb = sa + Cfg::kABytes;
                        if constexpr (MC) {
                            // this CTA's half of the weight k-block, delivered to both CTAs (and to both `full` barriers)
                            constexpr int kHalfRows = BN / 2;
                            tma_load_2d_multicast(sb + cta_rank * (kHalfRows * kBK * 2), &tmB, &full[stage], kb * kBK,
                                                  n_blk * BN + cta_rank * kHalfRows, (uint16_t)3, kEvictLast);
                        } else {
#pragma unroll
                            for (int c = 0; c < Cfg::kChunks; ++c)
                                tma_load_2d(sb + c * (Cfg::kChunkN * kBK * 2), &tmB, &full[stage], kb * kBK,
                                            n_blk * BN + c * Cfg::kChunkN, kEvictLast);
                        }
                    }
                    if (++stage == Cfg::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0 && (!PAIR || cta_rank == 0)) {   // pair: the leader issues for both SMs
            constexpr uint32_t idesc = make_idesc(PAIR ? 2 * kBM : kBM, Cfg::kChunkN, FMT);
            uint32_t stage = 0, phase = 0, local = 0, seg = 0, cur_n = 0xffffffffu;
            for (uint32_t tile = t_begin; tile < t_end; tile += t_step, ++local) {
                const uint32_t as = local % Cfg::kAccStages, aphase = (local / Cfg::kAccStages) & 1;
                if constexpr (RES) {
                    const uint32_t n_blk = tile_n(tile);
                    if (n_blk != cur_n) {
                        mbar_wait(b_full, seg & 1);
                        tc_fence_after();
                        cur_n = n_blk;
                        ++seg;
                    }
                }
                mbar_wait(&tmem_empty[as], aphase ^ 1);
                tc_fence_after();
                for (uint32_t kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(ring + stage * Cfg::kStageBytes);
                    const uint32_t sb = RES ? smem_u32(smem + kb * Cfg::kBBytes) : sa + Cfg::kABytes;
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        const uint64_t adesc = make_smem_desc(sa + k * 32);
#pragma unroll
                        for (int c = 0; c < Cfg::kChunks; ++c) {
                            const uint64_t bdesc = make_smem_desc(sb + c * (Cfg::kChunkN * kBK * 2) + k * 32);
                            if constexpr (PAIR)
                                umma_pair(tmem_base + as * BN, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                            else
                                umma(tmem_base + as * BN + c * Cfg::kChunkN, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                    }
                    if constexpr (PAIR)
                        umma_commit_pair(&empty[stage]);      // frees the stage in BOTH CTAs' rings
                    else if constexpr (MC || AMC)
                        umma_commit_multicast(&empty[stage], (uint16_t)3);
                    else
                        umma_commit(&empty[stage]);
                    if (++stage == Cfg::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if constexpr (PAIR)
                    umma_commit_pair(&tmem_full[as]);         // each CTA's epilogue reads its own 128 rows
                else
                    umma_commit(&tmem_full[as]);
                if constexpr (RES) {
                    // last tile that reads this weight slice: its completion frees the resident buffer
                    const uint32_t next = tile + t_step;
                    if (next < t_end && tile_n(next) != cur_n) umma_commit(b_empty);
                }
            }
        }
    } else {
        // ================= epilogue: warp % 4 selects the TMEM lane quarter, (warp - 2) / 4 the column group ======
        const uint32_t quarter = warp & 3;
        const uint32_t cg = (warp - 2) >> 2;
        constexpr int kCW = Cfg::kColsPerWarp;
        constexpr uint32_t kEpiThreads = Cfg::kEpiWarps * 32;
        const uint32_t etid = threadIdx.x - 64;
        if constexpr (EPI == EPI_BIAS_RES_LN) {
            // one set of per-column vectors for the whole kernel (BN == N, or this CTA's half of the row)
            const uint32_t cbase = SPLIT ? cta_rank * BN : 0;
            for (uint32_t i = etid; i < BN; i += kEpiThreads) {
                svec[i] = __ldg(p.bias + cbase + i);
                svec[BN + i] = __ldg(p.gamma + cbase + i);
                svec[2 * BN + i] = __ldg(p.beta + cbase + i);
            }
            bar_sync(1, kEpiThreads);
        } else {
            // the whole bias vector (N <= kMaxBiasN, checked by the launcher): no per-tile global load on the critical path
            for (uint32_t i = etid; i < p.N; i += kEpiThreads) svec[i] = __ldg(p.bias + i);
            bar_sync(1, kEpiThreads);
            if (etid == 0) tma_prefetch_desc(&tmO);
        }
        unsigned char *obuf = smem + Cfg::kOutOff + (warp - 2) * Cfg::kOutWarpBytes;
        uint32_t local = 0;
        for (uint32_t tile = t_begin; tile < t_end; tile += t_step, ++local) {
            const uint32_t m_blk = tile_m(tile), n_blk = tile_n(tile);
            const uint32_t as = local % Cfg::kAccStages, aphase = (local / Cfg::kAccStages) & 1;
            const float *sbias = EPI == EPI_BIAS_RES_LN ? svec : svec + n_blk * BN;
            const uint32_t row_in_tile = quarter * 32 + lane;
            const uint32_t row = m_blk * kBM + row_in_tile;
            const bool row_ok = row < p.M;
            const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + as * BN + cg * kCW;
            const uint32_t col0 = n_blk * BN + cg * kCW;
            uint16_t *orow = reinterpret_cast<uint16_t *>(p.out) + (size_t)row * p.ldo + col0;
            [[maybe_unused]] uint4 rr[EPI == EPI_BIAS_RES_LN ? kCW / 8 : 1];
            // coalesced mapping of the warp's [32 rows x 64 columns] block: 4 rows x 128 bytes per instruction
            const uint32_t crow = lane >> 3, cch = lane & 7;
            if constexpr (EPI == EPI_BIAS_RES_LN) {
                // the residual block, requested while the MMAs of the tile are still running
                if constexpr (Cfg::kLnStaged) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t r = m_blk * kBM + quarter * 32 + 4 * i + crow;
                        rr[i] = r < p.M ? *reinterpret_cast<const uint4 *>(reinterpret_cast<const uint16_t *>(p.residual) +
                                                                           (size_t)r * p.ldr + col0 + cch * 8)
                                        : make_uint4(0, 0, 0, 0);
                    }
                } else {
                    const uint16_t *rrow = reinterpret_cast<const uint16_t *>(p.residual) + (size_t)row * p.ldr + col0;
#pragma unroll
                    for (int j = 0; j < kCW / 8; ++j)
                        rr[j] = row_ok ? *reinterpret_cast<const uint4 *>(rrow + j * 8) : make_uint4(0, 0, 0, 0);
                }
            }
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            if constexpr (EPI == EPI_BIAS_RES_LN && Cfg::kLnStaged) {
                // transpose through the warp's staging tile (16-byte chunk index ^= row & 7): coalesced -> row per thread
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t r = 4 * i + crow;
                    *reinterpret_cast<uint4 *>(obuf + r * 128 + ((cch ^ (r & 7)) << 4)) = rr[i];
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j) rr[j] = *reinterpret_cast<const uint4 *>(obuf + lane * 128 + ((j ^ (lane & 7)) << 4));
                __syncwarp();
            }

            if constexpr (EPI == EPI_BIAS_RES_LN) {
                // pass 1: x = acc + bias + residual, kept in TMEM; partial row sums of this warp's columns
                float sum = 0.f, sq = 0.f;
#pragma unroll
                for (int c = 0; c < kCW / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c * 32, v);
                    tmem_ld_wait();
                    const uint32_t *rh = reinterpret_cast<const uint32_t *>(&rr[c * 4]);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float2 r2 = unpack16<FMT>(rh[j]);
                        const float2 b2 = *reinterpret_cast<const float2 *>(sbias + cg * kCW + c * 32 + 2 * j);
                        const float x0 = __uint_as_float(v[2 * j]) + b2.x + r2.x;
                        const float x1 = __uint_as_float(v[2 * j + 1]) + b2.y + r2.y;
                        sum += x0 + x1;
                        sq = fmaf(x0, x0, fmaf(x1, x1, sq));
                        v[2 * j] = __float_as_uint(x0);
                        v[2 * j + 1] = __float_as_uint(x1);
                    }
                    tmem_st32(taddr + c * 32, v);
                }
                float2 *st2 = stat + (local & 1) * (Cfg::kColGroups * kBM);
                st2[cg * kBM + row_in_tile] = make_float2(sum, sq);
                tmem_st_wait();
                bar_sync(1, Cfg::kEpiWarps * 32);
                sum = 0.f;
                sq = 0.f;
#pragma unroll
                for (int i = 0; i < Cfg::kColGroups; ++i) {
                    const float2 t2 = st2[i * kBM + row_in_tile];
                    sum += t2.x;
                    sq += t2.y;
                }
                if constexpr (SPLIT) {
                    // this CTA's row totals -> the peer (one thread per row sends), the peer's -> here
                    if (cg == 0) {
                        const uint32_t peer = cta_rank ^ 1u;
                        st_cluster_f32x2(mapa_shared(smem_u32(xstat + (local & 1) * kBM + row_in_tile), peer), sum, sq);
                        mbar_arrive_remote_release(mapa_shared(smem_u32(&x_full[local & 1]), peer));
                    }
                    mbar_wait_acquire_cluster(&x_full[local & 1], (local >> 1) & 1);
                    const float2 o2 = xstat[(local & 1) * kBM + row_in_tile];
                    sum += o2.x;
                    sq += o2.y;
                }
                constexpr float kInvN = 1.0f / (SPLIT ? 2 * BN : BN);
                const float mean = sum * kInvN;
                const float var = fmaxf(sq * kInvN - mean * mean, 0.f);
                const float rstd = rsqrtf(var + p.ln_eps);
#pragma unroll 1
                for (int c = 0; c < kCW / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c * 32, v);
                    tmem_ld_wait();
                    uint32_t o[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float2 g2 = *reinterpret_cast<const float2 *>(svec + BN + cg * kCW + c * 32 + 2 * j);
                        const float2 be2 = *reinterpret_cast<const float2 *>(svec + 2 * BN + cg * kCW + c * 32 + 2 * j);
                        const float y0 = (__uint_as_float(v[2 * j]) - mean) * rstd * g2.x + be2.x;
                        const float y1 = (__uint_as_float(v[2 * j + 1]) - mean) * rstd * g2.y + be2.y;
                        o[j] = pack16<FMT>(y0, y1);
                    }
                    if constexpr (Cfg::kLnStaged) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<uint4 *>(obuf + lane * 128 + (((c * 4 + j) ^ (lane & 7)) << 4)) =
                                make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                    } else if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<uint4 *>(orow + c * 32 + j * 8) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                    }
                }
                if constexpr (Cfg::kLnStaged) {
                    // row per thread -> coalesced: every store instruction writes 4 rows x 128 contiguous bytes
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t rl = 4 * i + crow, r = m_blk * kBM + quarter * 32 + rl;
                        const uint4 v4 = *reinterpret_cast<const uint4 *>(obuf + rl * 128 + ((cch ^ (rl & 7)) << 4));
                        if (r < p.M)
                            *reinterpret_cast<uint4 *>(reinterpret_cast<uint16_t *>(p.out) + (size_t)r * p.ldo + col0 + cch * 8) = v4;
                    }
                    __syncwarp();
                }
            } else {
                // two register buffers: the TMEM load of chunk c + 1 is in flight while chunk c is processed
                constexpr int kNC = kCW / 32;
                uint32_t vbuf[2][32];
                tmem_ld32(taddr, vbuf[0]);
#pragma unroll
                for (int c = 0; c < kNC; ++c) {
                    tmem_ld_wait();
                    if (c + 1 < kNC) tmem_ld32(taddr + (c + 1) * 32, vbuf[(c + 1) & 1]);
                    const uint32_t(&v)[32] = vbuf[c & 1];
                    uint32_t o[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float2 b2 = *reinterpret_cast<const float2 *>(sbias + cg * kCW + c * 32 + 2 * j);
                        float x0 = __uint_as_float(v[2 * j]), x1 = __uint_as_float(v[2 * j + 1]);
                        if constexpr (EPI == EPI_BIAS_GELU) {
                            bias_gelu_erf2(x0, x1, b2.x, b2.y);
                        } else if constexpr (EPI == EPI_BIAS_GELU_TANH) {
                            x0 = gelu_tanh(x0 + b2.x);
                            x1 = gelu_tanh(x1 + b2.y);
                        } else {
                            x0 += b2.x;
                            x1 += b2.y;
                        }
                        o[j] = pack16<FMT>(x0, x1);
                    }
                    // [32 rows x 32 columns] -> this warp's staging tile (64-byte swizzle) -> one TMA store.
                    // (A direct st.global from the accumulator layout writes 16 bytes to each of 32 rows per
                    // instruction -- 32 LSU wavefronts for 512 bytes.)
                    if (lane == 0) bulk_wait_read0();   // the previous store has finished reading the tile
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4 *>(obuf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                            make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmO, obuf, (int32_t)(col0 + c * 32), (int32_t)(m_blk * kBM + quarter * 32));
                        bulk_commit();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) {
                    if (cta_rank == 0)
                        mbar_arrive(&tmem_empty[as]);
                    else
                        mbar_arrive_remote_release(mapa_shared(smem_u32(&tmem_empty[as]), 0));
                } else {
                    mbar_arrive(&tmem_empty[as]);
                }
            }
        }
        if constexpr (EPI != EPI_BIAS_RES_LN) {
            if (lane == 0) bulk_wait_read0();
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (MC || SPLIT || PAIR) cluster_sync_all();   // no CTA leaves while the peer can still signal its barriers
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR)
            tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
        else
            tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

template <int BN, int EPI, int FMT, bool RES, bool MC, bool SPLIT = false, int CG = 0, bool AMC = false, bool PAIR = false>
static cudaError_t launch_cfg(const GemmParams &p, const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmO,
                              int sm_count, cudaStream_t st)
{
    using Cfg = GemmCfg<BN, RES, EPI == EPI_BIAS_RES_LN, CG, PAIR>;
    auto kern = gemm_tc_kernel<BN, EPI, FMT, RES, MC, SPLIT, CG, AMC, PAIR>;
    cudaError_t e = set_max_smem(kern, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    const uint32_t tiles_m = ceil_div<uint32_t>(p.M, kBM), tiles_n = p.N / BN;
    LaunchAttrs attrs;
    attrs.pdl();
    uint32_t grid;
    if constexpr (MC || SPLIT || PAIR) {
        const uint32_t units = SPLIT ? tiles_m : ceil_div<uint32_t>(tiles_m, 2) * tiles_n;
        const uint32_t clusters = std::min<uint32_t>(units, (uint32_t)sm_count / 2);
        grid = 2 * clusters;
        attrs.cluster(2);
    } else {
        grid = std::min<uint32_t>(tiles_m * tiles_n, (uint32_t)sm_count);
    }
    e = launch_ex(kern, dim3(grid), dim3(Cfg::kThreads), (size_t)Cfg::kSmemBytes, st, attrs, tmA, tmB, tmO, p);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

// resident-weights variant: K <= 384, N a multiple of 192, enough tiles per CTA to amortise the weight load
static bool use_resident(const GemmParams &p, int epi, int sm_count)
{
    if (epi == EPI_BIAS_RES_LN) return false;
    // measured on B200 (r1): with 144 KB of weights resident only 3 activation stages (48 KB) fit, and the ring then
    // cannot cover the TMA round trip; the streaming + multicast variant is faster.  Kept as an opt-in experiment.
    static const bool enabled = getenv("MX_GEMM_RESIDENT") != nullptr;
    if (!enabled) return false;
    if (p.K > (uint32_t)(kResMaxKB * kBK) || p.N % 192 != 0 || p.N > 2048) return false;
    const uint32_t n_tiles = ceil_div<uint32_t>(p.M, kBM) * (p.N / 192);
    return n_tiles >= 4u * (uint32_t)sm_count;
}

int gemm_tc_block_n(uint32_t N, int epi)
{
    // LayerNorm epilogue: the row lives in one CTA (N = 384) or in the two CTAs of a cluster (N = 384 or 768)
    if (epi == EPI_BIAS_RES_LN) return (N == 384 || N == 768) ? 384 : 0;
    if (N % 256 == 0) return 256;
    if (N % 192 == 0) return 192;
    if (N % 128 == 0) return 128;
    return 0;
}

cudaError_t launch_gemm_tc(const GemmParams &p, int epi, int sm_count, cudaStream_t st, const char **why)
{
    const bool res = epi != EPI_BIAS_GELU_TANH && use_resident(p, epi, sm_count);
    const int bn = res ? 192 : gemm_tc_block_n(p.N, epi);
    if (bn == 0 || p.K % 8 != 0 || p.lda % 8 != 0 || p.ldw % 8 != 0 || p.ldo % 8 != 0) {
        if (why) *why = "unsupported GEMM shape for the tcgen05 path";
        return cudaErrorInvalidValue;
    }
    if (epi != EPI_BIAS_RES_LN && p.N > (res ? 2048u : 4096u)) {
        if (why) *why = "N > 4096 is not supported by the bias staging";
        return cudaErrorInvalidValue;
    }
    // 2-CTA clusters with weight multicast (needs enough tile pairs to fill the chip).  Measured on B200 (r1,
    // profiles/README.md): it halves the weight traffic from L2 but the kernels are paced by the tensor pipe, not by
    // L2, and the lock-step of the two CTAs costs 3-6 % -- so it is opt-in (MX_GEMM_MULTICAST=1; the parity tests
    // run both settings).
    static const bool mc_on = getenv("MX_GEMM_MULTICAST") != nullptr;
    const bool mc = !res && mc_on && epi != EPI_BIAS_GELU_TANH && sm_count >= 2 && ceil_div<uint32_t>(p.M, 2 * kBM) * (p.N / bn) >= (uint32_t)sm_count / 2;
    // CTA pairs (cta_group::2, GemmCfg): parity-green, but MEASURED SLOWER than the single-CTA form on the encoder step
    // (r2, one box, A/B: 97.9 / 97.5 k against 103.6 / 103.0 k segments/s; GEMM time 1.90 against 1.77 ms) -- the
    // accumulator hand-over couples the two CTAs' epilogues -- so it is opt-in (MX_GEMM_PAIR=1)
    static const bool pair_on = getenv("MX_GEMM_PAIR") != nullptr && atoi(getenv("MX_GEMM_PAIR")) != 0;
    const bool pair = pair_on && !res && !mc && (epi == EPI_BIAS || epi == EPI_BIAS_GELU) && (bn == 192 || bn == 256) &&
                      sm_count >= 2 && ceil_div<uint32_t>(p.M, 2 * kBM) * (p.N / bn) >= (uint32_t)sm_count / 2 &&
                      p.N <= (bn == 192 ? 2304u : 3072u);
    CUtensorMap tmA, tmB, tmO;
    uint32_t chunk_rows = (mc || pair) ? bn / 2 : (bn > 256 ? bn / 2 : bn);
    if (epi == EPI_BIAS_RES_LN && !mc) chunk_rows = 192;   // whole-row 384 (2 chunks), split 2 x 192, split 2 x 384 (2 chunks each)
    // split-row LayerNorm variant with activation multicast (MX_GEMM_LN_AMC=1, opt-in until measured): each CTA of the
    // pair loads 64 of the tile's 128 rows
    static const bool amc_on = getenv("MX_GEMM_LN_AMC") != nullptr;
    static const bool ln_no_split = getenv("MX_GEMM_LN_NO_SPLIT") != nullptr;
    const bool amc = amc_on && epi == EPI_BIAS_RES_LN && !mc && sm_count >= 2 && (p.N == 768 || !ln_no_split);
    if (!make_tmap_k_major_16bit(&tmA, p.A, p.M, p.K, p.lda, amc ? kBM / 2 : kBM, p.fmt == 1) ||
        !make_tmap_k_major_16bit(&tmB, p.W, p.N, p.K, p.ldw, chunk_rows, p.fmt == 1) ||
        !make_tmap_store_32x32_16bit(&tmO, p.out, p.M, p.N, p.ldo, p.fmt == 1)) {
        if (why) *why = "cuTensorMapEncodeTiled failed";
        return cudaErrorInvalidValue;
    }
#define MX_GEMM(BN_, EPI_, RES_, MC_)                                                         \
    return p.fmt == 1 ? launch_cfg<BN_, EPI_, 1, RES_, MC_>(p, tmA, tmB, tmO, sm_count, st)   \
                      : launch_cfg<BN_, EPI_, 0, RES_, MC_>(p, tmA, tmB, tmO, sm_count, st)
#define MX_GEMM_MC(BN_, EPI_)                 \
    do {                                      \
        if (mc) {                             \
            MX_GEMM(BN_, EPI_, false, true);  \
        } else {                              \
            MX_GEMM(BN_, EPI_, false, false); \
        }                                     \
    } while (0)
    if (epi == EPI_BIAS_RES_LN) {
        // N = 768: split over a 2-CTA cluster (2 x 384).  N = 384: split (2 x 192, two accumulator stages) when there
        // are enough row tiles for the 74 clusters; MX_GEMM_LN_NO_SPLIT keeps the single-CTA whole-row variant.
        static const bool no_split = getenv("MX_GEMM_LN_NO_SPLIT") != nullptr;
        const bool can_split = sm_count >= 2;
        if (p.N == 768) {
            if (!can_split) {
                if (why) *why = "N = 768 LayerNorm epilogue needs a 2-CTA cluster";
                return cudaErrorInvalidValue;
            }
            if (amc)
                return p.fmt == 1 ? launch_cfg<384, EPI_BIAS_RES_LN, 1, false, false, true, 0, true>(p, tmA, tmB, tmO, sm_count, st)
                                  : launch_cfg<384, EPI_BIAS_RES_LN, 0, false, false, true, 0, true>(p, tmA, tmB, tmO, sm_count, st);
            return p.fmt == 1 ? launch_cfg<384, EPI_BIAS_RES_LN, 1, false, false, true>(p, tmA, tmB, tmO, sm_count, st)
                              : launch_cfg<384, EPI_BIAS_RES_LN, 0, false, false, true>(p, tmA, tmB, tmO, sm_count, st);
        }
        if (can_split && !no_split && !mc) {
            if (amc)
                return p.fmt == 1 ? launch_cfg<192, EPI_BIAS_RES_LN, 1, false, false, true, 0, true>(p, tmA, tmB, tmO, sm_count, st)
                                  : launch_cfg<192, EPI_BIAS_RES_LN, 0, false, false, true, 0, true>(p, tmA, tmB, tmO, sm_count, st);
            return p.fmt == 1 ? launch_cfg<192, EPI_BIAS_RES_LN, 1, false, false, true>(p, tmA, tmB, tmO, sm_count, st)
                              : launch_cfg<192, EPI_BIAS_RES_LN, 0, false, false, true>(p, tmA, tmB, tmO, sm_count, st);
        }
        MX_GEMM_MC(384, EPI_BIAS_RES_LN);
    }
    if (pair) {
        if (epi == EPI_BIAS_GELU && bn == 256)
            return p.fmt == 1 ? launch_cfg<256, EPI_BIAS_GELU, 1, false, false, false, 2, false, true>(p, tmA, tmB, tmO, sm_count, st)
                              : launch_cfg<256, EPI_BIAS_GELU, 0, false, false, false, 2, false, true>(p, tmA, tmB, tmO, sm_count, st);
        if (epi == EPI_BIAS_GELU)
            return p.fmt == 1 ? launch_cfg<192, EPI_BIAS_GELU, 1, false, false, false, 2, false, true>(p, tmA, tmB, tmO, sm_count, st)
                              : launch_cfg<192, EPI_BIAS_GELU, 0, false, false, false, 2, false, true>(p, tmA, tmB, tmO, sm_count, st);
        if (bn == 256)
            return p.fmt == 1 ? launch_cfg<256, EPI_BIAS, 1, false, false, false, 2, false, true>(p, tmA, tmB, tmO, sm_count, st)
                              : launch_cfg<256, EPI_BIAS, 0, false, false, false, 2, false, true>(p, tmA, tmB, tmO, sm_count, st);
        return p.fmt == 1 ? launch_cfg<192, EPI_BIAS, 1, false, false, false, 2, false, true>(p, tmA, tmB, tmO, sm_count, st)
                          : launch_cfg<192, EPI_BIAS, 0, false, false, false, 2, false, true>(p, tmA, tmB, tmO, sm_count, st);
    }
    if (epi == EPI_BIAS_GELU) {
        if (res) MX_GEMM(192, EPI_BIAS_GELU, true, false);
        // BN = 256 with 8 epilogue warps and four 48 KB stages instead of 16 warps and three (GemmCfg, CG): +3.4 % on the
        // whole encoder step (r1); MX_GEMM_EPI16=1 keeps the 16-warp form for A/B measurements
        static const bool epi8 = getenv("MX_GEMM_EPI16") == nullptr;
        if (bn == 256 && epi8 && !mc && p.N <= 3072)
            return p.fmt == 1 ? launch_cfg<256, EPI_BIAS_GELU, 1, false, false, false, 2>(p, tmA, tmB, tmO, sm_count, st)
                              : launch_cfg<256, EPI_BIAS_GELU, 0, false, false, false, 2>(p, tmA, tmB, tmO, sm_count, st);
        if (bn == 256) MX_GEMM_MC(256, EPI_BIAS_GELU);
        if (bn == 192) MX_GEMM_MC(192, EPI_BIAS_GELU);
        MX_GEMM_MC(128, EPI_BIAS_GELU);
    }
    if (epi == EPI_BIAS_GELU_TANH) {   // ALBERT only: the plain streaming variant
        if (bn == 256) MX_GEMM(256, EPI_BIAS_GELU_TANH, false, false);
        if (bn == 192) MX_GEMM(192, EPI_BIAS_GELU_TANH, false, false);
        MX_GEMM(128, EPI_BIAS_GELU_TANH, false, false);
    }
    if (res) MX_GEMM(192, EPI_BIAS, true, false);
    // BN = 192 with 8 epilogue warps and five 40 KB stages instead of 12 warps and four: the QKV projection 66 -> 58 us,
    // +1.7 % on the whole encoder step (r1); MX_GEMM_QKV12=1 keeps the 12-warp form for A/B measurements
    static const bool qkv8 = getenv("MX_GEMM_QKV12") == nullptr;
    if (bn == 192 && qkv8 && !mc && p.N <= 2304)
        return p.fmt == 1 ? launch_cfg<192, EPI_BIAS, 1, false, false, false, 2>(p, tmA, tmB, tmO, sm_count, st)
                          : launch_cfg<192, EPI_BIAS, 0, false, false, false, 2>(p, tmA, tmB, tmO, sm_count, st);
    if (bn == 256) MX_GEMM_MC(256, EPI_BIAS);
    if (bn == 192) MX_GEMM_MC(192, EPI_BIAS);
    MX_GEMM_MC(128, EPI_BIAS);
#undef MX_GEMM_MC
#undef MX_GEMM
}

}  // namespace mx
