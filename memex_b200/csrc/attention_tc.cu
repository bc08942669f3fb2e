// attention_tc.cu -- K6 on tcgen05: fused masked softmax attention for sequences of up to 256 keys.
//
//   ctx[b, q, h, :] = softmax_k( q . k / sqrt(dh) + padding mask ) v        k < lens[b]
//
// replaces BertSelfAttention's bmm -> softmax -> bmm (libtorch materialises [B, heads, S, S] under
// `model.encode(&segments)`, reference lib/libmemex/src/llm/embedding.rs:109).
//
// A UNIT is one (sequence, head, 128 query rows).  Both matrix products of a unit run on the tensor pipe:
//     S[128, keys] = Q K^T     one tcgen05.mma group, M = 128, N = keys (<= 256), accumulator in 256 TMEM columns
//     O[128, dh]   = P V       per 64-key chunk of P; V is the MN-major B operand exactly as it lies in qkv
// and the softmax in between is one THREAD PER QUERY ROW (thread = TMEM lane): the row max is thread-local, no
// shuffles; the row sum comes out of the tensor pipe as well (P times a tile of ones, into 16 more TMEM columns).
// P never leaves TMEM: it is written back (tcgen05.st, packed 16-bit pairs) over the score columns it came from and
// read by the tensor pipe as the A operand of P V and P 1 (the `.ts` MMA form) -- no shared-memory round trip and no
// generic -> async proxy fence (measured: the st.shared + fence.proxy.async version cost 25 us of 128 per layer).
//
// The op is bound by the exponentials (S x S x heads per sequence on 16 MUFU lanes per SM against 4 S^2 dh flops at
// 8192 flop/clk), so a CTA keeps TWO independent units in flight -- group g owns TMEM columns [256 g, 256 g + 256),
// its own Q / K / V / P buffers, four softmax warps, one TMA producer warp and one MMA issuer warp -- and the
// tensor pipe and the TMA loads of one group run under the other group's exponentials.  Slot layout: S f32 in columns
// [0, 256); P overlays [0, 128); O and l use [128, 128 + dh + 16), dead once the scores of those keys have been
// consumed (the products of the first chunks are held back until then).
//
// Persistent: grid = #SMs, units dealt round-robin (the two units of one (sequence, head) land in one CTA).
//
// Measured (r1, B200, MiniLM-L6, B = 256, S = 256): 130 us per layer against 141 us for the mma.sync kernel
// (attention_mma.cu) and a 43 us MUFU bound (scripts/ubench/pipes.cu: 16 ex2 / clk / SM); ncu puts the XU pipe at
// 37-43 % in both.  One softmax warp needs ~5.4 k cycles of issue time per unit (a MUFU holds its warp for 8-11
// cycles, plus FFMA / pack / TMEM load / proxy fence), and 512 TMEM columns hold only two 256-key score tiles, i.e.
// two such instruction streams per SM sub-partition.  Variants tried and measured slower: two threads per row
// (148-177 us), single group with ping-pong score slots + P kept in TMEM as the A operand + auxiliary max / drain
// warps (140-165 us; kept as experiments/attention_tc_pingpong.cu.txt), a rolled 8-key software pipeline (184 us).
// Also measured slower (r1): announcing a P chunk half a chunk late so that tcgen05.wait::st never stalls the softmax warp
// (+6 % kernel time: the last two chunks' P V products then queue up behind the final announcement and lengthen the unit's
// tail).  Also tried (r2): 128-key sub-units -- four score tiles in TMEM, four independent softmax streams per sub-partition,
// the two key blocks of a row combined through shared memory as in split-KV decoding; correct, but slower at the benchmark
// shape (768 threads leave 85 registers per thread, and the pair combination adds a hand-over per unit).  Kept, not compiled,
// as experiments/attention_tc4.cu.txt.
#include <cstdlib>

#include "common.cuh"
#include "encoder.cuh"
#include "tc.cuh"

namespace mx {

using namespace tc;

namespace {

constexpr int kQT = 128;        // query rows per unit = UMMA M = TMEM lanes
constexpr int kKC = 64;         // keys per P chunk: one 128-byte swizzle-atom row of 16-bit values
constexpr int kMaxKeys = 256;   // keys per unit: one UMMA N, 256 f32 TMEM columns
constexpr int kGroups = 2;

template <int DH>
struct AttCfg {
    static constexpr int kSoftmaxWarps = 4 * kGroups;           // thread = query row = TMEM lane
    static constexpr int kThreads = 32 * (kSoftmaxWarps + 2 * kGroups);
    static constexpr int kRowBytes = DH * 2;                    // 64 (SWIZZLE_64B) or 128 (SWIZZLE_128B)
    static constexpr int kQBytes = kQT * kRowBytes;
    static constexpr int kKBytes = kMaxKeys * kRowBytes;
    static constexpr int kVBufs = 2;                            // the next unit's V lands while this unit's P V runs
    static constexpr int kOCol = 128;                           // O inside the slot (P occupies [0, 128))
    static constexpr int kLCol = kOCol + DH;                    // row sums
    // O and l sit on the score columns [128, 128 + DH + 16): the products wait until those have been consumed
    static constexpr int kHoldChunk = (kOCol + DH + 16 - 1) / kKC;
    static constexpr int kQOff = 0;
    static constexpr int kKOff = kQBytes;
    static constexpr int kVOff = kKOff + kKBytes;
    static constexpr int kGroupBytes = kVOff + kVBufs * kKBytes;
    static constexpr int kOnesOff = kGroups * kGroupBytes;      // [16 rows][64 keys] of 1.0, K-major: B operand of the row sums
    static constexpr int kOnesBytes = 16 * 128;
    static constexpr int kBarOff = kOnesOff + kOnesBytes;
    static constexpr int kUsedBytes = kBarOff + 256;
    // + up to 1 KB to align the base by hand; head_dim 64 fills the SM (768 bytes of slack left: the kernel traps if the
    // dynamic window is not 256-byte aligned, in practice it starts 1 KB into the CTA's shared memory)
    static constexpr int kSmemBytes = kUsedBytes + 1024 <= 227 * 1024 ? kUsedBytes + 1024 : 227 * 1024;
    static constexpr uint32_t kLayout = DH <= 32 ? 4u : 2u;     // UMMA layout type: SWIZZLE_64B / SWIZZLE_128B
    static constexpr uint32_t kSbo = 8 * kRowBytes;             // 8-row groups: 512 / 1024 bytes apart
    static_assert(kUsedBytes <= 227 * 1024, "attention tiles do not fit");
    static_assert(kGroupBytes % 1024 == 0, "group buffers keep the 1024-byte alignment");
};

// per-group barrier slots
constexpr int kPRing = 8;   // "P chunk written" barriers: two units' worth, so a phase is always observed before its reuse
enum { B_QK_FULL = 0, B_V_FULL = 1 /* 2 */, B_S_FULL = 3, B_P_READY = 4 /* kPRing */, B_O_FULL = B_P_READY + kPRing,
       B_S_EMPTY, B_PER_GROUP };
static_assert(kGroups * B_PER_GROUP * 8 + 8 <= 256, "barrier block");

__device__ __forceinline__ uint64_t att_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                    // leading byte offset: not used by these shapes (one swizzle atom wide)
    d |= (uint64_t)(sbo_bytes >> 4) << 32;     // stride between 8-row groups
    d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
    d |= (uint64_t)layout << 61;
    return d;
}

__device__ __forceinline__ float ex2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <bool BF16>
__device__ __forceinline__ uint32_t pk2(float a, float b)
{
    if constexpr (BF16) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&h);
    } else {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&h);
    }
}

struct Unit {
    uint32_t b, h, q0, len;
    uint32_t row0;   // first row of the sequence in qkv / ctx: b S (padded layout) or cu[b] (packed layout)
    bool skip;       // every query row of the tile is padding
};
// The sequence length (and, packed, its first row) is the only thing a unit reads from global memory before its
// barriers; every role fetches it ONE UNIT AHEAD (meta_of for the next unit at the top of the current one), so its
// latency never sits between two units.
struct UnitMeta {
    int32_t raw_len, row0;
};
__device__ __forceinline__ UnitMeta meta_of(uint32_t u, uint32_t n_units, uint32_t n_qt, uint32_t heads, uint32_t S,
                                            const int32_t *lens, const int32_t *cu)
{
    UnitMeta m{0, 0};
    if (u < n_units) {
        const uint32_t b = u / (n_qt * heads);
        m.raw_len = __ldg(lens + b);
        m.row0 = cu ? __ldg(cu + b) : (int32_t)(b * S);
    }
    return m;
}
__device__ __forceinline__ Unit decode_unit(uint32_t u, UnitMeta m, uint32_t n_qt, uint32_t heads, uint32_t S)
{
    Unit r;
    const uint32_t qt = u % n_qt, bh = u / n_qt;
    r.h = bh % heads;
    r.b = bh / heads;
    r.q0 = qt * kQT;
    r.len = min((uint32_t)max(m.raw_len, 0), S);
    r.row0 = (uint32_t)m.row0;
    r.skip = r.q0 >= r.len;
    return r;
}

// maximum of 32 scores (keys key0 .. key0 + 31 of this thread's row; keys >= len are padding)
__device__ __forceinline__ float max32(const uint32_t (&v)[32], uint32_t key0, uint32_t len)
{
    float m0 = kNegInf, m1 = kNegInf, m2 = kNegInf, m3 = kNegInf;
    if (key0 + 32 <= len) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            m0 = fmaxf(m0, __uint_as_float(v[j]));
            m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
            m2 = fmaxf(m2, __uint_as_float(v[j + 2]));
            m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            m0 = fmaxf(m0, key0 + j < len ? __uint_as_float(v[j]) : kNegInf);
            m1 = fmaxf(m1, key0 + j + 1 < len ? __uint_as_float(v[j + 1]) : kNegInf);
            m2 = fmaxf(m2, key0 + j + 2 < len ? __uint_as_float(v[j + 2]) : kNegInf);
            m3 = fmaxf(m3, key0 + j + 3 < len ? __uint_as_float(v[j + 3]) : kNegInf);
        }
    }
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// exponentials of 32 scores of this thread's row -> 16 packed 16-bit pairs (keys >= len give 0)
template <bool BF16, bool NOEXP = false>
__device__ __forceinline__ void softmax32(const uint32_t (&v)[32], float sc, float msc, uint32_t key0, uint32_t len,
                                          uint32_t (&o)[16])
{
    if constexpr (NOEXP) {   // timing diagnostic only: everything but the MUFU
#pragma unroll
        for (int j = 0; j < 16; ++j)
            o[j] = pk2<BF16>(fmaf(__uint_as_float(v[2 * j]), sc, -msc), fmaf(__uint_as_float(v[2 * j + 1]), sc, -msc));
        return;
    }
    if (key0 + 32 <= len) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            o[j] = pk2<BF16>(ex2(fmaf(__uint_as_float(v[2 * j]), sc, -msc)), ex2(fmaf(__uint_as_float(v[2 * j + 1]), sc, -msc)));
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float p0 = key0 + 2 * j < len ? ex2(fmaf(__uint_as_float(v[2 * j]), sc, -msc)) : 0.f;
            const float p1 = key0 + 2 * j + 1 < len ? ex2(fmaf(__uint_as_float(v[2 * j + 1]), sc, -msc)) : 0.f;
            o[j] = pk2<BF16>(p0, p1);
        }
    }
}

// DIAG != 0 builds are TIMING DIAGNOSTICS with wrong results (MX_ATTN_DIAG, scripts/diag_attention.py):
//   1 no MUFU, 2 no max pass, 4 no P stores / proxy fence, 8 no pass-2 work at all, 16 no O read-out / store
template <bool BF16, int DH, int DIAG = 0>
__global__ void __launch_bounds__(AttCfg<DH>::kThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const int32_t *__restrict__ lens, uint16_t *__restrict__ ctx,
                    uint32_t B, uint32_t S, uint32_t H, uint32_t heads, float scale_log2e, const int32_t *__restrict__ cu)
{
    using Cfg = AttCfg<DH>;
    constexpr uint32_t kSoftmaxWarps = Cfg::kSoftmaxWarps;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    if ((uint32_t)(smem - smem_raw) + Cfg::kUsedBytes > (uint32_t)Cfg::kSmemBytes) __trap();
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::kBarOff);
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + kGroups * B_PER_GROUP);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n_qt = (S + kQT - 1) / kQT;
    const uint32_t n_units = B * heads * n_qt;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQKV);
        for (int g = 0; g < kGroups; ++g) {
            uint64_t *bg = bars + g * B_PER_GROUP;
            mbar_init(bg + B_QK_FULL, 1);
            mbar_init(bg + B_V_FULL, 1);
            mbar_init(bg + B_V_FULL + 1, 1);
            mbar_init(bg + B_S_FULL, 1);
            for (int i = 0; i < kPRing; ++i) mbar_init(bg + B_P_READY + i, 4);   // one arrive per softmax warp
            mbar_init(bg + B_O_FULL, 1);
            mbar_init(bg + B_S_EMPTY, 4);
        }
        fence_barrier_init();
    }
    // [16][64] tile of 1.0 (constant, so the swizzle does not matter): B operand that turns P into its row sums
    for (uint32_t i = threadIdx.x; i < Cfg::kOnesBytes / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(smem + Cfg::kOnesOff)[i] = BF16 ? 0x3f803f80u : 0x3c003c00u;
    fence_proxy_async_smem();
    if (warp == kSoftmaxWarps) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_trigger();
    pdl_wait();   // qkv is the previous kernel's output (lens / cu were copied in before the step)

    if (warp < kSoftmaxWarps) {
        // ================= softmax: thread = query row = TMEM lane =================
        const uint32_t g = warp >> 2, quarter = warp & 3;
        uint64_t *bg = bars + g * B_PER_GROUP;
        const uint32_t row = quarter * 32 + lane;
        const uint32_t t_s = tmem_base + ((quarter * 32u) << 16) + g * kMaxKeys;
        uint32_t it = 0, cc = 0;
        const bool packed = cu != nullptr;   // packed layout: rows of padding do not exist, nothing is written for them
        UnitMeta meta = meta_of(blockIdx.x * kGroups + g, n_units, n_qt, heads, S, lens, cu);
        for (uint32_t u = blockIdx.x * kGroups + g; u < n_units; u += kGroups * gridDim.x) {
            const Unit un = decode_unit(u, meta, n_qt, heads, S);
            meta = meta_of(u + kGroups * gridDim.x, n_units, n_qt, heads, S, lens, cu);
            const uint32_t q = un.q0 + row;
            uint16_t *orow = ctx + ((size_t)un.row0 + q) * H + (size_t)un.h * DH;
            if (un.skip) {
                // padding rows are written as zero: the following GEMMs stay finite
                if (q < S && !packed) {
#pragma unroll
                    for (int j = 0; j < DH / 8; ++j) *reinterpret_cast<uint4 *>(orow + j * 8) = make_uint4(0, 0, 0, 0);
                }
                continue;
            }
            const uint32_t len = un.len;
            const uint32_t nch = (len + kKC - 1) / kKC;
            const uint32_t nsteps = 2 * nch;   // 32-key steps
            mbar_wait(bg + B_S_FULL, it & 1);
            tc_fence_after();
            uint32_t va[32], vb[32];
            // ---- pass 1: row maximum over the valid keys (the load of step t + 1 is in flight during step t) ----
            float m = kNegInf;
            if constexpr (DIAG & 2) {
                m = 0.f;
            } else {
                tmem_ld32(t_s, va);
                for (uint32_t t = 0; t < nsteps; t += 2) {
                    tmem_ld_wait();
                    tmem_ld32(t_s + (t + 1) * 32, vb);
                    m = fmaxf(m, max32(va, t * 32, len));
                    tmem_ld_wait();
                    if (t + 2 < nsteps) tmem_ld32(t_s + (t + 2) * 32, va);
                    m = fmaxf(m, max32(vb, (t + 1) * 32, len));
                }
            }
            if constexpr (!(DIAG & 8)) tmem_ld32(t_s, va);   // first load of pass 2
            // key 0 < len always holds for a unit that is not skipped, so m is finite
            const float msc = m * scale_log2e;
            // ---- pass 2: p = 2^(s scale log2e - max) -> packed 16-bit pairs written back over the score columns (key k ->
            //      column k / 2, always behind this thread's read position); the tensor pipe reads them as the A operand ----
            for (uint32_t c = 0; c < nch; ++c, ++cc) {
                uint32_t o[16];
                if constexpr (!(DIAG & 8)) {
                    tmem_ld_wait();
                    tmem_ld32(t_s + c * kKC + 32, vb);
                    softmax32<BF16, (DIAG & 1) != 0>(va, scale_log2e, msc, c * kKC, len, o);
                    if constexpr (!(DIAG & 4)) tmem_st16(t_s + c * (kKC / 2), o);
                    tmem_ld_wait();
                    if (c + 1 < nch) tmem_ld32(t_s + (c + 1) * kKC, va);
                    softmax32<BF16, (DIAG & 1) != 0>(vb, scale_log2e, msc, c * kKC + 32, len, o);
                    if constexpr (!(DIAG & 4)) tmem_st16(t_s + c * (kKC / 2) + 16, o);
                    tmem_st_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bg + B_P_READY + cc % kPRing);
            }
            // ---- O = P V in columns [128, 128 + DH) of the tile, the row sums of P (P times ones) next to it ----
            mbar_wait(bg + B_O_FULL, it & 1);
            tc_fence_after();
            uint32_t ov[DH], sv[1];
#pragma unroll
            for (int c = 0; c < DH / 32; ++c) tmem_ld32(t_s + Cfg::kOCol + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&ov[c * 32]));
            tmem_ld1(t_s + Cfg::kLCol, sv);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bg + B_S_EMPTY);   // the tile may be overwritten by the next unit's Q K^T
            const float inv = q < len ? 1.0f / __uint_as_float(sv[0]) : 0.f;
            if ((packed ? q < len : q < S) && !(DIAG & 16)) {
#pragma unroll
                for (int j = 0; j < DH / 8; ++j) {
                    uint4 w;
                    w.x = pk2<BF16>(__uint_as_float(ov[8 * j]) * inv, __uint_as_float(ov[8 * j + 1]) * inv);
                    w.y = pk2<BF16>(__uint_as_float(ov[8 * j + 2]) * inv, __uint_as_float(ov[8 * j + 3]) * inv);
                    w.z = pk2<BF16>(__uint_as_float(ov[8 * j + 4]) * inv, __uint_as_float(ov[8 * j + 5]) * inv);
                    w.w = pk2<BF16>(__uint_as_float(ov[8 * j + 6]) * inv, __uint_as_float(ov[8 * j + 7]) * inv);
                    *reinterpret_cast<uint4 *>(orow + j * 8) = w;
                }
            }
            ++it;
        }
    } else if (warp < kSoftmaxWarps + kGroups) {
        // ================= TMA producer of group g =================
        if (lane == 0) {
            const uint32_t g = warp - kSoftmaxWarps;
            uint64_t *bg = bars + g * B_PER_GROUP;
            unsigned char *grp = smem + g * Cfg::kGroupBytes;
            uint32_t it = 0;
            UnitMeta meta = meta_of(blockIdx.x * kGroups + g, n_units, n_qt, heads, S, lens, cu);
            for (uint32_t u = blockIdx.x * kGroups + g; u < n_units; u += kGroups * gridDim.x) {
                const Unit un = decode_unit(u, meta, n_qt, heads, S);
                meta = meta_of(u + kGroups * gridDim.x, n_units, n_qt, heads, S, lens, cu);
                if (un.skip) continue;
                const uint32_t nch = (un.len + kKC - 1) / kKC;
                // packed layout: the boxes may run into the next sequence's rows (finite values; keys >= len are masked,
                // query rows >= len are not stored) or past the end of the buffer (TMA zero fill)
                const int32_t row0 = (int32_t)un.row0;
                const int32_t colq = (int32_t)(un.h * DH);
                // Q and K: free once the previous unit's Q K^T has completed
                if (it >= 1) mbar_wait(bg + B_S_FULL, (it - 1) & 1);
                mbar_arrive_expect_tx(bg + B_QK_FULL, (kQT + nch * kKC) * Cfg::kRowBytes);
#pragma unroll
                for (int i = 0; i < kQT / 64; ++i)
                    tma_load_2d(grp + Cfg::kQOff + i * 64 * Cfg::kRowBytes, &tmQKV, bg + B_QK_FULL, colq,
                                row0 + (int32_t)un.q0 + i * 64, kEvictFirst);
                for (uint32_t i = 0; i < nch; ++i)
                    tma_load_2d(grp + Cfg::kKOff + i * 64 * Cfg::kRowBytes, &tmQKV, bg + B_QK_FULL, (int32_t)H + colq,
                                row0 + (int32_t)i * 64, kEvictLast);
                // V buffer: free once the P V products of the unit that used it have completed
                const uint32_t vb = it % Cfg::kVBufs;
                if (it >= (uint32_t)Cfg::kVBufs) mbar_wait(bg + B_O_FULL, (it - Cfg::kVBufs) & 1);
                mbar_arrive_expect_tx(bg + B_V_FULL + vb, nch * kKC * Cfg::kRowBytes);
                for (uint32_t i = 0; i < nch; ++i)
                    tma_load_2d(grp + Cfg::kVOff + vb * Cfg::kKBytes + i * 64 * Cfg::kRowBytes, &tmQKV, bg + B_V_FULL + vb,
                                2 * (int32_t)H + colq, row0 + (int32_t)i * 64, kEvictLast);
                ++it;
            }
        }
    } else {
        // ================= MMA issuer of group g =================
        if (lane == 0) {
            const uint32_t g = warp - kSoftmaxWarps - kGroups;
            uint64_t *bg = bars + g * B_PER_GROUP;
            const uint32_t grp = smem_u32(smem + g * Cfg::kGroupBytes);
            const uint32_t t_s = tmem_base + g * kMaxKeys;
            constexpr uint32_t fmt = BF16 ? 1u : 0u;
            constexpr uint32_t idesc_pv = make_idesc(kQT, DH, fmt) | (1u << 16);   // B = V is MN-major
            constexpr uint32_t idesc_sum = make_idesc(kQT, 16, fmt);               // B = ones, K-major
            const uint32_t ones = smem_u32(smem + Cfg::kOnesOff);
            uint32_t it = 0, cc = 0;
            UnitMeta meta = meta_of(blockIdx.x * kGroups + g, n_units, n_qt, heads, S, lens, cu);
            for (uint32_t u = blockIdx.x * kGroups + g; u < n_units; u += kGroups * gridDim.x) {
                const Unit un = decode_unit(u, meta, n_qt, heads, S);
                meta = meta_of(u + kGroups * gridDim.x, n_units, n_qt, heads, S, lens, cu);
                if (un.skip) continue;
                const uint32_t nch = (un.len + kKC - 1) / kKC;
                mbar_wait(bg + B_QK_FULL, it & 1);
                if (it >= 1) mbar_wait(bg + B_S_EMPTY, (it - 1) & 1);
                tc_fence_after();
                // S = Q K^T : both operands K-major rows of DH elements
                const uint32_t idesc_s = make_idesc(kQT, nch * kKC, fmt);
#pragma unroll
                for (int k = 0; k < DH / 16; ++k)
                    umma(t_s, att_desc(grp + Cfg::kQOff + k * 32, Cfg::kSbo, Cfg::kLayout),
                         att_desc(grp + Cfg::kKOff + k * 32, Cfg::kSbo, Cfg::kLayout), idesc_s, k != 0 ? 1u : 0u);
                umma_commit(bg + B_S_FULL);
                const uint32_t vb = it % Cfg::kVBufs;
                const uint32_t va0 = grp + Cfg::kVOff + vb * Cfg::kKBytes;
                const uint32_t hold = nch >= 3 ? min(nch - 1, (uint32_t)Cfg::kHoldChunk) : 0u;
                uint32_t issued = 0;
                for (uint32_t c = 0; c < nch; ++c, ++cc) {
                    mbar_wait(bg + B_P_READY + cc % kPRing, (cc / kPRing) & 1);
                    if (c == 0) mbar_wait(bg + B_V_FULL + vb, (it / Cfg::kVBufs) & 1);
                    if (c < hold) continue;
                    tc_fence_after();
                    for (; issued <= c; ++issued) {
#pragma unroll
                        for (int k = 0; k < kKC / 16; ++k) {
                            const uint32_t pa = t_s + issued * (kKC / 2) + k * 8;
                            umma_ts(t_s + Cfg::kOCol, pa, att_desc(va0 + (issued * kKC + k * 16) * Cfg::kRowBytes, Cfg::kSbo, Cfg::kLayout),
                                    idesc_pv, (issued | (uint32_t)k) != 0 ? 1u : 0u);
                            umma_ts(t_s + Cfg::kLCol, pa, make_smem_desc(ones + k * 32), idesc_sum, (issued | (uint32_t)k) != 0 ? 1u : 0u);
                        }
                    }
                }
                umma_commit(bg + B_O_FULL);
                ++it;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kSoftmaxWarps) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// qkv [T, 3H] 16-bit as a 2-D tensor; box = [64 rows][DH columns], swizzle = the row size in bytes
bool make_tmap_qkv(CUtensorMap *out, const void *base, uint64_t T, uint64_t H, uint32_t dh, bool is_bf16)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {3 * H, T};
    cuuint64_t strides[1] = {3 * H * 2};
    cuuint32_t box[2] = {dh, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                    const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    dh * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <bool BF16, int DH>
cudaError_t launch_at(const void *qkv, const int32_t *lens, void *ctx, uint32_t B, uint32_t S, uint32_t H, uint32_t heads,
                      int sm_count, const int32_t *cu, uint32_t n_rows, cudaStream_t st)
{
    using Cfg = AttCfg<DH>;
    auto kern = attention_tc_kernel<BF16, DH>;
    if constexpr (BF16 && DH == 32) {
        static const int diag = getenv("MX_ATTN_DIAG") ? atoi(getenv("MX_ATTN_DIAG")) : 0;
        switch (diag) {
            case 1: kern = attention_tc_kernel<BF16, DH, 1>; break;
            case 2: kern = attention_tc_kernel<BF16, DH, 2>; break;
            case 3: kern = attention_tc_kernel<BF16, DH, 3>; break;
            case 4: kern = attention_tc_kernel<BF16, DH, 4>; break;
            case 7: kern = attention_tc_kernel<BF16, DH, 7>; break;
            case 10: kern = attention_tc_kernel<BF16, DH, 10>; break;
            case 26: kern = attention_tc_kernel<BF16, DH, 26>; break;
            default: break;
        }
    }
    cudaError_t e = set_max_smem(kern, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    CUtensorMap tm;
    if (!make_tmap_qkv(&tm, qkv, cu ? (uint64_t)n_rows : (uint64_t)B * S, H, DH, BF16)) return cudaErrorInvalidValue;
    const uint32_t n_units = B * heads * ceil_div<uint32_t>(S, kQT);
    const uint32_t grid = std::min<uint32_t>((uint32_t)sm_count, ceil_div<uint32_t>(n_units, kGroups));
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)DH);
    LaunchAttrs attrs;
    attrs.pdl();
    e = launch_ex(kern, dim3(grid), dim3(Cfg::kThreads), (size_t)Cfg::kSmemBytes, st, attrs, tm, lens, (uint16_t *)ctx, B, S, H,
                  heads, scale_log2e, cu);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace

bool attention_tc_supported(uint32_t S, uint32_t H, uint32_t heads)
{
    if (heads == 0 || H % heads != 0 || H % 8 != 0) return false;
    const uint32_t dh = H / heads;
    return (dh == 32 || dh == 64) && S >= 1 && S <= (uint32_t)kMaxKeys;
}

// 16-bit activations, head_dim 32 or 64, S <= 256 (attention_tc_supported); same contract as launch_attention_mma
cudaError_t launch_attention_tc(const void *qkv, const int32_t *lens_dev, void *ctx, int act, uint32_t B, uint32_t S,
                                uint32_t H, uint32_t heads, int sm_count, const int32_t *cu, uint32_t n_rows, cudaStream_t st)
{
    if (act == ACT_F32 || !attention_tc_supported(S, H, heads) || B == 0 || (cu && n_rows == 0)) return cudaErrorInvalidValue;
    const bool bf = act == ACT_BF16;
    if (H / heads == 32)
        return bf ? launch_at<true, 32>(qkv, lens_dev, ctx, B, S, H, heads, sm_count, cu, n_rows, st)
                  : launch_at<false, 32>(qkv, lens_dev, ctx, B, S, H, heads, sm_count, cu, n_rows, st);
    return bf ? launch_at<true, 64>(qkv, lens_dev, ctx, B, S, H, heads, sm_count, cu, n_rows, st)
              : launch_at<false, 64>(qkv, lens_dev, ctx, B, S, H, heads, sm_count, cu, n_rows, st);
}

}  // namespace mx
