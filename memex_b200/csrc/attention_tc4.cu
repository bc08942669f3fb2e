// attention_tc4.cu -- K6 on tcgen05, four-stream form (head_dim 32, up to 256 keys): the same math as attention_tc.cu,
// re-cut so that FOUR independent softmax instruction streams run per SM sub-partition instead of two.
//
//   ctx[b, q, h, :] = softmax_k( q . k / sqrt(dh) + padding mask ) v        k < lens[b]
//   (BertSelfAttention under `model.encode(&segments)`, reference lib/libmemex/src/llm/embedding.rs:109)
//
// Why: attention_tc.cu measures 130 us per MiniLM layer against a 43 us MUFU bound.  One softmax warp needs ~5 k issue
// cycles per 128 x 256 unit (a MUFU holds its warp for 8+ cycles; FFMA, pack, TMEM traffic and barrier round trips come on
// top), and 512 TMEM columns hold only two 256-key score tiles, i.e. two such streams per sub-partition.  Here a
// SUB-UNIT is (sequence, head, 128 query rows, 128-KEY BLOCK): its score tile takes 128 TMEM columns, so four groups fit.
// The two key blocks of a 256-key row are processed by a PAIR of groups independently -- each with its own block
// maximum, as in split-KV decoding -- and combined at the end through shared memory:
//     m = max(m_a, m_b);  O = (2^(m_a - m) O_a + 2^(m_b - m) O_b) / (2^(m_a - m) l_a + 2^(m_b - m) l_b)
// Sequences of up to 128 keys need no pairing: the four groups take four different units.
//
// Per group: 4 softmax warps (thread = query row = TMEM lane) + 1 control warp (TMA loads and MMA issue).  Slot layout
// (128 columns): S f32 [0, 128); P overlays [0, 64) as packed 16-bit pairs (tcgen05.st, read back by the tensor pipe as
// the A operand of P V and P 1, no shared-memory round trip); O [64, 96) and the row sums l [96, 112), written once the
// whole block has been consumed.  V is the MN-major B operand exactly as it lies in qkv.
#include "common.cuh"
#include "encoder.cuh"
#include "tc.cuh"

namespace mx {

using namespace tc;

namespace {

constexpr int kDH = 32;
constexpr int kQT = 128;        // query rows per sub-unit = UMMA M = TMEM lanes
constexpr int kKB = 128;        // keys per sub-unit = TMEM columns of its score tile
constexpr int kKC = 64;         // keys per P V product group (and per TMA box)
constexpr int kGroups = 4;
constexpr int kSoftmaxWarps = 4 * kGroups;
constexpr int kThreads = 32 * (kSoftmaxWarps + kGroups);
constexpr int kOCol = 64, kLCol = 96;

constexpr int kRowBytes = kDH * 2;                   // 64: SWIZZLE_64B rows
constexpr int kQBytes = kQT * kRowBytes;             // 8 KB
constexpr int kKBytes = kKB * kRowBytes;             // 8 KB
constexpr int kGroupBytes = kQBytes + 2 * kKBytes;   // Q | K | V
constexpr int kOnesOff = kGroups * kGroupBytes;      // [16][64] of 1.0, K-major: B operand of the row sums
constexpr int kOnesBytes = 16 * 128;
constexpr int kXchStride = 36;                       // floats per row: O[32], m, l, pad
constexpr int kXchOff = kOnesOff + kOnesBytes;       // f32 [2 pairs][2 key blocks][128 rows][36]
constexpr int kXchBytes = 2 * 2 * kQT * kXchStride * 4;
constexpr int kBarOff = kXchOff + kXchBytes;
constexpr int kSmemBytes = kBarOff + 256 + 1024;
constexpr uint32_t kLayout64 = 4u;                   // UMMA layout type SWIZZLE_64B
constexpr uint32_t kSbo = 8 * kRowBytes;
static_assert(kSmemBytes <= 227 * 1024, "attention tiles do not fit");
static_assert(kGroupBytes % 1024 == 0, "tiles keep the 1024-byte alignment");

enum { B_QK_FULL = 0, B_V_FULL, B_S_FULL, B_P_READY, B_O_FULL, B_S_EMPTY, B_PER_GROUP };

__device__ __forceinline__ uint64_t att4_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ float ex2f(float x)
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

// walk over the sub-units of one group; identical in the group's softmax warps and its control warp
struct Walk4 {
    uint32_t u, stride, n_units, n_qt, heads, S, kb;
    const int32_t *lens;
    int32_t raw_next;
    uint32_t b, h, q0, len, keys, k0, nkb;
    bool skip;       // nothing to compute for this group (padding tile, or the row has no second key block)
    bool zero_fill;  // ... and this group writes the tile's zero rows

    __device__ __forceinline__ void init(uint32_t g, uint32_t gpu, uint32_t n_units_, uint32_t n_qt_, uint32_t heads_, uint32_t S_,
                                         const int32_t *lens_)
    {
        n_units = n_units_, n_qt = n_qt_, heads = heads_, S = S_, lens = lens_;
        const uint32_t streams = kGroups / gpu;
        kb = g % gpu;
        u = blockIdx.x * streams + g / gpu;
        stride = gridDim.x * streams;
        raw_next = u < n_units ? __ldg(lens + u / (n_qt * heads)) : 0;
    }
    __device__ __forceinline__ bool valid() const { return u < n_units; }
    __device__ __forceinline__ void decode()
    {
        const uint32_t qt = u % n_qt, bh = u / n_qt;
        h = bh % heads;
        b = bh / heads;
        q0 = qt * kQT;
        len = min((uint32_t)max(raw_next, 0), S);
        nkb = (len + kKB - 1) / kKB;
        k0 = kb * kKB;
        keys = len > k0 ? min(len - k0, (uint32_t)kKB) : 0u;
        const bool padding = q0 >= len;
        skip = padding || keys == 0;
        zero_fill = padding && kb == 0;
        const uint32_t un = u + stride;
        raw_next = un < n_units ? __ldg(lens + un / (n_qt * heads)) : 0;
    }
    __device__ __forceinline__ void advance() { u += stride; }
};

__device__ __forceinline__ float max32(const uint32_t (&v)[32], uint32_t key0, uint32_t keys)
{
    float m0 = kNegInf, m1 = kNegInf, m2 = kNegInf, m3 = kNegInf;
    if (key0 + 32 <= keys) {
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
            m0 = fmaxf(m0, key0 + j < keys ? __uint_as_float(v[j]) : kNegInf);
            m1 = fmaxf(m1, key0 + j + 1 < keys ? __uint_as_float(v[j + 1]) : kNegInf);
            m2 = fmaxf(m2, key0 + j + 2 < keys ? __uint_as_float(v[j + 2]) : kNegInf);
            m3 = fmaxf(m3, key0 + j + 3 < keys ? __uint_as_float(v[j + 3]) : kNegInf);
        }
    }
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

template <bool BF16>
__device__ __forceinline__ void softmax32(const uint32_t (&v)[32], float sc, float msc, uint32_t key0, uint32_t keys, uint32_t (&o)[16])
{
    if (key0 + 32 <= keys) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            o[j] = pk2<BF16>(ex2f(fmaf(__uint_as_float(v[2 * j]), sc, -msc)), ex2f(fmaf(__uint_as_float(v[2 * j + 1]), sc, -msc)));
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float p0 = key0 + 2 * j < keys ? ex2f(fmaf(__uint_as_float(v[2 * j]), sc, -msc)) : 0.f;
            const float p1 = key0 + 2 * j + 1 < keys ? ex2f(fmaf(__uint_as_float(v[2 * j + 1]), sc, -msc)) : 0.f;
            o[j] = pk2<BF16>(p0, p1);
        }
    }
}

template <bool BF16>
__global__ void __launch_bounds__(kThreads, 1)
attention_tc4_kernel(const __grid_constant__ CUtensorMap tmQKV, const int32_t *__restrict__ lens, uint16_t *__restrict__ ctx,
                     uint32_t B, uint32_t S, uint32_t H, uint32_t heads, float scale_log2e)
{
    constexpr int DH = kDH;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kBarOff);
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + kGroups * B_PER_GROUP);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n_qt = (S + kQT - 1) / kQT;
    const uint32_t n_units = B * heads * n_qt;
    const uint32_t gpu = S > (uint32_t)kKB ? 2u : 1u;   // groups per unit: a pair splits the keys of a long row

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQKV);
        for (int g = 0; g < kGroups; ++g) {
            uint64_t *bg = bars + g * B_PER_GROUP;
            mbar_init(bg + B_QK_FULL, 1);
            mbar_init(bg + B_V_FULL, 1);
            mbar_init(bg + B_S_FULL, 1);
            mbar_init(bg + B_P_READY, 4);
            mbar_init(bg + B_O_FULL, 1);
            mbar_init(bg + B_S_EMPTY, 4);
        }
        fence_barrier_init();
    }
    for (uint32_t i = threadIdx.x; i < kOnesBytes / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(smem + kOnesOff)[i] = BF16 ? 0x3f803f80u : 0x3c003c00u;
    fence_proxy_async_smem();
    if (warp == kSoftmaxWarps) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp < kSoftmaxWarps) {
        // ================= softmax: thread = query row = TMEM lane of group g =================
        const uint32_t g = warp >> 2, quarter = warp & 3;
        uint64_t *bg = bars + g * B_PER_GROUP;
        const uint32_t row = quarter * 32 + lane;
        const uint32_t t_s = tmem_base + ((quarter * 32u) << 16) + g * kKB;
        float *xch = reinterpret_cast<float *>(smem + kXchOff) + (size_t)(g / 2) * (2 * kQT * kXchStride);   // [2][128][36]
        Walk4 w;
        w.init(g, gpu, n_units, n_qt, heads, S, lens);
        uint32_t it = 0;
        for (; w.valid(); w.advance()) {
            w.decode();
            const uint32_t q = w.q0 + row;
            uint16_t *orow = ctx + ((size_t)w.b * S + q) * H + (size_t)w.h * DH;
            if (w.skip) {
                if (w.zero_fill && q < S) {   // padding rows are written as zero: the following GEMMs stay finite
#pragma unroll
                    for (int j = 0; j < DH / 8; ++j) *reinterpret_cast<uint4 *>(orow + j * 8) = make_uint4(0, 0, 0, 0);
                }
                continue;
            }
            const uint32_t keys = w.keys;
            const uint32_t nsteps = (keys + 31) / 32;
            mbar_wait(bg + B_S_FULL, it & 1);
            tc_fence_after();
            uint32_t va[32], vb[32];
            // ---- pass 1: block maximum over the valid keys (the load of step t + 1 is in flight during step t) ----
            float m = kNegInf;
            tmem_ld32(t_s, va);
            for (uint32_t t = 0; t < nsteps; t += 2) {
                tmem_ld_wait();
                if (t + 1 < nsteps) tmem_ld32(t_s + (t + 1) * 32, vb);
                m = fmaxf(m, max32(va, t * 32, keys));
                if (t + 1 < nsteps) {
                    tmem_ld_wait();
                    if (t + 2 < nsteps) tmem_ld32(t_s + (t + 2) * 32, va);
                    m = fmaxf(m, max32(vb, (t + 1) * 32, keys));
                }
            }
            tmem_ld32(t_s, va);
            const float msc = m * scale_log2e;   // finite: the block has at least one valid key
            // ---- pass 2: p = 2^(s scale log2e - max) -> packed pairs over the score columns (behind the read position) ----
            const uint32_t nsteps_p = ((keys + kKC - 1) / kKC) * 2;   // the products consume whole 64-key chunks
            for (uint32_t t = 0; t < nsteps_p; t += 2) {
                uint32_t o[16];
                tmem_ld_wait();
                tmem_ld32(t_s + (t + 1) * 32, vb);
                softmax32<BF16>(va, scale_log2e, msc, t * 32, keys, o);
                tmem_st16(t_s + t * 16, o);
                tmem_ld_wait();
                if (t + 2 < nsteps_p) tmem_ld32(t_s + (t + 2) * 32, va);
                softmax32<BF16>(vb, scale_log2e, msc, (t + 1) * 32, keys, o);
                tmem_st16(t_s + (t + 1) * 16, o);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bg + B_P_READY);
            // ---- O = P V and l = P 1 of this key block ----
            mbar_wait(bg + B_O_FULL, it & 1);
            tc_fence_after();
            uint32_t ov[32], lv[1];
            tmem_ld32(t_s + kOCol, ov);
            tmem_ld1(t_s + kLCol, lv);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bg + B_S_EMPTY);   // the slot may take the next sub-unit's Q K^T
            const float l = __uint_as_float(lv[0]);
            if (w.nkb == 1) {
                const float inv = q < w.len ? 1.0f / l : 0.f;
                if (q < S) {
#pragma unroll
                    for (int j = 0; j < DH / 8; ++j) {
                        uint4 o4;
                        o4.x = pk2<BF16>(__uint_as_float(ov[8 * j]) * inv, __uint_as_float(ov[8 * j + 1]) * inv);
                        o4.y = pk2<BF16>(__uint_as_float(ov[8 * j + 2]) * inv, __uint_as_float(ov[8 * j + 3]) * inv);
                        o4.z = pk2<BF16>(__uint_as_float(ov[8 * j + 4]) * inv, __uint_as_float(ov[8 * j + 5]) * inv);
                        o4.w = pk2<BF16>(__uint_as_float(ov[8 * j + 6]) * inv, __uint_as_float(ov[8 * j + 7]) * inv);
                        *reinterpret_cast<uint4 *>(orow + j * 8) = o4;
                    }
                }
            } else {
                // ---- combine with the partner group's key block; this group finishes columns [16 kb, 16 kb + 16) ----
                const uint32_t kb = w.kb;
                float *mine = xch + ((size_t)kb * kQT + row) * kXchStride;
                const float *theirs = xch + ((size_t)(kb ^ 1u) * kQT + row) * kXchStride;
                bar_sync(1 + g / 2, 256);   // the partner has finished reading the previous unit's exchange
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4 *>(mine + 4 * j) = make_float4(__uint_as_float(ov[4 * j]), __uint_as_float(ov[4 * j + 1]),
                                                                            __uint_as_float(ov[4 * j + 2]), __uint_as_float(ov[4 * j + 3]));
                *reinterpret_cast<float2 *>(mine + 32) = make_float2(msc, l);
                bar_sync(1 + g / 2, 256);
                const float2 ml = *reinterpret_cast<const float2 *>(theirs + 32);
                const float mm = fmaxf(msc, ml.x);
                const float fa = ex2f(msc - mm), fb = ex2f(ml.x - mm);
                const float inv = q < w.len ? 1.0f / fmaf(fa, l, fb * ml.y) : 0.f;
                const float ca = fa * inv, cb = fb * inv;
                if (q < S) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const float4 t0 = *reinterpret_cast<const float4 *>(theirs + kb * 16 + 8 * j);
                        const float4 t1 = *reinterpret_cast<const float4 *>(theirs + kb * 16 + 8 * j + 4);
                        // own values of the same columns: ov[16 kb + 8 j ..] -- selected without dynamic register indexing
                        float x[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(kb ? ov[16 + 8 * j + e] : ov[8 * j + e]);
                        uint4 o4;
                        o4.x = pk2<BF16>(fmaf(ca, x[0], cb * t0.x), fmaf(ca, x[1], cb * t0.y));
                        o4.y = pk2<BF16>(fmaf(ca, x[2], cb * t0.z), fmaf(ca, x[3], cb * t0.w));
                        o4.z = pk2<BF16>(fmaf(ca, x[4], cb * t1.x), fmaf(ca, x[5], cb * t1.y));
                        o4.w = pk2<BF16>(fmaf(ca, x[6], cb * t1.z), fmaf(ca, x[7], cb * t1.w));
                        *reinterpret_cast<uint4 *>(orow + kb * 16 + j * 8) = o4;
                    }
                }
            }
            ++it;
        }
    } else {
        // ================= control warp of group g: TMA loads and MMA issue =================
        if (lane == 0) {
            const uint32_t g = warp - kSoftmaxWarps;
            uint64_t *bg = bars + g * B_PER_GROUP;
            unsigned char *grp = smem + g * kGroupBytes;
            const uint32_t gaddr = smem_u32(grp);
            const uint32_t t_s = tmem_base + g * kKB;
            constexpr uint32_t fmt = BF16 ? 1u : 0u;
            constexpr uint32_t idesc_pv = make_idesc(kQT, DH, fmt) | (1u << 16);   // B = V is MN-major
            constexpr uint32_t idesc_sum = make_idesc(kQT, 16, fmt);               // B = ones, K-major
            const uint32_t ones = smem_u32(smem + kOnesOff);
            Walk4 w, ahead;
            w.init(g, gpu, n_units, n_qt, heads, S, lens);
            ahead = w;
            auto next_ahead = [&]() -> bool {   // next sub-unit this group computes
                while (ahead.valid()) {
                    ahead.decode();
                    ahead.advance();
                    if (!ahead.skip) return true;
                }
                return false;
            };
            auto load_qk = [&](const Walk4 &x) {
                const uint32_t nch = (x.keys + kKC - 1) / kKC;
                const int32_t row0 = (int32_t)(x.b * S), colq = (int32_t)(x.h * DH);
                mbar_arrive_expect_tx(bg + B_QK_FULL, (kQT + nch * kKC) * kRowBytes);
#pragma unroll
                for (int i = 0; i < kQT / 64; ++i)
                    tma_load_2d(grp + i * 64 * kRowBytes, &tmQKV, bg + B_QK_FULL, colq, row0 + (int32_t)x.q0 + i * 64, kEvictLast);
                for (uint32_t i = 0; i < nch; ++i)
                    tma_load_2d(grp + kQBytes + i * 64 * kRowBytes, &tmQKV, bg + B_QK_FULL, (int32_t)H + colq,
                                row0 + (int32_t)(x.k0 + i * 64), kEvictLast);
            };
            auto load_v = [&](const Walk4 &x) {
                const uint32_t nch = (x.keys + kKC - 1) / kKC;
                const int32_t row0 = (int32_t)(x.b * S), colq = (int32_t)(x.h * DH);
                mbar_arrive_expect_tx(bg + B_V_FULL, nch * kKC * kRowBytes);
                for (uint32_t i = 0; i < nch; ++i)
                    tma_load_2d(grp + kQBytes + kKBytes + i * 64 * kRowBytes, &tmQKV, bg + B_V_FULL, 2 * (int32_t)H + colq,
                                row0 + (int32_t)(x.k0 + i * 64), kEvictLast);
            };
            bool have = next_ahead();
            if (have) {
                load_qk(ahead);
                load_v(ahead);
            }
            uint32_t it = 0;
            while (have) {
                const uint32_t nch = (ahead.keys + kKC - 1) / kKC;   // sub-unit `it`
                have = next_ahead();                                 // `ahead` now describes sub-unit it + 1
                mbar_wait(bg + B_QK_FULL, it & 1);
                if (it >= 1) mbar_wait(bg + B_S_EMPTY, (it - 1) & 1);
                tc_fence_after();
                const uint32_t idesc_s = make_idesc(kQT, nch * kKC, fmt);
#pragma unroll
                for (int k = 0; k < DH / 16; ++k)
                    umma(t_s, att4_desc(gaddr + k * 32, kSbo, kLayout64), att4_desc(gaddr + kQBytes + k * 32, kSbo, kLayout64), idesc_s,
                         k != 0 ? 1u : 0u);
                umma_commit(bg + B_S_FULL);
                // Q and K are free once the product has completed: request the next sub-unit's right away
                mbar_wait(bg + B_S_FULL, it & 1);
                if (have) load_qk(ahead);
                mbar_wait(bg + B_P_READY, it & 1);
                mbar_wait(bg + B_V_FULL, it & 1);
                tc_fence_after();
                for (uint32_t c = 0; c < nch; ++c) {
#pragma unroll
                    for (int k = 0; k < kKC / 16; ++k) {
                        const uint32_t pa = t_s + c * (kKC / 2) + k * 8;
                        umma_ts(t_s + kOCol, pa, att4_desc(gaddr + kQBytes + kKBytes + (c * kKC + k * 16) * kRowBytes, kSbo, kLayout64),
                                idesc_pv, (c | (uint32_t)k) != 0 ? 1u : 0u);
                        umma_ts(t_s + kLCol, pa, make_smem_desc(ones + k * 32), idesc_sum, (c | (uint32_t)k) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(bg + B_O_FULL);
                // V is free once these products have completed
                mbar_wait(bg + B_O_FULL, it & 1);
                if (have) load_v(ahead);
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

bool make_tmap_qkv4(CUtensorMap *out, const void *base, uint64_t T, uint64_t H, bool is_bf16)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {3 * H, T};
    cuuint64_t strides[1] = {3 * H * 2};
    cuuint32_t box[2] = {kDH, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                    const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <bool BF16>
cudaError_t launch_at4(const void *qkv, const int32_t *lens, void *ctx, uint32_t B, uint32_t S, uint32_t H, uint32_t heads,
                       int sm_count, cudaStream_t st)
{
    auto kern = attention_tc4_kernel<BF16>;
    cudaError_t e = set_max_smem(kern, kSmemBytes);
    if (e != cudaSuccess) return e;
    CUtensorMap tm;
    if (!make_tmap_qkv4(&tm, qkv, (uint64_t)B * S, H, BF16)) return cudaErrorInvalidValue;
    const uint32_t n_units = B * heads * ceil_div<uint32_t>(S, kQT);
    const uint32_t streams = S > (uint32_t)kKB ? 2u : 4u;
    const uint32_t grid = std::min<uint32_t>((uint32_t)sm_count, ceil_div<uint32_t>(n_units, streams));
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)kDH);
    kern<<<grid, kThreads, kSmemBytes, st>>>(tm, lens, (uint16_t *)ctx, B, S, H, heads, scale_log2e);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

bool attention_tc4_supported(uint32_t S, uint32_t H, uint32_t heads)
{
    if (heads == 0 || H % heads != 0 || H % 8 != 0) return false;
    return H / heads == (uint32_t)kDH && S >= 1 && S <= 2u * kKB;
}

cudaError_t launch_attention_tc4(const void *qkv, const int32_t *lens_dev, void *ctx, int act, uint32_t B, uint32_t S,
                                 uint32_t H, uint32_t heads, int sm_count, cudaStream_t st)
{
    if (act == ACT_F32 || !attention_tc4_supported(S, H, heads) || B == 0) return cudaErrorInvalidValue;
    return act == ACT_BF16 ? launch_at4<true>(qkv, lens_dev, ctx, B, S, H, heads, sm_count, st)
                           : launch_at4<false>(qkv, lens_dev, ctx, B, S, H, heads, sm_count, st);
}

}  // namespace mx
