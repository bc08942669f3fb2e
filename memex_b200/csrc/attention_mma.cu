// attention_mma.cu -- K6: fused masked softmax attention for the 16-bit encoder paths.
//
//   ctx[b, q, h, :] = softmax_k( q . k / sqrt(dh)  + padding mask ) v        k < lens[b]
//
// replaces BertSelfAttention's bmm -> softmax -> bmm (which materialises [B, h, S, S] through
// libtorch under `model.encode(&segments)`, reference lib/libmemex/src/llm/embedding.rs:109) with a
// flash-style kernel: the score tile never leaves registers.
//
// Why warp-level mma.sync here and not tcgen05: with head_dim 32 (MiniLM) the op is bound by the
// exp throughput of the SFU (S*S*heads exponentials per sequence against only 4*S*S*dh flops),
// the tensor pipe idles either way, and keeping S/P in the mma.sync register fragments avoids the
// TMEM -> register -> shared-memory round trip a tcgen05 formulation needs for P.
//
// One CTA = one (sequence, head) x 128 query rows; 8 warps x 16 rows.  K and V of the head sit in
// shared memory (row pitch dh + 8 elements: conflict-free ldmatrix); keys are consumed in blocks of
// 64 with an online softmax (running max / sum in f32, exp2 with the scale folded in).
#include "common.cuh"
#include "encoder.cuh"

namespace mx {

namespace {

constexpr int kAmQ = 128;       // query rows per CTA
constexpr int kAmWarps = 8;
constexpr int kAmKB = 64;       // keys per block

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
template <bool BF16>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    if constexpr (BF16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most `pending` of this thread's most recent groups are still in flight (immediate operand)
__device__ __forceinline__ void cp_async_wait_pending(uint32_t pending)
{
    switch (pending) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
        case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
    }
}
__device__ __forceinline__ float fast_exp2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b)
{
    if constexpr (BF16) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&h);
    } else {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&h);
    }
}

template <bool BF16, int DH>
__global__ void __launch_bounds__(kAmWarps * 32, DH <= 32 ? 3 : 1) attention_mma_kernel(const uint16_t *__restrict__ qkv,
                                                                      const int32_t *__restrict__ lens,
                                                                      uint16_t *__restrict__ ctx, uint32_t S, uint32_t H,
                                                                      uint32_t heads, float scale_log2e)
{
    constexpr int PITCH = DH + 8;          // elements; (DH * 2 + 16) bytes per row
    constexpr int CH = DH / 8;             // 16-byte chunks per row
    constexpr int KS = DH / 16;            // k-steps of q.k
    constexpr int NT = DH / 8;             // n-tiles of the output
    extern __shared__ __align__(16) uint16_t att_smem[];
    const uint32_t b = blockIdx.x / heads, h = blockIdx.x % heads;
    const uint32_t q0 = blockIdx.y * kAmQ;
    const uint32_t len = min((uint32_t)max(lens[b], 0), S);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t ldq = 3 * (size_t)H;
    const uint16_t *base = qkv + (size_t)b * S * ldq + (size_t)h * DH;
    uint16_t *obase = ctx + (size_t)b * S * H + (size_t)h * DH;

    if (q0 >= len) {
        // the whole chunk is padding: zero rows keep the following GEMMs finite
        for (uint32_t i = threadIdx.x; i < kAmQ * CH; i += blockDim.x) {
            const uint32_t qi = q0 + i / CH;
            if (qi < S) *reinterpret_cast<uint4 *>(obase + (size_t)qi * H + (i % CH) * 8) = make_uint4(0, 0, 0, 0);
        }
        return;
    }
    const uint32_t n_kb = (len + kAmKB - 1) / kAmKB;
    const uint32_t n_keys = n_kb * kAmKB;
    uint16_t *Qs = att_smem;                       // [kAmQ][PITCH]
    uint16_t *Ks = Qs + kAmQ * PITCH;              // [n_keys][PITCH]
    uint16_t *Vs = Ks + (size_t)n_keys * PITCH;    // [n_keys][PITCH]

    // asynchronous staging (cp.async, 16 bytes each): group 0 = Q + keys [0, 64), group g = keys [64 g, 64 g + 64).
    // Block kb of the main loop only waits for groups 0..kb, so the later loads overlap the first blocks' math.
    for (uint32_t i = threadIdx.x; i < kAmQ * CH; i += blockDim.x) {
        const uint32_t r = i / CH, c = i % CH, qi = q0 + r;
        uint16_t *dst = Qs + r * PITCH + c * 8;
        if (qi < S)
            cp_async16(dst, base + (size_t)qi * ldq + c * 8);
        else
            *reinterpret_cast<uint4 *>(dst) = make_uint4(0, 0, 0, 0);
    }
    for (uint32_t kb = 0; kb < n_kb; ++kb) {
        for (uint32_t i = threadIdx.x; i < kAmKB * CH; i += blockDim.x) {
            const uint32_t r = kb * kAmKB + i / CH, c = i % CH;
            uint16_t *kd = Ks + r * PITCH + c * 8, *vd = Vs + r * PITCH + c * 8;
            if (r < S) {
                cp_async16(kd, base + (size_t)r * ldq + H + c * 8);
                cp_async16(vd, base + (size_t)r * ldq + 2 * H + c * 8);
            } else {
                *reinterpret_cast<uint4 *>(kd) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4 *>(vd) = make_uint4(0, 0, 0, 0);
            }
        }
        cp_async_commit();
    }

    const uint32_t g = lane >> 2, t = lane & 3;
    const uint32_t row_a = q0 + warp * 16 + g, row_b = row_a + 8;     // the two rows this thread's fragments cover
    // a warp whose 16 rows are all padding still takes part in the block barriers below
    const bool active = q0 + warp * 16 < len;
    if (!active) {
        for (uint32_t i = lane; i < 16 * CH; i += 32) {
            const uint32_t qi = q0 + warp * 16 + i / CH;
            if (qi < S) *reinterpret_cast<uint4 *>(obase + (size_t)qi * H + (i % CH) * 8) = make_uint4(0, 0, 0, 0);
        }
    }

    uint32_t qf[KS][4];   // Q fragments: KS k-steps x (a0a1, a2a3, a4a5, a6a7), loaded once group 0 has landed

    float o[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) o[n][i] = 0.f;
    float m_a = kNegInf, m_b = kNegInf;
    // row sums of P come out of the tensor pipe too: P times a column of ones (every column of this 16 x 8
    // accumulator tile holds the row's sum); the SFU/ALU-bound softmax loop loses 32 adds per key block
    float lsum[4] = {0.f, 0.f, 0.f, 0.f};
    constexpr uint32_t kOnes = BF16 ? 0x3f803f80u : 0x3c003c00u;
    // per-lane shared-memory byte addresses of the ldmatrix rows (advanced by one key block per iteration)
    const uint32_t k_lane = (uint32_t)__cvta_generic_to_shared(Ks + (lane & 7) * PITCH + (lane >> 3) * 8);
    const uint32_t v_lane = (uint32_t)__cvta_generic_to_shared(Vs + ((lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (lane >> 4) * 8);

    for (uint32_t kb = 0; kb < n_kb; ++kb) {
        const uint32_t k0 = kb * kAmKB;
        cp_async_wait_pending(n_kb - 1 - kb);
        __syncthreads();
        if (!active) continue;
        if (kb == 0) {
            const uint32_t mi = lane >> 3;
            const uint32_t r = warp * 16 + (lane & 7) + (mi & 1) * 8;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
                ldmatrix_x4(qf[ks], (uint32_t)__cvta_generic_to_shared(Qs + r * PITCH + ks * 16 + (mi >> 1) * 8));
        }
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int i = 0; i < 4; ++i) s[j][i] = 0.f;
            // B fragments of K for keys k0 + 8j .. +7: matrices (keys, d 0-7) (keys, d 8-15) (keys, d 16-23) (keys, d 24-31)
#pragma unroll
            for (int kk = 0; kk < KS / 2; ++kk) {
                uint32_t kf[4];
                ldmatrix_x4(kf, k_lane + (k0 + 8 * j) * (PITCH * 2) + kk * 64);
                mma16816<BF16>(s[j], qf[2 * kk], kf[0], kf[1]);
                mma16816<BF16>(s[j], qf[2 * kk + 1], kf[2], kf[3]);
            }
        }
        // mask (only a block that straddles len has keys to mask), block row max of the raw scores
        if (k0 + kAmKB > len) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t key = k0 + 8 * j + 2 * t;
                const bool ok0 = key < len, ok1 = key + 1 < len;
                s[j][0] = ok0 ? s[j][0] : kNegInf;
                s[j][1] = ok1 ? s[j][1] : kNegInf;
                s[j][2] = ok0 ? s[j][2] : kNegInf;
                s[j][3] = ok1 ? s[j][3] : kNegInf;
            }
        }
        float mx_a = kNegInf, mx_b = kNegInf;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mx_a = fmaxf(mx_a, fmaxf(s[j][0], s[j][1]));
            mx_b = fmaxf(mx_b, fmaxf(s[j][2], s[j][3]));
        }
        mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1));
        mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
        mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1));
        mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
        // running maxima are kept in the exp2 domain (raw score * scale * log2 e; the scale is positive).
        // key k0 < len always holds inside the loop, so the new maxima are finite
        const float mn_a = fmaxf(m_a, mx_a * scale_log2e), mn_b = fmaxf(m_b, mx_b * scale_log2e);
        const float corr_a = fast_exp2(m_a - mn_a), corr_b = fast_exp2(m_b - mn_b);
        m_a = mn_a;
        m_b = mn_b;
        uint32_t pf[4][4];   // A fragments of P: 4 k-steps of 16 keys
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float p0 = fast_exp2(fmaf(s[j][0], scale_log2e, -mn_a)), p1 = fast_exp2(fmaf(s[j][1], scale_log2e, -mn_a));
            const float p2 = fast_exp2(fmaf(s[j][2], scale_log2e, -mn_b)), p3 = fast_exp2(fmaf(s[j][3], scale_log2e, -mn_b));
            pf[j >> 1][(j & 1) * 2 + 0] = pack2<BF16>(p0, p1);
            pf[j >> 1][(j & 1) * 2 + 1] = pack2<BF16>(p2, p3);
        }
        lsum[0] *= corr_a;
        lsum[2] *= corr_b;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            o[n][0] *= corr_a;
            o[n][1] *= corr_a;
            o[n][2] *= corr_b;
            o[n][3] *= corr_b;
        }
        // O += P V : k = keys (4 steps of 16), n = d (NT tiles of 8, two per ldmatrix)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int np = 0; np < NT / 2; ++np) {
                uint32_t vf[4];
                ldmatrix_x4_trans(vf, v_lane + (k0 + kk * 16) * (PITCH * 2) + np * 32);
                mma16816<BF16>(o[2 * np], pf[kk], vf[0], vf[1]);
                mma16816<BF16>(o[2 * np + 1], pf[kk], vf[2], vf[3]);
            }
            mma16816<BF16>(lsum, pf[kk], kOnes, kOnes);
        }
    }
    if (!active) return;
    const float inv_a = row_a < len ? 1.0f / lsum[0] : 0.f;
    const float inv_b = row_b < len ? 1.0f / lsum[2] : 0.f;
    // the warp's 16 x DH result goes through its own rows of the (now idle) Q tile so that global stores are
    // 16-byte pieces of contiguous rows instead of 4-byte fragments
    uint16_t *qrow = Qs + warp * 16 * PITCH;
    __syncwarp();
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        *reinterpret_cast<uint32_t *>(qrow + g * PITCH + n * 8 + 2 * t) = pack2<BF16>(o[n][0] * inv_a, o[n][1] * inv_a);
        *reinterpret_cast<uint32_t *>(qrow + (g + 8) * PITCH + n * 8 + 2 * t) = pack2<BF16>(o[n][2] * inv_b, o[n][3] * inv_b);
    }
    __syncwarp();
    for (uint32_t i = lane; i < 16 * CH; i += 32) {
        const uint32_t r = i / CH, c = i % CH, qi = q0 + warp * 16 + r;
        if (qi < S) *reinterpret_cast<uint4 *>(obase + (size_t)qi * H + c * 8) = *reinterpret_cast<const uint4 *>(qrow + r * PITCH + c * 8);
    }
}

template <bool BF16, int DH>
cudaError_t launch_am(const void *qkv, const int32_t *lens, void *ctx, uint32_t B, uint32_t S, uint32_t H, uint32_t heads,
                      cudaStream_t st)
{
    auto kern = attention_mma_kernel<BF16, DH>;
    const uint32_t s_pad = ceil_div<uint32_t>(S, kAmKB) * kAmKB;
    const size_t smem = (size_t)(kAmQ + 2 * s_pad) * (DH + 8) * 2;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = set_max_smem(kern, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(B * heads, ceil_div<uint32_t>(S, kAmQ));
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)DH);
    kern<<<grid, kAmWarps * 32, smem, st>>>((const uint16_t *)qkv, lens, (uint16_t *)ctx, S, H, heads, scale_log2e);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

// 16-bit activations only (ACT_BF16 / ACT_F16); H and the qkv pitch must keep 16-byte row alignment
cudaError_t launch_attention_mma(const void *qkv, const int32_t *lens_dev, void *ctx, int act, uint32_t B, uint32_t S,
                                 uint32_t H, uint32_t heads, cudaStream_t st)
{
    const uint32_t dh = H / heads;
    if (act == ACT_F32 || H % 8 != 0) return cudaErrorInvalidValue;
    const bool bf = act == ACT_BF16;
    switch (dh) {
        case 32: return bf ? launch_am<true, 32>(qkv, lens_dev, ctx, B, S, H, heads, st) : launch_am<false, 32>(qkv, lens_dev, ctx, B, S, H, heads, st);
        case 64: return bf ? launch_am<true, 64>(qkv, lens_dev, ctx, B, S, H, heads, st) : launch_am<false, 64>(qkv, lens_dev, ctx, B, S, H, heads, st);
        case 128: return bf ? launch_am<true, 128>(qkv, lens_dev, ctx, B, S, H, heads, st) : launch_am<false, 128>(qkv, lens_dev, ctx, B, S, H, heads, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace mx
