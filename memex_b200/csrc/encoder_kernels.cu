// encoder_kernels.cu -- the encoder's non-GEMM stages and the fp32 validation GEMM.
//
// Restates, op for op, what rust-bert's BertForSentenceEmbeddings + Pooling + Normalize run on
// libtorch for `model.encode(&segments)` (reference lib/libmemex/src/llm/embedding.rs:109):
//   BertEmbeddings : word + position + token_type(0) -> LayerNorm(eps)              (K4)
//   BertSelfAttention: softmax(q k^T / sqrt(dh) + padding mask) v                    (K6, SIMT form)
//   BertSelfOutput / BertOutput: LayerNorm(dense(x) + residual)                      (fp32 path)
//   Pooling(mean over attention_mask) -> Normalize(p=2)                              (K10)
// All statistics, softmax and pooling arithmetic is f32 regardless of the activation type.
#include "common.cuh"
#include "encoder.cuh"
#include "gemm.cuh"

namespace mx {

template <int ACT>
struct Act;
template <>
struct Act<ACT_F32> {
    using T = float;
    __device__ static __forceinline__ float ld(const T *p) { return *p; }
    __device__ static __forceinline__ void st(T *p, float v) { *p = v; }
};
template <>
struct Act<ACT_BF16> {
    using T = __nv_bfloat16;
    __device__ static __forceinline__ float ld(const T *p) { return __bfloat162float(*p); }
    __device__ static __forceinline__ void st(T *p, float v) { *p = __float2bfloat16_rn(v); }
};
template <>
struct Act<ACT_F16> {
    using T = __half;
    __device__ static __forceinline__ float ld(const T *p) { return __half2float(*p); }
    __device__ static __forceinline__ void st(T *p, float v) { *p = __float2half_rn(v); }
};

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

constexpr int kMaxHPerLane = 32;  // H <= 1024

// two-pass LayerNorm of a row held as vals[i] = row[lane + 32 i]
__device__ __forceinline__ void row_layer_norm(float (&vals)[kMaxHPerLane], uint32_t H, float eps, float &mean,
                                               float &rstd)
{
    const uint32_t lane = lane_id();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxHPerLane; ++i)
        if (lane + 32 * i < H) s += vals[i];
    mean = warp_sum(s) / (float)H;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxHPerLane; ++i)
        if (lane + 32 * i < H) {
            const float d = vals[i] - mean;
            q = fmaf(d, d, q);
        }
    rstd = rsqrtf(warp_sum(q) / (float)H + eps);
}

// ---------------------------------------------------------------------------------------------
// K4: embedding gather + LayerNorm, one warp per token
// ---------------------------------------------------------------------------------------------
template <int ACT>
__global__ void __launch_bounds__(256) embed_ln_kernel(const int32_t *__restrict__ ids, const float *__restrict__ word,
                                                       const float *__restrict__ pos, const float *__restrict__ type0,
                                                       const float *__restrict__ gamma, const float *__restrict__ beta,
                                                       float eps, typename Act<ACT>::T *__restrict__ x, uint32_t n_tokens,
                                                       uint32_t S, uint32_t H, uint32_t vocab, const int32_t *__restrict__ cu)
{
    const uint32_t t = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (t >= n_tokens) return;
    const uint32_t lane = lane_id();
    uint32_t orow = t;   // packed layout (cu != null): token i of sequence b goes to row cu[b] + i, padding is dropped
    if (cu) {
        const int32_t lo = cu[t / S], hi = cu[t / S + 1];
        if ((int32_t)(t % S) >= hi - lo) return;
        orow = (uint32_t)lo + t % S;
    }
    uint32_t id = (uint32_t)ids[t];
    if (id >= vocab) id = 0;  // out-of-vocabulary ids read the [PAD] row instead of faulting
    const float *wr = word + (size_t)id * H;
    const float *pr = pos + (size_t)(t % S) * H;
    float vals[kMaxHPerLane];
#pragma unroll
    for (int i = 0; i < kMaxHPerLane; ++i) {
        const uint32_t c = lane + 32 * i;
        vals[i] = c < H ? (wr[c] + type0[c]) + pr[c] : 0.f;   // HF: (inputs_embeds + token_type) + position
    }
    float mean, rstd;
    row_layer_norm(vals, H, eps, mean, rstd);
#pragma unroll
    for (int i = 0; i < kMaxHPerLane; ++i) {
        const uint32_t c = lane + 32 * i;
        if (c < H) Act<ACT>::st(x + (size_t)orow * H + c, (vals[i] - mean) * rstd * gamma[c] + beta[c]);
    }
}

// vectorised form for H = 128 * V (384, 768, ...): a lane owns 4 consecutive columns per 128-column group
template <int ACT, int V>
__global__ void __launch_bounds__(256) embed_ln_vec_kernel(const int32_t *__restrict__ ids, const float *__restrict__ word,
                                                           const float *__restrict__ pos, const float *__restrict__ type0,
                                                           const float *__restrict__ gamma, const float *__restrict__ beta,
                                                           float eps, typename Act<ACT>::T *__restrict__ x, uint32_t n_tokens,
                                                           uint32_t S, uint32_t vocab, const int32_t *__restrict__ cu)
{
    constexpr uint32_t H = 128 * V;
    const uint32_t t = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (t >= n_tokens) return;
    const uint32_t lane = lane_id();
    uint32_t orow = t;   // packed layout: see embed_ln_kernel
    if (cu) {
        const int32_t lo = cu[t / S], hi = cu[t / S + 1];
        if ((int32_t)(t % S) >= hi - lo) return;
        orow = (uint32_t)lo + t % S;
    }
    uint32_t id = (uint32_t)ids[t];
    if (id >= vocab) id = 0;
    const float4 *wr = reinterpret_cast<const float4 *>(word + (size_t)id * H);
    const float4 *pr = reinterpret_cast<const float4 *>(pos + (size_t)(t % S) * H);
    const float4 *ty = reinterpret_cast<const float4 *>(type0);
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float4 a = __ldg(wr + lane + 32 * i), b = __ldg(ty + lane + 32 * i), c = __ldg(pr + lane + 32 * i);
        v[i] = make_float4((a.x + b.x) + c.x, (a.y + b.y) + c.y, (a.z + b.z) + c.z, (a.w + b.w) + c.w);   // HF order
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) / (float)H;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
        q = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, fmaf(dw, dw, q))));
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)H + eps);
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma) + lane + 32 * i);
        const float4 b = __ldg(reinterpret_cast<const float4 *>(beta) + lane + 32 * i);
        typename Act<ACT>::T *o = x + (size_t)orow * H + (lane + 32 * i) * 4;
        const float y0 = (v[i].x - mean) * rstd * g.x + b.x, y1 = (v[i].y - mean) * rstd * g.y + b.y;
        const float y2 = (v[i].z - mean) * rstd * g.z + b.z, y3 = (v[i].w - mean) * rstd * g.w + b.w;
        if constexpr (ACT == ACT_F32) {
            *reinterpret_cast<float4 *>(o) = make_float4(y0, y1, y2, y3);
        } else {
            // one 8-byte store per lane instead of four 2-byte ones
            typename Act<ACT>::T tmp[4];
            Act<ACT>::st(tmp, y0);
            Act<ACT>::st(tmp + 1, y1);
            Act<ACT>::st(tmp + 2, y2);
            Act<ACT>::st(tmp + 3, y3);
            *reinterpret_cast<uint2 *>(o) = *reinterpret_cast<const uint2 *>(tmp);
        }
    }
}

cudaError_t launch_embed_ln(const int32_t *ids, const float *word, const float *pos, const float *type0,
                            const float *gamma, const float *beta, float eps, void *x, int act, uint32_t n_tokens,
                            uint32_t S, uint32_t H, uint32_t vocab, const int32_t *cu, cudaStream_t st)
{
    if (H > 32 * kMaxHPerLane) return cudaErrorInvalidValue;
    const unsigned grid = ceil_div<uint32_t>(n_tokens, 8);
#define MX_LV(A, V)                                                                                                     \
    embed_ln_vec_kernel<A, V><<<grid, 256, 0, st>>>(ids, word, pos, type0, gamma, beta, eps, (typename Act<A>::T *)x, \
                                                    n_tokens, S, vocab, cu)
    if (H == 384 || H == 768) {
        if (act == ACT_F32) { if (H == 384) MX_LV(ACT_F32, 3); else MX_LV(ACT_F32, 6); }
        else if (act == ACT_BF16) { if (H == 384) MX_LV(ACT_BF16, 3); else MX_LV(ACT_BF16, 6); }
        else { if (H == 384) MX_LV(ACT_F16, 3); else MX_LV(ACT_F16, 6); }
        count_launch();
        return cudaGetLastError();
    }
#undef MX_LV
#define MX_L(A)                                                                                              \
    embed_ln_kernel<A><<<grid, 256, 0, st>>>(ids, word, pos, type0, gamma, beta, eps, (typename Act<A>::T *)x, \
                                             n_tokens, S, H, vocab, cu)
    if (act == ACT_F32) MX_L(ACT_F32);
    else if (act == ACT_BF16) MX_L(ACT_BF16);
    else MX_L(ACT_F16);
#undef MX_L
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// K6, SIMT form: one CTA per (sequence, head, 64-query chunk); keys staged through shared memory in
// blocks of 128 with an online softmax; one warp works on 8 queries.
// ---------------------------------------------------------------------------------------------
constexpr int kAttQ = 64;      // queries per CTA
constexpr int kAttKB = 128;    // keys per staged block
constexpr int kAttWarps = 8;
constexpr int kAttQPW = kAttQ / kAttWarps;

template <int ACT, int DH>
__global__ void __launch_bounds__(kAttWarps * 32) attention_simt_kernel(const typename Act<ACT>::T *__restrict__ qkv,
                                                                       const int32_t *__restrict__ lens,
                                                                       typename Act<ACT>::T *__restrict__ ctx, uint32_t S,
                                                                       uint32_t H, uint32_t heads, float scale,
                                                                       const float *__restrict__ rel_bias, uint32_t bias_span)
{
    using T = typename Act<ACT>::T;
    constexpr int DPL = DH / 32;  // output dims per lane
    extern __shared__ __align__(16) float att_smem[];
    float *qs = att_smem;                        // [kAttQ][DH]
    float *ks = qs + kAttQ * DH;                 // [kAttKB][DH + 1]
    float *vs = ks + kAttKB * (DH + 1);          // [kAttKB][DH]

    const uint32_t b = blockIdx.x / heads, h = blockIdx.x % heads;
    const uint32_t q0 = blockIdx.y * kAttQ;
    const uint32_t len = min((uint32_t)max(lens[b], 0), S);
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const size_t ld = 3 * (size_t)H;
    const T *base = qkv + (size_t)b * S * ld + (size_t)h * DH;

    if (q0 >= len) {
        // whole chunk is padding: zero rows keep the following GEMMs finite
        for (uint32_t i = threadIdx.x; i < kAttQ * DH; i += blockDim.x) {
            const uint32_t qi = q0 + i / DH;
            if (qi < S) Act<ACT>::st(ctx + ((size_t)b * S + qi) * H + h * DH + i % DH, 0.f);
        }
        return;
    }
    for (uint32_t i = threadIdx.x; i < kAttQ * DH; i += blockDim.x) {
        const uint32_t qi = q0 + i / DH;
        qs[i] = qi < S ? Act<ACT>::ld(base + (size_t)qi * ld + i % DH) * scale : 0.f;
    }

    float m[kAttQPW], l[kAttQPW], acc[kAttQPW][DPL];
#pragma unroll
    for (int i = 0; i < kAttQPW; ++i) {
        m[i] = kNegInf;
        l[i] = 0.f;
#pragma unroll
        for (int d = 0; d < DPL; ++d) acc[i][d] = 0.f;
    }

    for (uint32_t k0 = 0; k0 < len; k0 += kAttKB) {
        const uint32_t kn = min((uint32_t)kAttKB, len - k0);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < kAttKB * DH; i += blockDim.x) {
            const uint32_t kj = i / DH, d = i % DH;
            float kv = 0.f, vv = 0.f;
            if (kj < kn) {
                kv = Act<ACT>::ld(base + (size_t)(k0 + kj) * ld + H + d);
                vv = Act<ACT>::ld(base + (size_t)(k0 + kj) * ld + 2 * H + d);
            }
            ks[kj * (DH + 1) + d] = kv;
            vs[kj * DH + d] = vv;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kAttQPW; ++i) {
            const float *q = qs + (warp * kAttQPW + i) * DH;
            float sc[kAttKB / 32];
            float bmax = kNegInf;
#pragma unroll
            for (int j = 0; j < kAttKB / 32; ++j) {
                const uint32_t kj = lane + 32 * j;
                float s = 0.f;
#pragma unroll 8
                for (int d = 0; d < DH; ++d) s = fmaf(q[d], ks[kj * (DH + 1) + d], s);
                // T5: additive relative position bias of this head, indexed by (key - query)
                if (rel_bias != nullptr && kj < kn)
                    s += rel_bias[(size_t)h * (2 * bias_span - 1) + (k0 + kj) + (bias_span - 1) - (q0 + warp * kAttQPW + i)];
                sc[j] = kj < kn ? s : kNegInf;
                bmax = fmaxf(bmax, sc[j]);
            }
            bmax = warp_max(bmax);
            const float mnew = fmaxf(m[i], bmax);
            const float corr = __expf(m[i] - mnew);   // m = -inf on the first block -> 0
            float psum = 0.f;
#pragma unroll
            for (int j = 0; j < kAttKB / 32; ++j) {
                sc[j] = __expf(sc[j] - mnew);
                psum += sc[j];
            }
            l[i] = l[i] * corr + warp_sum(psum);
            m[i] = mnew;
#pragma unroll
            for (int d = 0; d < DPL; ++d) acc[i][d] *= corr;
#pragma unroll
            for (int j = 0; j < kAttKB / 32; ++j) {
                for (int src = 0; src < 32; ++src) {
                    const float pj = __shfl_sync(0xffffffffu, sc[j], src);
                    const uint32_t kj = src + 32 * j;
#pragma unroll
                    for (int d = 0; d < DPL; ++d) acc[i][d] = fmaf(pj, vs[kj * DH + lane + 32 * d], acc[i][d]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kAttQPW; ++i) {
        const uint32_t qi = q0 + warp * kAttQPW + i;
        if (qi >= S) continue;
        const bool valid = qi < len;
        const float inv = valid ? 1.0f / l[i] : 0.f;
#pragma unroll
        for (int d = 0; d < DPL; ++d)
            Act<ACT>::st(ctx + ((size_t)b * S + qi) * H + h * DH + lane + 32 * d, valid ? acc[i][d] * inv : 0.f);
    }
}

template <int ACT, int DH>
static cudaError_t launch_att(const void *qkv, const int32_t *lens, void *ctx, uint32_t B, uint32_t S, uint32_t H,
                              uint32_t heads, cudaStream_t st, float scale, const float *rel_bias, uint32_t bias_span)
{
    using T = typename Act<ACT>::T;
    auto kern = attention_simt_kernel<ACT, DH>;
    const size_t smem = sizeof(float) * (kAttQ * DH + kAttKB * (DH + 1) + kAttKB * DH);
    if (smem > 48 * 1024) {
        cudaError_t e = set_max_smem(kern, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(B * heads, ceil_div<uint32_t>(S, kAttQ));
    kern<<<grid, kAttWarps * 32, smem, st>>>((const T *)qkv, lens, (T *)ctx, S, H, heads,
                                             scale > 0.f ? scale : 1.0f / sqrtf((float)DH), rel_bias, bias_span);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_attention_simt(const void *qkv, const int32_t *lens_dev, void *ctx, int act, uint32_t B, uint32_t S,
                                  uint32_t H, uint32_t heads, cudaStream_t st, float scale, const float *rel_bias,
                                  uint32_t bias_span)
{
    const uint32_t dh = H / heads;
    if (rel_bias != nullptr && S > bias_span) return cudaErrorInvalidValue;
#define MX_A(A)                                                                                                  \
    switch (dh) {                                                                                                \
        case 32: return launch_att<A, 32>(qkv, lens_dev, ctx, B, S, H, heads, st, scale, rel_bias, bias_span);   \
        case 64: return launch_att<A, 64>(qkv, lens_dev, ctx, B, S, H, heads, st, scale, rel_bias, bias_span);   \
        case 128: return launch_att<A, 128>(qkv, lens_dev, ctx, B, S, H, heads, st, scale, rel_bias, bias_span); \
        default: return cudaErrorInvalidValue;                                                                   \
    }
    if (act == ACT_F32) { MX_A(ACT_F32) }
    if (act == ACT_BF16) { MX_A(ACT_BF16) }
    MX_A(ACT_F16)
#undef MX_A
}

// ---------------------------------------------------------------------------------------------
// fp32 path: out = LayerNorm(y + residual), one warp per row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) add_ln_f32_kernel(const float *__restrict__ y, const float *__restrict__ residual,
                                                         const float *__restrict__ gamma, const float *__restrict__ beta,
                                                         float eps, float *__restrict__ out, uint32_t rows, uint32_t H)
{
    const uint32_t r = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const uint32_t lane = lane_id();
    float vals[kMaxHPerLane];
#pragma unroll
    for (int i = 0; i < kMaxHPerLane; ++i) {
        const uint32_t c = lane + 32 * i;
        vals[i] = c < H ? y[(size_t)r * H + c] + residual[(size_t)r * H + c] : 0.f;
    }
    float mean, rstd;
    row_layer_norm(vals, H, eps, mean, rstd);
#pragma unroll
    for (int i = 0; i < kMaxHPerLane; ++i) {
        const uint32_t c = lane + 32 * i;
        if (c < H) out[(size_t)r * H + c] = (vals[i] - mean) * rstd * gamma[c] + beta[c];
    }
}

cudaError_t launch_add_ln_f32(const float *y, const float *residual, const float *gamma, const float *beta, float eps,
                              float *out, uint32_t rows, uint32_t H, cudaStream_t st)
{
    if (H > 32 * kMaxHPerLane) return cudaErrorInvalidValue;
    add_ln_f32_kernel<<<ceil_div<uint32_t>(rows, 8), 256, 0, st>>>(y, residual, gamma, beta, eps, out, rows, H);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// K10: masked mean-pool + L2 normalise, one CTA per sequence
// ---------------------------------------------------------------------------------------------
// one CTA per sequence; the 8 warps split the tokens, every lane owns columns lane, lane + 32, ... (coalesced rows),
// partial sums meet in shared memory
template <int ACT>
__global__ void __launch_bounds__(256) pool_normalize_kernel(const typename Act<ACT>::T *__restrict__ x,
                                                             const int32_t *__restrict__ lens, float *__restrict__ out,
                                                             uint32_t S, uint32_t H, uint32_t normalize,
                                                             const int32_t *__restrict__ cu)
{
    __shared__ float part[8][1024];   // H <= 1024
    __shared__ float red[8];
    __shared__ float inv_s;
    const uint32_t b = blockIdx.x;
    const uint32_t len = min((uint32_t)max(lens[b], 0), S);
    const size_t row0 = cu ? (size_t)cu[b] : (size_t)b * S;   // packed layout: the sequence starts at row cu[b]
    const float denom = fmaxf((float)len, 1e-9f);   // sentence-transformers Pooling: clamp(sum_mask, 1e-9)
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    if constexpr (ACT != ACT_F32) {
        // 16-byte loads: a lane owns 8 consecutive columns per 256-column group (H <= 1024 -> 4 groups)
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[i][e] = 0.f;
        const uint32_t chunks = H / 8;   // H % 8 == 0 on the 16-bit paths
        // four tokens per trip: all their loads are issued before the first add (the loop is latency-bound otherwise)
        for (uint32_t t0 = warp; t0 < len; t0 += 32) {
            uint4 u[4][4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t t = t0 + 8 * k;
                const uint4 *row = reinterpret_cast<const uint4 *>(x + (row0 + min(t, len - 1)) * H);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t ch = lane + 32 * i;
                    u[k][i] = (ch < chunks && t < len) ? row[ch] : make_uint4(0, 0, 0, 0);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const typename Act<ACT>::T *h8 = reinterpret_cast<const typename Act<ACT>::T *>(&u[k][i]);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[i][e] += Act<ACT>::ld(h8 + e);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t ch = lane + 32 * i;
            if (ch < chunks)
#pragma unroll
                for (int e = 0; e < 8; ++e) part[warp][ch * 8 + e] = acc[i][e];
        }
    } else {
        float acc[kMaxHPerLane];
#pragma unroll
        for (int i = 0; i < kMaxHPerLane; ++i) acc[i] = 0.f;
        for (uint32_t t = warp; t < len; t += 8) {
            const typename Act<ACT>::T *row = x + (row0 + t) * H;
#pragma unroll
            for (int i = 0; i < kMaxHPerLane; ++i) {
                const uint32_t c = lane + 32 * i;
                if (c < H) acc[i] += Act<ACT>::ld(row + c);
            }
        }
#pragma unroll
        for (int i = 0; i < kMaxHPerLane; ++i) {
            const uint32_t c = lane + 32 * i;
            if (c < H) part[warp][c] = acc[i];
        }
    }
    __syncthreads();
    float sq = 0.f;
    float pooled[4];                                 // H <= 1024 with 256 threads
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t c = threadIdx.x + 256 * i;
        float s = 0.f;
        if (c < H) {
#pragma unroll
            for (int w = 0; w < 8; ++w) s += part[w][c];
        }
        pooled[i] = s / denom;
        sq = fmaf(pooled[i], pooled[i], sq);
    }
    sq = warp_sum(sq);
    if (lane == 0) red[warp] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        inv_s = normalize ? 1.0f / fmaxf(sqrtf(t), 1e-12f) : 1.0f;   // F.normalize(p=2, eps=1e-12)
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t c = threadIdx.x + 256 * i;
        if (c < H) out[(size_t)b * H + c] = pooled[i] * inv_s;
    }
}

cudaError_t launch_pool_normalize(const void *x, int act, const int32_t *lens_dev, float *out, uint32_t B, uint32_t S,
                                  uint32_t H, uint32_t normalize, const int32_t *cu, cudaStream_t st)
{
    if (H > 1024) return cudaErrorInvalidValue;
    if (act == ACT_F32)
        pool_normalize_kernel<ACT_F32><<<B, 256, 0, st>>>((const float *)x, lens_dev, out, S, H, normalize, cu);
    else if (act == ACT_BF16)
        pool_normalize_kernel<ACT_BF16><<<B, 256, 0, st>>>((const __nv_bfloat16 *)x, lens_dev, out, S, H, normalize, cu);
    else
        pool_normalize_kernel<ACT_F16><<<B, 256, 0, st>>>((const __half *)x, lens_dev, out, S, H, normalize, cu);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// sentence-transformers Dense module after pooling (DistiluseBaseMultilingualCased: 768 -> 512, Tanh):
//   out[b, :] = act(W pooled[b, :] + bias), optionally L2-normalised.  One CTA per sequence, the pooled row in shared
// memory, one warp per output column (coalesced f32 weight rows; the matrix is L2-resident: <= 4 MB).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dense_tail_kernel(const float *__restrict__ pooled, const float *__restrict__ W,
                                                         const float *__restrict__ bias, float *__restrict__ out,
                                                         uint32_t H, uint32_t N, uint32_t act, uint32_t normalize)
{
    __shared__ float row[1024];    // H <= 1024
    __shared__ float res[1024];    // N <= 1024
    __shared__ float red[8];
    __shared__ float inv_s;
    const uint32_t b = blockIdx.x, warp = threadIdx.x >> 5, lane = lane_id();
    for (uint32_t c = threadIdx.x; c < H; c += 256) row[c] = pooled[(size_t)b * H + c];
    __syncthreads();
    for (uint32_t n = warp; n < N; n += 8) {
        const float *w = W + (size_t)n * H;
        float acc = 0.f;
        for (uint32_t c = lane; c < H; c += 32) acc = fmaf(w[c], row[c], acc);
        acc = warp_sum(acc);
        if (lane == 0) {
            float v = acc + (bias ? bias[n] : 0.f);
            if (act == 1u) v = tanhf(v);
            res[n] = v;
        }
    }
    __syncthreads();
    float sq = 0.f;
    for (uint32_t n = threadIdx.x; n < N; n += 256) sq = fmaf(res[n], res[n], sq);
    sq = warp_sum(sq);
    if (lane == 0) red[warp] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        inv_s = normalize ? 1.0f / fmaxf(sqrtf(t), 1e-12f) : 1.0f;
    }
    __syncthreads();
    for (uint32_t n = threadIdx.x; n < N; n += 256) out[(size_t)b * N + n] = res[n] * inv_s;
}

cudaError_t launch_dense_tail(const float *pooled, const float *W, const float *bias, float *out, uint32_t B, uint32_t H,
                              uint32_t N, uint32_t act, uint32_t normalize, cudaStream_t st)
{
    if (H > 1024 || N > 1024 || N == 0) return cudaErrorInvalidValue;
    dense_tail_kernel<<<B, 256, 0, st>>>(pooled, W, bias, out, H, N, act, normalize);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
template <int ACT>
__global__ void convert_weight_kernel(const float *__restrict__ src, typename Act<ACT>::T *__restrict__ dst, uint64_t n)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) Act<ACT>::st(dst + i, src[i]);
}

cudaError_t launch_convert_weight(const float *src, void *dst, int act, uint64_t numel, cudaStream_t st)
{
    const unsigned grid = (unsigned)ceil_div<uint64_t>(numel, 256);
    if (act == ACT_F32)
        convert_weight_kernel<ACT_F32><<<grid, 256, 0, st>>>(src, (float *)dst, numel);
    else if (act == ACT_BF16)
        convert_weight_kernel<ACT_BF16><<<grid, 256, 0, st>>>(src, (__nv_bfloat16 *)dst, numel);
    else
        convert_weight_kernel<ACT_F16><<<grid, 256, 0, st>>>(src, (__half *)dst, numel);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// fp32 validation GEMM: out = epi(A . W^T), 64x64 tile, 256 threads x (4x4), K step 16
// ---------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(256) gemm_ref_kernel(GemmRefParams p)
{
    __shared__ float As[16][64 + 4];
    __shared__ float Ws[16][64 + 4];
    const uint32_t tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const uint32_t m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (uint32_t k0 = 0; k0 < p.K; k0 += 16) {
        for (uint32_t i = threadIdx.x; i < 64 * 16; i += 256) {
            const uint32_t r = i / 16, c = i % 16;
            As[c][r] = (m0 + r < p.M && k0 + c < p.K) ? p.A[(size_t)(m0 + r) * p.K + k0 + c] : 0.f;
            Ws[c][r] = (n0 + r < p.N && k0 + c < p.K) ? p.W[(size_t)(n0 + r) * p.K + k0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = As[kk][ty * 4 + i];
                w[i] = Ws[kk][tx * 4 + i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < p.M && n < p.N) {
                float v = acc[i][j] + p.bias[n];
                if (EPI == EPI_BIAS_GELU) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
                if (EPI == EPI_BIAS_GELU_TANH) v = 0.5f * v * (1.0f + tanhf(0.7978845608028654f * (v + 0.044715f * v * v * v)));
                p.out[(size_t)m * p.N + n] = v;
            }
        }
}

cudaError_t launch_gemm_ref(const GemmRefParams &p, int epi, cudaStream_t st)
{
    dim3 grid(ceil_div<uint32_t>(p.N, 64), ceil_div<uint32_t>(p.M, 64));
    if (epi == EPI_BIAS_GELU)
        gemm_ref_kernel<EPI_BIAS_GELU><<<grid, 256, 0, st>>>(p);
    else if (epi == EPI_BIAS_GELU_TANH)
        gemm_ref_kernel<EPI_BIAS_GELU_TANH><<<grid, 256, 0, st>>>(p);
    else
        gemm_ref_kernel<EPI_BIAS><<<grid, 256, 0, st>>>(p);
    count_launch();
    if (epi == EPI_BIAS_RES_LN) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        return launch_add_ln_f32(p.out, p.residual, p.gamma, p.beta, p.ln_eps, p.out, p.M, p.N, st);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// T5 encoder stack (SentenceT5Base, reference lib/libmemex/src/llm/embedding.rs:32,52): pre-RMSNorm layers around an f32
// residual stream.  xr [T, H] f32 is the stream; the GEMMs read / write the activation dtype.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) t5_embed_kernel(const int32_t *__restrict__ ids, const float *__restrict__ word,
                                                       float *__restrict__ xr, uint32_t n_tokens, uint32_t H, uint32_t vocab)
{
    const uint32_t t = blockIdx.x;
    if (t >= n_tokens) return;
    const int32_t id = min(max(ids[t], 0), (int32_t)vocab - 1);
    for (uint32_t c = threadIdx.x; c < H; c += blockDim.x) xr[(size_t)t * H + c] = word[(size_t)id * H + c];
}

cudaError_t launch_t5_embed(const int32_t *ids, const float *word, float *xr, uint32_t n_tokens, uint32_t H, uint32_t vocab,
                            cudaStream_t st)
{
    if (n_tokens == 0) return cudaSuccess;
    t5_embed_kernel<<<n_tokens, 256, 0, st>>>(ids, word, xr, n_tokens, H, vocab);
    count_launch();
    return cudaGetLastError();
}

// xr += delta (the previous sub-layer's output, may be null); out = xr / sqrt(mean(xr^2) + eps) * g   (T5LayerNorm: no mean
// subtraction, no bias).  One warp per row.
template <int ACT>
__global__ void __launch_bounds__(256) t5_add_rmsnorm_kernel(float *__restrict__ xr, const typename Act<ACT>::T *__restrict__ delta,
                                                             const float *__restrict__ g, float eps,
                                                             typename Act<ACT>::T *__restrict__ out, uint32_t rows, uint32_t H)
{
    const uint32_t row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = lane_id();
    if (row >= rows) return;
    float *x = xr + (size_t)row * H;
    float ss = 0.f;
    for (uint32_t c = lane; c < H; c += 32) {
        float v = x[c];
        if (delta != nullptr) {
            v += Act<ACT>::ld(delta + (size_t)row * H + c);
            x[c] = v;
        }
        ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    const float r = rsqrtf(ss / (float)H + eps);
    for (uint32_t c = lane; c < H; c += 32) Act<ACT>::st(out + (size_t)row * H + c, x[c] * r * g[c]);
}

cudaError_t launch_t5_add_rmsnorm(float *xr, const void *delta, const float *g, float eps, void *out, int act, uint32_t rows,
                                  uint32_t H, cudaStream_t st)
{
    if (rows == 0) return cudaSuccess;
    const uint32_t grid = ceil_div<uint32_t>(rows, 8);
    if (act == ACT_F32)
        t5_add_rmsnorm_kernel<ACT_F32><<<grid, 256, 0, st>>>(xr, (const float *)delta, g, eps, (float *)out, rows, H);
    else if (act == ACT_BF16)
        t5_add_rmsnorm_kernel<ACT_BF16><<<grid, 256, 0, st>>>(xr, (const __nv_bfloat16 *)delta, g, eps, (__nv_bfloat16 *)out, rows, H);
    else
        t5_add_rmsnorm_kernel<ACT_F16><<<grid, 256, 0, st>>>(xr, (const __half *)delta, g, eps, (__half *)out, rows, H);
    count_launch();
    return cudaGetLastError();
}

// gated feed-forward: a <- a * b  (a = gelu_new(wi_0 x) from the GEMM epilogue, b = wi_1 x)
template <int ACT>
__global__ void __launch_bounds__(256) gated_mul_kernel(typename Act<ACT>::T *__restrict__ a, const typename Act<ACT>::T *__restrict__ b,
                                                        uint64_t n)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) Act<ACT>::st(a + i, Act<ACT>::ld(a + i) * Act<ACT>::ld(b + i));
}

cudaError_t launch_gated_mul(void *a, const void *b, int act, uint64_t n, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)ceil_div<uint64_t>(n, 256);
    if (act == ACT_F32)
        gated_mul_kernel<ACT_F32><<<grid, 256, 0, st>>>((float *)a, (const float *)b, n);
    else if (act == ACT_BF16)
        gated_mul_kernel<ACT_BF16><<<grid, 256, 0, st>>>((__nv_bfloat16 *)a, (const __nv_bfloat16 *)b, n);
    else
        gated_mul_kernel<ACT_F16><<<grid, 256, 0, st>>>((__half *)a, (const __half *)b, n);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mx
