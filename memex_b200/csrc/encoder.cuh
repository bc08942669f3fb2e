// encoder.cuh -- launchers of the encoder's non-GEMM kernels (encoder_kernels.cu) and of the
// attention kernels.  ACT selects the activation storage type: 0 = f32 (validation path),
// 1 = bf16, 2 = f16 (tensor-core paths).
#pragma once
#include "common.cuh"

namespace mx {

enum ActType { ACT_F32 = 0, ACT_BF16 = 1, ACT_F16 = 2 };
inline size_t act_size(int act) { return act == ACT_F32 ? 4 : 2; }

// PACKED LAYOUT.  ids arrive padded [B, S]; when `cu` ([B + 1] prefix sums of the clamped lengths, device memory) is not
// null the activations are stored without the padding -- token i of sequence b in row cu[b] + i of [cu[B], H] -- so the
// GEMMs only see real tokens.  cu == nullptr is the padded layout (row b S + i).

// K4: x[row, :] = LayerNorm(word[ids[t]] + pos[t % S] + type[0]) for the padded token index t = b S + i
cudaError_t launch_embed_ln(const int32_t *ids, const float *word, const float *pos, const float *type0,
                            const float *gamma, const float *beta, float eps, void *x, int act, uint32_t n_tokens,
                            uint32_t S, uint32_t H, uint32_t vocab, const int32_t *cu, cudaStream_t st);

// K6 (CUDA-core version): ctx = softmax(q k^T / sqrt(dh) + mask(lens)) v, from the fused qkv buffer
// [T, 3H] (q | k | v, heads contiguous inside each).  Rows at or beyond lens[b] are written as zero.
// scale <= 0 selects 1 / sqrt(dh); rel_bias (may be null): [heads][2 bias_span - 1] f32 added to the score of (query, key)
// at index key - query + bias_span - 1 (T5's relative position bias; S <= bias_span)
cudaError_t launch_attention_simt(const void *qkv, const int32_t *lens_dev, void *ctx, int act, uint32_t B, uint32_t S,
                                  uint32_t H, uint32_t heads, cudaStream_t st, float scale = 0.f,
                                  const float *rel_bias = nullptr, uint32_t bias_span = 0);

// T5 stack: f32 residual stream xr [T, H]; see encoder_kernels.cu
cudaError_t launch_t5_embed(const int32_t *ids, const float *word, float *xr, uint32_t n_tokens, uint32_t H, uint32_t vocab,
                            cudaStream_t st);
cudaError_t launch_t5_add_rmsnorm(float *xr, const void *delta, const float *g, float eps, void *out, int act, uint32_t rows,
                                  uint32_t H, cudaStream_t st);
cudaError_t launch_gated_mul(void *a, const void *b, int act, uint64_t n, cudaStream_t st);

// K6 (tensor-core version, attention_mma.cu): same contract, 16-bit activations only
cudaError_t launch_attention_mma(const void *qkv, const int32_t *lens_dev, void *ctx, int act, uint32_t B, uint32_t S,
                                 uint32_t H, uint32_t heads, cudaStream_t st);

// K6 (tcgen05 version, attention_tc.cu): same contract; head_dim 32 or 64 and S <= 256 only
bool attention_tc_supported(uint32_t S, uint32_t H, uint32_t heads);
// (cu, n_rows): packed layout, qkv / ctx have n_rows = cu[B] rows; cu == nullptr: padded, B S rows
cudaError_t launch_attention_tc(const void *qkv, const int32_t *lens_dev, void *ctx, int act, uint32_t B, uint32_t S,
                                uint32_t H, uint32_t heads, int sm_count, const int32_t *cu, uint32_t n_rows, cudaStream_t st);

// fp32 path only: x = LayerNorm(y + residual) (the GEMM already added the bias)
cudaError_t launch_add_ln_f32(const float *y, const float *residual, const float *gamma, const float *beta, float eps,
                              float *out, uint32_t rows, uint32_t H, cudaStream_t st);

// K10: masked mean-pool over the first lens[b] tokens, optional L2 normalise -> out [B, H] f32
cudaError_t launch_pool_normalize(const void *x, int act, const int32_t *lens_dev, float *out, uint32_t B, uint32_t S,
                                  uint32_t H, uint32_t normalize, const int32_t *cu, cudaStream_t st);

// sentence-transformers Dense module after pooling: out[b, :] = act(W pooled[b, :] + bias) (act 1 = tanh), optional L2
// normalise; W [N, H] f32, bias may be null
cudaError_t launch_dense_tail(const float *pooled, const float *W, const float *bias, float *out, uint32_t B, uint32_t H,
                              uint32_t N, uint32_t act, uint32_t normalize, cudaStream_t st);

// f32 -> 16-bit weight conversion (rows of `cols`, written at out + row * ldo)
cudaError_t launch_convert_weight(const float *src, void *dst, int act, uint64_t numel, cudaStream_t st);

}  // namespace mx
