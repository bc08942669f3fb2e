// gemm.cuh -- interface of the encoder's dense-layer kernels (gemm_tc.cu: tcgen05; gemm_ref.cu:
// fp32 CUDA-core validation path).
#pragma once
#include "common.cuh"

namespace mx {

// EPI_BIAS_GELU_TANH: ALBERT's "gelu_new" (0.5 x (1 + tanh(sqrt(2 / pi) (x + 0.044715 x^3))))
enum GemmEpilogue { EPI_BIAS = 0, EPI_BIAS_GELU = 1, EPI_BIAS_RES_LN = 2, EPI_BIAS_GELU_TANH = 3 };

// out[M, N] = epi(A[M, K] . W[N, K]^T)
struct GemmParams {
    const void *A;              // [M, lda]  16-bit (bf16 or f16, see fmt)
    const void *W;              // [N, ldw]  nn.Linear layout [out, in]
    const float *bias;          // [N]
    const void *residual;       // [M, ldr]  (EPI_BIAS_RES_LN)
    const float *gamma, *beta;  // [N]       (EPI_BIAS_RES_LN)
    void *out;                  // [M, ldo]
    uint32_t M, N, K;
    uint32_t lda, ldw, ldr, ldo;
    float ln_eps;
    uint32_t fmt;               // 1 = bf16, 0 = f16 (the tcgen05 kind::f16 operand format codes)
};

int gemm_tc_block_n(uint32_t N, int epi);  // 0 = shape not supported by the tcgen05 path
cudaError_t launch_gemm_tc(const GemmParams &p, int epi, int sm_count, cudaStream_t st, const char **why);

// fp32 validation path: same contract in float
struct GemmRefParams {
    const float *A, *W, *bias, *residual, *gamma, *beta;
    float *out;
    uint32_t M, N, K;
    float ln_eps;
};
cudaError_t launch_gemm_ref(const GemmRefParams &p, int epi, cudaStream_t st);

}  // namespace mx
