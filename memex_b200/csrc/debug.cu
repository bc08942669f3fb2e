// debug.cu -- test-only entry points (include/memex_b200_debug.h): single kernels, device pointers.
#include "../../include/memex_b200_debug.h"
#include "common.cuh"
#include "encoder.cuh"
#include "gemm.cuh"

using namespace mx;

static thread_local std::string g_debug_error;

extern "C" {

const char *mx_debug_last_error(void) { return g_debug_error.c_str(); }

int32_t mx_debug_gemm(const void *A, const void *W, const float *bias, const void *residual, const float *gamma,
                      const float *beta, void *out, uint32_t M, uint32_t N, uint32_t K, uint32_t fmt, uint32_t epi,
                      float ln_eps, int32_t device)
{
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        g_debug_error = cudaGetErrorString(e);
        return MX_ERR_CONNECTION;
    }
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, device);
    GemmParams p{};
    p.A = A;
    p.W = W;
    p.bias = bias;
    p.residual = residual;
    p.gamma = gamma;
    p.beta = beta;
    p.out = out;
    p.M = M;
    p.N = N;
    p.K = K;
    p.lda = K;
    p.ldw = K;
    p.ldr = N;
    p.ldo = N;
    p.ln_eps = ln_eps;
    p.fmt = fmt;
    const char *why = nullptr;
    e = launch_gemm_tc(p, (int)epi, prop.multiProcessorCount, nullptr, &why);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        g_debug_error = std::string(cudaGetErrorString(e)) + (why ? std::string(" (") + why + ")" : "");
        return MX_ERR_ENCODE;
    }
    return MX_OK;
}

// host-only: the staging copy + finiteness check of the host-buffer search calls (no device needed)
int64_t mx_debug_copy_checking_finite(float *dst, const float *src, uint64_t rows, uint64_t dim)
{
    return copy_checking_finite(dst, src, (size_t)rows, (size_t)dim);
}

int32_t mx_debug_attention(const void *qkv, const int32_t *lens_dev, void *ctx, uint32_t B, uint32_t S, uint32_t H,
                           uint32_t heads, uint32_t fmt, uint32_t impl, int32_t device)
{
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) {
        const int act = fmt == 1 ? ACT_BF16 : ACT_F16;
        if (impl == 2) {
            cudaDeviceProp prop;
            e = cudaGetDeviceProperties(&prop, device);
            if (e == cudaSuccess)
                e = attention_tc_supported(S, H, heads)
                        ? launch_attention_tc(qkv, lens_dev, ctx, act, B, S, H, heads, prop.multiProcessorCount, nullptr, 0, nullptr)
                        : cudaErrorNotSupported;
        } else {
            e = impl == 0 ? launch_attention_simt(qkv, lens_dev, ctx, act, B, S, H, heads, nullptr)
                          : launch_attention_mma(qkv, lens_dev, ctx, act, B, S, H, heads, nullptr);
        }
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        g_debug_error = cudaGetErrorString(e);
        return MX_ERR_ENCODE;
    }
    return MX_OK;
}
}
