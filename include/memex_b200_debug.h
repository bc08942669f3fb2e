/*
 * memex_b200_debug.h -- test-only entry points of libmemex_b200.so.
 *
 * NOT part of the drop-in boundary (that is memex_b200.h): these expose single kernels so that
 * tests/ can check them in isolation against a plain reference.  All pointers are DEVICE pointers.
 */
#ifndef MEMEX_B200_DEBUG_H
#define MEMEX_B200_DEBUG_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* out[M,N] = epi(A[M,K] . W[N,K]^T) on the tcgen05 path.
 *   fmt 1 = bf16, 0 = f16 (A, W, residual, out);  bias/gamma/beta f32
 *   epi 0 = +bias, 1 = +bias, erf-GELU, 2 = LayerNorm(+bias +residual) (N must be 384 or 768)
 * returns 0 or a negative MX_ERR_* code; synchronous. */
int32_t mx_debug_gemm(const void *A, const void *W, const float *bias, const void *residual, const float *gamma,
                      const float *beta, void *out, uint32_t M, uint32_t N, uint32_t K, uint32_t fmt, uint32_t epi,
                      float ln_eps, int32_t device);
/* ctx[B*S, H] = attention(qkv[B*S, 3H]) with the padding mask lens_dev[B]; fmt as above;
 * impl 0 = CUDA-core kernel, 1 = mma.sync flash kernel (attention_mma.cu), 2 = tcgen05 kernel
 * (attention_tc.cu; head_dim 32 / 64 and S <= 256, else MX_ERR_ENCODE).  synchronous. */
int32_t mx_debug_attention(const void *qkv, const int32_t *lens_dev, void *ctx, uint32_t B, uint32_t S, uint32_t H,
                           uint32_t heads, uint32_t fmt, uint32_t impl, int32_t device);
const char *mx_debug_last_error(void);
/* HOST ONLY (runs without a device): the one-pass staging copy + finiteness check of mx_store_search / _submit and the
 * shard-group calls -- copies [rows, dim] f32 to dst and returns the first row holding a NaN or an infinity, or -1 */
int64_t mx_debug_copy_checking_finite(float *dst, const float *src, uint64_t rows, uint64_t dim);
/* rerank_kernel phase timestamps (%globaltimer ns of CTA 0: start, loads issued, sorted, merged + certified, entries
 * ready, folded, written) of the last search on a store created under MX_RERANK_PROF=1; `store` is an mx_store *. */
int32_t mx_debug_rerank_prof(void *store, uint64_t *out8);
/* scan_tc_kernel phase timestamps of CTA 0 ([0] start, [1] queries prepared, [2] sampled, [3] barrier passed + tau0, [4]
 * sampled tiles scanned again, [5] last tile done, [6] barrier passed; [8 + 4 i ..] the i-th real tile, i < 2: loop top, tau
 * loaded, accumulator ready, tile filtered) of the last tcgen05 scan on a store created under MX_SCAN_TC_PROF=1; 32 values */
int32_t mx_debug_scan_tc_prof(void *store, uint64_t *out32);
/* wall-clock phases (us, summed over `*calls` calls) of mx_store_search under MX_HOST_PROF=1: [0] finiteness check,
 * [1] device + staging buffers, [2] copy into pinned memory, [3] H2D enqueue, [4] kernel launches, [5] D2H enqueue,
 * [6] stream synchronise (= the device's work), [7] copy out; reset != 0 zeroes the counters */
int32_t mx_debug_host_prof(double *out12, uint64_t *calls, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif
