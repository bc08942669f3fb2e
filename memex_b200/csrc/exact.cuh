// exact.cuh -- the reference's distance arithmetic as a device function (shared by rerank.cu and the exact fallback scan).
//
//   hnsw_rs 0.1.20 DistCosine::eval  (called under reference storage/local.rs:76):
//       products a*b, a*a, b*b in f32, widened to f64, folded left to right in f64;
//       d = max(0, 1 - ab / sqrt(aa * bb)) as f32; either norm zero -> d = 0
// Every operation is the IEEE round-to-nearest one, evaluated in the reference's order -> bit-identical keys.
#pragma once
#include "common.cuh"

namespace mx {

__device__ __forceinline__ float load_elem(const void *rows, uint32_t dtype, size_t idx)
{
    return dtype == MX_DTYPE_F32 ? reinterpret_cast<const float *>(rows)[idx]
                                 : __half2float(reinterpret_cast<const __half *>(rows)[idx]);
}

// exact key of (query, row): cosine distance, or -dot for the dot metric
__device__ inline float exact_key(const float *q, const void *rows, uint32_t dtype, uint32_t metric, size_t row_off,
                                  uint32_t dim, double *aa_out)
{
    double ab = 0.0, aa = 0.0, bb = 0.0;
    if (dtype == MX_DTYPE_F32) {
        const float *r = reinterpret_cast<const float *>(rows) + row_off;
        for (uint32_t i = 0; i < dim; ++i) {
            const float a = q[i], b = r[i];
            ab = __dadd_rn(ab, (double)__fmul_rn(a, b));
            aa = __dadd_rn(aa, (double)__fmul_rn(a, a));
            bb = __dadd_rn(bb, (double)__fmul_rn(b, b));
        }
    } else {
        const __half *r = reinterpret_cast<const __half *>(rows) + row_off;
        for (uint32_t i = 0; i < dim; ++i) {
            const float a = q[i], b = __half2float(r[i]);
            ab = __dadd_rn(ab, (double)__fmul_rn(a, b));
            aa = __dadd_rn(aa, (double)__fmul_rn(a, a));
            bb = __dadd_rn(bb, (double)__fmul_rn(b, b));
        }
    }
    if (aa_out) *aa_out = aa;
    if (metric == MX_METRIC_DOT) return -(float)ab;
    if (aa > 0.0 && bb > 0.0) {
        const double du = __dsub_rn(1.0, __ddiv_rn(ab, __dsqrt_rn(__dmul_rn(aa, bb))));
        return (float)fmax(du, 0.0);
    }
    return 0.f;
}

}  // namespace mx
