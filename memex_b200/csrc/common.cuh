// common.cuh -- shared host/device helpers for libmemex_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/memex_b200.h"

namespace mx {

constexpr int kNumSMsDefault = 148;

extern std::atomic<uint64_t> g_launches;      // kernels launched by this library
extern thread_local std::string g_last_error;  // last create/load failure on this thread

// Every handle starts with this so mx_last_error(handle) works for both kinds.
struct HandleBase {
    uint32_t magic = 0;
    std::string last_error;
};
constexpr uint32_t kStoreMagic = 0x4d585354;     // "MXST"
constexpr uint32_t kEmbedderMagic = 0x4d58454d;  // "MXEM"

inline int32_t fail(HandleBase *h, int32_t code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h)
        h->last_error = buf;
    else
        g_last_error = buf;
    return code;
}

#define MX_CUDA(h, code, expr)                                                                  \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return ::mx::fail((h), (code), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                              \
    } while (0)

inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// CUDA-event timing of the library's own kernels on the launching stream.
struct KernelTimer {
    bool on = false;
    struct Span {
        cudaEvent_t a, b;
        int kind;
    };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> pool;
    double total_ms[2] = {0, 0};
    uint64_t launches[2] = {0, 0};

    cudaEvent_t get()
    {
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    // kind 0 = dominant kernel (scan / GEMM), 1 = everything else
    void begin(cudaStream_t st, int kind)
    {
        if (!on) return;
        Span s{get(), get(), kind};
        cudaEventRecord(s.a, st);
        spans.push_back(s);
    }
    void end(cudaStream_t st)
    {
        if (!on) return;
        cudaEventRecord(spans.back().b, st);
    }
    void collect()
    {
        for (Span &s : spans) {
            cudaEventSynchronize(s.b);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, s.a, s.b);
            total_ms[s.kind] += ms;
            launches[s.kind] += 1;
            pool.push_back(s.a);
            pool.push_back(s.b);
        }
        spans.clear();
    }
    void reset()
    {
        collect();
        total_ms[0] = total_ms[1] = 0;
        launches[0] = launches[1] = 0;
    }
    ~KernelTimer()
    {
        for (Span &s : spans) {
            cudaEventDestroy(s.a);
            cudaEventDestroy(s.b);
        }
        for (cudaEvent_t e : pool) cudaEventDestroy(e);
    }
};

template <typename T>
inline T ceil_div(T a, T b)
{
    return (a + b - 1) / b;
}

// One in-flight host-buffer search of the submit / collect pair (two per handle): its own pinned block {queries | ids |
// scores | counts} (the kernels store the answer straight into it), its own device copy of the queries, and the event the
// collect call waits for.
struct IoSlot {
    void *pinned = nullptr;
    void *dev = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;
    uint64_t ticket = 0;
    uint32_t nq = 0, k = 0;
    size_t off_i = 0, off_s = 0, off_c = 0;
    bool busy = false, empty = false;

    cudaError_t reserve(size_t bytes)
    {
        cudaError_t e = cudaSuccess;
        if (!done && (e = cudaEventCreateWithFlags(&done, cudaEventDisableTiming)) != cudaSuccess) return e;
        if (bytes <= cap) return cudaSuccess;
        if (pinned) cudaFreeHost(pinned);
        cudaFree(dev);
        pinned = dev = nullptr;
        cap = 0;
        if ((e = cudaMallocHost(&pinned, bytes)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&dev, bytes)) != cudaSuccess) return e;
        cap = bytes;
        return cudaSuccess;
    }
    void layout(size_t qb, uint32_t nq_, uint32_t k_)
    {
        nq = nq_;
        k = k_;
        off_i = (qb + 255) & ~(size_t)255;
        off_s = off_i + (((size_t)nq * k * sizeof(uint64_t) + 255) & ~(size_t)255);
        off_c = off_s + (((size_t)nq * k * sizeof(float) + 255) & ~(size_t)255);
    }
    size_t total() const { return off_c + (((size_t)nq * 4 + 255) & ~(size_t)255); }
    void copy_out(uint64_t *ids_out, float *scores_out, uint32_t *counts_out) const
    {
        const char *hp = static_cast<const char *>(pinned);
        __builtin_memcpy(ids_out, hp + off_i, (size_t)nq * k * sizeof(uint64_t));
        __builtin_memcpy(scores_out, hp + off_s, (size_t)nq * k * sizeof(float));
        __builtin_memcpy(counts_out, hp + off_c, (size_t)nq * sizeof(uint32_t));
    }
    void release()
    {
        if (pinned) cudaFreeHost(pinned);
        cudaFree(dev);
        if (done) cudaEventDestroy(done);
        pinned = dev = nullptr;
        done = nullptr;
        cap = 0;
        busy = false;
    }
};

// Copies [rows, dim] f32 into `dst` (the pinned staging buffer) and returns the first row holding a non-finite value, or
// -1.  One pass, exponent test on the bit patterns with an OR-accumulator per row so that the compiler vectorises it
// (the scalar isfinite loop with its early exit took 10 us per 64 x 384 block -- as long as the two kernel launches).
inline int64_t copy_checking_finite(float *dst, const float *src, size_t rows, size_t dim)
{
    for (size_t r = 0; r < rows; ++r) {
        const float *s = src + r * dim;
        float *d = dst + r * dim;
        uint32_t bad = 0;
        for (size_t i = 0; i < dim; ++i) {
            uint32_t b;
            __builtin_memcpy(&b, s + i, 4);
            bad |= ((b & 0x7f800000u) == 0x7f800000u) ? 1u : 0u;
            __builtin_memcpy(d + i, &b, 4);
        }
        if (bad) return (int64_t)r;
    }
    return -1;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device) instead of once per launch: the call takes
// ~1 us of host time, which the latency of a small search (two launches) and a 32-launch forward pass both pay for.
cudaError_t set_max_smem_impl(const void *func, int bytes);
template <typename Kern>
inline cudaError_t set_max_smem(Kern kern, int bytes)
{
    return set_max_smem_impl(reinterpret_cast<const void *>(kern), bytes);
}

// Programmatic dependent launch: a kernel launched with this attribute may be SCHEDULED while its predecessor in the
// stream is still running (its CTAs start on SMs the predecessor has vacated, run their prologue -- barrier init, TMEM
// allocation, descriptor prefetch -- and then block in pdl_wait() until the predecessor has completed and its writes are
// visible).  Takes the launch latency and the prologue of every kernel of a chain off the critical path.  MX_PDL=0 turns
// it off (A/B measurements).  A kernel launched WITH the attribute must call pdl_wait() before it reads anything the
// predecessor wrote or writes anything it read.
inline bool pdl_enabled()
{
    static const bool on = getenv("MX_PDL") == nullptr || atoi(getenv("MX_PDL")) != 0;
    return on;
}
struct LaunchAttrs {
    cudaLaunchAttribute a[3];
    unsigned n = 0;
    LaunchAttrs &pdl()
    {
        if (pdl_enabled()) {
            a[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            a[n].val.programmaticStreamSerializationAllowed = 1;
            ++n;
        }
        return *this;
    }
    LaunchAttrs &cluster(unsigned x)
    {
        a[n].id = cudaLaunchAttributeClusterDimension;
        a[n].val.clusterDim.x = x;
        a[n].val.clusterDim.y = 1;
        a[n].val.clusterDim.z = 1;
        ++n;
        return *this;
    }
    LaunchAttrs &cooperative()
    {
        a[n].id = cudaLaunchAttributeCooperative;
        a[n].val.cooperative = 1;
        ++n;
        return *this;
    }
};
template <typename Kern, typename... Args>
inline cudaError_t launch_ex(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, LaunchAttrs &attrs, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = attrs.n ? attrs.a : nullptr;
    cfg.numAttrs = attrs.n;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

}  // namespace mx

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace mx {

constexpr float kNegInf = -__builtin_huge_valf();
constexpr uint32_t kNoRow = 0xffffffffu;

// 128-bit streaming load: read-only path, do not allocate in L1 (rows are touched once)
__device__ __forceinline__ float4 ldg_stream_f4(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// programmatic dependent launch (see LaunchAttrs::pdl): no-ops in a kernel that was launched without the attribute
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// "a ranks before b" for approximate candidates: higher score first, ties -> lower row
__device__ __forceinline__ bool cand_before(float sa, uint32_t ra, float sb, uint32_t rb)
{
    return sa > sb || (sa == sb && ra < rb);
}

// A top-(32*E) list spread over the lanes of one warp: lane l holds E entries.  Unordered; the
// warp-uniform (thr_s, thr_r) is the entry that would be evicted next.
template <int E>
struct WarpTopK {
    float s[E];
    uint32_t r[E];
    float thr_s;
    uint32_t thr_r;

    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            s[e] = kNegInf;
            r[e] = kNoRow;
        }
        thr_s = kNegInf;
        thr_r = kNoRow;
    }
    __device__ __forceinline__ bool accepts(float v, uint32_t row) const
    {
        return cand_before(v, row, thr_s, thr_r);
    }
    // cheap pre-filter used in hot loops
    __device__ __forceinline__ bool maybe(float v) const { return v >= thr_s; }

    // warp-uniform call: (v, row) identical on all lanes, accepts(v,row) already true
    __device__ __forceinline__ void insert(float v, uint32_t row)
    {
        bool mine = false;
#pragma unroll
        for (int e = 0; e < E; ++e) mine |= (s[e] == thr_s && r[e] == thr_r);
        uint32_t holders = __ballot_sync(0xffffffffu, mine);
        int holder = __ffs(holders) - 1;
        if ((int)lane_id() == holder) {
            bool done = false;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                if (!done && s[e] == thr_s && r[e] == thr_r) {
                    s[e] = v;
                    r[e] = row;
                    done = true;
                }
            }
        }
        refresh();
    }
    // recompute the eviction entry: minimal score, ties -> maximal row
    __device__ __forceinline__ void refresh()
    {
        float ms = s[0];
        uint32_t mr = r[0];
#pragma unroll
        for (int e = 1; e < E; ++e) {
            if (cand_before(ms, mr, s[e], r[e])) {
                ms = s[e];
                mr = r[e];
            }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            float os = __shfl_xor_sync(0xffffffffu, ms, o);
            uint32_t orow = __shfl_xor_sync(0xffffffffu, mr, o);
            if (cand_before(ms, mr, os, orow)) {
                ms = os;
                mr = orow;
            }
        }
        thr_s = ms;
        thr_r = mr;
    }
    // offer one candidate per lane (flag = this lane has one); lanes are drained in lane order
    __device__ __forceinline__ void offer(bool flag, float v, uint32_t row)
    {
        uint32_t m = __ballot_sync(0xffffffffu, flag && accepts(v, row));
        while (m) {
            int src = __ffs(m) - 1;
            float cv = __shfl_sync(0xffffffffu, v, src);
            uint32_t cr = __shfl_sync(0xffffffffu, row, src);
            if (accepts(cv, cr)) insert(cv, cr);
            m &= m - 1;
            // the threshold moved: drop lanes that no longer qualify
            m &= __ballot_sync(0xffffffffu, flag && accepts(v, row));
        }
    }
};

}  // namespace mx
#endif
