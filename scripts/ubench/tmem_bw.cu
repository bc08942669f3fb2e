// tmem_bw.cu -- TMEM -> register read bandwidth per SM (tcgen05.ld.32x32b.x32), B200 sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
// One CTA per SM; W warps (W = 4, 8, 16) each read the 32 lanes of their quarter, all 512 columns, ITER times.
// `wait_each` = 1 waits after every load (latency-bound), 0 waits once per 16 loads (bandwidth-bound).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

template <int WAIT_EACH>
__global__ void k(uint32_t *out, long long *clk, int iters)
{
    __shared__ uint32_t tmem_ptr;
    const uint32_t warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_ptr + (((warp & 3) * 32u) << 16);
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int c = 0; c < 512; c += 64) {
            uint32_t a[32], b[32];
            tmem_ld32(base + c, a);
            if (WAIT_EACH) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tmem_ld32(base + c + 32, b);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= a[j] ^ b[j];
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_ptr) : "memory");
}

int main()
{
    uint32_t *out;
    long long *clk;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&clk, 148 * 8);
    const int iters = 200;
    for (int wait_each = 0; wait_each < 2; ++wait_each)
        for (int warps : {4, 8, 16}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (wait_each) k<1><<<148, warps * 32>>>(out, clk, iters);
                else k<0><<<148, warps * 32>>>(out, clk, iters);
                cudaDeviceSynchronize();
            }
            long long h[148];
            cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < 148; ++i) avg += h[i];
            avg /= 148;
            const double bytes = (double)iters * 512 * 128 * warps;   // per SM: each warp reads 32 lanes x 512 cols x 4 B
            printf("wait_each %d warps/SM %2d : %8.1f B/clk/SM, %7.1f clk per x32 load per warp\n", wait_each, warps, bytes / avg,
                   avg / (iters * 16.0));
        }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    return 0;
}
