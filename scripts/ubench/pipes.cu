// pipes.cu -- per-SM throughput of the instruction pipes the fused epilogues / softmax lean on (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
// Each kernel runs ITER dependent-free batches of 8 independent chains per thread; ops / clk / SM is reported
// for 4, 8, 16 and 32 warps per SM (one CTA per SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

constexpr int ITER = 4096;

template <int OP>
__global__ void k(float *out, float seed, long long *clk)
{
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = seed + threadIdx.x * 1e-6f + j * 1e-3f;
    unsigned long long acc2[4] = {0x3f8000003f800000ull, 0x3f8000013f800001ull, 0x3f8000023f800002ull, 0x3f8000033f800003ull};
    uint32_t pk = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < ITER; ++i) {
        if (OP == 0) {          // MUFU.EX2
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
        } else if (OP == 1) {   // FFMA
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[j]) : "f"(1.0001f), "f"(seed));
        } else if (OP == 2) {   // FFMA2 (two f32 per instruction)
#pragma unroll
            for (int j = 0; j < 4; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(acc2[j]) : "l"(0x3f8000003f800000ull));
        } else if (OP == 3) {   // F2FP pack (cvt.rn.bf16x2.f32)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t r;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[j]), "f"(x[(j + 1) & 7]));
                pk ^= r;
            }
        } else if (OP == 4) {   // ex2 + 3 FFMA per element interleaved (softmax-like mix)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[j]) : "f"(0.5f), "f"(seed));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
            }
        } else if (OP == 5) {   // tanh.approx (MUFU.TANH)
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[j]));
        } else if (OP == 6) {   // ex2.approx.f16x2 (two per instruction?)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t r = __float_as_uint(x[j]);
                asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r));
                x[j] = __uint_as_float(r);
            }
        } else if (OP == 7) {   // tanh.approx.bf16x2
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t r = __float_as_uint(x[j]);
                asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(r));
                x[j] = __uint_as_float(r);
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j];
    s += (float)(acc2[0] ^ acc2[1] ^ acc2[2] ^ acc2[3]) + pk;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int per_iter_ops)
{
    float *out;
    long long *clk;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&clk, 148 * 8);
    for (int warps : {4, 8, 16, 32}) {
        k<OP><<<148, warps * 32>>>(out, 0.5f, clk);
        cudaDeviceSynchronize();
        k<OP><<<148, warps * 32>>>(out, 0.5f, clk);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += h[i];
        avg /= 148;
        double ops = (double)ITER * per_iter_ops * warps * 32;
        printf("%-28s warps/SM %2d : %8.2f thread-ops/clk/SM\n", name, warps, ops / avg);
    }
    cudaFree(out);
    cudaFree(clk);
}

int main()
{
    run<0>("MUFU.EX2 f32", 8);
    run<1>("FFMA", 8);
    run<2>("FFMA2 (f32 FMAs)", 8);
    run<3>("F2FP.BF16 pack (packs)", 8);
    run<4>("FFMA+EX2 pairs (ex2 count)", 8);
    run<5>("MUFU.TANH f32", 8);
    run<6>("EX2 f16x2 (instr)", 8);
    run<7>("TANH bf16x2 (instr)", 8);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    return 0;
}
