// fp64lat.cu -- latency / issue rate of the building blocks of a sweep level step on one SM (B200).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fp64lat fp64lat.cu && ./fp64lat
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }

template <int KIND>
__global__ void k(double* out, long long* cyc, int iters, double a, double b)
{
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 0.0;
    __syncthreads();
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    float f0 = threadIdx.x;
    unsigned addr = smem_u32(sm) + 8 * (threadIdx.x & 31);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (KIND == 0) {            // dependent DFMA chain, 8 per iteration
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) x0 = fma(x0, a, b);
        } else if (KIND == 1) {     // 8 independent DFMA chains
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        } else if (KIND == 2) {     // dependent DADD chain
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) x0 = x0 + a;
        } else if (KIND == 3) {     // dependent FFMA chain
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) f0 = fmaf(f0, (float) a, (float) b);
        } else if (KIND == 4) {     // LDS -> STS round trip through shared memory (8 per iteration)
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                double v;
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
                unsigned long long u = __double_as_longlong(v) + 1;
                asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(u) : "memory");
            }
        } else if (KIND == 5) {     // bar.sync among all warps of the CTA, 8 per iteration
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) asm volatile("bar.sync 1, %0;" ::"r"(blockDim.x) : "memory");
        } else if (KIND == 6) {     // pointer chase through shared memory (pure LDS latency)
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) { unsigned nx; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(nx) : "r"(addr)); addr += nx; }
        } else if (KIND == 7) {     // __syncwarp
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) __syncwarp();
        } else if (KIND == 8) {     // shuffle chain (64-bit = 2 SHFL)
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) x0 = __shfl_sync(0xffffffffu, x0, (threadIdx.x + 1) & 31);
        } else if (KIND == 9) {     // STS -> bar.sync -> LDS of another warp's value -> DFMA (the level step skeleton), 8 per iteration
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr + 256 * (threadIdx.x >> 5)), "d"(x0) : "memory");
                asm volatile("bar.sync 1, %0;" ::"r"(blockDim.x) : "memory");
                double v;
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr + 256 * (((threadIdx.x >> 5) + 1) % (blockDim.x >> 5))));
                x0 = fma(v, a, b);
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + f0 + (double) addr;
}

template <int KIND>
void run(const char* name, int threads)
{
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) k<KIND><<<1, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-60s %4d threads: %7.1f cycles per op  (%s)\n", name, threads, (double) h / iters / 8, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int t : {32, 64, 128, 256}) {
        run<0>("dependent DFMA chain", t);
        run<1>("8 independent DFMA chains (per DFMA)", t);
        run<2>("dependent DADD chain", t);
        run<3>("dependent FFMA chain", t);
        run<4>("LDS -> STS same address chain", t);
        run<5>("bar.sync (all warps)", t);
        run<6>("LDS pointer chase", t);
        run<7>("__syncwarp", t);
        run<8>("64-bit shuffle chain", t);
        run<9>("STS -> bar.sync -> LDS -> DFMA", t);
    }
    return 0;
}
