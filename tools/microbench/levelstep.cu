// levelstep.cu -- what does one level of k_sweep cost on an SM, piece by piece?  One CTA, 8 consumer warps, each with one
// record per level (the c3 situation).  Pieces can be switched off with a bit mask to see what they contribute.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o levelstep levelstep.cu && ./levelstep
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void named_barrier(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void st_relaxed(double* p, double v) { asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ double2 lds2(unsigned a) { double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ double lds1(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ int4 ldsi4(unsigned a) { int4 v; asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void sts1(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

enum { OPERANDS = 1, BARRIER = 2, XLOADS = 4, MATH = 8, STS_ = 16, STG_ = 32, CODES = 64 };

__global__ void k(int mask, int nwork, int iters, double* gout, long long* cycles)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* d = reinterpret_cast<double*>(sm);
    for (int i = threadIdx.x; i < 20000; i += blockDim.x) d[i] = 1.0 + (i & 7) * 1e-3;
    int* ci = reinterpret_cast<int*>(sm + 160000);
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) ci[i] = 32 * ((i * 37) & 255);     // byte offsets of rows
    __syncthreads();
    const unsigned base = smem_u32(sm);
    const unsigned vals = base + 16 * lane + warp * 2304 * 2, xw = base + 100000, codes = base + 160000 + 16 * lane + warp * 512;
    double acc = 0.0, carry = 1.0;
    long long t0 = clock64();
    if (warp < 8) {
        const bool has = warp < nwork;
        for (int i = 0; i < iters; ++i) {
            double2 a01 = {1, 1}, a23 = {1, 1}, a45 = {1, 1}, a67 = {1, 1}; double a8 = 1, r0 = carry;
            int4 cd = {0, 32, 64, 96 + 8 * (lane % 3)};
            if (has) {
                const unsigned v = vals + (i & 1) * 2304;
                if (mask & CODES) cd = ldsi4(codes + (i & 3) * 4096);
                if (mask & OPERANDS) { a01 = lds2(v); a23 = lds2(v + 512); a45 = lds2(v + 1024); a67 = lds2(v + 1536); a8 = lds1(v + 2048 - 8 * lane); r0 = lds1(base + 90000 + 8 * lane + (i & 15) * 256); }
            }
            if (mask & BARRIER) named_barrier(1, 256);
            if (has) {
                double2 x0 = {1, 1}, x1 = {1, 1}, x2 = {1, 1}; double x02 = 1, x12 = 1, x22 = 1;
                if (mask & XLOADS) { x0 = lds2(xw + cd.x); x1 = lds2(xw + cd.y); x2 = lds2(xw + cd.z); x02 = lds1(xw + cd.x + 16); x12 = lds1(xw + cd.y + 16); x22 = lds1(xw + cd.z + 16); }
                if (mask & MATH) {
                    const double t0_ = fma(a23.x, x02, fma(a01.y, x0.y, a01.x * x0.x));
                    const double t1_ = fma(a45.y, x12, fma(a45.x, x1.y, a23.y * x1.x));
                    const double t2_ = fma(a8, x22, fma(a67.y, x2.y, a67.x * x2.x));
                    acc = ((r0 - t0_) - t1_) - t2_;
                } else acc = r0 + x0.x + x1.x + x2.x + x02 + x12 + x22 + a01.x + a23.x + a45.x + a67.x + a8;
                carry = acc * 1e-6;
                if (mask & STS_) sts1(xw + cd.w, acc);
                if (mask & STG_) st_relaxed(gout + ((i * 256 + warp * 32 + lane) & 65535), acc);
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    if (acc == 123.456) gout[0] = acc;
}

int main()
{
    double* gout; long long* cyc;
    cudaMalloc(&gout, 65536 * 8 + 64); cudaMalloc(&cyc, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 180000);
    struct { const char* name; int mask, nwork; } cases[] = {
        {"everything, 8 warps work", 127, 8}, {"everything, 4 warps work", 127, 4}, {"everything, 1 warp works", 127, 1},
        {"no barrier, 8 warps", 127 & ~BARRIER, 8}, {"no barrier, 1 warp", 127 & ~BARRIER, 1},
        {"no global store", 127 & ~STG_, 8}, {"no stores at all", 127 & ~(STG_ | STS_), 8},
        {"no operand/code loads", 127 & ~(OPERANDS | CODES), 8}, {"no x loads", 127 & ~XLOADS, 8}, {"no math", 127 & ~MATH, 8},
        {"barrier + x loads + math + sts only", BARRIER | XLOADS | MATH | STS_, 8}, {"barrier only", BARRIER, 8},
        {"barrier + sts", BARRIER | STS_, 8}, {"barrier + stg", BARRIER | STG_, 8}, {"barrier + x loads", BARRIER | XLOADS, 8}, {"barrier + operands", BARRIER | OPERANDS | CODES, 8},
    };
    for (auto& c : cases) {
        const int iters = 2000;
        for (int rep = 0; rep < 2; ++rep) k<<<1, 352, 180000>>>(c.mask, c.nwork, iters, gout, cyc);
        cudaDeviceSynchronize();
        long long h = 0;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-44s %8.1f cycles/level (%s)\n", c.name, (double) h / iters, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
