// pollrate.cu -- can a thread keep several polling loads of L2-resident lines in flight?  One warp issues K independent loads
// (different 128-byte lines, written by another kernel before: L2 hits, never in L1) of a given flavour and waits for all of
// them; cycles per round against K tells whether the flavour is pipelined or serialised.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o pollrate pollrate.cu && ./pollrate
#include <cstdio>
#include <cuda_runtime.h>
enum { RELAXED_GPU = 0, VOLATILE_ = 1, CG = 2, CV = 3, WEAK = 4, ACQUIRE = 5, RELAXED_SYS = 6, NFL };
static const char* names[] = {"ld.relaxed.gpu", "ld.volatile", "ld.global.cg", "ld.global.cv", "ld.global (weak, L1)", "ld.acquire.gpu", "ld.relaxed.sys"};

template <int F>
__device__ __forceinline__ double ld(const double* p)
{
    double v;
    if (F == RELAXED_GPU) asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    else if (F == VOLATILE_) asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    else if (F == CG) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    else if (F == CV) asm volatile("ld.global.cv.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    else if (F == WEAK) asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    else if (F == ACQUIRE) asm volatile("ld.acquire.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

template <int F, int K>
__global__ void k(const double* buf, int rounds, long long* cyc, double* sink)
{
    const int lane = threadIdx.x;
    double acc = 0.0;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        double v[K];
#pragma unroll
        for (int j = 0; j < K; ++j) v[j] = ld<F>(buf + ((size_t) ((r * K + j) * 32 + lane) * 16));      // a fresh line per lane and load
#pragma unroll
        for (int j = 0; j < K; ++j) acc += v[j];
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    sink[lane] = acc;
}
__global__ void fill(double* b, size_t n) { for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) b[i] = 1.0; }

template <int F, int K>
void run(double* buf, size_t n, long long* cyc, double* sink)
{
    const int rounds = 200;
    fill<<<1024, 256>>>(buf, n);                 // lines land in L2 (written by other SMs)
    cudaDeviceSynchronize();
    k<F, K><<<1, 32>>>(buf, rounds, cyc, sink);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-24s K = %d: %8.1f cycles per round, %7.1f per load  (%s)\n", names[F], K, (double) h / rounds, (double) h / rounds / K, cudaGetErrorString(cudaGetLastError()));
}
template <int F>
void all(double* buf, size_t n, long long* cyc, double* sink) { run<F, 1>(buf, n, cyc, sink); run<F, 2>(buf, n, cyc, sink); run<F, 4>(buf, n, cyc, sink); run<F, 8>(buf, n, cyc, sink); }

int main()
{
    const size_t n = (size_t) 200 * 8 * 32 * 16 + 1024;       // 6.5 MB: stays in L2
    double* buf; long long* cyc; double* sink;
    cudaMalloc(&buf, n * 8); cudaMalloc(&cyc, 8); cudaMalloc(&sink, 256);
    all<RELAXED_GPU>(buf, n, cyc, sink); all<VOLATILE_>(buf, n, cyc, sink); all<CG>(buf, n, cyc, sink); all<CV>(buf, n, cyc, sink);
    all<WEAK>(buf, n, cyc, sink); all<ACQUIRE>(buf, n, cyc, sink); all<RELAXED_SYS>(buf, n, cyc, sink);
    return 0;
}
