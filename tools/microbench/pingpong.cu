// pingpong.cu -- inter-SM signalling latency on B200: one thread in CTA 0 and one in CTA `peer`
// bounce a counter through global memory.  Reports the one-way latency (ns) per access flavour.
// Used to design the dataflow triangular solve.  Build: nvcc -arch=sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

enum Mode { RELAXED = 0, VOLATILE = 1, CG = 2, ATOMIC_POLL = 3, ACQREL = 4, DOUBLE_VAL = 5, NMODES };
static const char* names[] = {"ld/st.relaxed.gpu", "volatile", "ld.cg/st.cg", "atomicAdd(p,0) poll + atomicExch", "ld.acquire/st.release", "f64 value relaxed"};

template <int MODE>
__device__ __forceinline__ unsigned long long ld(const unsigned long long* p)
{
    unsigned long long v;
    if (MODE == RELAXED || MODE == DOUBLE_VAL) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    else if (MODE == VOLATILE) v = *(volatile const unsigned long long*) p;
    else if (MODE == CG) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    else if (MODE == ATOMIC_POLL) v = atomicAdd((unsigned long long*) p, 0ULL);
    else asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
template <int MODE>
__device__ __forceinline__ void st(unsigned long long* p, unsigned long long v)
{
    if (MODE == RELAXED || MODE == DOUBLE_VAL) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    else if (MODE == VOLATILE) *(volatile unsigned long long*) p = v;
    else if (MODE == CG) asm volatile("st.global.cg.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    else if (MODE == ATOMIC_POLL) atomicExch(p, v);
    else asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

template <int MODE>
__global__ void pingpong(unsigned long long* a, unsigned long long* b, int iters, int peer, long long* cycles, int* smids)
{
    if (threadIdx.x != 0) return;
    unsigned smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    if (blockIdx.x == 0) {
        smids[0] = smid;
        long long t0 = clock64();
        for (int i = 1; i <= iters; ++i) {
            st<MODE>(a, (unsigned long long) i);
            while (ld<MODE>(b) != (unsigned long long) i) {}
        }
        *cycles = clock64() - t0;
    } else if ((int) blockIdx.x == peer) {
        smids[1] = smid;
        for (int i = 1; i <= iters; ++i) {
            while (ld<MODE>(a) != (unsigned long long) i) {}
            st<MODE>(b, (unsigned long long) i);
        }
    }
}

// chain: CTA k waits for slot[k-1] then writes slot[k]  (a dependency chain across all SMs, like levels)
__global__ void chain(unsigned long long* slots, int n, long long* cycles)
{
    if (threadIdx.x != 0) return;
    int k = blockIdx.x;
    long long t0 = clock64();
    if (k > 0) {
        unsigned long long v;
        do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(slots + (k - 1) * 16) : "memory"); } while (v == 0);
    }
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(slots + k * 16), "l"(1ULL) : "memory");
    if (k == n - 1) *cycles = clock64() - t0;
}

template <int MODE>
static void run(int peer, int iters, unsigned long long* d, long long* dcyc, int* dsm, double ghz)
{
    cudaMemset(d, 0, 4096);
    pingpong<MODE><<<148, 32>>>(d, d + 64, iters, peer, dcyc, dsm);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0; int sm[2] = {0, 0};
    cudaMemcpy(&cyc, dcyc, sizeof cyc, cudaMemcpyDeviceToHost);
    cudaMemcpy(sm, dsm, sizeof sm, cudaMemcpyDeviceToHost);
    printf("%-36s peer CTA %3d (SM %3d <-> %3d): one-way %.0f cycles = %.0f ns  [%s]\n", names[MODE], peer, sm[0], sm[1],
           (double) cyc / iters / 2, (double) cyc / iters / 2 / ghz, cudaGetErrorString(e));
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz * 1e-6;
    printf("%s, %d SMs, clock %.3f GHz (nominal)\n", p.name, p.multiProcessorCount, ghz);
    unsigned long long* d; long long* dcyc; int* dsm;
    cudaMalloc(&d, 1 << 20); cudaMalloc(&dcyc, 8); cudaMalloc(&dsm, 8);
    const int iters = 2000;
    for (int peer : {1, 2, 37, 74, 110, 147}) {
        run<RELAXED>(peer, iters, d, dcyc, dsm, ghz);
        run<VOLATILE>(peer, iters, d, dcyc, dsm, ghz);
        run<CG>(peer, iters, d, dcyc, dsm, ghz);
        run<ATOMIC_POLL>(peer, iters, d, dcyc, dsm, ghz);
        run<ACQREL>(peer, iters, d, dcyc, dsm, ghz);
    }
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(d, 0, 1 << 20);
        chain<<<148, 32>>>(d, 148, dcyc);
        cudaDeviceSynchronize();
        long long cyc = 0; cudaMemcpy(&cyc, dcyc, sizeof cyc, cudaMemcpyDeviceToHost);
        printf("chain over 148 CTAs: %.0f cycles per hop = %.0f ns\n", (double) cyc / 147, (double) cyc / 147 / ghz);
    }
    return 0;
}
