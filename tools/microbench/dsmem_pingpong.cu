// dsmem_pingpong.cu -- hand-over latency between two CTAs of a thread-block cluster through distributed shared memory
// (st.shared::cluster into the peer's shared memory, the peer polls its OWN shared memory), against the L2 path of pingpong.cu.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dsmem_pingpong dsmem_pingpong.cu && ./dsmem_pingpong
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster(unsigned addr, unsigned long long v) { asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_local(unsigned addr)
{
    unsigned long long v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}

// rank 0 <-> rank `peer` of every cluster bounce a counter; 3 payload doubles travel with it when PAYLOAD (written before the flag)
template <bool PAYLOAD>
__global__ void k_pp(int iters, int peer, long long* cycles, int* smids)
{
    __shared__ __align__(16) unsigned long long box[8];
    cg::cluster_group cl = cg::this_cluster();
    const unsigned rank = cl.block_rank();
    if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) box[i] = 0;
    cl.sync();
    unsigned smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x == 0 && blockIdx.x < cl.num_blocks()) {
        const unsigned mine = smem_u32(box);
        if (rank == 0) {
            smids[0] = smid;
            const unsigned theirs = mapa(mine, peer);
            long long t0 = clock64();
            for (int i = 1; i <= iters; ++i) {
                if (PAYLOAD) { st_cluster(theirs + 8, i); st_cluster(theirs + 16, i); st_cluster(theirs + 24, i); }
                st_cluster(theirs, (unsigned long long) i);
                while (ld_local(mine) != (unsigned long long) i) {}
            }
            *cycles = clock64() - t0;
        } else if ((int) rank == peer) {
            smids[1] = smid;
            const unsigned theirs = mapa(mine, 0);
            for (int i = 1; i <= iters; ++i) {
                while (ld_local(mine) != (unsigned long long) i) {}
                if (PAYLOAD) { st_cluster(theirs + 8, i); st_cluster(theirs + 16, i); st_cluster(theirs + 24, i); }
                st_cluster(theirs, (unsigned long long) i);
            }
        }
    }
    cl.sync();
}

template <bool PAYLOAD>
static void run(int csize, int peer, int nclusters, double ghz, long long* dcyc, int* dsm)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize * nclusters); cfg.blockDim = dim3(32);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (csize > 8) cudaFuncSetAttribute(k_pp<PAYLOAD>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const int iters = 2000;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_pp<PAYLOAD>, iters, peer, dcyc, dsm);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    long long cyc = 0; int sm[2] = {0, 0};
    cudaMemcpy(&cyc, dcyc, sizeof cyc, cudaMemcpyDeviceToHost);
    cudaMemcpy(sm, dsm, sizeof sm, cudaMemcpyDeviceToHost);
    printf("cluster %2d x %3d, rank 0 <-> %2d (SM %3d <-> %3d)%s: one-way %.0f cycles = %.0f ns  [%s]\n", csize, nclusters, peer, sm[0], sm[1],
           PAYLOAD ? " +24 B payload" : "", (double) cyc / iters / 2, (double) cyc / iters / 2 / ghz, cudaGetErrorString(e));
    cudaGetLastError();
}

int main()
{
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    long long* dcyc; int* dsm;
    cudaMalloc(&dcyc, 8); cudaMalloc(&dsm, 8);
    for (int csize : {2, 4, 8, 16})
        for (int peer : {1, csize - 1}) {
            run<false>(csize, peer, 1, ghz, dcyc, dsm);
            run<true>(csize, peer, 1, ghz, dcyc, dsm);
            if (csize == 2) break;
        }
    // how many clusters of each size are co-resident with ~200 KB of shared memory per CTA is a separate question (occupancy API)
    for (int csize : {2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(csize * 148); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 200 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaFuncSetAttribute(k_pp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (csize > 8) cudaFuncSetAttribute(k_pp<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k_pp<false>, &cfg);
        printf("max co-resident clusters of %2d CTAs with 200 KB each: %d (%d SMs)  [%s]\n", csize, n, n * csize, cudaGetErrorString(e));
        cudaGetLastError();
    }
    return 0;
}
