// sweep2.cuh -- round-2 triangular sweeps: device-side descriptors and the stream fill kernel (the sweep kernel itself, k_sweep2,
// is in kernels.cuh next to the tails it shares with the round-1 kernel; schedule and rationale: sweep2.hpp).
//
// One persistent CTA per PART (pencil of grid lines), all resident, as in round 1; inside the CTA
//   consumer warps : the steps (level sets) of the part are cut into chunks of <= 32 rows (a lane per row) and the chunks of
//                    consecutive steps go round robin to the consumer warps.  A warp walks ITS OWN record stream: values (27 or 36
//                    doubles per block row), 16 bytes of codes per row, both fetched from global memory straight into registers
//                    right after the warp has finished its previous chunk, i.e. several step times before they are used -- the
//                    HBM latency hides behind the other warps' steps and nothing but the dependencies passes through shared
//                    memory.  What is left between two steps of a part is the dependent chain only: bar.sync (step l - 1
//                    complete) -> 6 shared loads of earlier rows -> 27 fma -> 2 shared stores -> bar.arrive.  The global stores
//                    of the result and the next fetch come after the arrive.
//   external rows  : rows owned by other parts travel through L2 (the result vector is armed with a NaN sentinel, the value is
//                    its own ready flag, as in round 1).  The lane that needs one polls it itself: the loads are issued with the
//                    operand fetch, checked -- and repeated while the sentinel is still there -- before the warp waits for the
//                    previous step, and their contribution is folded into the row's start value there, outside the dependent
//                    chain.  No helper warps, no parking ring, no published counters.
// Measured building blocks on B200 (tools/microbench/fp64lat.cu, sweep2_proto.cu): DFMA 8 cycles dependent / 2.3 issue, shared
// load 30, bar.sync 26-40, STS -> bar -> LDS -> DFMA 65-90; a 64-row step takes 310 cycles with 8 groups x 2 warps against 510
// with one group and 850 in the round-1 kernel; 148 CTAs stream at 6.3 TB/s (96 % of the measured copy bandwidth).
#pragma once

namespace b200 {

struct S2PartD { int ncw, nsteps, row0, nrows, stream0, pad0, pad1, pad2; };
struct S2StreamD { long long vals_off, code_off; int hdr_off, nrec; };
struct S2BuildD { long long vals_off; int src_off, cnt, first, pad; };
enum : int { S2D_FIRST = 1, S2D_LAST = 2, S2D_SYNC = 4, S2D_ARRIVE = 8, S2D_EXT = 16, S2D_MULTI = 32 };

constexpr int kS2Header = 128;         // shared: [9] timeout seen, [10] steps traced, [11] steps of the part that have started

struct Sweep2Args {
    const S2PartD* parts;
    const S2StreamD* streams;
    const int2* hdrs;
    const int4* codes;
    const double* vals;
    int window, ncw;                   // rows of the shared-memory window, consumer warps
};

__device__ __forceinline__ double2 ldg_stream_f64x2(const double* p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ldg_stream_s32x4(const int4* p)
{
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void sts_f64x2(unsigned a, double v, double w) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v), "d"(w) : "memory"); }
__device__ __forceinline__ long long globaltimer_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
constexpr long long kS2TimeoutNs = 4000000000LL;      // a wait that lasts 4 s is a deadlock, not a slow neighbour

// Scatter the BSR factor (p-space) into the record streams of one sweep: one warp per record, a lane per row.
//   lower: value 9 j + 3 c + e = L_j[c][e];  upper: (D^-1 U_j)[c][e] and value 27 + 3 c + e = (w D^-1)[c][e] (first record of a row)
template <bool LOWER>
__global__ void __launch_bounds__(256) k_fill_stream2(const S2BuildD* __restrict__ build, int nrec, const int* __restrict__ src,
                                                      const double* __restrict__ LU, double* __restrict__ vals, double relax)
{
    const int lane = threadIdx.x & 31;
    for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < nrec; c += (gridDim.x * blockDim.x) >> 5) {
        const S2BuildD b = build[c];
        if (lane >= b.cnt) continue;
        double* out = vals + b.vals_off + 2 * lane;
        const size_t ps = 2 * (size_t) b.cnt;              // doubles between two pairs of a row
        double inv[9];
        if (!LOWER) {
            const double* d = LU + (size_t) src[b.src_off + 3 * b.cnt + lane] * 9;
#pragma unroll
            for (int e = 0; e < 9; ++e) inv[e] = d[e];
        }
        double v[36];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int k = src[b.src_off + j * b.cnt + lane];
            double u[9];
#pragma unroll
            for (int e = 0; e < 9; ++e) u[e] = k >= 0 ? LU[(size_t) k * 9 + e] : 0.0;
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
#pragma unroll
                for (int e = 0; e < 3; ++e)
                    v[9 * j + 3 * cc + e] = LOWER ? u[3 * cc + e] : inv[3 * cc] * u[e] + inv[3 * cc + 1] * u[3 + e] + inv[3 * cc + 2] * u[6 + e];
        }
        v[27] = 0.0;
        if (!LOWER) {
#pragma unroll
            for (int e = 0; e < 9; ++e) v[27 + e] = b.first ? relax * inv[e] : 0.0;
        }
        constexpr int NP = LOWER ? 14 : 18;
#pragma unroll
        for (int k = 0; k < NP; ++k) *reinterpret_cast<double2*>(out + k * ps) = make_double2(v[2 * k], v[2 * k + 1]);
    }
}

// meta bits of a lane (rows that take several lanes): row within the record (5 bits) | position within the row << 5 | further lanes << 7,
// kept in the low three bits (always zero in a slot code) of the three dependency codes
__device__ __forceinline__ int s2d_meta(const int4& cd) { return (cd.x & 7) | ((cd.x >> 13) & 0x38) | ((cd.y & 7) << 6); }

// the operands of one record in registers
template <bool LOWER>
struct S2Ops {
    double2 v[LOWER ? 14 : 18];
    double r0, r1, r2;
    int4 cd;
};

}  // namespace b200
