// kernels.cuh -- hand-written sm_100a fp64 kernels of the ILU0-BiCGSTAB hot path.
//
// Everything on the device lives in "p-space": block rows renumbered level by level (analysis.hpp),
// so the triangular sweeps stream the factor linearly and every gather stays within +-1 level.
// All kernels are HBM/latency bound (18 flop per 76 B in the SpMV); no tensor cores are used.
//
// Reference semantics restated by each kernel (paths relative to opm/simulators/linalg/):
//   k_spmv            y = A x, b[3r+i] = sum_blk sum_j vals[blk*9+3i+j] x[3 col+j]  bda/openclKernels.cpp:155-221
//   k_ilu_factor*     left-looking block ILU0, stored inverse pivot                 ParallelOverlappingILU0.hpp:440-494,
//                                                                                   bda/openclKernels.cpp:480-617
//   k_trsv<true/false> unit-lower forward / upper backward + inverse-pivot multiply ParallelOverlappingILU0.hpp:867-901,
//                                                                                   bda/openclKernels.cpp:225-383
//   k_vec_*           BiCGSTAB vector updates and dot products                      bda/cusparseSolverBackend.cu:60-184
//   k_wells           y -= C^T (D^-1 (B x)), full perforation loop                  wells/StandardWell_impl.hpp:1251-1277,
//                                                                                   bda/WellContributions.cu:36-126
#pragma once
#include <climits>
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr unsigned long long kSentinel = 0x7FF8DEADBEEF0B20ULL;   // quiet-NaN payload: "not computed yet"
constexpr unsigned kFull = 0xffffffffu;
constexpr int kVecThreads = 256;
constexpr int kPadColD = (int) 0x80000000;   // == analysis.hpp kPadCol (INT_MIN)
constexpr int kMaxPartials = 4096;     // per reduced quantity

// Device-resident Krylov state: scalars never round-trip through the host inside the loop.
struct Scalars {
    double rho, rho_new, alpha, omega, h, tr, tt, norm, norm0, tol;
    double red[2];      // multi-GPU: raw local sums waiting for the all-reduce (k_finish consumes them)
    int it_half, done, converged, breakdown, first, max_half, trsv_timeout, singular;
    // deferred solution update: x += pend * y has not been applied yet (done by the idle CTAs of the next lower sweep, or
    // folded into the final scatter when the solve ends first)
    int pend_on, pad_;
    double pend;
};

__device__ __forceinline__ double ld_relaxed(const double* p)
{
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(double* p, double v)
{
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// Waits of the dataflow (rows of other parts, peers' flags) give up on ELAPSED TIME, not on spin counts: under MPS, time
// slicing or a debugger a count fires spuriously.  4 s without progress is a deadlock, not a slow neighbour.
constexpr long long kWaitTimeoutNs = 4000000000LL;
__device__ __forceinline__ long long wall_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// one register instead of two where registers are short (k_sweep2): the low word wraps every 4.29 s, differences of two reads
// taken less than that apart are exact
constexpr unsigned kWaitTimeoutLoNs = 3000000000u;
__device__ __forceinline__ unsigned wall_lo_ns() { unsigned t; asm volatile("mov.u32 %0, %%globaltimer_lo;" : "=r"(t)); return t; }
__device__ __forceinline__ bool is_sentinel(double v) { return (unsigned long long) __double_as_longlong(v) == kSentinel; }
__device__ __forceinline__ double sentinel() { return __longlong_as_double((long long) kSentinel); }

// Programmatic dependent launch (kernels of the BiCGSTAB iteration): let the next kernel of the stream be scheduled while this
// one runs, and wait for the previous one to complete (its stores visible) before touching anything it wrote.  Both are
// no-ops in a launch without the attribute.  A kernel launched WITH the attribute must call this before its first access.
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Deterministic grid reduction of NV doubles: warp shuffles -> shared memory -> one partial per
// block -> the LAST block to finish sums the partials in a fixed order.  Returns true in thread 0
// of that last block with the totals in out[].  Fixed grid => bit-reproducible results.
template <int NV>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double* __restrict__ partials, unsigned* ticket,
                                            double (&out)[NV])
{
    __shared__ double sm[NV][32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = warp_sum(v[i]);
        if (lane == 0) sm[i][warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double x = lane < nwarp ? sm[i][lane] : 0.0;
            x = warp_sum(x);
            if (lane == 0) partials[i * kMaxPartials + blockIdx.x] = x;
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = 0.0;
        for (int b = threadIdx.x; b < (int) gridDim.x; b += blockDim.x) x += __ldcg(partials + i * kMaxPartials + b);
        x = warp_sum(x);
        __syncthreads();
        if (lane == 0) sm[i][warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double x = lane < nwarp ? sm[i][lane] : 0.0;
            out[i] = warp_sum(x);
        }
    }
    if (threadIdx.x == 0) *ticket = 0;
    return threadIdx.x == 0;
}

// ---- permutation into p-space ---------------------------------------------------------------

// A_p[q] = stage[srcblk[q]]  (72-byte block gather; one thread per scalar)
__global__ void __launch_bounds__(256) k_permute_vals(const double* __restrict__ stage, const int* __restrict__ srcblk,
                                                      double* __restrict__ A, long long nnz)
{
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (long long) gridDim.x * blockDim.x) {
        long long q = i / 9;
        int e = (int) (i - q * 9);
        A[i] = __ldg(stage + (long long) __ldg(srcblk + q) * 9 + e);
    }
}

__global__ void __launch_bounds__(256) k_gather_vec(const double* __restrict__ in_nat, const int* __restrict__ perm,
                                                    double* __restrict__ out_p, int N)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        int q = i / 3;
        out_p[i] = in_nat[3 * perm[q] + (i - 3 * q)];
    }
}
// the solution, natural order, including a solution update that is still pending (Scalars::pend_on)
__global__ void __launch_bounds__(256) k_scatter_solution(const double* __restrict__ x_p, const double* __restrict__ y_p,
                                                          const int* __restrict__ perm, double* __restrict__ out_nat, int N,
                                                          const Scalars* __restrict__ S)
{
    const bool on = S->pend_on != 0;
    const double a = S->pend;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        int q = i / 3;
        out_nat[3 * perm[q] + (i - 3 * q)] = on ? x_p[i] + a * y_p[i] : x_p[i];
    }
}
__global__ void __launch_bounds__(256) k_scatter_vec(const double* __restrict__ in_p, const int* __restrict__ perm,
                                                     double* __restrict__ out_nat, int N)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        int q = i / 3;
        out_nat[3 * perm[q] + (i - 3 * q)] = in_p[i];
    }
}
__global__ void __launch_bounds__(256) k_fill(double* __restrict__ a, double v, int N)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) a[i] = v;
}

// Scalar epilogues of the three reducing vector phases.  Single GPU: run by the last block of the phase
// itself.  Multi-GPU: the phase leaves its raw local sums in S->red, an all-reduce sums them over the
// ranks in place, and k_finish runs the same epilogue -- every rank takes the same decisions from the
// same bits (cusparseSolverBackend.cu:92-176 for the formulas).
__device__ __forceinline__ void finish_init(Scalars* S, double rr, double tol, int max_half)
{
    S->norm0 = sqrt(rr); S->norm = S->norm0; S->rho_new = rr;
    S->rho = 1.0; S->alpha = 1.0; S->omega = 1.0; S->h = 0.0; S->tr = 0.0; S->tt = 0.0;
    S->tol = tol; S->it_half = 0; S->converged = 0; S->breakdown = 0; S->first = 1;
    S->max_half = max_half; S->trsv_timeout = 0; S->pend_on = 0; S->pend = 0.0;
    // Dune: norm0 already below the absolute floor -> converged with 0 iterations
    S->done = (S->norm0 < 1e-30) ? 1 : 0;
    if (S->done) S->converged = 1;
    if (S->singular) S->done = 1;      // factorisation failed: skip the Krylov loop
}
__device__ __forceinline__ void finish_xr1(Scalars* S, double rr)
{
    S->first = 0;
    S->norm = sqrt(rr); S->it_half += 1;
    if (S->norm < S->tol * S->norm0) { S->converged = 1; S->done = 1; }
}
__device__ __forceinline__ void finish_xr2(Scalars* S, double rr, double rtr)
{
    S->rho = S->rho_new; S->rho_new = rtr;
    S->norm = sqrt(rr); S->it_half += 1;
    if (S->norm < S->tol * S->norm0 || S->norm < 1e-30) { S->converged = 1; S->done = 1; }
    else if (fabs(S->rho_new) <= 1e-80 || fabs(S->omega) <= 1e-80 || !(S->norm == S->norm)) { S->breakdown = 1; S->done = 1; }      // the rho of the NEXT iteration (Dune: abs(rho) <= EPSILON right after it is computed)
    else if (S->it_half >= S->max_half) S->done = 1;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned* p, unsigned v)
{
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Multi-GPU all-reduce of a phase's scalars over peer-memory mailboxes (layout: k_allreduce_p2p below), run by ONE thread:
// the thread of the producing kernel that has just summed the local partials (the last block of its grid reduction), so the
// exchange needs no launch of its own.  Stores to the `world` mailboxes are posted, one system fence, the flags, then the
// contributions are added in rank order -- every rank gets the same bits.
struct MailD { unsigned* flags[64]; double* vals[64]; };   // mapped mail flags / values of every rank (own included)
struct DistRedD { const MailD* mail; unsigned* seq; int rank, world; double tol; int max_half; };      // mail == nullptr: not fused
template <int NV>
__device__ __noinline__ void mail_allreduce(const DistRedD D, Scalars* S, double (&v)[NV])
{
    const MailD* M = D.mail;
    const unsigned seq = *D.seq + 1u;
    *D.seq = seq;
    const size_t slot = (size_t) (seq & 1u) * 64;
    for (int r = 0; r < D.world; ++r) {
        double* dst = M->vals[r] + (slot + D.rank) * 4;
#pragma unroll
        for (int k = 0; k < NV; ++k) dst[k] = v[k];
    }
    __threadfence_system();
    for (int r = 0; r < D.world; ++r) st_relaxed_sys(M->flags[r] + slot + D.rank, seq);
    double t[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) t[k] = 0.0;
    for (int r = 0; r < D.world; ++r) {
        const unsigned* mine = M->flags[D.rank] + slot + r;
        long long spins = 0, t0 = 0;
        while (ld_acquire_sys(mine) != seq) {
            if ((++spins & 1023) == 0) {
                if (t0 == 0) t0 = wall_ns();
                else if (wall_ns() - t0 > kWaitTimeoutNs) { S->trsv_timeout = 1; break; }
            }
        }
        const double* src = M->vals[D.rank] + (slot + r) * 4;
#pragma unroll
        for (int k = 0; k < NV; ++k) t[k] += __ldcg(src + k);
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = t[k];
}

// PHASE 0: after k_init, 1: after k_vec_xr1, 2: after k_vec_xr2 (multi-GPU only, one thread)
template <int PHASE>
__global__ void k_finish(Scalars* S, double tol, int max_half)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (PHASE == 0) { finish_init(S, S->red[0], tol, max_half); return; }
    if (S->done) return;
    if (PHASE == 1) finish_xr1(S, S->red[0]);
    if (PHASE == 2) finish_xr2(S, S->red[0], S->red[1]);
}

// r = rt = P b, x = 0, arm the dataflow arrays, norm0 = ||r||, rho_new = <rt, r>
__global__ void __launch_bounds__(kVecThreads) k_init(const double* __restrict__ b_nat, const int* __restrict__ perm,
                                                      double* __restrict__ r, double* __restrict__ rt, double* __restrict__ x,
                                                      double* __restrict__ w, double* __restrict__ y, int N, Scalars* S,
                                                      double* partials, unsigned* ticket, double tol, int max_half, int dist)
{
    double acc[1] = {0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        int q = i / 3;
        double v = b_nat[3 * perm[q] + (i - 3 * q)];
        r[i] = v; rt[i] = v; x[i] = 0.0;
        w[i] = sentinel(); y[i] = sentinel();
        acc[0] += v * v;
    }
    double tot[1];
    if (grid_reduce<1>(acc, partials, ticket, tot)) {
        if (dist) { S->red[0] = tot[0]; S->done = 0; }
        else finish_init(S, tot[0], tol, max_half);
    }
}

// ---- block ILU0 -----------------------------------------------------------------------------

__device__ __forceinline__ bool inv3(const double* m, double* inv)
{
    // closed form of MatrixBlock.hpp:720-749 (Opm::Detail::Inverter<3>)
    double t4 = m[0] * m[4], t6 = m[0] * m[5], t8 = m[1] * m[3];
    double t10 = m[2] * m[3], t12 = m[1] * m[6], t14 = m[2] * m[6];
    double det = t4 * m[8] - t6 * m[7] - t8 * m[8] + t10 * m[7] + t12 * m[5] - t14 * m[4];
    double t17 = 1.0 / det;
    inv[0] = (m[4] * m[8] - m[5] * m[7]) * t17;
    inv[1] = -(m[1] * m[8] - m[2] * m[7]) * t17;
    inv[2] = (m[1] * m[5] - m[2] * m[4]) * t17;
    inv[3] = -(m[3] * m[8] - m[5] * m[6]) * t17;
    inv[4] = (m[0] * m[8] - t14) * t17;
    inv[5] = -(t6 - t10) * t17;
    inv[6] = (m[3] * m[7] - m[4] * m[6]) * t17;
    inv[7] = -(m[0] * m[7] - t12) * t17;
    inv[8] = (t4 - t8) * t17;
    return det != 0.0 && isfinite(det);
}

// One warp factorises one block row of the level (p-space rows rowlist[0..nrows)): copies the row of A
// into LU, eliminates the lower entries in ascending NATURAL column order (the order the entries of a
// p-space row are stored in; rows of earlier levels are final), inverts the pivot.  Lanes 0..8 own one
// scalar of the 3x3 block being produced.
__global__ void __launch_bounds__(256) k_ilu_factor_level(const int* __restrict__ prow, const int* __restrict__ pcol,
                                                          const int* __restrict__ pdiag, const double* __restrict__ A,
                                                          double* LU, const int* __restrict__ rowlist, int nrows, Scalars* S)
{
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= nrows) return;
    const int i = rowlist[wid];
    const int rs = prow[i], re = prow[i + 1], di = pdiag[i];
    for (int q = rs * 9 + lane; q < re * 9; q += 32) LU[q] = A[q];
    __syncwarp();
    const int er = lane / 3, ec = lane - 3 * er;    // scalar (er, ec) for lanes 0..8
    for (int kj = rs; kj < di; ++kj) {
        const int j = pcol[kj];
        const int dj = pdiag[j], je = prow[j + 1];
        // L_ij = A_ij * inv(A_jj)   (rightmultiply by the stored inverse)
        double lij = 0.0;
        if (lane < 9) {
            const double* a = LU + (size_t) kj * 9 + er * 3;
            const double* d = LU + (size_t) dj * 9 + ec;
            lij = a[0] * d[0] + a[1] * d[3] + a[2] * d[6];
        }
        __syncwarp();
        if (lane < 9) LU[(size_t) kj * 9 + lane] = lij;
        __syncwarp();
        // A_ik -= L_ij * U_jk for every k > j present in both rows
        for (int jk = dj + 1; jk < je; ++jk) {
            const int colk = pcol[jk];
            int ik = -1;
            for (int base = kj + 1; base < re; base += 32) {
                int cand = base + lane;
                unsigned m = __ballot_sync(kFull, cand < re && pcol[cand] == colk);
                if (m) { ik = base + __ffs(m) - 1; break; }
            }
            if (ik >= 0 && lane < 9) {
                const double* l = LU + (size_t) kj * 9 + er * 3;
                const double* u = LU + (size_t) jk * 9 + ec;
                LU[(size_t) ik * 9 + lane] -= l[0] * u[0] + l[1] * u[3] + l[2] * u[6];
            }
            __syncwarp();
        }
    }
    if (lane == 0) {
        double m[9], inv[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) m[q] = LU[(size_t) di * 9 + q];
        if (!inv3(m, inv)) S->singular = 1;
#pragma unroll
        for (int q = 0; q < 9; ++q) LU[(size_t) di * 9 + q] = inv[q];
    }
}

// The same factorisation with an elimination PLAN (analysis.hpp facOps): the kernel above walks the pattern with ~40
// dependent global loads per row (15 us per level whatever its size); here a warp first fetches everything it needs in
// two rounds -- its plan and its own row, then all upstream blocks (pivots and U blocks of earlier levels, final) -- into
// shared memory and eliminates there: two L2/HBM round trips per level instead of forty.
constexpr int kFacMaxRow = 16, kFacMaxOps = 48, kFacWarps = 8;
__global__ void __launch_bounds__(32 * kFacWarps) k_ilu_factor_plan(const int* __restrict__ prow, const int* __restrict__ pdiag,
                                                                   const int* __restrict__ facPtr, const int2* __restrict__ facOps,
                                                                   const double* __restrict__ A, double* LU,
                                                                   const int* __restrict__ rowlist, int nrows, Scalars* S)
{
    __shared__ double srow[kFacWarps][kFacMaxRow * 9];
    __shared__ double sup[kFacWarps][kFacMaxOps * 9];
    __shared__ int2 sop[kFacWarps][kFacMaxOps];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int wid = blockIdx.x * kFacWarps + w;
    // Programmatic dependent launch: the next level's grid may be scheduled as soon as every CTA of this one has got here,
    // and runs its prologue (plan and own row of A, which no level writes) while this level still eliminates; it touches LU
    // only after griddepcontrol.wait, i.e. after this grid has completed and its stores are visible.  Launched without the
    // attribute both instructions are no-ops.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (wid >= nrows) return;
    const int i = rowlist[wid];
    const int rs = prow[i], re = prow[i + 1], di = pdiag[i] - rs;
    const int o0 = facPtr[i], nops = facPtr[i + 1] - o0;
    double* row = srow[w];
    double* up = sup[w];
    for (int o = lane; o < nops; o += 32) sop[w][o] = facOps[o0 + o];
    for (int f = lane; f < (re - rs) * 9; f += 32) row[f] = A[(size_t) rs * 9 + f];
    __syncwarp();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int f = lane; f < nops * 9; f += 32) {
        const int o = f / 9;
        up[f] = LU[(size_t) sop[w][o].x * 9 + (f - 9 * o)];
    }
    __syncwarp();
    const int g = lane / 9, e = lane - 9 * g, er = e / 3, ec = e - 3 * er;      // three 3x3 products side by side
    int o = 0;
    while (o < nops) {
        const int code = sop[w][o].y;
        const int toff = code & 255, nupd = code >> 8;
        // L_ij = A_ij * inv(A_jj)
        double lij = 0.0;
        if (lane < 9) {
            const double* a = row + toff * 9 + er * 3;
            const double* d = up + o * 9 + ec;
            lij = a[0] * d[0] + a[1] * d[3] + a[2] * d[6];
        }
        __syncwarp();
        if (lane < 9) row[toff * 9 + lane] = lij;
        __syncwarp();
        // A_ik -= L_ij * U_jk for the blocks present in both rows: distinct targets, three at a time
        for (int u0 = 0; u0 < nupd; u0 += 3) {
            const int u = u0 + g;
            if (g < 3 && u < nupd) {
                const int tgt = -(sop[w][o + 1 + u].y + 1);
                const double* l = row + toff * 9 + er * 3;
                const double* uu = up + (o + 1 + u) * 9 + ec;
                row[tgt * 9 + e] -= l[0] * uu[0] + l[1] * uu[3] + l[2] * uu[6];
            }
            __syncwarp();
        }
        o += 1 + nupd;
    }
    if (lane == 0) {
        double inv[9];
        if (!inv3(row + di * 9, inv)) S->singular = 1;
#pragma unroll
        for (int k = 0; k < 9; ++k) row[di * 9 + k] = inv[k];
    }
    __syncwarp();
    for (int f = lane; f < (re - rs) * 9; f += 32) LU[(size_t) rs * 9 + f] = row[f];
}

// Round 2: the same elimination with THREE rows per warp -- nine lanes per row, one lane per scalar of a 3x3 block -- instead
// of one (where 9 of the 32 lanes did the arithmetic and a large level was bound by warp throughput: 5.8 us per level on C3's
// 7500-row levels against 2 us on small ones).  rec (built by the host, level order, every level padded to whole triples):
//   rec[p] = { row i (-1: padding), first block of the row, first op, blocks | diagonal offset << 8 | ops << 16 }
// The launches stay one per level set, chained by programmatic dependent launch: a single resident launch in which rows wait
// for rows through L2 (per-row flags with release/acquire, or the sentinel-armed factor as its own flag) was measured at
// 4.4-8 us per level -- the L2 hand-over between SMs costs more than a PDL-chained launch.
constexpr int kFac3MaxWarps = 16;
__global__ void __launch_bounds__(32 * kFac3MaxWarps) k_ilu_factor_plan3(const int4* __restrict__ rec, int ntriples, const int2* __restrict__ facOps,
                                                                     const double* __restrict__ A, double* LU, int maxRow, int maxOps, Scalars* S)
{
    extern __shared__ __align__(16) unsigned char fac3_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int slot = lane / 9, e = lane - 9 * slot, er = e / 3, ec = e - 3 * er;
    const bool lane_on = slot < 3;
    const size_t per = (size_t) (maxRow + maxOps) * 72 + (size_t) maxOps * 8;
    double* row = reinterpret_cast<double*>(fac3_smem + (size_t) (3 * w + (lane_on ? slot : 0)) * per);
    double* up = row + maxRow * 9;
    int2* sop = reinterpret_cast<int2*>(up + maxOps * 9);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int t = blockIdx.x * (blockDim.x >> 5) + w;
    if (t >= ntriples) return;
    const int4 r = lane_on ? __ldg(rec + 3 * (size_t) t + slot) : make_int4(-1, 0, 0, 0);
    const int i = r.x, rs = r.y, o0 = r.z, nblk = r.w & 255, di = (r.w >> 8) & 255, nops = r.w >> 16;
    // everything that does not depend on earlier levels: the plan and the row of A (no level writes them)
    for (int o = e; o < nops; o += 9) sop[o] = __ldg(facOps + o0 + o);
    for (int b0 = 0; b0 < nblk; b0 += 8) {              // eight loads in flight per lane, not one round trip per block
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = b0 + k < nblk ? __ldg(A + (size_t) (rs + b0 + k) * 9 + e) : 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (b0 + k < nblk) row[(b0 + k) * 9 + e] = v[k];
    }
    __syncwarp();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // upstream blocks: pivots and U blocks of the rows this one eliminates with (earlier levels, final)
    for (int ob = 0; ob < nops; ob += 8) {
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = ob + k < nops ? __ldcg(LU + (size_t) sop[ob + k].x * 9 + e) : 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (ob + k < nops) up[(ob + k) * 9 + e] = v[k];
    }
    __syncwarp();
    int o = 0;
    while (__any_sync(kFull, o < nops)) {
        const bool act = o < nops;
        const int code = act ? sop[o].y : 0;
        const int toff = code & 255, nupd = code >> 8;
        // L_ij = A_ij * inv(A_jj)
        double lij = 0.0;
        if (act) {
            const double* a = row + toff * 9 + er * 3;
            const double* d = up + o * 9 + ec;
            lij = a[0] * d[0] + a[1] * d[3] + a[2] * d[6];
        }
        __syncwarp();
        if (act) row[toff * 9 + e] = lij;
        __syncwarp();
        // A_ik -= L_ij * U_jk for the blocks present in both rows: distinct targets, this lane's scalar of each
        for (int u = 0; u < nupd; ++u) {
            const int tgt = -(sop[o + 1 + u].y + 1);
            const double* l = row + toff * 9 + er * 3;
            const double* uu = up + (o + 1 + u) * 9 + ec;
            row[tgt * 9 + e] -= l[0] * uu[0] + l[1] * uu[3] + l[2] * uu[6];
        }
        __syncwarp();
        if (act) o += 1 + nupd;
    }
    if (e == 0 && i >= 0) {
        double inv[9];
        if (!inv3(row + di * 9, inv)) S->singular = 1;
#pragma unroll
        for (int k = 0; k < 9; ++k) row[di * 9 + k] = inv[k];
    }
    __syncwarp();
    for (int bk = 0; bk < nblk; ++bk) LU[(size_t) (rs + bk) * 9 + e] = row[bk * 9 + e];
}

// ---- triangular solves, level-synchronous (opt-in multi-colour ordering only) ---------------------------------------------
//
// With the colour ordering (coloring.hpp) the level sets are the ~15 colours: tens of thousands of independent rows each, no
// chain to pipeline.  One launch per level (chained by programmatic dependent launch inside the iteration's graph), three
// lanes per block row straight from the BSR factor: lane c of a row owns component c.
//   lower:  w_i = d_i - sum_{j < i} L_ij w_j                     (unit diagonal)
//   upper:  y_i = relax * Dinv_i (w_i - sum_{j > i} U_ij y_j)    (inverse pivot stored, ParallelOverlappingILU0.hpp:867-901)
// Rows of earlier levels were written by earlier launches: plain loads.
template <bool LOWER>
__global__ void __launch_bounds__(256) k_trsv_level(const int* __restrict__ prow, const int* __restrict__ pcol, const int* __restrict__ pdiag,
                                                    const double* __restrict__ LU, const double* __restrict__ rhs, double* out,
                                                    const int* __restrict__ rowlist, int nrows, double relax, const Scalars* S, int check_done)
{
    pdl_enter();
    if (check_done && S->done) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int q = lane / 3, c = lane - 3 * q;                   // ten rows per warp, lanes 30 and 31 idle
    const int slot = (t >> 5) * 10 + q;
    const bool on = q < 10 && slot < nrows;
    double acc = 0.0;
    int i = 0, di = 0;
    if (on) {
        i = rowlist[slot];
        di = pdiag[i];
        const int k0 = LOWER ? prow[i] : di + 1, k1 = LOWER ? di : prow[i + 1];
        acc = rhs[3 * (size_t) i + c];
        for (int kb = k0; kb < k1; kb += 4) {                   // four blocks' loads in flight, summed in column order
            int cc[4];
            double a[4][3], x[4][3];
#pragma unroll
            for (int u = 0; u < 4; ++u) cc[u] = kb + u < k1 ? pcol[kb + u] : -1;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (cc[u] >= 0) {
                    const double* ap = LU + (size_t) (kb + u) * 9 + 3 * c;
                    const double* xp = out + 3 * (size_t) cc[u];
                    a[u][0] = ap[0]; a[u][1] = ap[1]; a[u][2] = ap[2];
                    x[u][0] = xp[0]; x[u][1] = xp[1]; x[u][2] = xp[2];
                }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (cc[u] >= 0) acc -= a[u][0] * x[u][0] + a[u][1] * x[u][1] + a[u][2] * x[u][2];
        }
    }
    if (LOWER) {
        if (on) out[3 * (size_t) i + c] = acc;
        return;
    }
    // y = relax * Dinv t: the three components of t sit in the three lanes of the row
    const int base = 3 * q;
    const double t0 = __shfl_sync(kFull, acc, base < 30 ? base : 0), t1 = __shfl_sync(kFull, acc, base < 30 ? base + 1 : 0),
                 t2 = __shfl_sync(kFull, acc, base < 30 ? base + 2 : 0);
    if (on) {
        const double* d = LU + (size_t) di * 9 + 3 * c;
        out[3 * (size_t) i + c] = relax * (d[0] * t0 + d[1] * t1 + d[2] * t2);
    }
}

// ---- triangular solves: pencil-pipelined sweeps -----------------------------------------------------
//
// Device mirrors of analysis.hpp's StageRef / PartRef / BuildRef (layout checked by static_assert in
// b200bda.cu).
struct StageD { long long meta_off, vals_off; int meta_ints, vals_doubles, g_lo, g_rows; };
struct PartD { int stage_begin, stage_end, row0, nrows; };
struct BuildD { long long vals_off; int src_off, count, first; };

// position of field f of this lane inside a record of the value stream (== analysis.hpp sweep_vidx)
template <bool LOWER>
__device__ __forceinline__ int vidx(int f, int lane) { return (LOWER && f == 8) ? 256 + lane : (f >> 1) * 64 + 2 * lane + (f & 1); }

// Scatter the BSR factor (p-space) into the record stream of one sweep (analysis.hpp): one warp per record, one
// lane per (row q, component comp) of the record.
//   lower: field 3 j + v = L[block j][comp][v]
//   upper: field 3 j + v = (D^-1 U)[block j][comp][v], field 9 + v = (w D^-1)[comp][v]  (first record of a row only)
template <bool LOWER>
__global__ void __launch_bounds__(256) k_fill_stream(const BuildD* __restrict__ build, int nrec, const int* __restrict__ src,
                                                     const double* __restrict__ LU, double* __restrict__ vals, double relax)
{
    const int lane = threadIdx.x & 31;
    const int q = lane / 3, comp = lane - 3 * q;
    for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < nrec; c += (gridDim.x * blockDim.x) >> 5) {
        const BuildD b = build[c];
        double* out = vals + b.vals_off;
        const bool act = q < b.count;
        double inv[3] = {0.0, 0.0, 0.0};
        if (!LOWER) {
            if (act) {
                const double* d = LU + (size_t) src[b.src_off + 3 * b.count + q] * 9 + comp * 3;
                inv[0] = d[0]; inv[1] = d[1]; inv[2] = d[2];
            }
            // z = w U^-1 y (ParallelOverlappingILU0.hpp:897-901: the back-substitution, then `v *= w`): with z' = w z the recurrence
            // is z'_i = (w D^-1) y_i - (D^-1 U_ij) z'_j, so only the pivot field carries the relaxation factor
#pragma unroll
            for (int v = 0; v < 3; ++v) out[vidx<LOWER>(9 + v, lane)] = b.first ? relax * inv[v] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int k = act ? src[b.src_off + j * b.count + q] : -1;
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                double x = 0.0;
                if (k >= 0) {
                    const double* u = LU + (size_t) k * 9;
                    x = LOWER ? u[comp * 3 + v] : inv[0] * u[v] + inv[1] * u[3 + v] + inv[2] * u[6 + v];
                }
                out[vidx<LOWER>(3 * j + v, lane)] = x;
            }
        }
    }
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// waits until the phase with the given parity has completed.  try_wait with a suspend-time hint parks the warp in
// hardware instead of spinning: a hot try_wait loop in the producer / helper warps steals issue slots from the
// consumer warp that shares their scheduler (+50 % on the level step, tools/microbench/smlat.cu).
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680) : "memory");
}
// the same for warps off the critical path (producer, helpers): sleep between two attempts
__device__ __forceinline__ void mbar_wait_relaxed(unsigned long long* bar, unsigned parity)
{
    unsigned ok = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680) : "memory");
        if (ok) break;
        __nanosleep(200);
    }
}
// 1-D bulk copy global -> shared through the TMA unit, completion counted in bytes on an mbarrier
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_barrier(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

constexpr int kSweepMaxSlots = 8;      // ring slots (full + empty mbarriers and the ext flags fit the 256-byte header)
constexpr int kSweepHeader = 256;
constexpr int kXs = 4;                 // doubles per row of the shared-memory value space (== analysis.hpp kXwinStride)
constexpr int kSweepTailPad = 512;     // lanes of a partly filled record may read (never use) a few rows past the last rhs copy

// The SpMV that follows an upper sweep, run by the sweep's own CTAs as their parts finish (k_sweep<..., SPMV >= 0>)
struct FusedSpmv {
    const int* sptr; const int* sover; const int* scol; const double* sval;      // sliced-ELL copy of A (SellPlan)
    const int* prow; const int* pcol; const double* A;                            // BSR arrays: rows beyond the slice width
    double* y;                     // result of the product; the input vector is the sweep's `out`
    const double* d1;              // dot-product partner (MODE 1: <d1, y>; MODE 2: <y, d1>, <y, y>)
    const int2* units;             // {first slice, slices}, in expected order of readiness (FusedPlan)
    const int* need_ptr; const int* need;
    int* sync;                     // [0] unit counter, [1] ticket, [2 + p] part p finished; zero at launch (the last CTA re-zeroes it)
    double* partials;              // 2 x nunits
    int ring_bytes;                // dynamic shared memory of the launch: ring of SELL slices once the part is swept
    long long* dbg;                // debugging aid (may be null): per part {part done, exit, units taken, time spent waiting for parts} [ns]
    int Nb, nunits;
    int chunk;                     // slots of a slice per chunk buffer (0: kTailChunk)
    int prefetch;                  // 1: the producer warp pulls the whole unit into L2 when the unit starts
};

// Deferred solution update run by idle CTAs of a lower sweep (k_sweep<true, ..., SPMV = 3>): x += pend * y, y re-armed
struct XUpdate {
    double* x; double* y;
    int* sync;                     // [0] chunk counter, [1] ticket; zero at launch (the last CTA re-zeroes it)
    int n;                         // doubles
};

}  // namespace b200
#include "sweep2.cuh"
namespace b200 {

struct SweepArgs {
    FusedSpmv f;
    XUpdate xu;
    Sweep2Args v2;        // round-2 schedule (k_sweep2)
    const StageD* stages;
    const PartD* parts;
    const int* meta;
    const double* vals;
    const double* rhs;
    double* out;
    double* rearm;        // may be null: vector re-armed with the sentinel row by row as it is consumed
    Scalars* S;
    int nparts, nslots, window, extWindow, metaCap, valsCap, rhsCap, nwarps, nhalo, check_done, helper_sleep;
    int early;            // helpers start a stage's fetch when the stage before it has been issued (needs ring room for nslots + 1 stages)
    int nowait;           // experiment: external rows are taken as they are (wrong results): what the parts could do unstarved
    long long* trace;     // debugging aid (may be null): per part and stage {wait begin, data landed, stage done, issued} in SM cycles
    int trace_cap;
};

// shared-memory accesses by 32-bit shared address: generic pointers into dynamic shared memory make the compiler rebuild
// the shared window base (S2R SR_CgaCtaId + LEA) next to every use, in the middle of the dependent part of a row
__device__ __forceinline__ double2 lds_f64x2(unsigned a)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_f64(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int4 lds_s32x4(unsigned a)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned a, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

__device__ __forceinline__ int ld_volatile_s32(const int* p)
{
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_s32(int* p, int v)
{
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// One triangular sweep.  One persistent CTA per PART (pencil of grid lines, analysis.hpp), all resident.
//   producer warp : walks the part's stages and fetches each one -- meta ints, record-major factor values,
//                   rhs rows, all contiguous in processing order -- with three bulk copies (TMA unit,
//                   UBLKCP) into a ring of shared-memory slots, several stages ahead of the consumers:
//                   HBM latency never sits on the dependency chain;
//   helper warps  : rows owned by OTHER parts are the only values that travel through L2 (every sweep finds
//                   `out` armed with a NaN sentinel, producers overwrite it with relaxed gpu-scope 8-byte
//                   stores: the value is its own ready flag).  Helper warp h takes the stages i = h (mod H):
//                   as soon as the stage's meta has landed it polls the stage's external rows, level group by
//                   level group, and parks them in the external ring of the shared-memory value space -- one
//                   L2 round trip per group, but several stages AHEAD of the consumers and off their critical
//                   path; it publishes a per-slot counter the consumers check (shared memory, ~30 cycles);
//   consumer warps: a record = <= 10 rows of one level, 3 lanes per row.  A lone warp issues in order at ~5 cycles
//                   per instruction (measured: tools/microbench/smlat.cu, profiles/), so the time between two
//                   level barriers is the INSTRUCTION COUNT of whatever sits between them.  Hence: the record
//                   layout makes every operand one load with an immediate offset (values in 16-byte pairs,
//                   32-byte value rows, precomputed byte offsets), the dependent part is 6 shared-memory reads
//                   of earlier rows (window, parked external rows or the zero row: one flat value space, no
//                   branches), 9 fma + 3 add and one store, and consecutive levels may go to different warp
//                   GROUPS so that one group's operand fetch overlaps another group's dependent part.
//                   The upper sweep is the same code: w D^-1 is folded into the stream (k_fill_stream).
// Parts process their rows in ascending (descending for U) global level, a topological order of the
// whole DAG, and a level only ever waits for rows of earlier levels, so the waits cannot cycle as long
// as every CTA is resident (grid <= SMs).
__device__ __forceinline__ int ld_acquire_gpu_s32(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_s32(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// Tail of an upper-sweep CTA: its part is finished while the wavefront is still crossing the other parts (the corner part
// is done after 50 of 240 us), so the SM would idle.  Instead it publishes "part done" and works through SpMV units (the
// product that follows the sweep in BiCGSTAB, y = A * out) whose rows and columns all lie in finished parts.  Only CTAs
// that have nothing left to sweep take units, so a sweeping part never shares its SM with the SpMV (a second grid on the
// same SMs slowed the sweep by 45 %).  With 11 warps per SM plain loads cannot keep enough bytes in flight (13 GB/s per SM
// measured), so the producer warp streams the sliced-ELL values and columns of the unit's slices with cp.async.bulk into
// per-warp rings over the (now free) dynamic shared memory and the other warps compute from there; only the gather of
// the input vector is a register load.  That vector was written by other SMs during this launch, but the gather may use
// L1 (with ld.cg every 8-byte load is its own L2 request and the tail is bound by the SM's request rate): a unit's `need`
// list names the owner of every row within 6 rows (a 128-byte line) of any row it touches, so a line fetched into L1 never
// holds a row that is still to be written.  Dot products: one partial per UNIT, summed in unit order by the last CTA to finish -- deterministic
// whoever took a unit.
constexpr int kMaxSweepParts = 1024, kTailChunk = 4, kTailBufBytes = kTailChunk * (2304 + 128), kTailMaxCons = 16;
template <int MODE>
__device__ __forceinline__ void fused_spmv_tail(const SweepArgs& P, int part, unsigned char* ring)
{
    const FusedSpmv& F = P.f;
    __shared__ int s_unit[2];
    __shared__ unsigned char s_done[kMaxSweepParts];
    __shared__ double s_red[2][32];
    __shared__ bool s_last;
    // every consumer warp owns two chunk buffers (kTailChunk slots of a slice each) and their barriers: a barrier is only ever
    // waited on by one warp, phase after phase, so the parity test cannot alias
    __shared__ __align__(8) unsigned long long s_full[kTailMaxCons][2], s_empty[kTailMaxCons][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int PW = P.nwarps;                                   // the sweep's producer warp keeps that job
    const int NC = nwarp - 1, ci = warp < PW ? warp : warp - 1;
    const int tc = max(1, min(F.chunk > 0 ? F.chunk : kTailChunk, kTailChunk));          // slots of a slice per chunk buffer
    const int bufBytes = tc * (2304 + 128);
    const int NR = min(min(NC, kTailMaxCons), F.ring_bytes / (2 * bufBytes));   // consumers fed through the ring
    const int NU = (NR > 0 && !(P.nowait & 32)) ? NR : NC;       // warps that take slices: a plain-load warp next to ring-fed ones would be the straggler
    for (int p = threadIdx.x; p < P.nparts; p += blockDim.x) s_done[p] = 0;
    if (threadIdx.x < 32) { s_red[0][threadIdx.x] = 0.0; s_red[1][threadIdx.x] = 0.0; }
    if (threadIdx.x == 0) {
        for (int k = 0; k < kTailMaxCons; ++k)
            for (int b2 = 0; b2 < 2; ++b2) { mbar_init(&s_full[k][b2], 1); mbar_init(&s_empty[k][b2], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();                       // every row of this part is stored; the sweep's shared memory is free
    long long t_wait = 0, t_full = 0, t_top = 0;
    int n_units = 0, n_chunks = 0;
    auto now_ns = []() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    const bool has_part = part < P.nparts;                     // (round-2 kernel: CTAs beyond the parts only run the tail)
    if (F.dbg && threadIdx.x == 0 && has_part) F.dbg[4 * part] = now_ns();
    if (threadIdx.x == 0) {
        if (has_part) {
            __threadfence();
            st_release_gpu_s32(F.sync + 2 + part, 1);
        }
        s_unit[0] = atomicAdd(F.sync, 1);
    }
    if (warp == PW) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const double* x = P.out;
    int cnt = 0;                           // chunks this warp has consumed (consumer) / this lane has issued for consumer `lane` (producer)
    for (int it = 0;; ++it) {
        const long long tt0 = F.dbg && threadIdx.x == 0 ? now_ns() : 0;
        __syncthreads();
        if (F.dbg && threadIdx.x == 0) t_top += now_ns() - tt0;
        const int u = s_unit[it & 1];
        if (u >= F.nunits) break;
        const int2 un = __ldg(F.units + u);
        if (warp == PW) {                                      // ---- producer: lane c feeds consumer c
            // The ring holds 2 x ~10 KB per consumer: at HBM latency that is ~40 GB/s per SM, what a full machine needs -- but
            // while the sweep still runs only the idle SMs work on the product and HBM has bandwidth to spare.  So the unit's
            // slices (consecutive in the SELL arrays, ~0.5 MB) are requested into L2 at once: the ring's copies then see L2
            // latency, and an idle SM takes units several times faster.  MEASURED: slower (C3: 243 -> 257 us per launch, the
            // requests compete with the operand streams of the parts that are still sweeping) -- off by default.
            if (F.prefetch) {
                for (int jj = lane; jj < un.y; jj += 32) {
                    const int q0 = __ldg(F.sptr + un.x + jj), qw = __ldg(F.sptr + un.x + jj + 1) - q0;
                    if (qw > 0) {
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(F.sval + (size_t) q0 * 288), "r"((unsigned) qw * 2304u) : "memory");
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(F.scol + (size_t) q0 * 32), "r"((unsigned) qw * 128u) : "memory");
                    }
                }
            }
            // one converged loop with non-blocking barrier tests: lanes that slept in private wait loops would serialise
            // (nine lanes x 200 ns of sleep per poll round = the whole ring starves)
            int j = lane, k0 = 0, s0 = 0, w = 0;
            auto next_slice = [&]() {
                for (; j < un.y; j += NU) {
                    s0 = __ldg(F.sptr + un.x + j); w = __ldg(F.sptr + un.x + j + 1) - s0; k0 = 0;
                    if (w > 0) return true;
                }
                return false;
            };
            bool active = lane < NR && next_slice();
            while (__any_sync(kFull, active)) {
                bool issued = false;
                if (active) {
                    const int b2 = cnt & 1, n = min(tc, w - k0);
                    if (cnt < 2 || mbar_test(&s_empty[lane][b2], ((cnt >> 1) - 1) & 1)) {
                        unsigned char* buf = ring + (size_t) (2 * lane + b2) * bufBytes;
                        mbar_expect_tx(&s_full[lane][b2], (unsigned) n * (2304 + 128));
                        bulk_g2s(buf, F.sval + (size_t) (s0 + k0) * 288, (unsigned) n * 2304, &s_full[lane][b2]);
                        bulk_g2s(buf + tc * 2304, F.scol + (size_t) (s0 + k0) * 32, (unsigned) n * 128, &s_full[lane][b2]);
                        ++cnt; k0 += tc; issued = true;
                        if (k0 >= w) { j += NU; active = next_slice(); }
                    }
                }
                if (!__any_sync(kFull, issued)) __nanosleep(64);
            }
            continue;
        }
        // ---- consumers
        if (threadIdx.x == 0) s_unit[(it + 1) & 1] = atomicAdd(F.sync, 1);     // claim ahead
        const int n0 = __ldg(F.need_ptr + u), nn = __ldg(F.need_ptr + u + 1) - n0;
        const long long tw0 = F.dbg && threadIdx.x == 0 ? now_ns() : 0;
        for (int k = ci * 32 + lane; k < nn; k += NC * 32) {
            const int p = __ldg(F.need + n0 + k);
            if (!s_done[p]) {
                int spins = 0;
                while (ld_acquire_gpu_s32(F.sync + 2 + p) == 0) {
                    __nanosleep(400);
                    if ((++spins & 255) == 0 && (spins > (1 << 17) || *((volatile int*) &P.S->trsv_timeout))) { P.S->trsv_timeout = 1; break; }
                }
                s_done[p] = 1;
            }
        }
        named_barrier(2, NC * 32);
        if (F.dbg && threadIdx.x == 0) { t_wait += now_ns() - tw0; ++n_units; }
        double acc0 = 0.0, acc1 = 0.0;
        for (int j = ci; j < un.y && ci < NU; j += NU) {
            const int slice = un.x + j;
            const int s0 = __ldg(F.sptr + slice), s1 = __ldg(F.sptr + slice + 1), w = s1 - s0;
            const int row = 32 * slice + lane;
            double y0 = 0.0, y1 = 0.0, y2 = 0.0;
            if (ci < NR) {
                for (int k0 = 0; k0 < w; k0 += tc, ++cnt) {
                    const int b2 = cnt & 1, n = min(tc, w - k0);
                    const long long tf0 = F.dbg && threadIdx.x == 0 ? now_ns() : 0;
                    mbar_wait(&s_full[ci][b2], (cnt >> 1) & 1);
                    if (F.dbg && threadIdx.x == 0) { t_full += now_ns() - tf0; ++n_chunks; }
                    const unsigned char* buf = ring + (size_t) (2 * ci + b2) * bufBytes;
                    const double* vs = reinterpret_cast<const double*>(buf) + lane;
                    const int* cs = reinterpret_cast<const int*>(buf + tc * 2304) + lane;
                    double xr[kTailChunk][3];
#pragma unroll
                    for (int k = 0; k < kTailChunk; ++k)
                        if (k < n) {
                            const double* xx = x + 3 * (size_t) cs[32 * k];
                            xr[k][0] = xx[0]; xr[k][1] = xx[1]; xr[k][2] = xx[2];
                        }
#pragma unroll
                    for (int k = 0; k < kTailChunk; ++k)
                        if (k < n) {
                            const double* v = vs + 288 * k;
                            y0 += v[0] * xr[k][0] + v[32] * xr[k][1] + v[64] * xr[k][2];
                            y1 += v[96] * xr[k][0] + v[128] * xr[k][1] + v[160] * xr[k][2];
                            y2 += v[192] * xr[k][0] + v[224] * xr[k][1] + v[256] * xr[k][2];
                        }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_empty[ci][b2]);
                }
            } else {
                const double* v = F.sval + (size_t) s0 * 288 + lane;
                const int* c = F.scol + (size_t) s0 * 32 + lane;
#pragma unroll 2
                for (int k = s0; k < s1; ++k, v += 288, c += 32) {
                    const double* xx = x + 3 * (size_t) __ldg(c);
                    const double x0 = xx[0], x1 = xx[1], x2 = xx[2];
                    y0 += __ldg(v) * x0 + __ldg(v + 32) * x1 + __ldg(v + 64) * x2;
                    y1 += __ldg(v + 96) * x0 + __ldg(v + 128) * x1 + __ldg(v + 160) * x2;
                    y2 += __ldg(v + 192) * x0 + __ldg(v + 224) * x1 + __ldg(v + 256) * x2;
                }
            }
            if (__ldg(F.sover + slice) && row < F.Nb) {
                for (int k = __ldg(F.prow + row) + w, ke = __ldg(F.prow + row + 1); k < ke; ++k) {
                    const double* a = F.A + (size_t) k * 9;
                    const double* xx = x + 3 * (size_t) __ldg(F.pcol + k);
                    const double x0 = xx[0], x1 = xx[1], x2 = xx[2];
                    y0 += __ldg(a) * x0 + __ldg(a + 1) * x1 + __ldg(a + 2) * x2;
                    y1 += __ldg(a + 3) * x0 + __ldg(a + 4) * x1 + __ldg(a + 5) * x2;
                    y2 += __ldg(a + 6) * x0 + __ldg(a + 7) * x1 + __ldg(a + 8) * x2;
                }
            }
            if (row < F.Nb) {
                double* yy = F.y + 3 * (size_t) row;
                yy[0] = y0; yy[1] = y1; yy[2] = y2;
                const double* dd = F.d1 + 3 * (size_t) row;
                const double e0 = __ldg(dd), e1 = __ldg(dd + 1), e2 = __ldg(dd + 2);
                if (MODE == 1) acc0 += e0 * y0 + e1 * y1 + e2 * y2;
                if (MODE == 2) { acc0 += y0 * e0 + y1 * e1 + y2 * e2; acc1 += y0 * y0 + y1 * y1 + y2 * y2; }
            }
        }
        acc0 = warp_sum(acc0);
        if (MODE == 2) acc1 = warp_sum(acc1);
        if (lane == 0) { s_red[0][warp] = acc0; s_red[1][warp] = acc1; }
        named_barrier(2, NC * 32);
        if (threadIdx.x == 0) {
            double a = 0.0, b = 0.0;
            for (int w = 0; w < nwarp; ++w) { a += s_red[0][w]; b += s_red[1][w]; }
            F.partials[u] = a;
            if (MODE == 2) F.partials[F.nunits + u] = b;
        }
    }
    if (threadIdx.x == 0) {
        if (F.dbg && has_part) { F.dbg[4 * part + 1] = now_ns(); F.dbg[4 * part + 2] = n_units; F.dbg[4 * part + 3] = t_wait;
            if (part % 12 == 0) printf("    part %d warp 0: %d units, %d chunks, %.1f us waiting for chunks, %.1f us at the unit barrier, %.1f us waiting for parts, NR %d\n", part, n_units, n_chunks, t_full * 1e-3, t_top * 1e-3, t_wait * 1e-3, NR); }
        __threadfence();
        s_last = atomicAdd(F.sync + 1, 1) == (int) gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (F.dbg && threadIdx.x == 0) {
        long long first = LLONG_MAX, last = 0, end = 0, waited = 0, busy = 0;
        for (int p = 0; p < P.nparts; ++p) {
            const long long d = __ldcg(F.dbg + 4 * p), e = __ldcg(F.dbg + 4 * p + 1);
            first = min(first, d); last = max(last, d); end = max(end, e); waited += __ldcg(F.dbg + 4 * p + 3); busy += e - d;
        }
        printf("fused sweep+spmv: first part done at 0, last part done +%.1f us, kernel end +%.1f us; SpMV CTA time %.1f us x CTA, of which %.1f waiting for parts\n",
               (last - first) * 1e-3, (end - first) * 1e-3, busy * 1e-3 / P.nparts, waited * 1e-3 / P.nparts);
        for (int p = 0; p < P.nparts; p += max(1, P.nparts / 12))
            printf("  part %3d: done +%.1f us, exit +%.1f us, %lld units, waited %.1f us\n", p, (__ldcg(F.dbg + 4 * p) - first) * 1e-3,
                   (__ldcg(F.dbg + 4 * p + 1) - first) * 1e-3, __ldcg(F.dbg + 4 * p + 2), __ldcg(F.dbg + 4 * p + 3) * 1e-3);
    }
    double a = 0.0, b = 0.0;
    for (int u = threadIdx.x; u < F.nunits; u += blockDim.x) {
        a += __ldcg(F.partials + u);
        if (MODE == 2) b += __ldcg(F.partials + F.nunits + u);
    }
    a = warp_sum(a); b = warp_sum(b);
    __syncthreads();
    if (lane == 0) { s_red[0][warp] = a; s_red[1][warp] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = 0.0; b = 0.0;
        for (int w = 0; w < nwarp; ++w) { a += s_red[0][w]; b += s_red[1][w]; }
        if (MODE == 1) P.S->h = a;
        if (MODE == 2) { P.S->tr = a; P.S->tt = b; }
    }
    for (int k = threadIdx.x; k < 2 + P.nparts; k += blockDim.x) F.sync[k] = 0;     // every other CTA has left: ready for the next launch
}

// Tail of a LOWER-sweep CTA: BiCGSTAB's solution updates x += alpha y / x += omega y have no consumer before the solve
// ends, so k_vec_xr1 / k_vec_xr2 only record the coefficient (Scalars::pend) and the CTAs of the next lower sweep apply
// it once their part is swept (and re-arm y with the sentinel for the upper sweep that follows): 96 bytes per block row
// leave the serial part of the iteration.  The work does not depend on the sweep at all; chunks come from a counter.
constexpr int kXUpdChunk = 8192;           // doubles per claim
__device__ __forceinline__ void xupdate_tail(const SweepArgs& P, double pend)
{
    const XUpdate& X = P.xu;
    __shared__ int s_chunk[2];
    __shared__ bool s_lastx;
    const int nchunks = (X.n + kXUpdChunk - 1) / kXUpdChunk;
    const double sent = sentinel();
    if (threadIdx.x == 0) s_chunk[0] = atomicAdd(X.sync, 1);
    for (int it = 0;; ++it) {
        __syncthreads();
        const int c = s_chunk[it & 1];
        if (c >= nchunks) break;
        if (threadIdx.x == 0) s_chunk[(it + 1) & 1] = atomicAdd(X.sync, 1);
        const int i0 = c * kXUpdChunk, i1 = min(X.n, i0 + kXUpdChunk);
        const int npair = (i1 - i0) >> 1;                      // i0 is even: 16-byte accesses
        double2* __restrict__ x2 = reinterpret_cast<double2*>(X.x + i0);
        double2* __restrict__ y2 = reinterpret_cast<double2*>(X.y + i0);
        // all loads of a batch before its first store: 11 warps per SM need the bytes in flight
        constexpr int kB = 6;
        for (int k0 = threadIdx.x; k0 < npair; k0 += kB * blockDim.x) {
            double2 xv[kB], yv[kB];
#pragma unroll
            for (int b = 0; b < kB; ++b) {
                const int k = k0 + b * blockDim.x;
                if (k < npair) { yv[b] = y2[k]; xv[b] = x2[k]; }
            }
#pragma unroll
            for (int b = 0; b < kB; ++b) {
                const int k = k0 + b * blockDim.x;
                if (k < npair) {
                    xv[b].x += pend * yv[b].x; xv[b].y += pend * yv[b].y;
                    x2[k] = xv[b];
                    y2[k] = make_double2(sent, sent);
                }
            }
        }
        if (threadIdx.x == 0 && ((i1 - i0) & 1)) { X.x[i1 - 1] += pend * X.y[i1 - 1]; X.y[i1 - 1] = sent; }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        s_lastx = atomicAdd(X.sync + 1, 1) == (int) gridDim.x - 1;
        if (s_lastx) { P.S->pend_on = 0; X.sync[0] = 0; X.sync[1] = 0; }     // every CTA read pend_on when the kernel started
    }
}

// SPMV: 3 = lower sweep whose CTAs then apply the pending solution update (xupdate_tail)
// SPMV: -1 = sweep only; 1 / 2 = the CTA goes on with the SpMV that follows (fused_spmv_tail<SPMV>)
constexpr int kFusedMaxThreads = 384;
template <bool LOWER, bool REARM, bool TRACE, int SPMV = -1>
__global__ void __launch_bounds__(SPMV >= 0 ? kFusedMaxThreads : 896) k_sweep(const SweepArgs P)
{
    extern __shared__ __align__(128) unsigned char sweep_smem[];
    pdl_enter();
    if (P.check_done && P.S->done) return;
    const int part = blockIdx.x;
    if (part >= P.nparts) return;
    const bool xupd = SPMV == 3 && P.S->pend_on != 0;
    const double xupd_coef = SPMV == 3 ? P.S->pend : 0.0;
    const PartD pr = P.parts[part];
    unsigned long long* full = reinterpret_cast<unsigned long long*>(sweep_smem);
    unsigned long long* empty = full + kSweepMaxSlots;
    int* ext_ready = reinterpret_cast<int*>(sweep_smem + 128);
    int* issued = reinterpret_cast<int*>(sweep_smem + 192);         // stages the producer has issued (monotonic)
    unsigned char* xwin = sweep_smem + kSweepHeader;
    const int W = P.window, EW = P.extWindow, zrow = W + EW;
    unsigned char* slots = xwin + 8 * kXs * (size_t) (zrow + 2);
    const size_t metaBytes = (size_t) P.metaCap * 4, valsBytes = (size_t) P.valsCap * 8, rhsBytes = (size_t) P.rhsCap * 24;
    const size_t slotBytes = metaBytes + valsBytes + rhsBytes;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NW = P.nwarps, NH = P.nhalo;
    const int nslots = P.nslots;
    constexpr int NF = LOWER ? 9 : 12;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nslots; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NW + 1); }
        for (int s = 0; s <= nslots; ++s) ext_ready[s] = 0;
        *issued = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x < 2 * kXs) reinterpret_cast<double*>(xwin)[kXs * (size_t) zrow + threadIdx.x] = 0.0;
    __syncthreads();
    const int nst = pr.stage_end - pr.stage_begin;

    if (warp == NW) {                               // ---- producer ----
        StageD cur = {0, 0, 0, 0, 0, 0}, nxt = cur;
        if (lane < nst) nxt = P.stages[pr.stage_begin + lane];
        for (int b0 = 0; b0 < nst; b0 += 32) {
            cur = nxt;
            if (b0 + 32 + lane < nst) nxt = P.stages[pr.stage_begin + b0 + 32 + lane];
            const int n = min(32, nst - b0);
            for (int k = 0; k < n; ++k) {
                const long long meta_off = __shfl_sync(kFull, cur.meta_off, k), vals_off = __shfl_sync(kFull, cur.vals_off, k);
                const unsigned bm = (unsigned) __shfl_sync(kFull, cur.meta_ints, k) * 4, bv = (unsigned) __shfl_sync(kFull, cur.vals_doubles, k) * 8;
                const int g_lo = __shfl_sync(kFull, cur.g_lo, k);
                const unsigned br = (unsigned) __shfl_sync(kFull, cur.g_rows, k) * 24;
                if (lane == 0) {
                    const int i = b0 + k, s = i % nslots;
                    if (i >= nslots) mbar_wait_relaxed(empty + s, ((i / nslots) - 1) & 1);
                    if (TRACE && i < P.trace_cap) P.trace[((size_t) part * P.trace_cap + i) * 4 + 3] = clock64();
                    unsigned char* base = slots + (size_t) s * slotBytes;
                    mbar_expect_tx(full + s, bm + bv + br);
                    bulk_g2s(base, P.meta + meta_off, bm, full + s);
                    if (bv) bulk_g2s(base + metaBytes, P.vals + vals_off, bv, full + s);
                    bulk_g2s(base + metaBytes + valsBytes, P.rhs + 3 * (size_t) g_lo, br, full + s);
                    st_volatile_s32(issued, i + 1);
                }
            }
        }
    } else if (warp > NW) {                         // ---- helpers: external rows ----
        // The stage's external rows are listed in the order the levels need them.  A helper keeps a WINDOW of
        // kHelperWindow x 32 rows in flight (one row per lane and sub-batch, three 8-byte loads each), parks every row as
        // soon as it has arrived and publishes the length of the finished PREFIX of the list: the consumers of a level
        // wait for `prefix >= rows needed up to this level`.  Polling one level group at a time would cap a part at one
        // level per L2 round trip (0.6-1 us) -- slower than the part upstream produces them, so that the lag grows by a
        // constant per level and per hop (measured: 2 us per level at the far corner of the c3 grid, profiles/).
        constexpr int kHelperWindow = 2;
        const int h = warp - NW - 1;
        double* ring = reinterpret_cast<double*>(xwin) + kXs * (size_t) W;
        // The list of a stage's external rows is static data: the helper reads it from the GLOBAL copy of the stage's meta
        // block (descriptor and header one stage of its own ahead), so it can start fetching stage i as soon as stage i - 1
        // has been ISSUED by the producer instead of when stage i has landed -- with the list read from the landed stage the rows arrived
        // 1-2 us into the stage and every consumer of the stage's first levels waited for them (+20 % per part).
        // Flow control: stage i - 1 issued => stage i - 1 - nslots is finished, so at most nslots + 1 stages have rows in
        // the ring (capacity checked by the host) and the ready words are indexed modulo nslots + 1.
        StageD sd = P.early && h < nst ? P.stages[pr.stage_begin + h] : StageD{0, 0, 0, 0, 0, 0};
        for (int i = h; i < nst; i += NH) {
            const int s = i % nslots, rdy = i % (nslots + 1);
            const int* m;
            if (P.early) m = P.meta + sd.meta_off;
            else {                                         // small rings: the list of the landed stage, as before
                mbar_wait_relaxed(full + s, (i / nslots) & 1);
                m = reinterpret_cast<const int*>(slots + (size_t) s * slotBytes);
            }
            const int next = m[4], ext_base = m[8];
            const int* extl = m + m[5];
            int first_rows[kHelperWindow];
#pragma unroll
            for (int k = 0; k < kHelperWindow; ++k) first_rows[k] = 32 * k + lane < next ? extl[32 * k + lane] : 0;
            if (P.early) {
                if (i + NH < nst) sd = P.stages[pr.stage_begin + i + NH];
                // (a counter, not the full barrier of stage i - 1: that barrier may be two phases further when a late helper
                // looks, and a parity wait would then never return)
                int spins = 0;
                while (ld_volatile_s32(issued) < i) {
                    __nanosleep(100);
                    if ((++spins & 1023) == 0 && (spins > (1 << 22) || *((volatile int*) &P.S->trsv_timeout))) { P.S->trsv_timeout = 1; break; }
                }
            }
            const int tag = (i + 1) << 16;
            int published = 0;
            long long* t3 = TRACE && i < P.trace_cap ? P.trace + (size_t) 2 * 148 * P.trace_cap * 4 + ((size_t) part * P.trace_cap + i) * 4 : nullptr;
            if (TRACE && t3 && lane == 0) { t3[0] = clock64(); t3[2] = next; }
            for (int e0 = 0; e0 < next; e0 += 32 * kHelperWindow) {
                const int wend = min(e0 + 32 * kHelperWindow, next);
                bool done[kHelperWindow];
                const double* xp[kHelperWindow];
#pragma unroll
                for (int k = 0; k < kHelperWindow; ++k) {
                    const int idx = e0 + 32 * k + lane;
                    done[k] = idx >= wend;
                    xp[k] = P.out + 3 * (size_t) (done[k] ? 0 : (e0 == 0 ? first_rows[k] : extl[idx]));
                }
                int spins = 0;
                while (true) {
                    double x[kHelperWindow][3];
#pragma unroll
                    for (int k = 0; k < kHelperWindow; ++k)
                        if (!done[k]) { x[k][0] = ld_relaxed(xp[k]); x[k][1] = ld_relaxed(xp[k] + 1); x[k][2] = ld_relaxed(xp[k] + 2); }
#pragma unroll
                    for (int k = 0; k < kHelperWindow; ++k)
                        if (!done[k] && (P.nowait || !(is_sentinel(x[k][0]) || is_sentinel(x[k][1]) || is_sentinel(x[k][2])))) {
                            double* d = ring + kXs * (size_t) ((ext_base + e0 + 32 * k + lane) & (EW - 1));
                            d[0] = x[k][0]; d[1] = x[k][1]; d[2] = x[k][2];
                            done[k] = true;
                        }
                    int prefix = e0;
                    bool all = true;
#pragma unroll
                    for (int k = 0; k < kHelperWindow; ++k) {
                        const unsigned bm = __ballot_sync(kFull, done[k]);
                        if (all) {
                            if (bm == kFull) prefix += 32;
                            else { prefix += __ffs(~bm) - 1; all = false; }
                        }
                    }
                    prefix = min(prefix, wend);
                    if (prefix > published) {
                        __threadfence_block();
                        __syncwarp();
                        if (lane == 0) st_volatile_s32(ext_ready + rdy, tag + prefix);
                        published = prefix;
                    }
                    if (prefix >= wend) break;
                    if ((++spins & 255) == 0 && (spins > (1 << 20) || *((volatile int*) &P.S->trsv_timeout))) {
                        P.S->trsv_timeout = 1;
                        if (lane == 0) st_volatile_s32(ext_ready + rdy, tag + next);
                        break;
                    }
                    if (P.helper_sleep) __nanosleep(P.helper_sleep);
                }
            }
            __syncwarp();
            if (TRACE && t3 && lane == 0) t3[1] = clock64();
            mbar_wait_relaxed(full + s, (i / nslots) & 1);          // arrive in the phase of stage i
            if (lane == 0) mbar_arrive(empty + s);
        }
    } else {
    // ---- consumers ----
    const int q = lane / 3, comp = lane - 3 * q;
    const int nthreads = NW * 32;
    const int lane_row = LOWER ? lane : comp - 3 * q;          // index of this lane's entry relative to 3 * g0 (rhs and out)
    const int lane_rhs = LOWER ? lane : -3 * q;                // first rhs entry this lane reads, relative to 3 * g0
    double* const out_lane = P.out + lane_row;
    double* const rearm_lane = REARM ? P.rearm + lane_row : nullptr;
    const bool tracing = TRACE && threadIdx.x == 0;
    const unsigned xw = smem_u32(xwin);
    const unsigned slots32 = smem_u32(slots);
    double carry = 0.0;
    int s = 0;
    unsigned parity = 0;
    for (int i = 0; i < nst; ++i) {
        if (tracing && i < P.trace_cap) P.trace[((size_t) part * P.trace_cap + i) * 4 + 0] = clock64();
        mbar_wait(full + s, parity);
        if (tracing && i < P.trace_cap) P.trace[((size_t) part * P.trace_cap + i) * 4 + 1] = clock64();
        const unsigned sb = slots32 + (unsigned) (s * slotBytes);
        const int4 h0 = lds_s32x4(sb);                                 // {ngroups, nrecords, g_lo, rhs rows}
        const int4 h1 = lds_s32x4(sb + 16);                            // {next, off_ext, off_wl, off_items}
        const unsigned wl = sb + 4 * (unsigned) h1.z;
        int tb, te, tail;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(tb) : "r"(wl + 4 * warp));
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(te) : "r"(wl + 4 * warp + 4));
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(tail) : "r"(wl + 4 * (NW + 1 + warp)));
        unsigned it_a = sb + 4 * (unsigned) h1.w + 16 * (unsigned) tb;                                  // items
        unsigned cd_a = sb + 4 * (unsigned) (h1.w + 4 * h0.y) + 16 * lane + 512 * (unsigned) tb;         // codes
        unsigned v_a = sb + (unsigned) metaBytes + 16 * lane + (unsigned) tb * (NF * 256);               // value pairs
        const unsigned v8_off = 2048 + 8 * lane - 16 * lane;                                             // lower: the ninth value
        const unsigned rr = sb + (unsigned) (metaBytes + valsBytes) + 8 * lane_rhs;
        const int tag = (i + 1) << 16, rdy = i % (nslots + 1);
        long long tp_last = TRACE ? clock64() : 0, tp_pre = 0, tp_bar = 0, tp_post = 0;
        for (int t = tb; t < te; ++t, it_a += 16, cd_a += 512, v_a += NF * 256) {
            // ---- operands of record t: nothing here depends on another row
            const int4 item = lds_s32x4(it_a);                // {rhs byte offset, first | last << 1 | count << 4 | barriers << 16, ext_need, 3 g0}
            const int4 cd = lds_s32x4(cd_a);                  // byte offsets in xwin of the dependencies and of the result
            const double2 a01 = lds_f64x2(v_a);
            const double2 a23 = lds_f64x2(v_a + 512);
            const double2 a45 = lds_f64x2(v_a + 1024);
            const double2 a67 = lds_f64x2(v_a + 1536);
            double a8, r0, r1 = 0.0, r2 = 0.0, i0 = 0.0, i1 = 0.0, i2 = 0.0;
            if constexpr (LOWER) {
                a8 = lds_f64(v_a + v8_off);
                r0 = lds_f64(rr + item.x);
            } else {
                const double2 a89 = lds_f64x2(v_a + 2048);
                const double2 aAB = lds_f64x2(v_a + 2560);
                a8 = a89.x; i0 = a89.y; i1 = aAB.x; i2 = aAB.y;
                r0 = lds_f64(rr + item.x); r1 = lds_f64(rr + item.x + 8); r2 = lds_f64(rr + item.x + 16);
            }
            long long tc0 = 0, tc1 = 0;
            if (tracing) { tc0 = clock64(); tp_pre += tc0 - tp_last; }
            // ---- level barriers
            if (item.y >> 16) {
                named_barrier(1, nthreads);
                for (int b = (item.y >> 16) - 1; b > 0; --b) named_barrier(1, nthreads);
            }
            if (item.z) {                                     // external rows of this level: parked by a helper warp
                int spins = 0;
                while (ld_volatile_s32(ext_ready + rdy) - (tag + item.z) < 0) {
                    if ((++spins & 4095) == 0 && *((volatile int*) &P.S->trsv_timeout)) break;
                }
                __threadfence_block();
            }
            if (tracing) { tc1 = clock64(); tp_bar += tc1 - tc0; }
            // ---- dependent part
            const double2 x0 = lds_f64x2(xw + cd.x);
            const double2 x1 = lds_f64x2(xw + cd.y);
            const double2 x2 = lds_f64x2(xw + cd.z);
            const double x02 = lds_f64(xw + cd.x + 16);
            const double x12 = lds_f64(xw + cd.y + 16);
            const double x22 = lds_f64(xw + cd.z + 16);
            double acc;
            if constexpr (LOWER) acc = r0;
            else acc = fma(i2, r2, fma(i1, r1, i0 * r0));
            if (!(item.y & 1)) acc = carry;                   // continuation record of a long row
            const double t0 = fma(a23.x, x02, fma(a01.y, x0.y, a01.x * x0.x));
            const double t1 = fma(a45.y, x12, fma(a45.x, x1.y, a23.y * x1.x));
            const double t2 = fma(a8, x22, fma(a67.y, x2.y, a67.x * x2.x));
            acc = ((acc - t0) - t1) - t2;
            carry = acc;
            if (cd.w >= 0) {
                sts_f64(xw + cd.w, acc);
                st_relaxed(out_lane + item.w, acc);
                if (REARM) rearm_lane[item.w] = sentinel();
            }
            if (tracing) { tp_last = clock64(); tp_post += tp_last - tc1; }
        }
        for (int b = 0; b < tail; ++b) named_barrier(1, nthreads);
        __syncwarp();
        if (tracing && i < P.trace_cap) {
            P.trace[((size_t) part * P.trace_cap + i) * 4 + 2] = clock64();
            // second half of the trace buffer: per stage {records of warp 0, cycles before / in / after the barriers}
            long long* t2 = P.trace + (size_t) 148 * P.trace_cap * 4 + ((size_t) part * P.trace_cap + i) * 4;
            t2[0] = te - tb; t2[1] = tp_pre; t2[2] = tp_bar; t2[3] = tp_post;
        }
        if (lane == 0) mbar_arrive(empty + s);
        if (++s == nslots) { s = 0; parity ^= 1; }
    }
    }
    if constexpr (SPMV == 1 || SPMV == 2) fused_spmv_tail<SPMV>(P, part, sweep_smem);
    if constexpr (SPMV == 3) {
        if (xupd) xupdate_tail(P, xupd_coef);
    }
}

// ---- round-2 sweep kernel (schedule and rationale: sweep2.hpp, sweep2.cuh) ---------------------------------------------------
// SPMV as k_sweep: -1 sweep only, 1 / 2 go on with the SpMV that follows (fused_spmv_tail), 3 lower sweep + pending x update.
constexpr int kS2Threads = 512;        // 15 consumer warps + an idle one (tails only) at 128 registers: the register file is handed out in
                                       // units that make 136 and 144 registers x 15 / 14 warps not fit (occupancy query: 0)
// ML: the schedule has rows that take several lanes (more than three dependencies: fault connections, wells folded into the
// matrix); the other variant carries none of that code (its address masks, lane meta and shuffles cost 15 % on the 1 M-cell grid).
template <bool LOWER, bool REARM, int SPMV = -1, bool TRACE = false, bool ML = false>
__global__ void __launch_bounds__(kS2Threads) k_sweep2(const SweepArgs P)
{
    extern __shared__ __align__(128) unsigned char sweep_smem[];
    pdl_enter();
    if (P.check_done && P.S->done) return;
    const int part = blockIdx.x;
    // CTAs beyond the parts (launched only by the variants with a tail, when the schedule has fewer parts than the device has
    // SMs): nothing to sweep, they go straight to the tail -- the SpMV / x update then runs on every SM, not on the parts' SMs only
    const bool has_part = part < P.nparts;
    if (!has_part && SPMV == -1) return;
    const bool xupd = SPMV == 3 && P.S->pend_on != 0;
    const double xupd_coef = SPMV == 3 ? P.S->pend : 0.0;
    const Sweep2Args& V = P.v2;
    S2PartD pr = {};
    if (has_part) pr = V.parts[part];
    int* hp = reinterpret_cast<int*>(sweep_smem);
    const int W = V.window;
    double* xyp = reinterpret_cast<double*>(sweep_smem + kS2Header);      // window of recent rows, components 0 and 1: 16 bytes per slot
    double* zp = xyp + 2 * (size_t) (W + 1);                                // component 2: 8 bytes per slot; slot W is the all-zero row
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 16) hp[threadIdx.x] = 0;
    if (TRACE && threadIdx.x == 0) { P.trace[8 * part] = globaltimer_ns(); for (int k = 1; k < 8; ++k) P.trace[8 * part + k] = 0; }
    if (threadIdx.x == 32) { xyp[2 * (size_t) W] = 0.0; xyp[2 * (size_t) W + 1] = 0.0; zp[W] = 0.0; }
    __syncthreads();
    constexpr int NP = LOWER ? 14 : 18;

    if (warp < pr.ncw) {
        // ---- consumers: warp w walks stream w of the part
        const S2StreamD sd = V.streams[pr.stream0 + warp];
        const int nrec = sd.nrec;
        const int2* __restrict__ hdrs = V.hdrs + sd.hdr_off;
        const double* vp = V.vals + sd.vals_off + 2 * lane;
        const int4* cp = V.codes + sd.code_off + lane;
        const unsigned xy = smem_u32(xyp), z = smem_u32(zp);
        const int zcode = 8 * W;
        const int2 hzero = make_int2(0, 0);
        int2 hb = lane < nrec ? __ldg(hdrs + lane) : hzero;                // headers of the records 32 b + lane
        int2 hbn = 32 + lane < nrec ? __ldg(hdrs + 32 + lane) : hzero;
        auto header = [&](int j) { return make_int2(__shfl_sync(kFull, hb.x, j), __shfl_sync(kFull, hb.y, j)); };
        S2Ops<LOWER> o;
#pragma unroll
        for (int k = 0; k < NP; ++k) o.v[k] = make_double2(0.0, 0.0);
        o.r0 = o.r1 = o.r2 = 0.0;
        int my_step = -1;                  // step of the current record (headers carry the distance to the warp's previous record)
        const int poll_lead = max(1, P.helper_sleep);
        // issue the loads of the record described by hd (nothing waits for them here) and advance the stream
        auto fetch = [&](const int2& hd) {
            const int cnt = hd.y & 63;
            if (lane < cnt) {
                const double* v = vp;
                const size_t stride = 2 * (size_t) cnt;
#pragma unroll
                for (int k = 0; k < NP; ++k, v += stride)
                    if (!TRACE || !(P.nowait & 8)) o.v[k] = ldg_stream_f64x2(v);
                o.cd = ldg_stream_s32x4(cp);
                if ((hd.y & (S2D_FIRST << 6)) && (!TRACE || !(P.nowait & 16))) {
                    // (a record with rows that take several lanes: the row of the lane comes with the codes; its further lanes start from zero)
                    const bool multi = ML && (hd.y & (S2D_MULTI << 6));
                    const int meta = multi ? s2d_meta(o.cd) : 0;
                    const int g0 = hd.x & 0x1fffffff, ord = multi ? (meta & 31) : lane;
                    if (!(meta & (3 << 5))) {                                    // first lane of its row
                        const double* r = P.rhs + 3 * (size_t) (LOWER ? g0 + ord : g0 - ord);
                        o.r0 = r[0]; o.r1 = r[1]; o.r2 = r[2];
                    } else { o.r0 = 0.0; o.r1 = 0.0; o.r2 = 0.0; }
                }
            } else o.cd = make_int4(zcode | (zcode << 16), zcode, -1, -1);
            vp += (size_t) (2 * NP) * cnt;
            cp += cnt;
        };
        // The loads of a record are issued when the warp has finished its previous one, (warps / chunks per step) steps ahead of
        // their use -- with 15 warps and 3 chunks per step that is 5 step times, and a warp needs HBM latency + its own work
        // (~1500 + 700 cycles) per record, i.e. >= 440 cycles per step.  So the bytes of the record AFTER the next one are
        // pulled into L2 at the same time (no registers needed): the loads that follow find them there.
        const unsigned char* pf = reinterpret_cast<const unsigned char*>(V.vals + sd.vals_off);      // start of the record after the next
        const int pfd = max(2, min(P.early, 8));       // records ahead of the current one whose bytes are pulled into L2
        auto prefetch_l2 = [&](int idx, int cur) {     // cur: record whose header batch is in hb
            if (idx >= nrec) return;
            const int j = idx & 31;
            const int y = __shfl_sync(kFull, (idx >> 5) == (cur >> 5) ? hb.y : hbn.y, j);                        // header batch of idx
            const unsigned bytes = 16u * NP * (unsigned) (y & 63);
            if (lane == 0 && bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pf), "r"(bytes) : "memory");
            pf += bytes;
        };
        int2 h = header(0);
        my_step += (int) (((unsigned) h.x >> 30) | (((unsigned) h.y >> 30) << 2));
        if (nrec > 0) { for (int k = 0; k < pfd; ++k) prefetch_l2(k, 0); fetch(h); }
        double y0 = 0.0, y1 = 0.0, y2 = 0.0;
        // debugging aid: where the cycles of this warp go (parts 0 and nparts / 2 of a traced launch)
        const bool prof = TRACE && P.trace_cap < 0 && (part == 0 || part == P.nparts / 2);      // trace_cap < 0: per-warp cycle accounting as well (spills: slow)
        long long pc[6] = {0, 0, 0, 0, 0, 0}, c0 = prof ? clock64() : 0;
        for (int i = 0; i < nrec; ++i) {
            const int cnt = h.y & 63, flags = (h.y >> 6) & 63;
            // everything the dependent part needs besides the dependencies themselves is computed BEFORE the barrier: a lone warp
            // issues an instruction every 4-5 cycles, so every instruction between bar.sync and bar.arrive is on the level's chain
            const unsigned cmask = ML ? 0xfff8u : 0xffffu;
            const unsigned d0 = (unsigned) o.cd.x & cmask, d1 = ML ? ((unsigned) o.cd.x >> 16) & cmask : (unsigned) o.cd.x >> 16, d2 = (unsigned) o.cd.y & cmask, oc = (unsigned) o.cd.y >> 16;
            const unsigned ax0 = xy + 2 * d0, ax1 = xy + 2 * d1, ax2 = xy + 2 * d2, az0 = z + d0, az1 = z + d1, az2 = z + d2;
            const unsigned sx = xy + 2 * oc, sz = z + oc;
            const bool multi = ML && (flags & S2D_MULTI);
            const int meta = multi ? s2d_meta(o.cd) : 0;                       // rows that take several lanes: row, position and width of the lane
            const bool store = (flags & S2D_LAST) && lane < cnt && !(meta & (3 << 5));
            const int bar_in = (h.y >> 12) & 15, bar_in_n = 32 * ((h.y >> 20) & 31), bar_out = (h.y >> 16) & 15, bar_out_n = 32 * ((h.y >> 25) & 31);
            const double* v = reinterpret_cast<const double*>(o.v);
            // start value: nothing here depends on another row of the part
            if (flags & S2D_FIRST) {
                if constexpr (LOWER) { y0 = o.r0; y1 = o.r1; y2 = o.r2; }
                else {
                    y0 = fma(v[29], o.r2, fma(v[28], o.r1, v[27] * o.r0));
                    y1 = fma(v[32], o.r2, fma(v[31], o.r1, v[30] * o.r0));
                    y2 = fma(v[35], o.r2, fma(v[34], o.r1, v[33] * o.r0));
                }
            }
            if (prof) { const long long c = clock64(); pc[0] += c - c0; c0 = c; }      // fetch issue, header, start value
            if (flags & S2D_EXT) {
                // External dependencies (rows of other parts: the value is its own ready flag).  The rows a step needs are one
                // level behind it, i.e. their owners produce them about when this part runs the step before: polling them from
                // the moment the warp is free (several steps early) only loads L2 and the SM's load pipe (15 warps x 5 rows x 3
                // loads per round trip and SM).  So the warp sleeps until the part is two steps away from its record, then polls
                // until the sentinel is gone, and folds the rows' contribution into the start value -- all before it waits for
                // the previous step, outside the dependent chain.
                double xa0 = 0.0, xa1 = 0.0, xa2 = 0.0, xb0 = 0.0, xb1 = 0.0, xb2 = 0.0;
                if ((P.nowait & 3) < 2) {
                    while (ld_volatile_s32(hp + 11) <= my_step - poll_lead) __nanosleep(64);      // until step my_step - poll_lead has started
                    int spins = 0;
                    unsigned t0 = 0;
                    bool wa = o.cd.z >= 0, wb = o.cd.w >= 0;
                    while (true) {
                        if (wa) { const double* x = P.out + 3 * (size_t) o.cd.z; xa0 = ld_relaxed(x); xa1 = ld_relaxed(x + 1); xa2 = ld_relaxed(x + 2); }
                        if (wb) { const double* x = P.out + 3 * (size_t) o.cd.w; xb0 = ld_relaxed(x); xb1 = ld_relaxed(x + 1); xb2 = ld_relaxed(x + 2); }
                        wa = wa && (is_sentinel(xa0) || is_sentinel(xa1) || is_sentinel(xa2));
                        wb = wb && (is_sentinel(xb0) || is_sentinel(xb1) || is_sentinel(xb2));
                        if (!__any_sync(kFull, wa || wb)) break;
                        if ((++spins & 255) == 0) {
                            // (elapsed time, not a spin count: kWaitTimeoutNs without the row is a deadlock, not a slow neighbour)
                            if (t0 == 0) t0 = wall_lo_ns();
                            else if (ld_volatile_s32(hp + 9) || *((volatile int*) &P.S->trsv_timeout) || wall_lo_ns() - t0 > kWaitTimeoutLoNs) {
                                P.S->trsv_timeout = 1; st_volatile_s32(hp + 9, 1);
                                break;
                            }
                        }
                    }
                }
                if (o.cd.z >= 0) {
                    y0 -= fma(v[20], xa2, fma(v[19], xa1, v[18] * xa0));
                    y1 -= fma(v[23], xa2, fma(v[22], xa1, v[21] * xa0));
                    y2 -= fma(v[26], xa2, fma(v[25], xa1, v[24] * xa0));
                }
                if (o.cd.w >= 0) {
                    y0 -= fma(v[11], xb2, fma(v[10], xb1, v[9] * xb0));
                    y1 -= fma(v[14], xb2, fma(v[13], xb1, v[12] * xb0));
                    y2 -= fma(v[17], xb2, fma(v[16], xb1, v[15] * xb0));
                }
            }
            if (prof) { const long long c = clock64(); pc[2] += c - c0; c0 = c; }      // external rows
            const bool lead = (h.x >> 29) & 1;                              // this record counts the part's steps
            const bool tr = TRACE && lead && lane == 0;                     // debugging aid: timeline of the part's steps
            if (flags & S2D_SYNC) bar_sync_n(bar_in, bar_in_n);                        // the previous step is complete
            if (prof) {      // (bar.sync does not block at issue: a shared load behind it does)
                if (ld_volatile_s32(hp + 9) == 12345) pc[5]++;
                const long long c = clock64(); pc[1] += c - c0; c0 = c;                // waiting for the previous step
            }
            if (tr) {                                                     // stores only: nothing here waits for global memory
                const int st = ld_volatile_s32(hp + 10);                  // LEAD records of a part are its steps in order (one warp at a time)
                st_volatile_s32(hp + 10, st + 1);
                P.trace[8 * part + 5] = st + 1;
                if (st < 256) P.trace[8192 + 2 * (256 * part + st)] = globaltimer_ns();
            }
            // ---- dependent part
            const double2 a0 = lds_f64x2(ax0), a1 = lds_f64x2(ax1), a2 = lds_f64x2(ax2);
            const double b0 = lds_f64(az0), b1 = lds_f64(az1), b2 = lds_f64(az2);
            if (prof) {      // the shared loads have landed
                if (__double_as_longlong(a0.x + a1.x + a2.x + b0 + b1 + b2) == 0x7ff123456789abcdLL) pc[5]++;
                const long long c = clock64(); pc[5] += c - c0; c0 = c;
            }
            const double p0 = fma(v[2], b0, fma(v[1], a0.y, v[0] * a0.x));
            const double p1 = fma(v[5], b0, fma(v[4], a0.y, v[3] * a0.x));
            const double p2 = fma(v[8], b0, fma(v[7], a0.y, v[6] * a0.x));
            const double q0 = fma(v[11], b1, fma(v[10], a1.y, v[9] * a1.x));
            const double q1 = fma(v[14], b1, fma(v[13], a1.y, v[12] * a1.x));
            const double q2 = fma(v[17], b1, fma(v[16], a1.y, v[15] * a1.x));
            const double s0 = fma(v[20], b2, fma(v[19], a2.y, v[18] * a2.x));
            const double s1 = fma(v[23], b2, fma(v[22], a2.y, v[21] * a2.x));
            const double s2 = fma(v[26], b2, fma(v[25], a2.y, v[24] * a2.x));
            y0 = ((y0 - p0) - q0) - s0; y1 = ((y1 - p1) - q1) - s1; y2 = ((y2 - p2) - q2) - s2;
            if (multi && (flags & S2D_LAST)) {
                // rows that take several lanes: add the partial sums of the row's lanes (pairs (0, 1) and (2, 3), then 0 += 2)
                const int k = (meta >> 5) & 3, nsec = (meta >> 7) & 3;
                double t0 = __shfl_down_sync(kFull, y0, 1), t1 = __shfl_down_sync(kFull, y1, 1), t2 = __shfl_down_sync(kFull, y2, 1);
                if (!(k & 1) && k + 1 <= nsec) { y0 += t0; y1 += t1; y2 += t2; }
                t0 = __shfl_down_sync(kFull, y0, 2); t1 = __shfl_down_sync(kFull, y1, 2); t2 = __shfl_down_sync(kFull, y2, 2);
                if (k == 0 && nsec >= 2) { y0 += t0; y1 += t1; y2 += t2; }
            }
            if (prof) { if (__double_as_longlong(y0) == 0x7ff123456789abcdLL) pc[5]++; const long long c = clock64(); pc[3] += c - c0; c0 = c; }   // dependencies + fma
            if (store) { sts_f64x2(sx, y0, y1); sts_f64(sz, y2); }
            if (flags & S2D_ARRIVE) bar_arrive_n(bar_out, bar_out_n);                  // this warp's share of the step is in shared memory
            if (store && (!TRACE || !(P.nowait & 4))) {
                const int g0 = h.x & 0x1fffffff, ord = multi ? (meta & 31) : lane;
                const size_t gi = 3 * (size_t) (LOWER ? g0 + ord : g0 - ord);
                st_relaxed(P.out + gi, y0); st_relaxed(P.out + gi + 1, y1); st_relaxed(P.out + gi + 2, y2);
                if (REARM) { P.rearm[gi] = sentinel(); P.rearm[gi + 1] = sentinel(); P.rearm[gi + 2] = sentinel(); }
            }
            if (prof) { const long long c = clock64(); pc[4] += c - c0; c0 = c; }      // stores, arrive
            if (lead && lane == 0) st_volatile_s32(hp + 11, ld_volatile_s32(hp + 11) + 1);      // steps of the part that have started
            // ---- operands of the next record of this warp: several steps ahead of their use
            if (i + 1 < nrec) {
                const int j = (i + 1) & 31;
                if (j == 0) { hb = hbn; hbn = i + 33 + lane < nrec ? __ldg(hdrs + i + 33 + lane) : hzero; }
                h = header(j);
                my_step += (int) (((unsigned) h.x >> 30) | (((unsigned) h.y >> 30) << 2));
                fetch(h);
                prefetch_l2(i + pfd, i + 1);
            }
        }
        if (prof && lane == 0) {
            long long* o2 = P.trace + 8192 + 2 * 256 * 1024 + (part == 0 ? 0 : 256) + 8 * warp;
            for (int k = 0; k < 6; ++k) o2[k] = pc[k];
            o2[6] = nrec; o2[7] = clock64() - c0;
        }
    }
    if constexpr (SPMV == 1 || SPMV == 2) fused_spmv_tail<SPMV>(P, part, sweep_smem);
    if constexpr (SPMV == 3) {
        if (xupd) xupdate_tail(P, xupd_coef);
    }
}

// ---- BSR SpMV with fused dot products ----------------------------------------------------------

// 3 lanes per block row (lane = scalar row).  MODE 0: y = A x.  MODE 1: + h = <d1, y>.
// MODE 2: + tr = <y, d1>, tt = <y, y>.  The totals land in S (raw sums; alpha/omega are derived by
// the consumers so that the well kernel can still correct them).
template <int MODE>
__global__ void __launch_bounds__(kVecThreads) k_spmv(const int* __restrict__ prow, const int* __restrict__ pcol,
                                                      const double* __restrict__ A, const double* __restrict__ x,
                                                      double* __restrict__ y, const double* __restrict__ d1, int N,
                                                      Scalars* S, double* partials, unsigned* ticket)
{
    pdl_enter();
    if (MODE != 0 && S->done) return;
    double acc[MODE == 2 ? 2 : 1] = {0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const int row = i / 3, comp = i - 3 * row;
        const int ks = prow[row], ke = prow[row + 1];
        double s = 0.0;
        for (int k = ks; k < ke; ++k) {
            const double* a = A + (size_t) k * 9 + comp * 3;
            const double* xx = x + 3 * (size_t) pcol[k];
            s += a[0] * xx[0] + a[1] * xx[1] + a[2] * xx[2];
        }
        y[i] = s;
        if (MODE == 1) acc[0] += d1[i] * s;
        if (MODE == 2) { acc[0] += s * d1[i]; acc[1] += s * s; }
    }
    if (MODE != 0) {
        double tot[MODE == 2 ? 2 : 1];
        if (grid_reduce<(MODE == 2 ? 2 : 1)>(acc, partials, ticket, tot)) {
            if (MODE == 1) S->h = tot[0];
            if (MODE == 2) { S->tr = tot[0]; S->tt = tot[1]; }
        }
    }
}

// The same product from the sliced-ELL copy (analysis.hpp SellPlan): one lane per block row, a warp per slice of 32 rows,
// every value load is 256 contiguous bytes.  The BSR kernel above needs ~5 L1 wavefronts per 128 useful bytes (3 lanes per
// row, 72-byte blocks) and is bound by the LSU data pipe (79 % busy at 72 % of the HBM peak); this layout needs one.
// Blocks beyond the slice width (over[slice] != 0) are read from the BSR arrays by the lane that owns the row.
template <int MODE>
__global__ void __launch_bounds__(kVecThreads) k_spmv_sell(const int* __restrict__ sptr, const int* __restrict__ sover,
                                                           const int* __restrict__ scol, const double* __restrict__ sval,
                                                           const int* __restrict__ prow, const int* __restrict__ pcol,
                                                           const double* __restrict__ A, const double* __restrict__ x,
                                                           double* __restrict__ y, const double* __restrict__ d1, int Nb, int nslices,
                                                           Scalars* S, double* partials, unsigned* ticket)
{
    pdl_enter();
    if (MODE != 0 && S->done) return;
    double acc[MODE == 2 ? 2 : 1] = {0.0};
    const int lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int slice = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; slice < nslices; slice += nw) {
        const int s0 = __ldg(sptr + slice), s1 = __ldg(sptr + slice + 1);
        const int row = 32 * slice + lane;
        const double* v = sval + (size_t) s0 * 288 + lane;
        const int* c = scol + (size_t) s0 * 32 + lane;
        double y0 = 0.0, y1 = 0.0, y2 = 0.0;
#pragma unroll 2
        for (int k = s0; k < s1; ++k, v += 288, c += 32) {
            const double* xx = x + 3 * (size_t) __ldg(c);
            const double x0 = xx[0], x1 = xx[1], x2 = xx[2];
            y0 += v[0] * x0 + v[32] * x1 + v[64] * x2;
            y1 += v[96] * x0 + v[128] * x1 + v[160] * x2;
            y2 += v[192] * x0 + v[224] * x1 + v[256] * x2;
        }
        if (__ldg(sover + slice) && row < Nb) {
            for (int k = prow[row] + (s1 - s0), ke = prow[row + 1]; k < ke; ++k) {
                const double* a = A + (size_t) k * 9;
                const double* xx = x + 3 * (size_t) pcol[k];
                const double x0 = xx[0], x1 = xx[1], x2 = xx[2];
                y0 += a[0] * x0 + a[1] * x1 + a[2] * x2;
                y1 += a[3] * x0 + a[4] * x1 + a[5] * x2;
                y2 += a[6] * x0 + a[7] * x1 + a[8] * x2;
            }
        }
        if (row < Nb) {
            double* yy = y + 3 * (size_t) row;
            yy[0] = y0; yy[1] = y1; yy[2] = y2;
            if (MODE != 0) {
                const double* dd = d1 + 3 * (size_t) row;
                const double e0 = dd[0], e1 = dd[1], e2 = dd[2];
                if (MODE == 1) acc[0] += e0 * y0 + e1 * y1 + e2 * y2;
                if (MODE == 2) { acc[0] += y0 * e0 + y1 * e1 + y2 * e2; acc[1] += y0 * y0 + y1 * y1 + y2 * y2; }
            }
        }
    }
    if (MODE != 0) {
        double tot[MODE == 2 ? 2 : 1];
        if (grid_reduce<(MODE == 2 ? 2 : 1)>(acc, partials, ticket, tot)) {
            if (MODE == 1) S->h = tot[0];
            if (MODE == 2) { S->tr = tot[0]; S->tt = tot[1]; }
        }
    }
}

// sval[(slot * 9 + e) * 32 + lane] = stage[src[slot * 32 + lane] * 9 + e]   (0 for padding); one thread per output double
__global__ void __launch_bounds__(256) k_fill_sell(const double* __restrict__ stage, const int* __restrict__ src,
                                                   double* __restrict__ sval, long long n)
{
    for (long long o = (long long) blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (long long) gridDim.x * blockDim.x) {
        const long long slot = o / 288;
        const int rem = (int) (o - slot * 288), e = rem >> 5, lane = rem & 31;
        const int b = __ldg(src + slot * 32 + lane);
        sval[o] = b >= 0 ? __ldg(stage + (long long) b * 9 + e) : 0.0;
    }
}

// ---- BiCGSTAB vector phases ------------------------------------------------------------------

// p = r + beta (p - omega v), beta = (rho_new/rho)(alpha/omega); first pass: p = r.
__global__ void __launch_bounds__(kVecThreads) k_vec_p(const double* __restrict__ r, double* __restrict__ p,
                                                       const double* __restrict__ v, int N, Scalars* S)
{
    pdl_enter();
    if (S->done) return;
    const bool first = S->first != 0;
    const double omega = S->omega;
    const double beta = first ? 0.0 : (S->rho_new / S->rho) * (S->alpha / omega);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
        p[i] = first ? r[i] : (p[i] - omega * v[i]) * beta + r[i];
}

// x += alpha y; r -= alpha v; norm = ||r||; y is re-armed (its last reader).  alpha = rho_new / h.
__global__ void __launch_bounds__(kVecThreads) k_vec_xr1(double* __restrict__ x, double* __restrict__ y, double* __restrict__ r,
                                                         const double* __restrict__ v, int N, Scalars* S, double* partials,
                                                         unsigned* ticket, int dist, int defer, const DistRedD D)
{
    pdl_enter();
    if (S->done) return;
    const double h = S->h;
    if (fabs(h) < 1e-80) {        // Dune: SolverAbort "abs(h) < EPSILON"
        if (blockIdx.x == 0 && threadIdx.x == 0) { S->breakdown = 1; S->done = 1; }
        return;
    }
    const double alpha = S->rho_new / h;
    double acc[1] = {0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        if (!defer) {
            x[i] += alpha * y[i];
            y[i] = sentinel();
        }
        const double rr = r[i] - alpha * v[i];
        r[i] = rr;
        acc[0] += rr * rr;
    }
    double tot[1];
    if (grid_reduce<1>(acc, partials, ticket, tot)) {
        S->alpha = alpha;
        if (defer) { S->pend = alpha; S->pend_on = 1; }
        if (dist && D.mail) { mail_allreduce<1>(D, S, tot); finish_xr1(S, tot[0]); }
        else if (dist) S->red[0] = tot[0];
        else finish_xr1(S, tot[0]);
    }
}

// x += omega y; r -= omega t; norm = ||r||; rho <- rho_new <- <rt, r>.  omega = tr / tt.
__global__ void __launch_bounds__(kVecThreads) k_vec_xr2(double* __restrict__ x, double* __restrict__ y, double* __restrict__ r,
                                                         const double* __restrict__ t, const double* __restrict__ rt, int N,
                                                         Scalars* S, double* partials, unsigned* ticket, int dist, int defer,
                                                         const DistRedD D)
{
    pdl_enter();
    if (S->done) return;
    const double omega = S->tr / S->tt;
    double acc[2] = {0.0, 0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        if (!defer) {
            x[i] += omega * y[i];
            y[i] = sentinel();
        }
        const double rr = r[i] - omega * t[i];
        r[i] = rr;
        acc[0] += rr * rr;
        acc[1] += rt[i] * rr;
    }
    double tot[2];
    if (grid_reduce<2>(acc, partials, ticket, tot)) {
        S->omega = omega;
        if (defer) { S->pend = omega; S->pend_on = 1; }
        if (dist && D.mail) { mail_allreduce<2>(D, S, tot); finish_xr2(S, tot[0], tot[1]); }
        else if (dist) { S->red[0] = tot[0]; S->red[1] = tot[1]; }
        else finish_xr2(S, tot[0], tot[1]);
    }
}

// ---- standard wells -------------------------------------------------------------------------------

// y -= C^T (D^-1 (B x)) for all standard wells, one CTA.  Phase 1: one warp per well reduces
// z1 = sum_p B_p x[col_p] over ALL perforations (8 perforations x 4 well equations per pass),
// z2 = D^-1 z1.  Phase 2: one thread per (unique perforated cell, component) gathers every
// contribution to that cell (no atomics, deterministic, two wells may share a cell) and patches
// the dot products the SpMV epilogue took before the wells were applied.
template <int MODE>
__global__ void __launch_bounds__(1024) k_wells(int nwells, const unsigned* __restrict__ wptr, const int* __restrict__ Bcols,
                                                const double* __restrict__ B, const double* __restrict__ C,
                                                const double* __restrict__ Dinv, int ncells, const int* __restrict__ ucell,
                                                const int* __restrict__ uptr, const int* __restrict__ ublock,
                                                const int* __restrict__ uwell, double* z2g, const double* __restrict__ x,
                                                double* y, const double* __restrict__ d1, Scalars* S)
{
    pdl_enter();
    if (MODE != 0 && S->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    // z2 of the first kWellSmem wells stays in shared memory (no global write -> read round trip between the phases)
    constexpr int kWellSmem = 1024;
    __shared__ double z2s[kWellSmem * 4];
    // operands of this thread's first phase-2 item that do not depend on phase 1: issued now, they arrive while the phase-1
    // chain (well pointer -> column -> x) runs
    const int t0 = threadIdx.x, u0 = t0 / 3, c0 = t0 - 3 * u0;
    const bool has0 = t0 < 3 * ncells;
    int pe0 = 0, pe1 = 0;
    size_t pidx = 0;
    double pold = 0.0, pd1 = 0.0;
    if (has0) {
        pe0 = uptr[u0]; pe1 = uptr[u0 + 1];
        pidx = 3 * (size_t) ucell[u0] + c0;
        pold = y[pidx];
        if (MODE != 0) pd1 = d1[pidx];
    }
    // phase 1: one LANE per perforation (all four well equations), two wells per warp and pass so that their loads overlap:
    // the kernel is a chain of dependent global loads (column -> x), not arithmetic
    for (int w0 = warp; w0 < nwells; w0 += 2 * nwarp) {
        const int w1 = w0 + nwarp;
        const bool two = w1 < nwells;
        const unsigned b0 = wptr[w0], e0 = wptr[w0 + 1], b1 = two ? wptr[w1] : 0u, e1 = two ? wptr[w1 + 1] : 0u;
        double z[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
        for (unsigned off = 0; off < max(e0 - b0, e1 - b1); off += 32) {
            const unsigned p[2] = {b0 + off + lane, b1 + off + lane};
            const bool on[2] = {p[0] < e0, p[1] < e1};
            double xx[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
            const double* bp[2] = {B + (size_t) p[0] * 12, B + (size_t) p[1] * 12};
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (on[h]) {
                    const double* xp = x + 3 * (size_t) Bcols[p[h]];
                    xx[h][0] = xp[0]; xx[h][1] = xp[1]; xx[h][2] = xp[2];
                }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    if (on[h]) z[h][r] += bp[h][3 * r] * xx[h][0] + bp[h][3 * r + 1] * xx[h][1] + bp[h][3 * r + 2] * xx[h][2];
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h == 1 && !two) break;
            const int w = h ? w1 : w0;
#pragma unroll
            for (int r = 0; r < 4; ++r) z[h][r] = warp_sum(z[h][r]);
            if (lane < 4) {
                const double* dd = Dinv + (size_t) w * 16 + lane * 4;
                const double z2 = dd[0] * z[h][0] + dd[1] * z[h][1] + dd[2] * z[h][2] + dd[3] * z[h][3];
                if (w < kWellSmem) z2s[w * 4 + lane] = z2;
                else z2g[w * 4 + lane] = z2;
            }
        }
    }
    __syncthreads();
    double acc[2] = {0.0, 0.0};
    for (int t = threadIdx.x; t < 3 * ncells; t += blockDim.x) {
        const int u = t / 3, c = t - 3 * u;
        double delta = 0.0;
        const bool pre = t == t0;
        for (int e = pre ? pe0 : uptr[u], ee = pre ? pe1 : uptr[u + 1]; e < ee; ++e) {
            const double* cb = C + (size_t) ublock[e] * 12 + c;
            const int w = uwell[e];
            const double* zz = w < kWellSmem ? z2s + 4 * w : z2g + 4 * w;
            delta += cb[0] * zz[0] + cb[3] * zz[1] + cb[6] * zz[2] + cb[9] * zz[3];
        }
        const size_t idx = pre ? pidx : 3 * (size_t) ucell[u] + c;
        const double old = pre ? pold : y[idx], now = old - delta;
        y[idx] = now;
        const double dd1 = MODE != 0 ? (pre ? pd1 : d1[idx]) : 0.0;
        if (MODE == 1) acc[0] += dd1 * (now - old);
        if (MODE == 2) { acc[0] += dd1 * (now - old); acc[1] += now * now - old * old; }
    }
    if (MODE != 0) {
        __shared__ double sm[2][32];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            double v = warp_sum(acc[i]);
            if (lane == 0) sm[i][warp] = v;
        }
        __syncthreads();
        if (warp == 0) {
            double a = lane < nwarp ? sm[0][lane] : 0.0, b = lane < nwarp ? sm[1][lane] : 0.0;
            a = warp_sum(a); b = warp_sum(b);
            if (lane == 0) {
                if (MODE == 1) S->h += a;
                if (MODE == 2) { S->tr += a; S->tt += b; }
            }
        }
    }
}

// The same apply with every index chain resolved on the host (the wells are re-uploaded for every solve anyway): used when
// the wells fit one CTA's registers and shared memory (<= 1024 perforations, <= 3072 (cell, component) items, <= 128 wells;
// C3: 1000 / 3000 / 50), k_wells otherwise.  k_wells is a chain of dependent global loads (well pointer -> column -> x, then
// list pointer -> block -> C); here a thread owns ONE perforation in phase 1 and up to three (cell, component) items in
// phase 2, and everything that does not depend on x or z2 -- column, B block, D^-1 row, the item's y index, its four C
// entries, its well -- is loaded at the top of the kernel from arrays laid out in thread order.  What is left on the
// critical path: one gather of x, two shared-memory steps, the y update and the block reduction.
//   phase 1: part[p][r] = (B_p x_p)[r]; thread (w, r) adds the parts of well w in perforation order (deterministic, the
//            oracle's order) and applies D^-1.
//   phase 2: item t = (unique cell u, component c): y -= sum_e C_e[:, c] . z2[well(e)]; the first contribution comes from
//            the pre-gathered itemC, further ones (two wells in one cell) walk the generic lists.
struct WellsFlatD {
    int nwells, nblocks, nitems;
    const unsigned* wptr;     // [nwells + 1]
    const int* bcol;          // [nblocks] p-space column of B_p
    const double* B;          // [nblocks][4][3]
    const double* Dinv;       // [nwells][4][4]
    const int4* item;         // [nitems] {3 * cell + c, contributions, first list entry, 4 * well of the first contribution}
    const double* itemC;      // [nitems][4] C entries of the first contribution: C[k][c], k = 0..3
    // generic lists for items with more than one contribution
    const double* C;
    const int* ublock;
    const int* uwell;
};

template <int MODE>
__global__ void __launch_bounds__(1024) k_wells_flat(WellsFlatD W, const double* __restrict__ x, double* y,
                                                     const double* __restrict__ d1, Scalars* S)
{
    pdl_enter();
    constexpr int kMaxWells = 128, kItems = 3;
    __shared__ double part[1024 * 4];
    __shared__ double z1s[kMaxWells * 4];
    __shared__ double z2s[kMaxWells * 4];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarp = blockDim.x >> 5;
    // ---- loads that depend on nothing computed here
    const bool hasp = t < W.nblocks;
    int col = 0;
    double b[12];
    if (hasp) {
        col = W.bcol[t];
        const double2* bp = reinterpret_cast<const double2*>(W.B + (size_t) t * 12);
#pragma unroll
        for (int i = 0; i < 6; ++i) { const double2 v = bp[i]; b[2 * i] = v.x; b[2 * i + 1] = v.y; }
    }
    const bool hasw = t < 4 * W.nwells;
    double dinv[4] = {0.0, 0.0, 0.0, 0.0};
    unsigned wb = 0, we = 0;
    if (hasw) {
        const int w = t >> 2;
        wb = W.wptr[w]; we = W.wptr[w + 1];
        const double2* dp = reinterpret_cast<const double2*>(W.Dinv + (size_t) t * 4);
        const double2 v0 = dp[0], v1 = dp[1];
        dinv[0] = v0.x; dinv[1] = v0.y; dinv[2] = v1.x; dinv[3] = v1.y;
    }
    int4 it[kItems];
    double ic[kItems][4], yold[kItems], dd1[kItems];
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const int q = t + k * 1024;
        it[k] = make_int4(0, 0, 0, 0);
        yold[k] = 0.0; dd1[k] = 0.0;
        if (q < W.nitems) {
            it[k] = W.item[q];
            const double2* cp = reinterpret_cast<const double2*>(W.itemC + (size_t) q * 4);
            const double2 v0 = cp[0], v1 = cp[1];
            ic[k][0] = v0.x; ic[k][1] = v0.y; ic[k][2] = v1.x; ic[k][3] = v1.y;
        }
    }
    const int done = MODE != 0 ? S->done : 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k)
        if (t + k * 1024 < W.nitems) {
            yold[k] = y[it[k].x];
            if (MODE != 0) dd1[k] = d1[it[k].x];
        }
    if (done) return;
    // ---- phase 1
    if (hasp) {
        const double* xp = x + 3 * (size_t) col;
        const double x0 = xp[0], x1 = xp[1], x2 = xp[2];
#pragma unroll
        for (int r = 0; r < 4; ++r) part[t * 4 + r] = b[3 * r] * x0 + b[3 * r + 1] * x1 + b[3 * r + 2] * x2;
    }
    __syncthreads();
    if (hasw) {
        const int r = t & 3;
        double z = 0.0;
        for (unsigned p = wb; p < we; ++p) z += part[p * 4 + r];
        z1s[t] = z;
    }
    __syncthreads();
    if (hasw) {
        const double* zz = z1s + (t & ~3);
        z2s[t] = dinv[0] * zz[0] + dinv[1] * zz[1] + dinv[2] * zz[2] + dinv[3] * zz[3];
    }
    __syncthreads();
    // ---- phase 2
    double acc[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        if (t + k * 1024 >= W.nitems) continue;
        const double* zz = z2s + it[k].w;
        double delta = ic[k][0] * zz[0] + ic[k][1] * zz[1] + ic[k][2] * zz[2] + ic[k][3] * zz[3];
        if (it[k].y > 1) {
            const int c = it[k].x % 3;
            for (int e = it[k].z + 1, ee = it[k].z + it[k].y; e < ee; ++e) {
                const double* cb = W.C + (size_t) W.ublock[e] * 12 + c;
                const double* z2 = z2s + 4 * W.uwell[e];
                delta += cb[0] * z2[0] + cb[3] * z2[1] + cb[6] * z2[2] + cb[9] * z2[3];
            }
        }
        const double now = yold[k] - delta;
        y[it[k].x] = now;
        if (MODE == 1) acc[0] += dd1[k] * (now - yold[k]);
        if (MODE == 2) { acc[0] += dd1[k] * (now - yold[k]); acc[1] += now * now - yold[k] * yold[k]; }
    }
    if (MODE != 0) {
        __shared__ double sm[2][32];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            double v = warp_sum(acc[i]);
            if (lane == 0) sm[i][warp] = v;
        }
        __syncthreads();
        if (warp == 0) {
            double a = lane < nwarp ? sm[0][lane] : 0.0, bsum = lane < nwarp ? sm[1][lane] : 0.0;
            a = warp_sum(a); bsum = warp_sum(bsum);
            if (lane == 0) {
                if (MODE == 1) S->h += a;
                if (MODE == 2) { S->tr += a; S->tt += bsum; }
            }
        }
    }
}

// The same apply on a CLUSTER of 8 CTAs x 128 threads (round 2): one SM cannot pull the ~600 KB of cold sectors of a
// 1000-perforation well set faster than ~15 us, eight can.  Thread T = 128 * rank + t of the cluster does exactly what thread T
// of k_wells_flat does; what that kernel keeps in one CTA's shared memory is read here through distributed shared memory from
// the CTA that owns it (perforation p: CTA p / 128; well row 4 w + r: CTA (4 w + r) / 128), eight loads in flight per thread, in the
// same summation order -- y is bit-identical to k_wells_flat, the dot-product patch is summed in a different (fixed) order.
constexpr int kWellClCtas = 8, kWellClThreads = 128;
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_dsmem_f64(const double* local, unsigned rank)
{
    unsigned a = smem_u32(local), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
    return v;
}
template <int MODE>
__global__ void __cluster_dims__(kWellClCtas, 1, 1) __launch_bounds__(kWellClThreads)
k_wells_cluster(WellsFlatD W, const double* __restrict__ x, double* y, const double* __restrict__ d1, Scalars* S)
{
    pdl_enter();
    constexpr int kItems = 3, kAll = kWellClCtas * kWellClThreads;      // == the 1024 threads of k_wells_flat
    __shared__ double part[kWellClThreads * 4];
    __shared__ double z1s[kWellClThreads];
    __shared__ double z2s[kWellClThreads];
    __shared__ double sm[2][kWellClThreads / 32];
    const unsigned rank = cluster_ctarank();
    const int tl = threadIdx.x, t = (int) rank * kWellClThreads + tl, lane = tl & 31, warp = tl >> 5;
    // ---- loads that depend on nothing computed here
    const bool hasp = t < W.nblocks;
    int col = 0;
    double b[12];
    if (hasp) {
        col = W.bcol[t];
        const double2* bp = reinterpret_cast<const double2*>(W.B + (size_t) t * 12);
#pragma unroll
        for (int i = 0; i < 6; ++i) { const double2 v = bp[i]; b[2 * i] = v.x; b[2 * i + 1] = v.y; }
    }
    const bool hasw = t < 4 * W.nwells;
    double dinv[4] = {0.0, 0.0, 0.0, 0.0};
    unsigned wb = 0, we = 0;
    if (hasw) {
        const int w = t >> 2;
        wb = W.wptr[w]; we = W.wptr[w + 1];
        const double2* dp = reinterpret_cast<const double2*>(W.Dinv + (size_t) t * 4);
        const double2 v0 = dp[0], v1 = dp[1];
        dinv[0] = v0.x; dinv[1] = v0.y; dinv[2] = v1.x; dinv[3] = v1.y;
    }
    int4 it[kItems];
    double ic[kItems][4], yold[kItems], dd1[kItems];
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const int q = t + k * kAll;
        it[k] = make_int4(0, 0, 0, 0);
        yold[k] = 0.0; dd1[k] = 0.0;
        if (q < W.nitems) {
            it[k] = W.item[q];
            const double2* cp = reinterpret_cast<const double2*>(W.itemC + (size_t) q * 4);
            const double2 v0 = cp[0], v1 = cp[1];
            ic[k][0] = v0.x; ic[k][1] = v0.y; ic[k][2] = v1.x; ic[k][3] = v1.y;
        }
    }
    const int done = MODE != 0 ? S->done : 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k)
        if (t + k * kAll < W.nitems) {
            yold[k] = y[it[k].x];
            if (MODE != 0) dd1[k] = d1[it[k].x];
        }
    if (done) return;                      // (every thread of every CTA alike: no cluster barrier has been entered yet)
    // ---- phase 1
    if (hasp) {
        const double* xp = x + 3 * (size_t) col;
        const double x0 = xp[0], x1 = xp[1], x2 = xp[2];
#pragma unroll
        for (int r = 0; r < 4; ++r) part[tl * 4 + r] = b[3 * r] * x0 + b[3 * r + 1] * x1 + b[3 * r + 2] * x2;
    }
    cluster_sync_all();
    if (hasw) {
        const int r = t & 3;
        double z = 0.0;
        for (unsigned p0 = wb; p0 < we; p0 += 8) {
            double v[8];
#pragma unroll
            for (unsigned k = 0; k < 8; ++k) {
                const unsigned p = p0 + k;
                v[k] = p < we ? ld_dsmem_f64(part + (p & (kWellClThreads - 1)) * 4 + r, p / kWellClThreads) : 0.0;
            }
#pragma unroll
            for (unsigned k = 0; k < 8; ++k)
                if (p0 + k < we) z += v[k];
        }
        z1s[tl] = z;
    }
    __syncthreads();                       // the four rows of a well sit in one CTA (128 is a multiple of 4)
    if (hasw) {
        const double* zz = z1s + (tl & ~3);
        z2s[tl] = dinv[0] * zz[0] + dinv[1] * zz[1] + dinv[2] * zz[2] + dinv[3] * zz[3];
    }
    cluster_sync_all();
    // ---- phase 2
    double acc[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        if (t + k * kAll >= W.nitems) continue;
        const int zi = it[k].w;                                  // 4 * well: its four rows are thread rows zi .. zi + 3 of the cluster
        const unsigned owner = (unsigned) zi / kWellClThreads;
        const double* zl = z2s + (zi & (kWellClThreads - 1));
        const double z0 = ld_dsmem_f64(zl, owner), z1 = ld_dsmem_f64(zl + 1, owner), z2v = ld_dsmem_f64(zl + 2, owner), z3 = ld_dsmem_f64(zl + 3, owner);
        double delta = ic[k][0] * z0 + ic[k][1] * z1 + ic[k][2] * z2v + ic[k][3] * z3;
        if (it[k].y > 1) {
            const int c = it[k].x % 3;
            for (int e = it[k].z + 1, ee = it[k].z + it[k].y; e < ee; ++e) {
                const double* cb = W.C + (size_t) W.ublock[e] * 12 + c;
                const int zj = 4 * W.uwell[e];
                const unsigned ow = (unsigned) zj / kWellClThreads;
                const double* zq = z2s + (zj & (kWellClThreads - 1));
                delta += cb[0] * ld_dsmem_f64(zq, ow) + cb[3] * ld_dsmem_f64(zq + 1, ow) + cb[6] * ld_dsmem_f64(zq + 2, ow) + cb[9] * ld_dsmem_f64(zq + 3, ow);
            }
        }
        const double now = yold[k] - delta;
        y[it[k].x] = now;
        if (MODE == 1) acc[0] += dd1[k] * (now - yold[k]);
        if (MODE == 2) { acc[0] += dd1[k] * (now - yold[k]); acc[1] += now * now - yold[k] * yold[k]; }
    }
    if (MODE != 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const double v = warp_sum(acc[i]);
            if (lane == 0) sm[i][warp] = v;
        }
    }
    cluster_sync_all();                    // every CTA's partial sums are in its shared memory
    if (MODE != 0 && rank == 0 && tl == 0) {
        double a = 0.0, bsum = 0.0;
        for (unsigned q = 0; q < (unsigned) kWellClCtas; ++q)
            for (int w = 0; w < kWellClThreads / 32; ++w) { a += ld_dsmem_f64(&sm[0][w], q); bsum += ld_dsmem_f64(&sm[1][w], q); }
        if (MODE == 1) S->h += a;
        if (MODE == 2) { S->tr += a; S->tt += bsum; }
    }
    cluster_sync_all();                    // no CTA leaves while its shared memory may still be read
}

// ---- multisegment wells ---------------------------------------------------------------------------
//
// y -= C^T (D^-1 (B x)) for multisegment wells (bda/MultisegmentWellContribution.cpp:70-110).  The reference copies x and y
// to the host in every operator apply, runs an UMFPACK solve per well and copies y back (bda/WellContributions.cu:167-187).
// Here D^-1 is formed ONCE per solve on the host (where the reference runs umfpack_di_numeric) as a dense M x M matrix
// (M = 4 segments), so the apply is two small launches on the solver's stream and the vectors never leave the device.
//
// k_mswell_z: one CTA per well.  z1 = B x (a thread per well equation of a segment), then z2 = D^-1 z1 (a warp per row of
// D^-1, lanes stride the row: coalesced, fixed summation order).
struct MsWellsD {
    int nwells;
    const int* zoff;          // [nwells + 1] first entry of a well in z1 / z2 (4 per segment)
    const int* rowoff;        // [nwells + 1] first segment (block row) of a well
    const long long* doff;    // [nwells] first entry of a well's dense D^-1 (row-major M x M)
    const int* rowptr;        // [segments + 1] blocks of a segment
    const int* bcol;          // [blocks] perforated cell (p-space)
    const double* B;          // [blocks][4][3]
    const double* C;          // [blocks][4][3]
    const double* Dinv;
    double* z1;
    double* z2;
    // gather lists of phase 2: unique perforated cells, their blocks and the z2 offset of each block's segment
    int ncells;
    const int* ucell;
    const int* uptr;
    const int* ublock;
    const int* uz;
};

__global__ void __launch_bounds__(256) k_mswell_z(MsWellsD W, const double* __restrict__ x, const Scalars* S, int check_done)
{
    pdl_enter();
    if (check_done && S->done) return;
    const int w = blockIdx.x;
    const int z0 = W.zoff[w], M = W.zoff[w + 1] - z0, r0 = W.rowoff[w];
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        const int row = r0 + (i >> 2), j = i & 3;
        double acc = 0.0;
        for (int blk = W.rowptr[row], be = W.rowptr[row + 1]; blk < be; ++blk) {
            const double* bb = W.B + (size_t) blk * 12 + j * 3;
            const double* xx = x + 3 * (size_t) W.bcol[blk];
            acc += bb[0] * xx[0] + bb[1] * xx[1] + bb[2] * xx[2];
        }
        W.z1[z0 + i] = acc;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const double* D = W.Dinv + W.doff[w];
    const double* z1 = W.z1 + z0;
    for (int i = warp; i < M; i += nwarp) {
        const double* dr = D + (size_t) i * M;
        double acc = 0.0;
        for (int k = lane; k < M; k += 32) acc += dr[k] * z1[k];
        acc = warp_sum(acc);
        if (lane == 0) W.z2[z0 + i] = acc;
    }
}

// k_mswell_y: one CTA; a thread per (unique perforated cell, component) gathers every contribution to that cell (no
// atomics, deterministic; wells may share a cell) and patches the dot products the SpMV epilogue took before the wells
// were applied, exactly as phase 2 of k_wells.
template <int MODE>
__global__ void __launch_bounds__(1024) k_mswell_y(MsWellsD W, double* y, const double* __restrict__ d1, Scalars* S)
{
    pdl_enter();
    if (MODE != 0 && S->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    double acc[2] = {0.0, 0.0};
    for (int t = threadIdx.x; t < 3 * W.ncells; t += blockDim.x) {
        const int u = t / 3, c = t - 3 * u;
        double delta = 0.0;
        for (int e = W.uptr[u], ee = W.uptr[u + 1]; e < ee; ++e) {
            const double* cb = W.C + (size_t) W.ublock[e] * 12 + c;
            const double* zz = W.z2 + W.uz[e];
            delta += cb[0] * zz[0] + cb[3] * zz[1] + cb[6] * zz[2] + cb[9] * zz[3];
        }
        const size_t idx = 3 * (size_t) W.ucell[u] + c;
        const double old = y[idx], now = old - delta;
        y[idx] = now;
        const double dd1 = MODE != 0 ? d1[idx] : 0.0;
        if (MODE == 1) acc[0] += dd1 * (now - old);
        if (MODE == 2) { acc[0] += dd1 * (now - old); acc[1] += now * now - old * old; }
    }
    if (MODE != 0) {
        __shared__ double sm[2][32];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            double v = warp_sum(acc[i]);
            if (lane == 0) sm[i][warp] = v;
        }
        __syncthreads();
        if (warp == 0) {
            double a = lane < nwarp ? sm[0][lane] : 0.0, b = lane < nwarp ? sm[1][lane] : 0.0;
            a = warp_sum(a); b = warp_sum(b);
            if (lane == 0) {
                if (MODE == 1) S->h += a;
                if (MODE == 2) { S->tr += a; S->tt += b; }
            }
        }
    }
}

// ---- multi-GPU: halo exchange over NVLink peer memory -----------------------------------------------
//
// Row-slab partition, ghosts numbered last (ISTLSolverEbos.hpp:171-180, findOverlapRowsAndColumns.hpp:119-139).
// Every rank owns a receive block [flags | 2 x 3 n_ghost doubles] that its neighbours map through CUDA IPC.
// k_halo_push writes the boundary entries of the SpMV input straight into the neighbours' receive blocks
// (st over NVLink, no staging, no NCCL) and then raises the neighbour's flag to the epoch of this
// exchange; k_spmv_ghost, which applies the few owned x ghost blocks after the big owned x owned SpMV
// has run, is the only kernel that waits for the flags -- the transfer hides behind the owned SpMV.
// Reference semantics: WellModelGhostLastMatrixAdapter::apply multiplies the interior rows with the
// full (owned + ghost) vector (WellOperators.hpp:200-214) after the communicator's copyOwnerToAll.
struct HaloPeerD {
    double* recv;                // neighbour's receive block: first double of MY section, parity 0
    unsigned* flag;              // neighbour's flag slot for me
    long long parity_stride;     // doubles between the neighbour's two parity buffers
    int send_begin, send_end;    // my entries of send_prow
};


// grid (blocks per neighbour, neighbours).  tickets: one counter per neighbour (self-resetting).
__global__ void __launch_bounds__(256) k_halo_push(const HaloPeerD* __restrict__ peers, const int* __restrict__ send_prow,
                                                   const double* __restrict__ y, const unsigned* epoch_ctr, unsigned* tickets, Scalars* S,
                                                   int check_done)
{
    if (check_done && S->done) return;
    // The exchange's epoch lives on the device (one past the last COMPLETED exchange; k_spmv_ghost, which consumes the exchange,
    // bumps it): no kernel argument changes from launch to launch, so the iteration can be replayed from a CUDA graph on several
    // GPUs too.  Every rank runs the same launch sequence and takes the same `done` decisions: the counters stay in step.
    const unsigned epoch = *epoch_ctr + 1u;
    const HaloPeerD P = peers[blockIdx.y];
    double* dst = P.recv + (epoch & 1u) * P.parity_stride;
    const int n3 = 3 * (P.send_end - P.send_begin);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += gridDim.x * blockDim.x) {
        const int e = i / 3, c = i - 3 * e;
        dst[i] = y[3 * (size_t) send_prow[P.send_begin + e] + c];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(tickets + blockIdx.y, 1u) == gridDim.x - 1) {
            tickets[blockIdx.y] = 0;
            __threadfence_system();
            st_release_sys(P.flag, epoch);
        }
    }
}

// y[row] += sum over the ghost blocks of the row of A_og x_ghost, for the boundary rows only; patches the dot
// products the owned SpMV took (MODE as k_spmv).  3 lanes per boundary row.  Block values are read from the
// staging copy of the caller's array (gsrc = block index there).
template <int MODE>
__global__ void __launch_bounds__(kVecThreads) k_spmv_ghost(int nrows, const int* __restrict__ grow, const int* __restrict__ gptr,
                                                            const int* __restrict__ gcol, const int* __restrict__ gsrc,
                                                            const double* __restrict__ stage, const double* __restrict__ ghost_x0,
                                                            long long ghost_parity_stride, const unsigned* flags, int nneigh,
                                                            unsigned* epoch_ctr, unsigned* bump_ticket, double* y,
                                                            const double* __restrict__ d1, Scalars* S, double* partials,
                                                            unsigned* ticket, int check_done, const DistRedD D)
{
    if (check_done && S->done) return;
    const unsigned epoch = *epoch_ctr + 1u;                       // the exchange k_halo_push has just started (see there)
    const double* __restrict__ ghost_x = ghost_x0 + (size_t) (epoch & 1u) * ghost_parity_stride;
    if (threadIdx.x < nneigh) {
        long long spins = 0, t0 = 0;
        while ((int) (ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) {
            if ((++spins & 1023) == 0) {
                if (t0 == 0) t0 = wall_ns();
                else if (wall_ns() - t0 > kWaitTimeoutNs || *((volatile int*) &S->trsv_timeout)) { S->trsv_timeout = 1; break; }
            }
        }
    }
    __syncthreads();
    double acc[2] = {0.0, 0.0};
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < 3 * nrows; t += gridDim.x * blockDim.x) {
        const int u = t / 3, c = t - 3 * u;
        double delta = 0.0;
        for (int e = gptr[u]; e < gptr[u + 1]; ++e) {
            const double* a = stage + (size_t) gsrc[e] * 9 + c * 3;
            const double* xx = ghost_x + 3 * (size_t) gcol[e];
            delta += a[0] * __ldcg(xx) + a[1] * __ldcg(xx + 1) + a[2] * __ldcg(xx + 2);
        }
        const size_t idx = 3 * (size_t) grow[u] + c;
        const double old = y[idx], now = old + delta;
        y[idx] = now;
        if (MODE == 1) acc[0] += d1[idx] * delta;
        if (MODE == 2) { acc[0] += d1[idx] * delta; acc[1] += now * now - old * old; }
    }
    if (MODE != 0) {
        double tot[2];
        if (grid_reduce<2>(acc, partials, ticket, tot)) {
            *epoch_ctr = epoch;                                    // last block: the exchange is consumed
            if (MODE == 1) tot[0] += S->h;
            if (MODE == 2) { tot[0] += S->tr; tot[1] += S->tt; }
            if (D.mail) {                                          // this launch completes the local sums: all-reduce them here
                if (MODE == 1) { double one[1] = {tot[0]}; mail_allreduce<1>(D, S, one); tot[0] = one[0]; }
                else mail_allreduce<2>(D, S, tot);
            }
            if (MODE == 1) S->h = tot[0];
            if (MODE == 2) { S->tr = tot[0]; S->tt = tot[1]; }
        }
    } else {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(bump_ticket, 1u) == gridDim.x - 1) { *bump_ticket = 0; *epoch_ctr = epoch; }
        }
    }
}

// ---- multi-GPU: all-reduce of the Krylov scalars over peer memory ---------------------------------------
//
// Layout of every rank's IPC-exported block: [0, 256) halo flags | [256, 768) mail flags u32[2][64] |
// [768, 4864) mail values double[2][64][4] | [kHaloRecvOffset, ...) halo receive buffers.
// One tiny kernel per reduction (`world` threads): thread r stores this rank's partial sums into rank r's mailbox
// (slot = parity of the sequence number, row = my rank) and raises r's flag to the sequence number, then waits for
// rank r's contribution in its own mailbox; thread 0 adds the `world` contributions in rank order -- every rank gets
// the same bits -- and runs the scalar epilogue of the phase.  One launch and one NVLink round trip (~5 us) against an
// NCCL all-reduce kernel plus a finish kernel (~20 us); nothing else travels between the GPUs per iteration.
// PHASE 0: after k_init (red[0]); 1: after the first SpMV (h); 2: after k_vec_xr1 (red[0]); 3: after the second SpMV
// (tr, tt); 4: after k_vec_xr2 (red[0..1]); 5: singular flag of the factorisation (max).
constexpr int kHaloRecvOffset = 4864;

template <int PHASE>
__global__ void __launch_bounds__(64) k_allreduce_p2p(const MailD M, int rank, int world, unsigned* seq_ctr, Scalars* S, double tol, int max_half)
{
    __shared__ double got[64][4];
    __shared__ unsigned s_seq;
    const int r = threadIdx.x;
    // the sequence number lives on the device (no kernel argument changes between launches: graph replay); it advances with
    // every exchange, here or in a producer's epilogue (mail_allreduce), on every rank alike
    if (PHASE != 0 && PHASE != 5 && S->done) return;         // (every rank takes this branch alike: done comes from reduced sums)
    if (r == 0) { s_seq = *seq_ctr + 1u; *seq_ctr = s_seq; }
    __syncthreads();
    const unsigned seq = s_seq;
    const int par = seq & 1u;
    if (r < world) {
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        if (PHASE == 0 || PHASE == 2) v[0] = S->red[0];
        if (PHASE == 1) v[0] = S->h;
        if (PHASE == 3) { v[0] = S->tr; v[1] = S->tt; }
        if (PHASE == 4) { v[0] = S->red[0]; v[1] = S->red[1]; }
        if (PHASE == 5) v[0] = (double) S->singular;
        double* dst = M.vals[r] + ((size_t) par * 64 + rank) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) dst[k] = v[k];
        __threadfence_system();
        st_release_sys(M.flags[r] + par * 64 + rank, seq);
        const unsigned* mine = M.flags[rank] + par * 64 + r;
        long long spins = 0, t0 = 0;
        while (ld_acquire_sys(mine) != seq) {
            if ((++spins & 1023) == 0) {
                if (t0 == 0) t0 = wall_ns();
                else if (wall_ns() - t0 > kWaitTimeoutNs) { S->trsv_timeout = 1; break; }
            }
        }
        const double* src = M.vals[rank] + ((size_t) par * 64 + r) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) got[r][k] = __ldcg(src + k);
    }
    __syncthreads();
    if (r != 0) return;
    double t[4] = {0.0, 0.0, 0.0, 0.0};
    for (int q = 0; q < world; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) t[k] += got[q][k];
    if (PHASE == 0) finish_init(S, t[0], tol, max_half);
    if (PHASE == 1) S->h = t[0];
    if (PHASE == 2) finish_xr1(S, t[0]);
    if (PHASE == 3) { S->tr = t[0]; S->tt = t[1]; }
    if (PHASE == 4) finish_xr2(S, t[0], t[1]);
    if (PHASE == 5) S->singular = t[0] > 0.0 ? 1 : 0;
}

// write-only sweep over a buffer larger than L2 (timing hygiene between measured launches)
__global__ void __launch_bounds__(256) k_flush_l2(double* __restrict__ buf, long long n, double v)
{
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) buf[i] = v;
}

}  // namespace b200
