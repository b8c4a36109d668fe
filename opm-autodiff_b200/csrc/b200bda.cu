// b200bda.cu -- solver object + C ABI (include/b200bda.h) of the B200-native ILU0-BiCGSTAB backend.
//
// State machine of b200_solve_system == cusparseSolverBackend<3>::solve_system
// (bda/cusparseSolverBackend.cu:480-499): first call initialize + copy_system_to_gpu + analyse_matrix,
// later calls update_system_on_gpu (values and rhs only); every call: new ILU0, BiCGSTAB, sync.
// There is no CPU fallback: without a CUDA device every computing entry point fails.
#include "../../include/b200bda.h"

#include <dlfcn.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "analysis.hpp"
#include "sweep2.hpp"
#include "coloring.hpp"
#include "kernels.cuh"

namespace b200 {

static thread_local std::string g_last_error;

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define CUDA_OK(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            throw CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(e__) + " (" +   \
                            __FILE__ + ":" + std::to_string(__LINE__) + ")");                      \
    } while (0)

static double wall()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    void alloc(size_t count)
    {
        if (count <= n && p) return;
        release();
        CUDA_OK(cudaMalloc((void**) &p, std::max<size_t>(count, 1) * sizeof(T)));
        n = count;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr; n = 0;
    }
    ~DevBuf() { release(); }
};

enum Kind { K_PERMUTE, K_INIT, K_FACTOR, K_SLICES, K_LOWER, K_UPPER, K_SPMV, K_WELL, K_VEC_P, K_VEC_XR1, K_VEC_XR2,
            K_UNPERMUTE, K_MISC, K_HALO_PUSH, K_SPMV_GHOST, K_ALLREDUCE, K_FINISH, K_UPPER_SPMV, K_COUNT };
static const char* kKindNames[K_COUNT] = {"permute", "init", "ilu_factor", "ilu_stream", "ilu_lower", "ilu_upper", "spmv",
                                          "well_apply", "vec_p", "vec_xr1", "vec_xr2", "unpermute", "misc",
                                          "halo_push", "spmv_ghost", "allreduce", "finish", "ilu_upper_spmv"};

// ---- NCCL, bound at run time ------------------------------------------------------------------------
// Only the multi-GPU entry points need NCCL, and a Python host already has torch's libnccl.so.2 mapped:
// dlopen by soname binds to that copy (one NCCL per process) or, in a plain C++ host, to the system one.
struct NcclId { char internal[128]; };
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    void load()
    {
        if (handle) return;
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) throw std::runtime_error(std::string("cannot load libnccl.so.2: ") + dlerror());
        auto sym = [&](const char* n) {
            void* f = dlsym(handle, n);
            if (!f) throw std::runtime_error(std::string("libnccl.so.2 lacks ") + n);
            return f;
        };
        GetUniqueId = (int (*)(NcclId*)) sym("ncclGetUniqueId");
        CommInitRank = (int (*)(void**, int, NcclId, int)) sym("ncclCommInitRank");
        CommDestroy = (int (*)(void*)) sym("ncclCommDestroy");
        AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)) sym("ncclAllReduce");
        GetErrorString = (const char* (*)(int)) sym("ncclGetErrorString");
    }
};
static NcclApi g_nccl;
constexpr int kNcclInt32 = 2, kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2;
#define NCCL_OK(expr)                                                                              \
    do {                                                                                           \
        int r__ = (expr);                                                                          \
        if (r__ != 0)                                                                              \
            throw CudaError(std::string(#expr) + " failed: " + g_nccl.GetErrorString(r__));       \
    } while (0)

// Multi-GPU state of one rank (row slab with ghosts numbered last).
struct Dist {
    bool enabled = false;
    int rank = 0, world = 1;
    void* comm = nullptr;
    // halo plan (host)
    bool have_halo = false;
    int n_ghost = 0, nneigh = 0;
    std::vector<int> neigh_rank, send_ptr, send_rows, recv_ptr;
    // owned x ghost coupling (built by analyse): boundary rows in p-space
    int gnrows = 0;
    long long gnblocks = 0;
    // peers
    std::vector<void*> rank_base;           // IPC-mapped block of every rank (own block: the local pointer)
    MailD mail{};                           // mail flags / values of every rank
    bool mail_ready = false;
    bool use_p2p_allreduce = true;
    std::vector<HaloPeerD> peers;
    bool peers_ready = false;
};

struct KStat {
    long long launches = 0;
    double ms = 0.0;
};

// Host-side well container: mirrors the three-phase fill of Opm::WellContributions
// (bda/WellContributions.cpp:152-259).  Device copies live in the solver (persistent across solves).
struct Wells {
    bool allocated = false;
    unsigned dim = 0, dim_wells = 0, num_blocks = 0, num_std_wells = 0;
    unsigned num_blocks_so_far = 0, num_std_wells_so_far = 0;
    std::vector<unsigned> val_pointers;
    std::vector<int> Ccols, Bcols;
    std::vector<double> Cnnzs, Dnnzs, Bnnzs;
    // multisegment wells (bda/WellContributions.hpp:85,92): one entry per addMultisegmentWellContribution
    struct MsWell {
        unsigned Mb = 0;                       // segments = block rows of B, C and D
        std::vector<unsigned> Brows, Bcols;    // blocked CSR pattern shared by B and C
        std::vector<double> B, C;              // 4x3 blocks, row-major [well eq][cell eq]
        std::vector<double> Dinv;              // dense (4 Mb) x (4 Mb), row-major
    };
    std::vector<MsWell> ms;
};

// Inverse of the scalar CSC matrix D of a multisegment well as a dense row-major M x M matrix, on the host: LU with partial
// pivoting that skips structural zeros through per-row column extents (D is a block tree, mostly a chain of segments: the
// factors stay banded, O(M bw^2)), then one forward / backward substitution per column of the identity (O(M^2 bw)).  The
// reference factorises D with UMFPACK at the same point (MultisegmentWellContribution.cpp:56-57) and solves on the host in
// every operator apply; with the explicit inverse the apply is a dense mat-vec on the device.
static void invert_csc(int M, const int* colptr, const int* rowidx, const double* vals, std::vector<double>& inv)
{
    std::vector<double> A((size_t) M * M, 0.0);
    std::vector<int> lo(M, M), hi(M, -1), perm(M);
    for (int c = 0; c < M; ++c) {
        if (colptr[c + 1] < colptr[c]) throw std::runtime_error("multisegment well: D column pointers must ascend");
        for (int q = colptr[c]; q < colptr[c + 1]; ++q) {
            const int r = rowidx[q];
            if (r < 0 || r >= M) throw std::runtime_error("multisegment well: D row index out of range");
            A[(size_t) r * M + c] += vals[q];
            lo[r] = std::min(lo[r], c); hi[r] = std::max(hi[r], c);
        }
    }
    for (int i = 0; i < M; ++i) perm[i] = i;
    for (int k = 0; k < M; ++k) {
        int pr = -1;
        double best = 0.0;
        for (int i = k; i < M; ++i) {
            if (lo[i] > k) continue;
            const double a = std::fabs(A[(size_t) i * M + k]);
            if (a > best) { best = a; pr = i; }
        }
        if (pr < 0) throw std::runtime_error("multisegment well: matrix D is singular");
        if (pr != k) {
            std::swap_ranges(A.begin() + (size_t) k * M, A.begin() + (size_t) (k + 1) * M, A.begin() + (size_t) pr * M);
            std::swap(lo[k], lo[pr]); std::swap(hi[k], hi[pr]); std::swap(perm[k], perm[pr]);
        }
        const double* ak = &A[(size_t) k * M];
        const double piv = 1.0 / ak[k];
        const int hk = hi[k];
        for (int i = k + 1; i < M; ++i) {
            if (lo[i] > k) continue;
            double* ai = &A[(size_t) i * M];
            if (ai[k] == 0.0) continue;
            const double f = ai[k] * piv;
            ai[k] = f;
            for (int c = k + 1; c <= hk; ++c) ai[c] -= f * ak[c];
            hi[i] = std::max(hi[i], hk);
        }
    }
    // columns of the inverse: L U x = P e_j
    inv.assign((size_t) M * M, 0.0);
    std::vector<double> y(M);
    for (int q = 0; q < M; ++q) {
        const int j = perm[q];                 // (P e_j) is the unit vector e_q
        for (int i = 0; i < q; ++i) y[i] = 0.0;
        y[q] = 1.0;
        for (int i = q + 1; i < M; ++i) {
            const double* ai = &A[(size_t) i * M];
            double sacc = 0.0;
            for (int c = std::max(lo[i], q); c < i; ++c) sacc += ai[c] * y[c];
            y[i] = -sacc;
        }
        for (int i = M - 1; i >= 0; --i) {
            const double* ai = &A[(size_t) i * M];
            double sacc = y[i];
            for (int c = i + 1; c <= hi[i]; ++c) sacc -= ai[c] * y[c];
            y[i] = sacc / ai[i];
        }
        for (int i = 0; i < M; ++i) inv[(size_t) i * M + j] = y[i];
    }
}

struct Solver {
    int verbosity = 0, maxit = 200, device = 0;
    double tolerance = 1e-2, relaxation = 1.0;
    bool pin_host = false, use_graph = true, profile = false;
    int lookahead = 2;
    // triangular sweeps: parts (0 = one per SM), consumer warps per CTA, ring slots, bytes per stage, window rows
    int sweep_parts = 0, sweep_warps = 8, sweep_groups = 1, sweep_helpers = 2, sweep_slots = 2, sweep_stage_bytes = 0, sweep_window = 0, sweep_ext_window = 0, sweep_helper_sleep = 0;
    // round-2 sweeps (k_sweep2): consumer warps (G x WG per part), helper warps, forced group count / group width (0 = automatic)
    int sweep_v2 = 1, s2_cw = 15, s2_helpers = 1, s2_poll_lead = 15, s2_prefetch = 2;      // s2_helpers: warps beyond the consumers (they only work in the tails)
    bool v2 = false, s2_mlL = false, s2_mlU = false;      // s2_ml*: the sweep's schedule has rows that take several lanes
    Sweep2Plan L2, U2;

    cudaStream_t stream = nullptr;
    cudaStream_t stream_aux = nullptr;         // multi-GPU: the halo push runs beside the well apply (fork / join by events)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int num_sms = 0;
    int vec_blocks = 0;
    int spmv_blocks_cap = kMaxPartials;
    size_t smem_optin = 0, sweep_smem = 0;
    int sweep_metaCap = 0, sweep_valsCap = 0, sweep_rhsCap = 0, sweep_extCap = 0;
    bool sweep_trace = false;
    static constexpr int kTraceCap = 1024;
    DevBuf<long long> d_trace;

    bool analysed = false, have_system = false, have_factor = false;
    bool poisoned = false;                     // a dataflow wait timed out: the solver refuses further solves
    int N = 0, Nb = 0;
    long long nnz = 0, nnzb = 0;              // owned x owned part (what ILU0 and the big SpMV see)
    long long nnz_stage = 0;                   // the caller's array (owned x (owned + ghost))
    Analysis an;

    Dist dist;
    DevBuf<unsigned char> d_halo;              // [flags: 64 x u32 | 2 x 3 n_ghost doubles], IPC-exported
    DevBuf<HaloPeerD> d_peers;
    DevBuf<int> d_send_prow, d_grow, d_gptr, d_gcol, d_gsrc;
    DevBuf<unsigned> d_push_tickets;
    DevBuf<MailD> d_mail;                      // device copy of dist.mail for the all-reduce run inside producer kernels
    DevBuf<unsigned> d_dist_ctr;               // device-side counters: [0] all-reduce sequence number, [1] halo epoch, [2] ticket

    DevBuf<int> d_prow, d_pcol, d_pdiag, d_srcblk, d_perm, d_flevRows, d_facPtr, d_facOps;
    DevBuf<int> d_sellPtr, d_sellOver, d_sellCol, d_sellSrc;     // sliced-ELL copy of A for the SpMV (analysis.hpp SellPlan)
    DevBuf<double> d_sellVal;
    int sell_slices = 0;
    long long sell_slots = 0;
    // Size-dependent features: option value 0 = off, 1 = automatic (on from kBigRows block rows: below that their fixed costs
    // -- a second gather per solve, tail barriers and tickets, extra global reads per stage -- outweigh what they save; the
    // Norne-size system has 44 k rows and an L2-resident matrix), 2 = on.
    static constexpr int kBigRows = 100000;
    bool feature_on(int opt) const { return opt == 2 || (opt == 1 && Nb >= kBigRows); }
    int spmv_sell = 1;                 // option: SpMV from the sliced-ELL copy (else BSR kernel, 3 lanes per row)
    // SpMV run by the upper sweep's CTAs as their parts finish (kernels.cuh fused_spmv_tail)
    DevBuf<int> d_fUnits, d_fNeedPtr, d_fNeed, d_fSync;
    DevBuf<double> d_fPartials;
    int fused_units = 0;
    int fuse_spmv = 1;                 // option
    bool fac_plan = false;
    // one-launch factorisation (k_ilu_factor_flow): a record per row in level order, a ready flag per row
    DevBuf<int> d_facRec;              // k_ilu_factor_plan3: a record per row, level order, levels padded to whole triples
    std::vector<int> fac3_ptr;         // first triple of every level (+ end)
    int fac_warps = 4;                 // option: warps per CTA of k_ilu_factor_plan3 (1..16), read at upload
    int fac_rows3 = 1;                 // option: 1 = three rows per warp (k_ilu_factor_plan3), 0 = one (k_ilu_factor_plan)
    size_t fac3_smem = 0;
    bool fac3_ready() const { return fac_plan && fac_rows3 && !fac3_ptr.empty(); }
    cudaGraphExec_t fac_graph_exec = nullptr;
    DevBuf<StageD> d_stagesL, d_stagesU;
    DevBuf<PartD> d_partsL, d_partsU;
    DevBuf<BuildD> d_buildL, d_buildU;
    DevBuf<int> d_metaL, d_metaU, d_srcL, d_srcU;
    DevBuf<double> d_valL, d_valU;
    DevBuf<S2PartD> d_s2partsL, d_s2partsU;
    DevBuf<S2StreamD> d_s2streamsL, d_s2streamsU;
    DevBuf<S2BuildD> d_s2buildL, d_s2buildU;
    DevBuf<int> d_s2hdrsL, d_s2hdrsU, d_s2codesL, d_s2codesU, d_s2srcL, d_s2srcU;
    DevBuf<double> d_stage, d_bstage, d_A, d_LU;
    DevBuf<double> d_x, d_r, d_rt, d_p, d_v, d_t, d_y, d_w, d_xnat, d_tmp1, d_tmp2;
    DevBuf<double> d_partials;
    DevBuf<unsigned> d_ticket;
    DevBuf<Scalars> d_S;
    Scalars* h_S = nullptr;       // pinned
    DevBuf<double> d_flush;

    // wells (device, p-space columns)
    int nwells = 0, nwblocks = 0, nucells = 0;
    DevBuf<unsigned> d_wptr;
    DevBuf<int> d_Bcols, d_ucell, d_uptr, d_ublock, d_uwell;
    DevBuf<double> d_B, d_C, d_Dinv, d_z2;
    // standard wells, index chains resolved on the host (k_wells_flat): used when the wells fit one CTA
    int wells_flat = 1;                // option: 0 never, 1 when they fit
    int wells_cluster = 1;             // option: the flat apply on a cluster of 8 CTAs (k_wells_cluster) instead of one CTA
    bool flat_ok = false;
    DevBuf<int4> d_item;
    DevBuf<double> d_itemC;
    WellsFlatD flatD{};
    // host copy of the last uploaded well structure (see upload_wells)
    bool wells_structure_valid = false;
    std::vector<unsigned> h_wptr;
    std::vector<int> h_wBcols, h_wCcols, h_ucell, h_uptr, h_ublock, h_uwell;
    // multisegment wells (device, p-space columns); ms_epoch changes whenever they are uploaded (graph signature)
    int nms = 0, ms_blocks = 0, ms_rows = 0, ms_ncells = 0, ms_epoch = 0;
    long long ms_dinv_entries = 0;
    DevBuf<int> d_msZoff, d_msRowoff, d_msRowptr, d_msBcol, d_msUcell, d_msUptr, d_msUblock, d_msUz;
    DevBuf<long long> d_msDoff;
    DevBuf<double> d_msB, d_msC, d_msDinv, d_msZ1, d_msZ2;
    MsWellsD msD{};

    // pinned-host registration of the caller's arrays
    const void* reg_vals = nullptr; size_t reg_vals_bytes = 0;
    const void* reg_b = nullptr; size_t reg_b_bytes = 0;
    const void* reg_x = nullptr; size_t reg_x_bytes = 0;      // the caller's solution vector (get_result)
    const void* cand_vals = nullptr; const void* cand_b = nullptr; const void* cand_x = nullptr;   // seen once, not yet pinned
    std::vector<std::pair<void*, size_t>> host_regs;          // b200_host_register: page-locked on the caller's explicit request

    KStat stats[K_COUNT];
    long long launch_count = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
    std::vector<std::pair<int, int>> ev_used;   // (kind, pool index)
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;

    ~Solver()
    {
        for (size_t r = 0; r < dist.rank_base.size(); ++r) if (dist.rank_base[r] && (int) r != dist.rank) cudaIpcCloseMemHandle(dist.rank_base[r]);
        if (dist.comm) g_nccl.CommDestroy(dist.comm);
        if (fac_graph_exec) cudaGraphExecDestroy(fac_graph_exec);
        if (iter_graph_exec) cudaGraphExecDestroy(iter_graph_exec);
        for (auto& r : host_regs) cudaHostUnregister(r.first);
        if (reg_vals) cudaHostUnregister((void*) reg_vals);
        if (reg_x) cudaHostUnregister((void*) reg_x);
        if (reg_b) cudaHostUnregister((void*) reg_b);
        for (auto& e : ev_pool) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        for (cudaEvent_t e : {ev_a, ev_b, ev_c, ev_d, ev_t0, ev_t1}) if (e) cudaEventDestroy(e);
        if (h_S) cudaFreeHost(h_S);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (stream_aux) cudaStreamDestroy(stream_aux);
        if (stream) cudaStreamDestroy(stream);
    }

    // ---- launch bookkeeping ------------------------------------------------------------------
    bool counting = true;         // false while a launch sequence is being captured into a graph (counted per replay instead)
    int prof_begin(int kind)
    {
        if (!counting) return -1;
        stats[kind].launches++;
        launch_count++;
        if (!profile) return -1;
        if (ev_used.size() >= ev_pool.size()) {
            cudaEvent_t a, b;
            CUDA_OK(cudaEventCreate(&a)); CUDA_OK(cudaEventCreate(&b));
            ev_pool.emplace_back(a, b);
        }
        int idx = (int) ev_used.size();
        ev_used.emplace_back(kind, idx);
        CUDA_OK(cudaEventRecord(ev_pool[idx].first, stream));
        return idx;
    }
    void prof_end(int idx)
    {
        if (idx >= 0) CUDA_OK(cudaEventRecord(ev_pool[idx].second, stream));
    }
    void prof_collect()
    {
        if (!profile) return;
        CUDA_OK(cudaStreamSynchronize(stream));
        for (auto& u : ev_used) {
            float ms = 0.f;
            CUDA_OK(cudaEventElapsedTime(&ms, ev_pool[u.second].first, ev_pool[u.second].second));
            stats[u.first].ms += ms;
        }
        ev_used.clear();
    }

    double alg_bytes(int kind) const
    {
        const double nb = (double) Nb, nz = (double) nnzb;
        double nnzL = 0;
        if (analysed) nnzL = (double) ((nnzb - Nb) / 2);
        switch (kind) {
            case K_SPMV: return 76.0 * nz + 52.0 * nb;
            case K_LOWER: return 76.0 * nnzL + 52.0 * nb;
            case K_UPPER: return 76.0 * (nz - nnzL) + 52.0 * nb;
            case K_UPPER_SPMV: return 76.0 * (nz - nnzL) + 52.0 * nb + 76.0 * nz + 52.0 * nb;     // the sweep and the product it also runs
            case K_FACTOR: return (148.0 * nz + 8.0 * nb) / std::max(1, an.nflev);      // per level launch
            case K_SLICES: return 0.5 * (72.0 * nz + 76.0 * nz + 52.0 * nb);             // factor read, stream written (two fills)
            case K_VEC_P: return 96.0 * nb;
            case K_VEC_XR1: return (defer_now() ? 72.0 : 144.0) * nb;      // deferred x update: r, v read, r written
            case K_VEC_XR2: return (defer_now() ? 96.0 : 168.0) * nb;
            case K_WELL: return 272.0 * nwblocks + 128.0 * nwells + 272.0 * ms_blocks + 8.0 * (double) ms_dinv_entries;
            case K_PERMUTE: return (sell_slices ? 296.0 : 148.0) * nz;
            default: return 0.0;
        }
    }

    // ---- setup ---------------------------------------------------------------------------------
    void init_device()
    {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw CudaError(std::string("no CUDA device available (") + cudaGetErrorString(e) +
                            "); this backend has no CPU fallback");
        if (device >= ndev) throw CudaError("device id " + std::to_string(device) + " out of range");
        CUDA_OK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            throw CudaError(std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                            std::to_string(prop.minor) + "; this library is built for sm_100a only");
        num_sms = prop.multiProcessorCount;
        CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CUDA_OK(cudaStreamCreateWithFlags(&stream_aux, cudaStreamNonBlocking));
        CUDA_OK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        CUDA_OK(cudaMallocHost((void**) &h_S, sizeof(Scalars)));
        memset(h_S, 0, sizeof(Scalars));
        d_S.alloc(1);
        d_partials.alloc(3 * kMaxPartials);
        d_ticket.alloc(1);
        CUDA_OK(cudaMemsetAsync(d_ticket.p, 0, sizeof(unsigned), stream));
        CUDA_OK(cudaMemsetAsync(d_S.p, 0, sizeof(Scalars), stream));
        for (cudaEvent_t* e2 : {&ev_a, &ev_b, &ev_c, &ev_d}) CUDA_OK(cudaEventCreate(e2));
        smem_optin = prop.sharedMemPerBlockOptin;
        vec_blocks = std::min(num_sms * 8, kMaxPartials);
        spmv_blocks_cap = std::min(num_sms * 16, kMaxPartials);     // 16 CTAs of 256 threads per SM, two waves: 7 % faster than one block per 256 rows
        if (verbosity > 0)
            fprintf(stderr, "[b200bda] device %d: %s, %d SMs, %zu B smem/CTA, vec grid %d x %d\n", device, prop.name,
                    num_sms, smem_optin, vec_blocks, kVecThreads);
    }

    // Part count of the round-2 sweeps below the size where every SM gets a part: measured, not guessed.  The best count depends
    // on how the parts factor into strips x bands for the grid at hand (100 x 100 x 13: 55 us with 44 parts, 69 with 40, 65 with
    // 72), so the analysis builds the plan for a few counts around the rule's value and times one lower + one upper sweep of
    // each on the device (zero factor values: the dataflow, which is all that is timed, does not depend on them).  Once per
    // pattern; ~0.1-0.3 s per candidate at these sizes.  Different counts give the same factors and solves up to rounding (a
    // row adds its other-part dependencies first), so `sweep_autotune = 0` pins the rule's value for bit-reproducible runs
    // across processes.
    // Opt-in multi-colour ordering (coloring.hpp; the reference's --opencl-ilu-reorder=graph_coloring): ILU0 of the colour-permuted
    // matrix, level-synchronous sweeps (k_trsv_level, one launch per colour).  A different preconditioner: never the default.
    int reorder = 0;                   // option: 0 natural order (level scheduling), 1 graph colouring; set before the first solve
    unsigned reorder_seed = 1;         // option: seed of the colouring's random weights (the reference seeds from random_device)
    bool level_sweeps = false;         // the sweeps run level by level (set by the analysis with reorder = 1)
    b200::ColorOrder color;
    int sweep_autotune = 1;            // option
    int autotune_parts(const int* r_, const int* c_, AnalysisOptions opt, int p0)
    {
        std::vector<int> cand;
        for (double f : {1.0, 0.62, 0.72, 0.82, 0.91, 1.12, 1.3}) {
            const int p = std::max(8, std::min(num_sms, (int) std::lround(f * p0)));
            if (std::find(cand.begin(), cand.end(), p) == cand.end()) cand.push_back(p);
        }
        const int threads = sweep_threads();
        DevBuf<double> rhs, out;
        rhs.alloc(N + 8); out.alloc(N + 8);
        CUDA_OK(cudaMemsetAsync(rhs.p, 0, sizeof(double) * (N + 8), stream));
        double armed;
        { const unsigned long long bits = b200::kSentinel; std::memcpy(&armed, &bits, sizeof armed); }
        cudaEvent_t e0, e1;
        CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
        int best = p0;
        double best_us = 1e30, p0_us = 1e30;
        for (int p : cand) {
            opt.parts = p;
            const b200::Analysis a = b200::analyse(Nb, r_, c_, opt);
            Sweep2Options o2;
            o2.consumerWarps = s2_cw;
            Sweep2Plan L, U;
            build_sweep2_plans(a, r_, c_, o2, L, U);
            const size_t smem = kS2Header + 24 * (size_t) (a.window + 1);
            if (smem > smem_optin || a.nparts > num_sms) continue;
            struct Dev { DevBuf<S2PartD> parts; DevBuf<S2StreamD> streams; DevBuf<int> hdrs, codes; DevBuf<double> vals; } dl, du;
            auto upl = [&](Sweep2Plan& P2, Dev& d) {
                d.parts.alloc(P2.parts.size()); d.streams.alloc(P2.streams.size());
                d.hdrs.alloc(P2.hdrs.size() + 4 * 64); d.codes.alloc(P2.codes.size() + 4); d.vals.alloc((size_t) P2.nvals + 2);
                CUDA_OK(cudaMemcpyAsync(d.parts.p, P2.parts.data(), sizeof(S2Part) * P2.parts.size(), cudaMemcpyHostToDevice, stream));
                CUDA_OK(cudaMemcpyAsync(d.streams.p, P2.streams.data(), sizeof(S2Stream) * P2.streams.size(), cudaMemcpyHostToDevice, stream));
                CUDA_OK(cudaMemcpyAsync(d.hdrs.p, P2.hdrs.data(), sizeof(int) * P2.hdrs.size(), cudaMemcpyHostToDevice, stream));
                CUDA_OK(cudaMemcpyAsync(d.codes.p, P2.codes.data(), sizeof(int) * P2.codes.size(), cudaMemcpyHostToDevice, stream));
                CUDA_OK(cudaMemsetAsync(d.vals.p, 0, sizeof(double) * ((size_t) P2.nvals + 2), stream));
            };
            upl(L, dl); upl(U, du);
            auto args = [&](Dev& d, bool ml) {
                SweepArgs sa = {};
                sa.rhs = rhs.p; sa.out = out.p; sa.rearm = nullptr; sa.S = d_S.p;
                sa.nparts = a.nparts; sa.window = a.window; sa.nwarps = s2_cw; sa.helper_sleep = s2_poll_lead; sa.early = s2_prefetch;
                sa.v2.parts = d.parts.p; sa.v2.streams = d.streams.p; sa.v2.hdrs = reinterpret_cast<const int2*>(d.hdrs.p);
                sa.v2.codes = reinterpret_cast<const int4*>(d.codes.p); sa.v2.vals = d.vals.p; sa.v2.window = a.window; sa.v2.ncw = s2_cw;
                (void) ml;
                return sa;
            };
            const bool mlL = L.nMultiLaneRecords > 0, mlU = U.nMultiLaneRecords > 0;
            auto launch = [&](bool lower) {
                const SweepArgs sa = args(lower ? dl : du, lower ? mlL : mlU);
                k_fill<<<blocks_for(N, 256, num_sms * 8), 256, 0, stream>>>(out.p, armed, N);
                CUDA_OK(cudaEventRecord(e0, stream));
                auto go = [&](auto kern) {
                    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
                    kern<<<a.nparts, threads, smem, stream>>>(sa);
                };
                if (lower) { if (mlL) go(k_sweep2<true, false, -1, false, true>); else go(k_sweep2<true, false>); }
                else { if (mlU) go(k_sweep2<false, false, -1, false, true>); else go(k_sweep2<false, false>); }
                CUDA_OK(cudaEventRecord(e1, stream));
                CUDA_OK(cudaEventSynchronize(e1));
                CUDA_OK(cudaGetLastError());
                float ms = 0.f;
                CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
                return 1e3 * (double) ms;
            };
            double us = 1e30;
            for (int rep = 0; rep < 4; ++rep) {
                const double t = launch(true) + launch(false);
                if (rep > 0) us = std::min(us, t);          // the first pass warms the caches
            }
            CUDA_OK(cudaMemcpy(h_S, d_S.p, sizeof(Scalars), cudaMemcpyDeviceToHost));
            if (h_S->trsv_timeout) throw std::runtime_error("triangular-sweep dataflow wait timed out while tuning the part count");
            if (verbosity > 0) fprintf(stderr, "[b200bda] part-count tuning: %d parts (%d asked) %.1f us per lower + upper sweep\n", a.nparts, p, us);
            if (p == p0) p0_us = us;
            if (us < best_us) { best_us = us; best = p; }
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (best != p0 && best_us > 0.97 * p0_us) best = p0;       // within the noise: keep the rule's value
        return best;
    }

    void analyse(int N_, long long nnz_, const int* rows, const int* cols)
    {
        Nb = N_ / 3; N = N_; nnz_stage = nnz_;
        const long long nnzb_in = nnz_ / 9;
        if (rows[Nb] != nnzb_in) throw std::runtime_error("rows[Nb] != nnz / 9");
        AnalysisOptions opt;
        // Parts of the sweeps (one CTA each, every CTA resident; checked below).  Automatic: an SM per part from ~270 k block rows;
        // smaller systems take fewer parts of ~1800 rows -- every part boundary on the dependency path costs a hand-over through
        // L2 (~1 us), a level inside a part 0.15-0.3 us (measured on B200: 27 k rows best with 16-24 parts, 44 k with 24, 110 k
        // with 48-74, 262 k with 148; the round-1 kernel takes an SM per part at every size)
        v2 = sweep_v2 != 0;
        opt.parts = sweep_parts > 0 ? std::min(sweep_parts, 8 * num_sms) : (v2 ? std::max(8, std::min(num_sms, (int) ((Nb + 900) / 1800))) : num_sms);
        // Stage size of the sweeps' TMA ring.  A part's consumers start when its first stage has landed, so a part that is
        // only one or two stages long (Norne size: 80 KB of factor per part) neither overlaps load and compute nor hands its
        // face rows over early; large parts (C3: 2 MB) want the largest stage that fits (per-stage hand-shakes amortised).
        // Measured on B200: C2 10.95 ms per solve at 80 KB, 10.42 at 24 KB, 10.51 at 16 KB, 12.2 at 8 KB; C3 35.8 ms at 80 KB,
        // 37.0 at 64 KB, 41.5 at 48 KB.  Automatic (option value 0): a quarter of the lower sweep's bytes per part.
        int stage_bytes = sweep_stage_bytes;
        if (stage_bytes <= 0) {
            const double nnzL = 0.5 * (double) (nnzb_in - Nb);
            const double part_bytes = (76.0 * nnzL + 52.0 * (double) Nb) / std::max(1, opt.parts);
            stage_bytes = (int) std::min(81920.0, std::max(16384.0, 4096.0 * std::ceil(part_bytes / 4.0 / 4096.0)));
        }
        opt.stageBytes = stage_bytes;
        opt.window = sweep_window;
        opt.extWindow = sweep_ext_window > 0 ? sweep_ext_window : 512;
        opt.buildStreams = !v2;
        opt.warps = sweep_warps;
        opt.groups = sweep_groups;
        auto build_v2 = [&](const int* r_, const int* c_) {
            Sweep2Options o2;
            s2_helpers = std::max(1, std::min(s2_helpers, 8));
            s2_cw = std::max(1, std::min(s2_cw, kS2Threads / 32 - 1));
            s2_cw = std::min(s2_cw, kS2MaxWarps);
            o2.consumerWarps = s2_cw;
            L2 = Sweep2Plan(); U2 = Sweep2Plan();
            build_sweep2_plans(an, r_, c_, o2, L2, U2);
        };
        const bool tune = v2 && sweep_autotune && sweep_parts <= 0 && opt.parts < num_sms && Nb >= 4096 && reorder == 0;
        level_sweeps = false;
        if (reorder != 0 && dist.enabled) throw std::runtime_error("the colour ordering is not available on several ranks");
        if (reorder != 0) {
            // colour the rows, permute the PATTERN (P A P^T) and analyse that: the level sets of the permuted matrix are the
            // colours.  The value, right-hand-side and solution permutations cost nothing extra -- they are composed into the
            // maps the device gathers / scatters with anyway (srcblk, perm).
            color = b200::graph_coloring(Nb, rows, cols, reorder_seed);
            std::vector<int> crows, ccols, csrc;
            b200::permute_pattern(Nb, rows, cols, color, crows, ccols, csrc);
            opt.buildStreams = false;
            opt.parts = 1;
            an = b200::analyse(Nb, crows.data(), ccols.data(), opt);
            for (auto& b : an.srcblk) b = csrc[b];                                  // p-space block -> block of the caller's array
            for (int q = 0; q < Nb; ++q) an.perm[q] = color.fromOrder[an.perm[q]];    // p-space row -> the caller's row
            for (int q = 0; q < Nb; ++q) an.iperm[an.perm[q]] = q;
            level_sweeps = true;
            if (verbosity > 0) fprintf(stderr, "[b200bda] graph colouring: %d colours, %d level sets of the permuted matrix\n", color.ncolors, an.nflev);
        } else if (!dist.enabled) {
            if (tune) opt.parts = autotune_parts(rows, cols, opt, opt.parts);
            an = b200::analyse(Nb, rows, cols, opt);
            if (v2) build_v2(rows, cols);
        } else {
            // Row slab of a partitioned matrix: columns >= Nb are ghosts (numbered last).  The preconditioner is
            // block-Jacobi ILU0 on the owned x owned block, exactly what the reference's parallel ILU0 does
            // (ghost_last_bilu0_decomposition, ParallelOverlappingILU0.hpp:440-494 with interiorSize = owned rows);
            // the owned x ghost blocks only enter the operator (k_spmv_ghost).
            if (!dist.have_halo) throw std::runtime_error("multi-GPU solver: b200_dist_set_halo must precede the first solve");
            std::vector<int> sq_rows((size_t) Nb + 1, 0), sq_cols, sq_src, grow_nat, gptr(1, 0), gcol, gsrc;
            sq_cols.reserve(nnzb_in); sq_src.reserve(nnzb_in);
            for (int r = 0; r < Nb; ++r) {
                bool any = false;
                for (int k = rows[r]; k < rows[r + 1]; ++k) {
                    const int c = cols[k];
                    if (c < 0 || c >= Nb + dist.n_ghost) throw std::runtime_error("column index outside owned + ghost range");
                    if (c < Nb) { sq_cols.push_back(c); sq_src.push_back(k); }
                    else { gcol.push_back(c - Nb); gsrc.push_back(k); any = true; }
                }
                sq_rows[r + 1] = (int) sq_cols.size();
                if (any) { grow_nat.push_back(r); gptr.push_back((int) gcol.size()); }
            }
            if (tune) opt.parts = autotune_parts(sq_rows.data(), sq_cols.data(), opt, opt.parts);
            an = b200::analyse(Nb, sq_rows.data(), sq_cols.data(), opt);
            if (v2) build_v2(sq_rows.data(), sq_cols.data());
            for (auto& b : an.srcblk) b = sq_src[b];            // p-space block -> block of the caller's array
            dist.gnrows = (int) grow_nat.size();
            dist.gnblocks = (long long) gcol.size();
            std::vector<int> grow_p(grow_nat.size()), send_prow(dist.send_rows.size());
            for (size_t i = 0; i < grow_nat.size(); ++i) grow_p[i] = an.iperm[grow_nat[i]];
            for (size_t i = 0; i < send_prow.size(); ++i) {
                if (dist.send_rows[i] < 0 || dist.send_rows[i] >= Nb) throw std::runtime_error("halo send row outside the owned range");
                send_prow[i] = an.iperm[dist.send_rows[i]];
            }
            auto upv = [&](DevBuf<int>& d, const std::vector<int>& h) {
                d.alloc(h.size());
                if (!h.empty()) CUDA_OK(cudaMemcpyAsync(d.p, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice, stream));
            };
            upv(d_grow, grow_p); upv(d_gptr, gptr); upv(d_gcol, gcol); upv(d_gsrc, gsrc); upv(d_send_prow, send_prow);
            CUDA_OK(cudaStreamSynchronize(stream));           // host vectors above are temporaries
        }
        nnzb = an.nnzb; nnz = 9 * nnzb;
        static_assert(kXs == kXwinStride, "device/host value-space stride mismatch");
        static_assert(sizeof(StageD) == sizeof(StageRef) && sizeof(PartD) == sizeof(PartRef) && sizeof(BuildD) == sizeof(BuildRef),
                      "device/host sweep descriptor mismatch");
        auto up = [&](auto& dbuf, const auto& hvec) {
            dbuf.alloc(hvec.size());
            if (!hvec.empty())
                CUDA_OK(cudaMemcpyAsync(dbuf.p, hvec.data(), sizeof(hvec[0]) * hvec.size(), cudaMemcpyHostToDevice, stream));
        };
        up(d_prow, an.prow); up(d_pcol, an.pcol); up(d_pdiag, an.pdiag); up(d_srcblk, an.srcblk); up(d_perm, an.perm);
        up(d_flevRows, an.flevRows);
        sell_slices = 0; sell_slots = 0;
        if (feature_on(spmv_sell)) {
            const b200::SellPlan sp = b200::build_sell(Nb, an.prow, an.pcol, an.srcblk);
            sell_slices = sp.nslices; sell_slots = sp.ptr[sp.nslices];
            up(d_sellPtr, sp.ptr); up(d_sellOver, sp.over); up(d_sellCol, sp.col); up(d_sellSrc, sp.src);
            d_sellVal.alloc((size_t) sell_slots * 288);
            CUDA_OK(cudaStreamSynchronize(stream));            // sp is a temporary
        }
        fac_plan = an.facMaxRow <= kFacMaxRow && an.facMaxOps <= kFacMaxOps;
        if (fac_plan) { up(d_facPtr, an.facPtr); up(d_facOps, an.facOps); }
        fac3_ptr.clear();
        if (fac_plan) {
            std::vector<int> recs;
            recs.reserve((size_t) 4 * (Nb + 3 * an.nflev));
            fac3_ptr.push_back(0);
            for (int l = 0; l < an.nflev; ++l) {
                for (int p = an.flevPtr[l]; p < an.flevPtr[l + 1]; ++p) {
                    const int i = an.flevRows[p], rs = an.prow[i];
                    recs.push_back(i); recs.push_back(rs); recs.push_back(an.facPtr[i]);
                    recs.push_back((an.prow[i + 1] - rs) | ((an.pdiag[i] - rs) << 8) | ((an.facPtr[i + 1] - an.facPtr[i]) << 16));
                }
                while ((recs.size() / 4) % 3) { recs.push_back(-1); recs.push_back(0); recs.push_back(0); recs.push_back(0); }
                fac3_ptr.push_back((int) (recs.size() / 12));
            }
            up(d_facRec, recs);
            fac3_smem = (size_t) fac_warps * 3 * ((size_t) (an.facMaxRow + an.facMaxOps) * 72 + (size_t) an.facMaxOps * 8);
            CUDA_OK(cudaFuncSetAttribute(k_ilu_factor_plan3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) fac3_smem));
            CUDA_OK(cudaStreamSynchronize(stream));            // recs is a temporary
        }
        auto upraw = [&](void* d, const void* h, size_t bytes) { if (bytes) CUDA_OK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, stream)); };
        const bool want_fused = feature_on(fuse_spmv) && sell_slices && an.nparts <= kMaxSweepParts;
        const size_t smem_limit = smem_optin - (want_fused ? 2560 : 0);      // static shared memory of the fused SpMV tail
        int occ = 8;
        int threads = 0;
        auto prep = [&](auto kern) {
            CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sweep_smem));
            int o = 0;
            CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, threads, sweep_smem));
            occ = std::min(occ, o);
        };
        if (level_sweeps) {
            // no pencil schedule: the sweeps run level by level from the BSR factor, SpMV and vector kernels are the separate ones
            defer_ok = false; fused_units = 0; occ = 1; sweep_smem = 0;
            d_valL.alloc(2); d_valU.alloc(2);
        } else if (v2) {
            static_assert(sizeof(S2PartD) == sizeof(S2Part) && sizeof(S2StreamD) == sizeof(S2Stream) &&
                          sizeof(S2BuildD) == sizeof(S2Build), "device/host round-2 sweep descriptor mismatch");
            auto up2 = [&](Sweep2Plan& P2, DevBuf<S2PartD>& dp, DevBuf<S2StreamD>& ds, DevBuf<S2BuildD>& dbu, DevBuf<int>& dh,
                           DevBuf<int>& dc, DevBuf<int>& dsr, DevBuf<double>& dv) {
                dp.alloc(P2.parts.size()); ds.alloc(P2.streams.size()); dbu.alloc(std::max<size_t>(1, P2.build.size()));
                upraw(dp.p, P2.parts.data(), sizeof(S2Part) * P2.parts.size());
                upraw(ds.p, P2.streams.data(), sizeof(S2Stream) * P2.streams.size());
                upraw(dbu.p, P2.build.data(), sizeof(S2Build) * P2.build.size());
                dh.alloc(P2.hdrs.size() + 4 * 64); dc.alloc(P2.codes.size() + 4); dsr.alloc(P2.src.size() + 1);
                upraw(dh.p, P2.hdrs.data(), sizeof(int) * P2.hdrs.size());
                upraw(dc.p, P2.codes.data(), sizeof(int) * P2.codes.size());
                upraw(dsr.p, P2.src.data(), sizeof(int) * P2.src.size());
                dv.alloc((size_t) P2.nvals + 2);
            };
            up2(L2, d_s2partsL, d_s2streamsL, d_s2buildL, d_s2hdrsL, d_s2codesL, d_s2srcL, d_valL);
            up2(U2, d_s2partsU, d_s2streamsU, d_s2buildU, d_s2hdrsU, d_s2codesU, d_s2srcU, d_valU);
            CUDA_OK(cudaStreamSynchronize(stream));
            // the schedule stays on the host only as long as the analysis needs it (the build lists are large)
            for (Sweep2Plan* P2 : {&L2, &U2}) { std::vector<int>().swap(P2->src); std::vector<int>().swap(P2->codes); std::vector<int>().swap(P2->hdrs); std::vector<int>().swap(P2->stepChunks); }
            threads = sweep_threads();
            const size_t need = kS2Header + 24 * (size_t) (an.window + 1);        // window of recent rows + the zero row
            // the SpMV tail streams SELL slices through the same dynamic shared memory: two chunk buffers per ring-fed consumer warp
            // (the shared-memory carve-out is taken from the L1: with the whole 227 KB handed to the tail's ring the sweeps ran 40 % slower --
            // their strided rhs loads and record headers live in L1 -- so only `fuse_ring_warps` consumers are ring-fed)
            const size_t tail = want_fused ? std::min<size_t>(smem_limit, (size_t) 2 * std::max(1, std::min(fuse_chunk, kTailChunk)) * (2304 + 128) * std::min({kTailMaxCons, threads / 32 - 1, std::max(1, fuse_ring_warps)})) : 0;
            sweep_smem = std::max(need, tail);
            if (sweep_smem > smem_limit) throw std::runtime_error("value space of the triangular sweeps does not fit the shared memory");
            prep(k_sweep2<true, false>); prep(k_sweep2<true, true>); prep(k_sweep2<false, false>); prep(k_sweep2<false, true>);
            prep(k_sweep2<true, false, -1, true, true>); prep(k_sweep2<true, true, -1, true, true>); prep(k_sweep2<false, false, -1, true, true>); prep(k_sweep2<false, true, -1, true, true>);
            prep(k_sweep2<true, false, -1, false, true>); prep(k_sweep2<true, true, -1, false, true>); prep(k_sweep2<false, false, -1, false, true>); prep(k_sweep2<false, true, -1, false, true>);
            s2_mlL = L2.nMultiLaneRecords > 0; s2_mlU = U2.nMultiLaneRecords > 0;
            defer_ok = false;
            if (defer_x != 0) {
                prep(k_sweep2<true, false, 3>); prep(k_sweep2<true, false, 3, false, true>);
                d_xSync.alloc(2);
                CUDA_OK(cudaMemsetAsync(d_xSync.p, 0, sizeof(int) * 2, stream));
                defer_ok = true;
            }
            fused_units = 0;
            if (want_fused) {
                prep(k_sweep2<false, true, 1>); prep(k_sweep2<false, true, 2>); prep(k_sweep2<false, true, 1, false, true>); prep(k_sweep2<false, true, 2, false, true>);
                const b200::FusedPlan fp = b200::build_fused(Nb, an.prow, an.pcol, an.partPtr, an.flevPtr, an.flevRows, std::max(1, fuse_unit_slices) * (threads / 32 - 1));
                fused_units = (int) fp.units.size() / 2;
                up(d_fUnits, fp.units); up(d_fNeedPtr, fp.needPtr); up(d_fNeed, fp.need);
                d_fSync.alloc(2 + an.nparts); d_fPartials.alloc((size_t) 2 * fused_units);
                CUDA_OK(cudaMemsetAsync(d_fSync.p, 0, sizeof(int) * (2 + an.nparts), stream));
                CUDA_OK(cudaStreamSynchronize(stream));            // fp is a temporary
            }
            if (verbosity > 0)
                fprintf(stderr, "[b200bda] round-2 sweeps: %d parts, %d consumer warps, %d threads, %zu B smem (window %d rows); "
                                "L: %lld records (up to %d chunks per step, %lld warp-steps with several records), %lld external dependencies, %.1f MB; U: %lld records, %.1f MB\n",
                        an.nparts, s2_cw, threads, sweep_smem, an.window, L2.nrecords, L2.maxChunks, L2.nmulti, L2.nExternal,
                        L2.nvals * 8e-6, U2.nrecords, U2.nvals * 8e-6);
        } else {
        up(d_metaL, an.L.meta); up(d_metaU, an.U.meta); up(d_srcL, an.L.src); up(d_srcU, an.U.src);
        d_stagesL.alloc(an.L.stages.size()); d_stagesU.alloc(an.U.stages.size());
        d_partsL.alloc(an.nparts); d_partsU.alloc(an.nparts);
        d_buildL.alloc(an.L.build.size()); d_buildU.alloc(an.U.build.size());
        upraw(d_stagesL.p, an.L.stages.data(), sizeof(StageRef) * an.L.stages.size());
        upraw(d_stagesU.p, an.U.stages.data(), sizeof(StageRef) * an.U.stages.size());
        upraw(d_partsL.p, an.L.parts.data(), sizeof(PartRef) * an.nparts);
        upraw(d_partsU.p, an.U.parts.data(), sizeof(PartRef) * an.nparts);
        upraw(d_buildL.p, an.L.build.data(), sizeof(BuildRef) * an.L.build.size());
        upraw(d_buildU.p, an.U.build.data(), sizeof(BuildRef) * an.U.build.size());
        d_valL.alloc((size_t) an.L.nvals + 2); d_valU.alloc((size_t) an.U.nvals + 2);
        // sweep launch shape: one CTA per part, ring of `slots` stages + the row window in shared memory
        sweep_metaCap = std::max(an.L.maxMetaInts, an.U.maxMetaInts);
        sweep_valsCap = std::max(an.L.maxValsDoubles, an.U.maxValsDoubles);
        sweep_rhsCap = std::max(an.L.maxRhsRows, an.U.maxRhsRows);
        sweep_extCap = std::max(an.L.maxExtRows, an.U.maxExtRows);
        const size_t slotBytes = (size_t) sweep_metaCap * 4 + (size_t) sweep_valsCap * 8 + (size_t) sweep_rhsCap * 24;
        const size_t fixedBytes = kSweepHeader + (size_t) (an.window + an.extWindow + 2) * 8 * kXs + kSweepTailPad;
        sweep_slots = std::max(2, std::min(sweep_slots, kSweepMaxSlots - 1));      // nslots + 1 ready words
        // helpers fetch one stage ahead of the ring when nslots + 1 stages of external rows fit the ring (sweep_early)
        // the stages in flight must fit the shared memory of an SM and their external rows the external ring
        while (sweep_slots > 2 && (fixedBytes + sweep_slots * slotBytes > smem_limit || (long long) sweep_slots * sweep_extCap > an.extWindow)) --sweep_slots;
        sweep_smem = fixedBytes + sweep_slots * slotBytes;
        sweep_early = feature_on(sweep_early_opt) && (long long) (sweep_slots + 1) * sweep_extCap <= an.extWindow;
        if ((long long) sweep_slots * sweep_extCap > an.extWindow)
            throw std::runtime_error("external-row ring of the triangular sweeps too small (" + std::to_string(sweep_extCap) + " rows per stage)");
        sweep_helpers = std::max(1, std::min({sweep_helpers, 27 - sweep_warps, sweep_slots}));   // a helper must never run a whole ring ahead
        if (sweep_smem > smem_limit)
            throw std::runtime_error("a block row is too long for the shared-memory ring of the triangular sweeps (" +
                                     std::to_string(slotBytes) + " B per stage)");
        threads = sweep_threads();
        prep(k_sweep<true, false, false>); prep(k_sweep<true, true, false>); prep(k_sweep<false, false, false>); prep(k_sweep<false, true, false>);
        prep(k_sweep<true, false, true>); prep(k_sweep<true, true, true>); prep(k_sweep<false, false, true>); prep(k_sweep<false, true, true>);
        defer_ok = false;
        // deferred solution updates pay at every size (C2: 10.15 -> 10.00 ms per solve): automatic = on
        if (defer_x != 0 && threads <= kFusedMaxThreads) {
            prep(k_sweep<true, false, false, 3>);
            d_xSync.alloc(2);
            CUDA_OK(cudaMemsetAsync(d_xSync.p, 0, sizeof(int) * 2, stream));
            defer_ok = true;
        }
        fused_units = 0;
        if (want_fused && threads <= kFusedMaxThreads) {
            prep(k_sweep<false, true, false, 1>); prep(k_sweep<false, true, false, 2>);
            const b200::FusedPlan fp = b200::build_fused(Nb, an.prow, an.pcol, an.partPtr, an.flevPtr, an.flevRows, std::max(1, fuse_unit_slices) * (threads / 32 - 1));
            fused_units = (int) fp.units.size() / 2;
            up(d_fUnits, fp.units); up(d_fNeedPtr, fp.needPtr); up(d_fNeed, fp.need);
            d_fSync.alloc(2 + an.nparts); d_fPartials.alloc((size_t) 2 * fused_units);
            CUDA_OK(cudaMemsetAsync(d_fSync.p, 0, sizeof(int) * (2 + an.nparts), stream));
            CUDA_OK(cudaStreamSynchronize(stream));            // fp is a temporary
        }
        }
        if (occ < 1) throw CudaError("triangular-sweep kernel does not fit on an SM");
        if (!level_sweeps && an.nparts > occ * num_sms)
            throw std::runtime_error("triangular sweeps: " + std::to_string(an.nparts) + " parts cannot all be resident (" +
                                     std::to_string(occ) + " CTAs per SM fit); lower sweep_parts or the stage size");
        if (verbosity > 0 && !v2)
            fprintf(stderr, "[b200bda] analysis: Nb %d nnzb %lld, %d reference levels, %d lines, %d strips, %d parts; "
                            "L: %zu stages %lld chunks (%lld window / %lld global deps), U: %zu stages; sweep grid %d x %d, "
                            "%d slots + window %d rows = %zu B smem; %lld external rows L (max %d per stage)\n",
                    Nb, (long long) nnzb, an.nlev, an.nlines, an.nstrips, an.nparts, an.L.stages.size(), an.L.nchunks, an.L.nWindow,
                    an.L.nExternal, an.U.stages.size(), an.nparts, threads, sweep_slots, sweep_window, sweep_smem,
                    an.L.nExtRows, sweep_extCap);
        if (verbosity > 0 && v2)
            fprintf(stderr, "[b200bda] analysis: Nb %d nnzb %lld, %d reference levels, %d lines, %d strips, %d parts\n", Nb, (long long) nnzb, an.nlev, an.nlines, an.nstrips, an.nparts);
        d_stage.alloc(nnz_stage); d_bstage.alloc(N); d_A.alloc(nnz); d_LU.alloc(nnz);
        // + 8 doubles: the sweeps' 16-byte aligned rhs copies may read one row past the end
        for (DevBuf<double>* v : {&d_x, &d_r, &d_rt, &d_p, &d_v, &d_t, &d_y, &d_w, &d_xnat, &d_tmp1, &d_tmp2}) {
            v->alloc(N + 8);
            CUDA_OK(cudaMemsetAsync(v->p, 0, sizeof(double) * (N + 8), stream));
        }
        CUDA_OK(cudaMemsetAsync(d_xnat.p, 0, sizeof(double) * N, stream));
        CUDA_OK(cudaMemsetAsync(d_v.p, 0, sizeof(double) * N, stream));
        CUDA_OK(cudaMemsetAsync(d_p.p, 0, sizeof(double) * N, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        analysed = true;
    }

    // Option pin_host = 1 (off by default: the library does not own the caller's memory): page-locks a caller's array once it
    // has been handed in on two consecutive calls (Flow's matrix, rhs and solution storage persists over the Newton steps).
    // A caller that passes a fresh buffer every time never pays for the registration -- pinning and unpinning a new 12 MB
    // vector per call cost more than the pageable copy it replaces.  Contract of the option: a buffer handed in stays
    // allocated until it is replaced by another one, released with b200_host_unregister, or the solver is destroyed.  The
    // explicit route (no heuristics) is b200_host_register / b200_host_unregister.
    void maybe_register(const void*& reg, size_t& reg_bytes, const void*& cand, const void* ptr, size_t bytes)
    {
        if (!pin_host) return;
        if (reg == ptr && reg_bytes == bytes) return;
        if (reg) { cudaHostUnregister((void*) reg); cudaGetLastError(); reg = nullptr; reg_bytes = 0; }
        if (cand != ptr) { cand = ptr; return; }     // first sighting: pageable copy this time
        cudaError_t e = cudaHostRegister((void*) ptr, bytes, cudaHostRegisterDefault);
        if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return; }
        if (e != cudaSuccess) { cudaGetLastError(); return; }   // pageable copy still works
        reg = ptr; reg_bytes = bytes;
    }

    // Multisegment wells: everything the two apply kernels need, in one piece (persistent device buffers).
    void upload_mswells(const Wells* w)
    {
        nms = 0; ms_blocks = 0; ms_rows = 0; ms_ncells = 0; ms_dinv_entries = 0;
        if (!w || w->ms.empty()) return;       // (the iteration graph's signature carries the well count)
        // (several ranks: a well lives inside one rank, as the standard wells do -- its apply and the patch of the local dot
        //  products need no exchange; the all-reduce that follows carries the patched sums)
        const MsWellsD before = msD;
        nms = (int) w->ms.size();
        std::vector<int> zoff(nms + 1, 0), rowoff(nms + 1, 0), rowptr(1, 0), bcol, uz_of_block;
        std::vector<long long> doff(nms, 0);
        std::vector<double> B, Cv, Dinv;
        for (int i = 0; i < nms; ++i) {
            const Wells::MsWell& m = w->ms[i];
            zoff[i + 1] = zoff[i] + 4 * (int) m.Mb;
            rowoff[i + 1] = rowoff[i] + (int) m.Mb;
            doff[i] = (long long) Dinv.size();
            Dinv.insert(Dinv.end(), m.Dinv.begin(), m.Dinv.end());
            const int base = (int) bcol.size();
            for (unsigned r = 0; r < m.Mb; ++r) {
                for (unsigned q = m.Brows[r]; q < m.Brows[r + 1]; ++q) {
                    if (m.Bcols[q] >= (unsigned) Nb) throw std::runtime_error("multisegment well column index out of range");
                    bcol.push_back(an.iperm[m.Bcols[q]]);
                    uz_of_block.push_back(zoff[i] + 4 * (int) r);
                }
                rowptr.push_back(base + (int) m.Brows[r + 1]);
            }
            B.insert(B.end(), m.B.begin(), m.B.end());
            Cv.insert(Cv.end(), m.C.begin(), m.C.end());
        }
        ms_blocks = (int) bcol.size(); ms_rows = rowoff[nms]; ms_dinv_entries = (long long) Dinv.size();
        std::vector<int> order(ms_blocks), ucell, uptr, ublock(ms_blocks), uz(ms_blocks);
        for (int p = 0; p < ms_blocks; ++p) order[p] = p;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return bcol[a] < bcol[b]; });
        for (int e = 0; e < ms_blocks; ++e) {
            const int p = order[e];
            if (e == 0 || bcol[p] != ucell.back()) { ucell.push_back(bcol[p]); uptr.push_back(e); }
            ublock[e] = p; uz[e] = uz_of_block[p];
        }
        uptr.push_back(ms_blocks);
        ms_ncells = (int) ucell.size();
        d_msZoff.alloc(nms + 1); d_msRowoff.alloc(nms + 1); d_msDoff.alloc(nms); d_msRowptr.alloc(ms_rows + 1);
        d_msBcol.alloc(ms_blocks); d_msUcell.alloc(ms_ncells); d_msUptr.alloc(ms_ncells + 1); d_msUblock.alloc(ms_blocks);
        d_msUz.alloc(ms_blocks); d_msB.alloc((size_t) ms_blocks * 12); d_msC.alloc((size_t) ms_blocks * 12);
        d_msDinv.alloc(Dinv.size()); d_msZ1.alloc(zoff[nms]); d_msZ2.alloc(zoff[nms]);
        auto H2D = [&](void* d, const void* h, size_t bytes) { if (bytes) CUDA_OK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, stream)); };
        H2D(d_msZoff.p, zoff.data(), sizeof(int) * (nms + 1));
        H2D(d_msRowoff.p, rowoff.data(), sizeof(int) * (nms + 1));
        H2D(d_msDoff.p, doff.data(), sizeof(long long) * nms);
        H2D(d_msRowptr.p, rowptr.data(), sizeof(int) * (ms_rows + 1));
        H2D(d_msBcol.p, bcol.data(), sizeof(int) * ms_blocks);
        H2D(d_msUcell.p, ucell.data(), sizeof(int) * ms_ncells);
        H2D(d_msUptr.p, uptr.data(), sizeof(int) * (ms_ncells + 1));
        H2D(d_msUblock.p, ublock.data(), sizeof(int) * ms_blocks);
        H2D(d_msUz.p, uz.data(), sizeof(int) * ms_blocks);
        H2D(d_msB.p, B.data(), sizeof(double) * B.size());
        H2D(d_msC.p, Cv.data(), sizeof(double) * Cv.size());
        H2D(d_msDinv.p, Dinv.data(), sizeof(double) * Dinv.size());
        CUDA_OK(cudaStreamSynchronize(stream));   // the host vectors above are temporaries
        msD.nwells = nms; msD.zoff = d_msZoff.p; msD.rowoff = d_msRowoff.p; msD.doff = d_msDoff.p; msD.rowptr = d_msRowptr.p;
        msD.bcol = d_msBcol.p; msD.B = d_msB.p; msD.C = d_msC.p; msD.Dinv = d_msDinv.p; msD.z1 = d_msZ1.p; msD.z2 = d_msZ2.p;
        msD.ncells = ms_ncells; msD.ucell = d_msUcell.p; msD.uptr = d_msUptr.p; msD.ublock = d_msUblock.p; msD.uz = d_msUz.p;
        // Flow rebuilds its WellContributions every Newton step: the captured iteration graph (kernel arguments = msD) stays
        // valid as long as the buffers and counts are the same -- the values were refreshed in place above
        if (memcmp(&before, &msD, sizeof(MsWellsD)) != 0) ++ms_epoch;
    }

    void upload_wells(const Wells* w)
    {
        upload_mswells(w);
        nwells = 0; nwblocks = 0; nucells = 0;
        if (!w || w->num_std_wells == 0) return;
        if (w->num_std_wells_so_far != w->num_std_wells || w->num_blocks_so_far != w->num_blocks)
            throw std::runtime_error("WellContributions incomplete: every well needs addMatrix C, D and B");
        nwells = (int) w->num_std_wells; nwblocks = (int) w->num_blocks;
        auto H2D = [&](void* d, const void* h, size_t bytes) { if (bytes) CUDA_OK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, stream)); };
        // The perforation pattern rarely changes between Newton steps: the index side (p-space columns, per-cell gather
        // lists, item table) is rebuilt and uploaded only when val_pointers / Bcols / Ccols differ from the last call
        // (the well container's structure stays on the device, the values are refreshed in one piece).
        const bool same = wells_structure_valid && w->val_pointers == h_wptr && w->Bcols == h_wBcols && w->Ccols == h_wCcols;
        if (!same) {
            std::vector<int> bc(nwblocks), cc(nwblocks);
            for (int p = 0; p < nwblocks; ++p) {
                if (w->Bcols[p] < 0 || w->Bcols[p] >= Nb || w->Ccols[p] < 0 || w->Ccols[p] >= Nb)
                    throw std::runtime_error("well column index out of range");
                bc[p] = an.iperm[w->Bcols[p]];
                cc[p] = an.iperm[w->Ccols[p]];
            }
            // unique perforated cells (by C column) with their contribution lists
            std::vector<int> order(nwblocks), well_of(nwblocks);
            for (int wi = 0; wi < nwells; ++wi)
                for (unsigned p = w->val_pointers[wi]; p < w->val_pointers[wi + 1]; ++p) well_of[p] = wi;
            for (int p = 0; p < nwblocks; ++p) order[p] = p;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cc[a] < cc[b]; });
            h_ucell.clear(); h_uptr.clear(); h_ublock.assign(nwblocks, 0); h_uwell.assign(nwblocks, 0);
            for (int e = 0; e < nwblocks; ++e) {
                int p = order[e];
                if (e == 0 || cc[p] != h_ucell.back()) { h_ucell.push_back(cc[p]); h_uptr.push_back(e); }
                h_ublock[e] = p; h_uwell[e] = well_of[p];
            }
            h_uptr.push_back(nwblocks);
            wells_structure_valid = false;
            h_wptr = w->val_pointers; h_wBcols = w->Bcols; h_wCcols = w->Ccols;
            nucells = (int) h_ucell.size();
            d_wptr.alloc(nwells + 1); d_Bcols.alloc(nwblocks); d_ucell.alloc(nucells); d_uptr.alloc(nucells + 1);
            d_ublock.alloc(nwblocks); d_uwell.alloc(nwblocks);
            d_B.alloc((size_t) nwblocks * 12); d_C.alloc((size_t) nwblocks * 12); d_Dinv.alloc((size_t) nwells * 16);
            d_z2.alloc((size_t) nwells * 4);
            H2D(d_wptr.p, w->val_pointers.data(), sizeof(unsigned) * (nwells + 1));
            H2D(d_Bcols.p, bc.data(), sizeof(int) * nwblocks);
            H2D(d_ucell.p, h_ucell.data(), sizeof(int) * nucells);
            H2D(d_uptr.p, h_uptr.data(), sizeof(int) * (nucells + 1));
            H2D(d_ublock.p, h_ublock.data(), sizeof(int) * nwblocks);
            H2D(d_uwell.p, h_uwell.data(), sizeof(int) * nwblocks);
            CUDA_OK(cudaStreamSynchronize(stream));   // bc is a temporary
        }
        nucells = (int) h_ucell.size();
        H2D(d_B.p, w->Bnnzs.data(), sizeof(double) * nwblocks * 12);
        H2D(d_C.p, w->Cnnzs.data(), sizeof(double) * nwblocks * 12);
        H2D(d_Dinv.p, w->Dnnzs.data(), sizeof(double) * nwells * 16);
        // k_wells_flat: one int4 + four C entries per (unique cell, component), in thread order
        const int nitems = 3 * nucells;
        flat_ok = nwblocks <= 1024 && nitems <= 3072 && nwells <= 128 && 3ll * Nb < (1ll << 31);
        std::vector<int4> item;
        std::vector<double> itemC;
        if (flat_ok) {
            itemC.resize((size_t) nitems * 4);
            if (!same) item.resize(nitems);
            for (int u = 0; u < nucells; ++u)
                for (int c = 0; c < 3; ++c) {
                    const int t = 3 * u + c, e0 = h_uptr[u];
                    if (!same) item[t] = make_int4(3 * h_ucell[u] + c, h_uptr[u + 1] - h_uptr[u], e0, 4 * h_uwell[e0]);
                    for (int k = 0; k < 4; ++k) itemC[(size_t) t * 4 + k] = w->Cnnzs[(size_t) h_ublock[e0] * 12 + 3 * k + c];
                }
            d_item.alloc(nitems); d_itemC.alloc((size_t) nitems * 4);
            if (!same) H2D(d_item.p, item.data(), sizeof(int4) * nitems);
            H2D(d_itemC.p, itemC.data(), sizeof(double) * nitems * 4);
            flatD.nwells = nwells; flatD.nblocks = nwblocks; flatD.nitems = nitems; flatD.wptr = d_wptr.p; flatD.bcol = d_Bcols.p;
            flatD.B = d_B.p; flatD.Dinv = d_Dinv.p; flatD.item = d_item.p; flatD.itemC = d_itemC.p; flatD.C = d_C.p;
            flatD.ublock = d_ublock.p; flatD.uwell = d_uwell.p;
        }
        CUDA_OK(cudaStreamSynchronize(stream));   // the host vectors above are temporaries
        wells_structure_valid = true;
    }

    // H2D of values + rhs (+ pattern and analysis on the first call)
    double upload(int N_, int nnz_, int dim, const double* vals, const int* rows, const int* cols, const double* b,
                  const Wells* wells, double* t_analysis)
    {
        if (dim != 3) throw std::runtime_error("only 3x3 blocks are supported (dim == 3), as BdaBridge.cpp:207-211");
        if (N_ <= 0 || N_ % 3 != 0 || nnz_ <= 0 || nnz_ % 9 != 0) throw std::runtime_error("N must be 3*Nb and nnz 9*nnzb");
        if (!vals || !b) throw std::runtime_error("null vals / b");
        *t_analysis = 0.0;
        if (!analysed) {
            if (!rows || !cols) throw std::runtime_error("null rows / cols on the first call");
            double t0 = wall();
            analyse(N_, nnz_, rows, cols);
            *t_analysis = wall() - t0;
        } else if (N_ != N || nnz_ != nnz_stage) {
            throw std::runtime_error("sparsity pattern changed after the first call (fixed, cusparseSolverBackend.cu:312)");
        }
        maybe_register(reg_vals, reg_vals_bytes, cand_vals, vals, sizeof(double) * nnz_stage);
        maybe_register(reg_b, reg_b_bytes, cand_b, b, sizeof(double) * N);
        CUDA_OK(cudaEventRecord(ev_a, stream));
        CUDA_OK(cudaMemcpyAsync(d_stage.p, vals, sizeof(double) * nnz_stage, cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaMemcpyAsync(d_bstage.p, b, sizeof(double) * N, cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaEventRecord(ev_b, stream));
        upload_wells(wells);
        have_system = true; have_factor = false;
        CUDA_OK(cudaEventSynchronize(ev_b));
        float ms = 0.f;
        CUDA_OK(cudaEventElapsedTime(&ms, ev_a, ev_b));
        return ms * 1e-3;
    }

    // ---- kernels ------------------------------------------------------------------------------
    int blocks_for(long long n, int threads, int cap) const
    {
        long long b = (n + threads - 1) / threads;
        return (int) std::max<long long>(1, std::min<long long>(b, cap));
    }

    void permute_values()
    {
        int id = prof_begin(K_PERMUTE);
        k_permute_vals<<<blocks_for(nnz, 256, num_sms * 16), 256, 0, stream>>>(d_stage.p, d_srcblk.p, d_A.p, nnz);
        if (sell_slices)
            k_fill_sell<<<blocks_for(sell_slots * 288, 256, num_sms * 16), 256, 0, stream>>>(d_stage.p, d_sellSrc.p, d_sellVal.p, sell_slots * 288);
        prof_end(id);
    }

    // ILU0 of the resident A: one kernel per level set, chained by programmatic dependent launch, + the two stream fills.
    // The level launches only depend on the pattern, so they are captured once into a CUDA graph and replayed per solve
    // (298 launches for C3).  With per-kernel profiling on, the chain is timed as ONE entry (events between the levels would
    // break the programmatic chain and time the event gaps, not the kernels): launches = levels, ms = the whole chain.
    void factorize_levels()
    {
        CUDA_OK(cudaMemsetAsync(&d_S.p->singular, 0, sizeof(int), stream));
        for (int l = 0; l < an.nflev; ++l) {
            int row0 = an.flevPtr[l], nrows = an.flevPtr[l + 1] - row0;
            if (fac3_ready()) {
                const int t0 = fac3_ptr[l], nt = fac3_ptr[l + 1] - t0;
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((nt + fac_warps - 1) / fac_warps); cfg.blockDim = dim3(32 * fac_warps); cfg.stream = stream;
                cfg.dynamicSmemBytes = fac3_smem;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = (fac_pdl && l > 0) ? 1 : 0;
                CUDA_OK(cudaLaunchKernelEx(&cfg, k_ilu_factor_plan3, reinterpret_cast<const int4*>(d_facRec.p) + 3 * (size_t) t0, nt,
                                           reinterpret_cast<const int2*>(d_facOps.p), (const double*) d_A.p, d_LU.p, an.facMaxRow, an.facMaxOps, d_S.p));
            } else if (fac_plan && fac_pdl && l > 0) {
                // programmatic dependent launch on the previous level: the launch gap and the prologue hide behind it
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((nrows + kFacWarps - 1) / kFacWarps); cfg.blockDim = dim3(32 * kFacWarps); cfg.stream = stream;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                CUDA_OK(cudaLaunchKernelEx(&cfg, k_ilu_factor_plan, (const int*) d_prow.p, (const int*) d_pdiag.p, (const int*) d_facPtr.p,
                                           reinterpret_cast<const int2*>(d_facOps.p), (const double*) d_A.p, d_LU.p,
                                           (const int*) (d_flevRows.p + row0), nrows, d_S.p));
            } else if (fac_plan)
                k_ilu_factor_plan<<<(nrows + kFacWarps - 1) / kFacWarps, 32 * kFacWarps, 0, stream>>>(
                    d_prow.p, d_pdiag.p, d_facPtr.p, reinterpret_cast<const int2*>(d_facOps.p), d_A.p, d_LU.p, d_flevRows.p + row0, nrows, d_S.p);
            else
                k_ilu_factor_level<<<(nrows + 7) / 8, 256, 0, stream>>>(d_prow.p, d_pcol.p, d_pdiag.p, d_A.p, d_LU.p, d_flevRows.p + row0, nrows, d_S.p);
        }
    }
    void fill_streams()
    {
        if (level_sweeps) return;          // the level-synchronous sweeps read the BSR factor itself
        if (v2) {
            int id2 = prof_begin(K_SLICES);
            if (!L2.build.empty())
                k_fill_stream2<true><<<blocks_for((long long) L2.build.size() * 32, 256, num_sms * 8), 256, 0, stream>>>(
                    d_s2buildL.p, (int) L2.build.size(), d_s2srcL.p, d_LU.p, d_valL.p, 1.0);
            prof_end(id2);
            id2 = prof_begin(K_SLICES);
            if (!U2.build.empty())
                k_fill_stream2<false><<<blocks_for((long long) U2.build.size() * 32, 256, num_sms * 8), 256, 0, stream>>>(
                    d_s2buildU.p, (int) U2.build.size(), d_s2srcU.p, d_LU.p, d_valU.p, relaxation);
            prof_end(id2);
            return;
        }
        int id = prof_begin(K_SLICES);
        k_fill_stream<true><<<blocks_for((long long) an.L.build.size() * 32, 256, num_sms * 8), 256, 0, stream>>>(
            d_buildL.p, (int) an.L.build.size(), d_srcL.p, d_LU.p, d_valL.p, 1.0);
        prof_end(id);
        id = prof_begin(K_SLICES);
        k_fill_stream<false><<<blocks_for((long long) an.U.build.size() * 32, 256, num_sms * 8), 256, 0, stream>>>(
            d_buildU.p, (int) an.U.build.size(), d_srcU.p, d_LU.p, d_valU.p, relaxation);
        prof_end(id);
    }
    void factorize()
    {
        const int id = prof_begin(K_FACTOR);
        stats[K_FACTOR].launches += an.nflev - 1; launch_count += an.nflev - 1;
        if (!use_graph) {
            factorize_levels();
        } else {
            if (!fac_graph_exec) {
                cudaGraph_t g = nullptr;
                CUDA_OK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
                try {
                    factorize_levels();
                } catch (...) {
                    cudaStreamEndCapture(stream, &g);
                    if (g) cudaGraphDestroy(g);
                    throw;
                }
                CUDA_OK(cudaStreamEndCapture(stream, &g));
                CUDA_OK(cudaGraphInstantiate(&fac_graph_exec, g, 0));
                cudaGraphDestroy(g);
            }
            CUDA_OK(cudaGraphLaunch(fac_graph_exec, stream));
        }
        prof_end(id);
        fill_streams();
        have_factor = true;
    }

    SweepArgs sweep_args(bool lower, const double* rhs, double* out, double* rearm, bool check_done) const
    {
        SweepArgs a = {};
        a.stages = lower ? d_stagesL.p : d_stagesU.p;
        a.parts = lower ? d_partsL.p : d_partsU.p;
        a.meta = lower ? d_metaL.p : d_metaU.p;
        a.vals = lower ? d_valL.p : d_valU.p;
        a.rhs = rhs; a.out = out; a.rearm = rearm; a.S = d_S.p;
        a.nparts = an.nparts; a.nslots = sweep_slots; a.window = an.window;
        a.metaCap = sweep_metaCap; a.valsCap = sweep_valsCap; a.rhsCap = sweep_rhsCap; a.extWindow = an.extWindow;
        a.nwarps = sweep_warps; a.nhalo = sweep_helpers; a.helper_sleep = sweep_helper_sleep;
        a.check_done = check_done ? 1 : 0;
        a.nowait = sweep_nowait; a.early = sweep_early ? 1 : 0;
        if (sweep_trace && an.nparts > 148) throw std::runtime_error("sweep_trace: the trace buffer holds 148 parts");
        a.trace = sweep_trace ? d_trace.p : nullptr;
        a.trace_cap = kTraceCap;
        if (v2) {
            a.v2.parts = lower ? d_s2partsL.p : d_s2partsU.p;
            a.v2.streams = lower ? d_s2streamsL.p : d_s2streamsU.p;
            a.v2.hdrs = reinterpret_cast<const int2*>(lower ? d_s2hdrsL.p : d_s2hdrsU.p);
            a.v2.codes = reinterpret_cast<const int4*>(lower ? d_s2codesL.p : d_s2codesU.p);
            a.v2.vals = lower ? d_valL.p : d_valU.p;
            a.v2.window = an.window; a.v2.ncw = s2_cw;
            a.nwarps = s2_cw;              // the SpMV tail's producer warp = the warp after the consumers
            a.helper_sleep = s2_poll_lead;  // steps before its own a warp starts to poll external rows
            a.early = s2_prefetch;          // records ahead whose bytes are pulled into L2
        }
        return a;
    }
    int sweep_threads() const { return v2 ? std::min(kS2Threads, (s2_cw + s2_helpers) * 32) : (sweep_warps + 1 + sweep_helpers) * 32; }
    // Launch of a kernel of the BiCGSTAB iteration (all of them start with pdl_enter()): with iter_pdl the launch carries the
    // programmatic-stream-serialization attribute, so its CTAs are scheduled while the previous kernel drains.
    int iter_pdl = 0;                  // option (measured: slower)
    template <class... KArgs, class... Args>
    void launch_iter(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args)
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = iter_pdl ? 1 : 0;
        cfg.attrs = at; cfg.numAttrs = 1;
        CUDA_OK(cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...));
    }
    template <bool LOWER>
    void launch_sweep(const SweepArgs& a)
    {
        const bool rearm = a.rearm != nullptr, trace = a.trace != nullptr;
        auto go = [&](auto kern) { launch_iter(kern, dim3(an.nparts), dim3(sweep_threads()), sweep_smem, a); };
        if (v2) {
            const bool ml = LOWER ? s2_mlL : s2_mlU;
            if (trace) { if (rearm) go(k_sweep2<LOWER, true, -1, true, true>); else go(k_sweep2<LOWER, false, -1, true, true>); }
            else if (ml) { if (rearm) go(k_sweep2<LOWER, true, -1, false, true>); else go(k_sweep2<LOWER, false, -1, false, true>); }
            else { if (rearm) go(k_sweep2<LOWER, true>); else go(k_sweep2<LOWER, false>); }
            return;
        }
        if (trace) { if (rearm) go(k_sweep<LOWER, true, true>); else go(k_sweep<LOWER, false, true>); }
        else { if (rearm) go(k_sweep<LOWER, true, false>); else go(k_sweep<LOWER, false, false>); }
    }
    // lower sweep whose CTAs then apply the pending solution update x += pend * y (kernels.cuh xupdate_tail)
    void trsv_lower_xupdate(const double* rhs, double* out)
    {
        int id = prof_begin(K_LOWER);
        SweepArgs a = sweep_args(true, rhs, out, nullptr, true);
        a.xu.x = d_x.p; a.xu.y = d_y.p; a.xu.sync = d_xSync.p; a.xu.n = N;
        if (v2 && s2_mlL) launch_iter(k_sweep2<true, false, 3, false, true>, dim3(tail_grid()), dim3(sweep_threads()), sweep_smem, a);
        else if (v2) launch_iter(k_sweep2<true, false, 3>, dim3(tail_grid()), dim3(sweep_threads()), sweep_smem, a);
        else launch_iter(k_sweep<true, false, false, 3>, dim3(an.nparts), dim3(sweep_threads()), sweep_smem, a);
        prof_end(id);
    }
    // round-2 sweeps with a tail (SpMV, x update): one CTA per SM even when the schedule has fewer parts -- the CTAs beyond the
    // parts skip the sweep and work on the tail (each CTA fills an SM: 512 threads x 128 registers, so all are resident)
    // Measured (one B200): 45 parts (130 k rows) 8.82 -> 8.38 ms per solve, 56 parts (110 k) 9.73 -> 9.18, 86 parts (250 k)
    // 11.87 -> 12.06: automatic (1) = only when the parts take at most half of the SMs; 2 = always, 0 = never.
    int tail_all_sms = 1;              // option
    int tail_grid() const { return (tail_all_sms == 2 || (tail_all_sms == 1 && 2 * an.nparts <= num_sms)) ? std::max(an.nparts, num_sms) : an.nparts; }
    bool defer_now() const { return defer_ok && !sweep_trace; }
    int defer_x = 1;                   // option
    bool defer_ok = false;             // the tail kernel fits (set by the analysis)
    DevBuf<int> d_xSync;
    // level-synchronous sweep (colour ordering): one launch per level set, ascending for L, descending for U
    template <bool LOWER>
    void trsv_levels(const double* rhs, double* out, bool check_done)
    {
        for (int n = 0; n < an.nflev; ++n) {
            const int l = LOWER ? n : an.nflev - 1 - n;
            const int row0 = an.flevPtr[l], nrows = an.flevPtr[l + 1] - row0;
            const int warps = (nrows + 9) / 10;
            launch_iter(k_trsv_level<LOWER>, dim3((warps + 7) / 8), dim3(256), 0, d_prow.p, d_pcol.p, d_pdiag.p, d_LU.p, rhs, out,
                        d_flevRows.p + row0, nrows, LOWER ? 1.0 : relaxation, d_S.p, check_done ? 1 : 0);
        }
    }
    void trsv_lower(const double* rhs, double* out, bool check_done)
    {
        int id = prof_begin(K_LOWER);
        if (level_sweeps) { trsv_levels<true>(rhs, out, check_done); prof_end(id); return; }
        launch_sweep<true>(sweep_args(true, rhs, out, nullptr, check_done));
        prof_end(id);
    }
    void trsv_upper(const double* rhs, double* out, double* rearm, bool check_done)
    {
        int id = prof_begin(K_UPPER);
        if (level_sweeps) { trsv_levels<false>(rhs, out, check_done); prof_end(id); return; }
        launch_sweep<false>(sweep_args(false, rhs, out, rearm, check_done));
        prof_end(id);
    }
    // upper sweep + the SpMV that follows it in one launch: y = A * out, dots as spmv<MODE>
    template <int MODE>
    void trsv_upper_spmv(const double* rhs, double* out, double* rearm, double* y, const double* d1)
    {
        int id = prof_begin(K_UPPER_SPMV);
        SweepArgs a = sweep_args(false, rhs, out, rearm, true);
        a.f.sptr = d_sellPtr.p; a.f.sover = d_sellOver.p; a.f.scol = d_sellCol.p; a.f.sval = d_sellVal.p;
        a.f.prow = d_prow.p; a.f.pcol = d_pcol.p; a.f.A = d_A.p; a.f.y = y; a.f.d1 = d1;
        a.f.units = reinterpret_cast<const int2*>(d_fUnits.p); a.f.need_ptr = d_fNeedPtr.p; a.f.need = d_fNeed.p;
        a.f.sync = d_fSync.p; a.f.partials = d_fPartials.p; a.f.Nb = Nb; a.f.nunits = fused_units;
        a.f.dbg = nullptr; a.f.ring_bytes = (int) sweep_smem; a.f.chunk = fuse_chunk; a.f.prefetch = fuse_prefetch;
        if (fuse_debug > 0) { --fuse_debug; d_fDbg.alloc((size_t) 4 * an.nparts); a.f.dbg = d_fDbg.p; }
        if (v2 && s2_mlU) launch_iter(k_sweep2<false, true, MODE, false, true>, dim3(tail_grid()), dim3(sweep_threads()), sweep_smem, a);
        else if (v2) launch_iter(k_sweep2<false, true, MODE>, dim3(tail_grid()), dim3(sweep_threads()), sweep_smem, a);
        else launch_iter(k_sweep<false, true, false, MODE>, dim3(an.nparts), dim3(sweep_threads()), sweep_smem, a);
        prof_end(id);
    }
    bool fused_now() const { return fused_units > 0 && !sweep_trace; }
    int fuse_debug = 0, sweep_nowait = 0;
    int fac_pdl = 1;                   // option: factorisation levels chained by programmatic dependent launch
    bool sweep_early = false;
    int sweep_early_opt = 1;           // option "sweep_early"
    int fuse_unit_slices = 2;          // option: SELL slices per consumer warp and unit
    int fuse_chunk = 4;                // option: slots of a SELL slice per chunk buffer of the SpMV tail (1..4)
    int fuse_prefetch = 0;             // option: the tail's producer warp pulls a unit's SELL slices into L2 when the unit starts
                                       // (measured on C3: 243 -> 257 us per upper sweep + SpMV launch -- the requests compete with the sweeps' own)
    int fuse_ring_warps = 8;           // option (round-2 sweeps): warps of the SpMV tail fed through the shared-memory ring (19 KB each)
    DevBuf<long long> d_fDbg;
    template <int MODE>
    void spmv(const double* x, double* y, const double* d1)
    {
        int id = prof_begin(K_SPMV);
        if (sell_slices)
            launch_iter(k_spmv_sell<MODE>, dim3(blocks_for(32LL * sell_slices, kVecThreads, spmv_blocks_cap)), dim3(kVecThreads), 0,
                        d_sellPtr.p, d_sellOver.p, d_sellCol.p, d_sellVal.p, d_prow.p, d_pcol.p, d_A.p, x, y, d1, Nb, sell_slices, d_S.p,
                        d_partials.p, d_ticket.p);
        else
            launch_iter(k_spmv<MODE>, dim3(blocks_for(N, kVecThreads, spmv_blocks_cap)), dim3(kVecThreads), 0, d_prow.p, d_pcol.p, d_A.p, x, y,
                        d1, N, d_S.p, d_partials.p, d_ticket.p);
        prof_end(id);
    }
    // multisegment wells first, then the standard wells: the order of WellContributions::apply (WellContributions.cu:167-193)
    template <int MODE>
    void mswells_apply(const double* x, double* y, const double* d1)
    {
        if (nms == 0) return;
        int id = prof_begin(K_WELL);
        launch_iter(k_mswell_z, dim3(nms), dim3(256), 0, msD, x, d_S.p, MODE != 0 ? 1 : 0);
        launch_iter(k_mswell_y<MODE>, dim3(1), dim3(1024), 0, msD, y, d1, d_S.p);
        prof_end(id);
    }
    template <int MODE>
    void wells_apply(const double* x, double* y, const double* d1)
    {
        mswells_apply<MODE>(x, y, d1);
        if (nwells == 0) return;
        int id = prof_begin(K_WELL);
        if (flat_ok && wells_flat && wells_cluster) {
            launch_iter(k_wells_cluster<MODE>, dim3(kWellClCtas), dim3(kWellClThreads), 0, flatD, x, y, d1, d_S.p);
            prof_end(id);
            return;
        }
        if (flat_ok && wells_flat) {
            launch_iter(k_wells_flat<MODE>, dim3(1), dim3(1024), 0, flatD, x, y, d1, d_S.p);
            prof_end(id);
            return;
        }
        launch_iter(k_wells<MODE>, dim3(1), dim3(1024), 0, nwells, d_wptr.p, d_Bcols.p, d_B.p, d_C.p, d_Dinv.p, nucells, d_ucell.p, d_uptr.p,
                                              d_ublock.p, d_uwell.p, d_z2.p, x, y, d1, d_S.p);
        prof_end(id);
    }

    // ---- multi-GPU pieces (no-ops on a single GPU) ------------------------------------------------
    void allreduce(void* buf, int count, int dtype, int op)
    {
        if (!dist.enabled || dist.world == 1) return;
        int id = prof_begin(K_ALLREDUCE);
        NCCL_OK(g_nccl.AllReduce(buf, buf, (size_t) count, dtype, op, dist.comm, stream));
        prof_end(id);
    }
    void allreduce_sum(double* buf, int count) { allreduce(buf, count, kNcclFloat64, kNcclSum); }
    template <int PHASE>
    void finish()
    {
        if (!dist.enabled) return;
        int id = prof_begin(K_FINISH);
        k_finish<PHASE><<<1, 32, 0, stream>>>(d_S.p, tolerance, 2 * maxit);
        prof_end(id);
    }
    // Sum the scalars of one Krylov phase over the ranks and run its epilogue.  PHASE as k_allreduce_p2p.
    // Peer-memory mailboxes when every rank is mapped (default), else NCCL + k_finish.
    // The exchange of a phase runs in the epilogue of the kernel that completes its local sums (mail_allreduce: k_vec_xr1,
    // k_vec_xr2, and k_spmv_ghost on ranks that have ghost columns); a launch of its own only where there is no such kernel.
    int fuse_allreduce = 1;            // option
    bool mail_on() const { return dist.enabled && dist.world > 1 && dist.use_p2p_allreduce && dist.mail_ready; }
    bool ghost_on() const { return dist.enabled && dist.nneigh > 0 && dist.gnrows > 0; }
    DistRedD dist_red() const
    {
        DistRedD D = {};
        if (mail_on() && fuse_allreduce) { D.mail = d_mail.p; D.seq = d_dist_ctr.p; D.rank = dist.rank; D.world = dist.world; D.tol = tolerance; D.max_half = 2 * maxit; }
        return D;
    }
    template <int PHASE>
    void reduce_phase()
    {
        if (!dist.enabled) return;
        if (mail_on() && fuse_allreduce && (PHASE == 2 || PHASE == 4 || ((PHASE == 1 || PHASE == 3) && ghost_on()))) return;
        if (dist.world > 1 && dist.use_p2p_allreduce && dist.mail_ready) {
            int id = prof_begin(K_ALLREDUCE);
            k_allreduce_p2p<PHASE><<<1, 64, 0, stream>>>(dist.mail, dist.rank, dist.world, d_dist_ctr.p, d_S.p, tolerance, 2 * maxit);
            prof_end(id);
            return;
        }
        if (PHASE == 0) { allreduce_sum(d_S.p->red, 1); finish<0>(); }
        if (PHASE == 1) allreduce_sum(&d_S.p->h, 1);
        if (PHASE == 2) { allreduce_sum(d_S.p->red, 1); finish<1>(); }
        if (PHASE == 3) allreduce_sum(&d_S.p->tr, 2);
        if (PHASE == 4) { allreduce_sum(d_S.p->red, 2); finish<2>(); }
        if (PHASE == 5) allreduce(&d_S.p->singular, 1, kNcclInt32, kNcclMax);
    }
    // boundary entries of y (p-space) -> the neighbours' receive blocks; one epoch per exchange.  The push only reads y and the
    // exchange counters, so it runs on a side stream beside the kernels that follow the sweep (the well apply, or the separate
    // SpMV) and is joined before k_spmv_ghost, the consumer of the exchange (halo_join).
    int halo_side = 1;                 // option
    bool halo_forked = false;
    void halo_push(const double* y, bool check_done)
    {
        if (!dist.enabled) return;
        if (dist.nneigh == 0) return;
        if (!dist.peers_ready) throw std::runtime_error("multi-GPU solver: peers not connected (b200_dist_connect_peer)");
        int maxsend = 0;
        for (int n = 0; n < dist.nneigh; ++n) maxsend = std::max(maxsend, dist.send_ptr[n + 1] - dist.send_ptr[n]);
        const int bx = std::max(1, std::min(128, (3 * maxsend + 511) / 512));
        cudaStream_t st = stream;
        halo_forked = halo_side && !profile;
        if (halo_forked) {
            CUDA_OK(cudaEventRecord(ev_fork, stream));
            CUDA_OK(cudaStreamWaitEvent(stream_aux, ev_fork, 0));
            st = stream_aux;
        }
        int id = prof_begin(K_HALO_PUSH);
        k_halo_push<<<dim3(bx, dist.nneigh), 256, 0, st>>>(d_peers.p, d_send_prow.p, y, d_dist_ctr.p + 1, d_push_tickets.p, d_S.p,
                                                            check_done ? 1 : 0);
        prof_end(id);
        if (halo_forked) CUDA_OK(cudaEventRecord(ev_join, stream_aux));
    }
    void halo_join()
    {
        if (!halo_forked) return;
        CUDA_OK(cudaStreamWaitEvent(stream, ev_join, 0));
        halo_forked = false;
    }
    const double* ghost_x0() const { return reinterpret_cast<const double*>(d_halo.p + kHaloRecvOffset); }      // parity 0; parity 1 follows 3 n_ghost doubles later
    template <int MODE>
    void spmv_ghost(double* y, const double* d1, bool check_done)
    {
        if (!dist.enabled || dist.nneigh == 0 || dist.gnrows == 0) return;
        int id = prof_begin(K_SPMV_GHOST);
        const int blocks = blocks_for(3ll * dist.gnrows, kVecThreads, num_sms * 2);
        k_spmv_ghost<MODE><<<blocks, kVecThreads, 0, stream>>>(dist.gnrows, d_grow.p, d_gptr.p, d_gcol.p, d_gsrc.p, d_stage.p, ghost_x0(),
                                                               3ll * dist.n_ghost, reinterpret_cast<const unsigned*>(d_halo.p), dist.nneigh,
                                                               d_dist_ctr.p + 1, d_dist_ctr.p + 2, y,
                                                               d1, d_S.p, d_partials.p, d_ticket.p, check_done ? 1 : 0,
                                                               MODE != 0 ? dist_red() : DistRedD{});
        prof_end(id);
    }

    void enqueue_iteration()
    {
        const int dm = dist.enabled ? 1 : 0;
        int id;
        id = prof_begin(K_VEC_P);
        launch_iter(k_vec_p, dim3(vec_blocks), dim3(kVecThreads), 0, d_r.p, d_p.p, d_v.p, N, d_S.p);
        prof_end(id);
        if (defer_now()) trsv_lower_xupdate(d_p.p, d_w.p); else trsv_lower(d_p.p, d_w.p, true);
        if (fused_now()) {
            trsv_upper_spmv<1>(d_w.p, d_y.p, d_w.p, d_v.p, d_rt.p);
            halo_push(d_y.p, true);       // multi-GPU: the neighbours' boundary rows wait for it in spmv_ghost
        } else {
            trsv_upper(d_w.p, d_y.p, d_w.p, true);
            halo_push(d_y.p, true);
            spmv<1>(d_y.p, d_v.p, d_rt.p);
        }
        wells_apply<1>(d_y.p, d_v.p, d_rt.p);
        halo_join();
        spmv_ghost<1>(d_v.p, d_rt.p, true);
        reduce_phase<1>();
        id = prof_begin(K_VEC_XR1);
        launch_iter(k_vec_xr1, dim3(vec_blocks), dim3(kVecThreads), 0, d_x.p, d_y.p, d_r.p, d_v.p, N, d_S.p, d_partials.p, d_ticket.p, dm, defer_now() ? 1 : 0, dist_red());
        prof_end(id);
        reduce_phase<2>();
        if (defer_now()) trsv_lower_xupdate(d_r.p, d_w.p); else trsv_lower(d_r.p, d_w.p, true);
        if (fused_now()) {
            trsv_upper_spmv<2>(d_w.p, d_y.p, d_w.p, d_t.p, d_r.p);
            halo_push(d_y.p, true);
        } else {
            trsv_upper(d_w.p, d_y.p, d_w.p, true);
            halo_push(d_y.p, true);
            spmv<2>(d_y.p, d_t.p, d_r.p);
        }
        wells_apply<2>(d_y.p, d_t.p, d_r.p);
        halo_join();
        spmv_ghost<2>(d_t.p, d_r.p, true);
        reduce_phase<3>();
        id = prof_begin(K_VEC_XR2);
        launch_iter(k_vec_xr2, dim3(vec_blocks), dim3(kVecThreads), 0, d_x.p, d_y.p, d_r.p, d_t.p, d_rt.p, N, d_S.p, d_partials.p, d_ticket.p, dm, defer_now() ? 1 : 0, dist_red());
        prof_end(id);
        reduce_phase<4>();
    }

    // One BiCGSTAB iteration: on a single GPU the 10-12 launches are replayed from a CUDA graph (the launch sequence
    // depends on nothing but the pointers; the convergence logic lives on the device), otherwise launched one by one.
    struct IterSig { const void* a[8]; int n[6]; bool operator!=(const IterSig& o) const { return memcmp(this, &o, sizeof *this) != 0; } };
    cudaGraphExec_t iter_graph_exec = nullptr;
    IterSig iter_sig{};
    void run_iteration()
    {
        // (several GPUs: graph replay needs the peer-memory all-reduce -- NCCL calls are not captured here)
        if (!use_graph || profile || sweep_trace || (dist.enabled && dist.world > 1 && !(dist.use_p2p_allreduce && dist.mail_ready))) { enqueue_iteration(); return; }
        IterSig sig{};
        sig.a[0] = d_B.p; sig.a[1] = d_C.p; sig.a[2] = d_Dinv.p; sig.a[3] = d_ucell.p; sig.a[4] = d_wptr.p; sig.a[5] = d_z2.p;
        sig.n[0] = nwells; sig.n[1] = nucells; sig.n[2] = nwblocks; sig.n[3] = sweep_helper_sleep + (fused_now() ? 1 << 20 : 0) + (defer_now() ? 1 << 21 : 0);
        sig.n[4] = nms; sig.n[5] = ms_epoch;
        sig.a[6] = d_item.p; sig.a[7] = d_itemC.p;
        sig.n[3] += (flat_ok && wells_flat) ? 1 << 22 : 0;
        sig.n[3] += wells_cluster ? 1 << 25 : 0;
        sig.n[3] += level_sweeps ? 1 << 26 : 0;
        sig.n[3] += (mail_on() && fuse_allreduce) ? 1 << 23 : 0;
        sig.n[3] += halo_side ? 1 << 24 : 0;
        if (!iter_graph_exec || sig != iter_sig) {
            if (iter_graph_exec) { cudaGraphExecDestroy(iter_graph_exec); iter_graph_exec = nullptr; }
            cudaGraph_t g = nullptr;
            CUDA_OK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
            counting = false;
            try {
                enqueue_iteration();
            } catch (...) {
                counting = true;
                cudaStreamEndCapture(stream, &g);
                if (g) cudaGraphDestroy(g);
                throw;
            }
            counting = true;
            CUDA_OK(cudaStreamEndCapture(stream, &g));
            CUDA_OK(cudaGraphInstantiate(&iter_graph_exec, g, 0));
            cudaGraphDestroy(g);
            iter_sig = sig;
        }
        CUDA_OK(cudaGraphLaunch(iter_graph_exec, stream));
        stats[K_VEC_P].launches++; stats[K_VEC_XR1].launches++; stats[K_VEC_XR2].launches++;
        stats[K_LOWER].launches += 2;
        if (fused_now()) stats[K_UPPER_SPMV].launches += 2;
        else { stats[K_UPPER].launches += 2; stats[K_SPMV].launches += 2; }
        if (nwells) stats[K_WELL].launches += 2;
        if (nms) stats[K_WELL].launches += 4;
        launch_count += (fused_now() ? 7 : 9) + (nwells ? 2 : 0) + (nms ? 4 : 0);
        if (dist.enabled && dist.world > 1) {      // the multi-GPU kernels of the replayed iteration
            const int push = dist.nneigh > 0 ? 2 : 0, ghost = ghost_on() ? 2 : 0;
            const int ar = fuse_allreduce ? (ghost_on() ? 0 : 2) : 4;      // all-reduce launches of their own (the others run inside their producers)
            stats[K_HALO_PUSH].launches += push; stats[K_SPMV_GHOST].launches += ghost; stats[K_ALLREDUCE].launches += ar;
            launch_count += push + ghost + ar;
        }
    }

    // permutation + ILU0 + BiCGSTAB on the resident system
    void solve_resident(b200_result* res)
    {
        if (!have_system) throw std::runtime_error("no system uploaded");
        if (poisoned) throw std::runtime_error("this solver saw a dataflow wait time out and is unusable: create a new one");
        const double t0 = wall();
        // verbosity >= 3: per-phase timings as the reference's backends print them (BdaSolver.hpp:47-51,
        // openclSolverBackend.cpp:451-459): every kernel of this solve is timed with CUDA events (no graph replay)
        const bool phase_times = verbosity >= 3 && !profile;
        double ms_before[K_COUNT];
        if (phase_times) { profile = true; for (int k = 0; k < K_COUNT; ++k) ms_before[k] = stats[k].ms; }
        CUDA_OK(cudaEventRecord(ev_a, stream));
        permute_values();
        factorize();
        CUDA_OK(cudaEventRecord(ev_b, stream));
        reduce_phase<5>();                                           // a failed pivot on any rank stops all of them
        int id = prof_begin(K_INIT);
        k_init<<<vec_blocks, kVecThreads, 0, stream>>>(d_bstage.p, d_perm.p, d_r.p, d_rt.p, d_x.p, d_w.p, d_y.p, N, d_S.p,
                                                        d_partials.p, d_ticket.p, tolerance, 2 * maxit, dist.enabled ? 1 : 0);
        prof_end(id);
        reduce_phase<0>();
        // the sweep tails leave their counters and flags zeroed; once per solve in case a launch was cut short
        if (fused_units) CUDA_OK(cudaMemsetAsync(d_fSync.p, 0, sizeof(int) * (2 + an.nparts), stream));
        if (defer_ok) CUDA_OK(cudaMemsetAsync(d_xSync.p, 0, sizeof(int) * 2, stream));
        int enq = 0;
        while (true) {
            // `lookahead` iterations per convergence read-back: once the device has set `done` the kernels of the
            // remaining ones return at once
            const int batch = std::max(1, std::min(lookahead, maxit - enq));
            for (int b = 0; b < batch; ++b) run_iteration();
            enq += batch;
            CUDA_OK(cudaMemcpyAsync(h_S, d_S.p, sizeof(Scalars), cudaMemcpyDeviceToHost, stream));
            CUDA_OK(cudaStreamSynchronize(stream));
            if (verbosity > 1)
                fprintf(stderr, "[b200bda] it %.1f norm %.6e\n", 0.5 * h_S->it_half, h_S->norm);
            if (h_S->done || h_S->singular || h_S->trsv_timeout || enq >= maxit) break;
        }
        id = prof_begin(K_UNPERMUTE);
        k_scatter_solution<<<blocks_for(N, 256, num_sms * 8), 256, 0, stream>>>(d_x.p, d_y.p, d_perm.p, d_xnat.p, N, d_S.p);
        prof_end(id);
        CUDA_OK(cudaEventRecord(ev_c, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        CUDA_OK(cudaGetLastError());
        prof_collect();
        float ms_f = 0.f, ms_k = 0.f;
        CUDA_OK(cudaEventElapsedTime(&ms_f, ev_a, ev_b));
        CUDA_OK(cudaEventElapsedTime(&ms_k, ev_b, ev_c));
        const Scalars& S = *h_S;
        const double it = std::min(0.5 * S.it_half, (double) maxit);
        res->it = it;
        res->iterations = (int) it;                                   // cusparseSolverBackend.cu:172
        res->norm0 = S.norm0; res->norm = S.norm;
        res->reduction = S.norm0 > 0.0 ? S.norm / S.norm0 : 0.0;      // :173
        res->conv_rate = it > 0.0 ? std::pow(res->reduction, 1.0 / it) : 0.0;
        res->converged = S.converged;
        res->breakdown = S.breakdown;
        res->num_levels = an.nlev;
        res->t_factor = ms_f * 1e-3; res->t_krylov = ms_k * 1e-3;
        res->elapsed = wall() - t0;
        if (verbosity > 0)
            fprintf(stderr, "[b200bda] converged %d, it %.1f, reduction %.3e, factor %.3f ms, krylov %.3f ms\n", S.converged, it,
                    res->reduction, ms_f, ms_k);
        if (phase_times) {
            profile = false;
            auto d = [&](int k) { return (stats[k].ms - ms_before[k]) * 1e-3; };
            const double prec = d(K_LOWER) + d(K_UPPER) + d(K_UPPER_SPMV), spmv_s = d(K_SPMV) + d(K_SPMV_GHOST) + d(K_HALO_PUSH), well = d(K_WELL);
            const double dec = d(K_PERMUTE) + d(K_FACTOR) + d(K_SLICES);
            const double rest = d(K_VEC_P) + d(K_VEC_XR1) + d(K_VEC_XR2) + d(K_INIT) + d(K_UNPERMUTE) + d(K_ALLREDUCE) + d(K_FINISH) + d(K_MISC);
            fprintf(stderr, "b200Solver::create_preconditioner: %g s\n"
                            "b200Solver::ilu_apply:   %g s%s\n"
                            "wellContributions::apply:  %g s\n"
                            "b200Solver::spmv:        %g s\n"
                            "b200Solver::rest:        %g s\n"
                            "b200Solver::total_solve: %g s\n",
                    dec, prec, d(K_UPPER_SPMV) > 0.0 ? " (the upper sweep's launch also runs the SpMV that follows it)" : "", well, spmv_s, rest,
                    res->elapsed);
        }
        if (S.trsv_timeout) {
            // the dataflow counters (and, on several GPUs, the exchange epochs of the ranks) may be out of step now
            poisoned = true;
            throw std::runtime_error("triangular-solve dataflow wait timed out");
        }
    }

    // natural-order host vector -> p-space device vector and back (kernel-level API)
    void to_device_p(const double* h, double* d_p_out)
    {
        CUDA_OK(cudaMemcpyAsync(d_tmp1.p, h, sizeof(double) * N, cudaMemcpyHostToDevice, stream));
        k_gather_vec<<<blocks_for(N, 256, num_sms * 8), 256, 0, stream>>>(d_tmp1.p, d_perm.p, d_p_out, N);
        launch_count++;
    }
    void to_host_nat(const double* d_p_in, double* h)
    {
        k_scatter_vec<<<blocks_for(N, 256, num_sms * 8), 256, 0, stream>>>(d_p_in, d_perm.p, d_tmp1.p, N);
        launch_count++;
        CUDA_OK(cudaMemcpyAsync(h, d_tmp1.p, sizeof(double) * N, cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        CUDA_OK(cudaGetLastError());
    }
    void fill(double* d, double v)
    {
        k_fill<<<blocks_for(N, 256, num_sms * 8), 256, 0, stream>>>(d, v, N);
        launch_count++;
    }
    void ensure_factor()
    {
        if (!have_system) throw std::runtime_error("no system uploaded");
        if (!have_factor) { permute_values(); factorize(); }
    }
};

template <class F>
static b200_status guarded(F&& f, b200_status on_error = B200_UNKNOWN_ERROR)
{
    try {
        return f();
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return on_error;
    } catch (...) {
        g_last_error = "unknown exception";
        return on_error;
    }
}

}  // namespace b200

using namespace b200;

struct b200_solver : b200::Solver {};
struct b200_wells : b200::Wells {};

extern "C" {

const char* b200_last_error(void) { return g_last_error.c_str(); }
const char* b200_version(void) { return "b200bda 0.1 (sm_100a)"; }

int b200_device_available(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return 0; }
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { cudaGetLastError(); return 0; }
    return p.major == 10 ? 1 : 0;
}

b200_solver* b200_create(int verbosity, int maxit, double tolerance, unsigned int device_id)
{
    b200_solver* s = nullptr;
    try {
        s = new b200_solver();
        s->verbosity = verbosity; s->maxit = maxit; s->tolerance = tolerance; s->device = (int) device_id;
        s->init_device();
        return s;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        delete s;
        return nullptr;
    }
}

void b200_destroy(b200_solver* s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    delete s;
}

b200_status b200_set_option(b200_solver* s, const char* key, double value)
{
    return guarded([&]() -> b200_status {
        if (!s || !key) throw std::runtime_error("null argument");
        std::string k(key);
        if (k == "relaxation") s->relaxation = value;
        else if (k == "tolerance") s->tolerance = value;
        else if (k == "maxit") s->maxit = (int) value;
        else if (k == "verbosity") s->verbosity = (int) value;
        else if (k == "pin_host") s->pin_host = value != 0.0;
        else if (k == "use_graph") s->use_graph = value != 0.0;
        else if (k == "lookahead") s->lookahead = std::max(1, (int) value);
        else if (k == "profile") s->profile = value != 0.0;
        else if (k == "wells_flat") s->wells_flat = value != 0.0 ? 1 : 0;
        else if (k == "wells_cluster") s->wells_cluster = value != 0.0 ? 1 : 0;
        else if (k == "spmv_sell") { if (s->analysed) throw std::runtime_error("spmv_sell must be set before the first solve"); s->spmv_sell = std::max(0, std::min(2, (int) value)); }
        else if (k == "fuse_spmv") { if (s->analysed) throw std::runtime_error("fuse_spmv must be set before the first solve"); s->fuse_spmv = std::max(0, std::min(2, (int) value)); }
        else if (k == "fuse_debug") s->fuse_debug = (int) value;
        else if (k == "sweep_nowait") s->sweep_nowait = (int) value;
        else if (k == "s2_prefetch") s->s2_prefetch = std::max(2, std::min(8, (int) value));
        else if (k == "s2_poll_lead") s->s2_poll_lead = std::max(0, (int) value);
        else if (k == "fuse_chunk") { if (s->analysed) throw std::runtime_error("fuse_chunk must be set before the first solve"); s->fuse_chunk = std::max(1, std::min(4, (int) value)); }
        else if (k == "fuse_ring_warps") { if (s->analysed) throw std::runtime_error("fuse_ring_warps must be set before the first solve"); s->fuse_ring_warps = std::max(1, (int) value); }
        else if (k == "sweep_v2" || k == "s2_cw" || k == "s2_helpers") {
            if (s->analysed) throw std::runtime_error("the sweep schedule must be set before the first solve");
            const int v = (int) value;
            if (k == "sweep_v2") s->sweep_v2 = v != 0;
            else if (k == "s2_cw") s->s2_cw = std::max(1, v);
            else s->s2_helpers = std::max(1, v);
        }
        else if (k == "iter_pdl") { s->iter_pdl = value != 0.0; if (s->iter_graph_exec) { cudaGraphExecDestroy(s->iter_graph_exec); s->iter_graph_exec = nullptr; } }
        else if (k == "fuse_allreduce") s->fuse_allreduce = (int) value;
        else if (k == "fuse_prefetch") { s->fuse_prefetch = (int) value; if (s->iter_graph_exec) { cudaGraphExecDestroy(s->iter_graph_exec); s->iter_graph_exec = nullptr; } }
        else if (k == "tail_all_sms") { s->tail_all_sms = (int) value; if (s->iter_graph_exec) { cudaGraphExecDestroy(s->iter_graph_exec); s->iter_graph_exec = nullptr; } }
        else if (k == "sweep_autotune") s->sweep_autotune = (int) value;
        else if (k == "reorder") { if (s->analysed) throw std::runtime_error("reorder must be set before the first solve"); s->reorder = value != 0.0 ? 1 : 0; }
        else if (k == "reorder_seed") { if (s->analysed) throw std::runtime_error("reorder_seed must be set before the first solve"); s->reorder_seed = (unsigned) value; }
        else if (k == "halo_side") s->halo_side = (int) value;
        else if (k == "fac_warps") s->fac_warps = std::max(1, std::min(kFac3MaxWarps, (int) value));
        else if (k == "fac_rows3") { s->fac_rows3 = (int) value; if (s->fac_graph_exec) { cudaGraphExecDestroy(s->fac_graph_exec); s->fac_graph_exec = nullptr; } }
        else if (k == "fac_pdl") { s->fac_pdl = value != 0.0; if (s->fac_graph_exec) { cudaGraphExecDestroy(s->fac_graph_exec); s->fac_graph_exec = nullptr; } }
        else if (k == "sweep_early") { if (s->analysed) throw std::runtime_error("sweep_early must be set before the first solve"); s->sweep_early_opt = std::max(0, std::min(2, (int) value)); }
        else if (k == "defer_x") { if (s->analysed) throw std::runtime_error("defer_x must be set before the first solve"); s->defer_x = std::max(0, std::min(2, (int) value)); }
        else if (k == "fuse_unit_slices") s->fuse_unit_slices = (int) value;
        else if (k == "spmv_blocks") s->spmv_blocks_cap = std::max(1, std::min((int) value, kMaxPartials));
        else if (k == "p2p_allreduce") s->dist.use_p2p_allreduce = value != 0.0;
        else if (k == "sweep_helper_sleep") s->sweep_helper_sleep = std::max(0, (int) value);
        else if (k == "sweep_trace") {
            s->sweep_trace = value != 0.0;
            if (s->sweep_trace) { s->d_trace.alloc((size_t) 3 * 148 * 4 * b200::Solver::kTraceCap); CUDA_OK(cudaMemset(s->d_trace.p, 0, sizeof(long long) * s->d_trace.n)); }
        }
        else if (k == "sweep_parts" || k == "sweep_warps" || k == "sweep_groups" || k == "sweep_helpers" || k == "sweep_slots" || k == "sweep_stage_bytes" || k == "sweep_window" || k == "sweep_ext_window") {
            if (s->analysed) throw std::runtime_error(k + " must be set before the first solve");
            const int v = (int) value;
            if (k == "sweep_parts") s->sweep_parts = std::max(0, v);
            else if (k == "sweep_warps") s->sweep_warps = std::min(26, std::max(1, v));
            else if (k == "sweep_groups") s->sweep_groups = std::max(1, v);
            else if (k == "sweep_helpers") s->sweep_helpers = std::min(8, std::max(1, v));
            else if (k == "sweep_slots") s->sweep_slots = std::min(kSweepMaxSlots, std::max(2, v));
            else if (k == "sweep_stage_bytes") s->sweep_stage_bytes = v <= 0 ? 0 : std::max(1024, v);
            else if (k == "sweep_ext_window") s->sweep_ext_window = v;
            else s->sweep_window = v;
        }
        else throw std::runtime_error("unknown option '" + k + "'");
        return B200_SUCCESS;
    });
}

static b200_status solve_common(b200_solver* s, b200_result* res, double t_analysis, double t_copy)
{
    memset(res, 0, sizeof *res);
    s->solve_resident(res);
    res->t_analysis = t_analysis; res->t_copy = t_copy;
    if (s->h_S->singular) {
        g_last_error = "ILU0: singular or non-finite pivot block";
        return B200_CREATE_PRECONDITIONER_FAILED;
    }
    return B200_SUCCESS;
}

b200_status b200_solve_system(b200_solver* s, int N, int nnz, int dim, const double* vals, const int* rows, const int* cols,
                              const double* b, b200_wells* wells, b200_result* res)
{
    if (!s || !res) { g_last_error = "null solver / result"; return B200_UNKNOWN_ERROR; }
    const double t0 = wall();
    bool in_analysis = !s->analysed;
    b200_status st = guarded([&]() -> b200_status {
        CUDA_OK(cudaSetDevice(s->device));
        double t_an = 0.0;
        double t_copy = s->upload(N, nnz, dim, vals, rows, cols, b, wells, &t_an);
        in_analysis = false;
        if (s->verbosity >= 3) {      // cusparseSolverBackend.cu:340-344, openclSolverBackend.cpp:625-629
            if (t_an > 0.0) fprintf(stderr, "b200Solver::analyse_matrix(): %g s\n", t_an);
            fprintf(stderr, "b200Solver::copy_system_to_gpu(): %g s\n", t_copy);
        }
        b200_status r = solve_common(s, res, t_an, t_copy);
        res->elapsed = wall() - t0;
        return r;
    });
    if (st == B200_UNKNOWN_ERROR && in_analysis && s->analysed == false && dim == 3 && vals && b && rows && cols)
        return B200_ANALYSIS_FAILED;
    return st;
}

b200_status b200_upload_system(b200_solver* s, int N, int nnz, int dim, const double* vals, const int* rows, const int* cols,
                               const double* b, b200_wells* wells)
{
    if (!s) { g_last_error = "null solver"; return B200_UNKNOWN_ERROR; }
    return guarded([&]() -> b200_status {
        CUDA_OK(cudaSetDevice(s->device));
        double t_an = 0.0;
        s->upload(N, nnz, dim, vals, rows, cols, b, wells, &t_an);
        return B200_SUCCESS;
    });
}

b200_status b200_solve_resident(b200_solver* s, b200_result* res)
{
    if (!s || !res) { g_last_error = "null solver / result"; return B200_UNKNOWN_ERROR; }
    return guarded([&]() -> b200_status {
        CUDA_OK(cudaSetDevice(s->device));
        return solve_common(s, res, 0.0, 0.0);
    });
}

// Explicit page-locking of caller-owned buffers (the caller knows their lifetime; the library does not).
b200_status b200_host_register(b200_solver* s, void* ptr, size_t bytes)
{
    if (!s || !ptr || !bytes) { g_last_error = "null argument"; return B200_UNKNOWN_ERROR; }
    return guarded([&]() -> b200_status {
        CUDA_OK(cudaSetDevice(s->device));
        for (auto& r : s->host_regs) if (r.first == ptr) { if (r.second == bytes) return B200_SUCCESS; throw std::runtime_error("b200_host_register: pointer already registered with another size"); }
        cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
        if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return B200_SUCCESS; }      // pinned by someone else: nothing to own
        if (e != cudaSuccess) { cudaGetLastError(); throw CudaError(std::string("cudaHostRegister: ") + cudaGetErrorString(e)); }
        s->host_regs.emplace_back(ptr, bytes);
        return B200_SUCCESS;
    });
}
b200_status b200_host_unregister(b200_solver* s, void* ptr)
{
    if (!s || !ptr) { g_last_error = "null argument"; return B200_UNKNOWN_ERROR; }
    return guarded([&]() -> b200_status {
        CUDA_OK(cudaSetDevice(s->device));
        CUDA_OK(cudaStreamSynchronize(s->stream));
        for (size_t i = 0; i < s->host_regs.size(); ++i)
            if (s->host_regs[i].first == ptr) {
                cudaHostUnregister(ptr); cudaGetLastError();
                s->host_regs.erase(s->host_regs.begin() + i);
                return B200_SUCCESS;
            }
        // buffers the pin_host heuristic registered
        for (const void** r : {&s->reg_vals, &s->reg_b, &s->reg_x})
            if (*r == ptr) { cudaHostUnregister(ptr); cudaGetLastError(); *r = nullptr; return B200_SUCCESS; }
        for (const void** c : {&s->cand_vals, &s->cand_b, &s->cand_x}) if (*c == ptr) *c = nullptr;
        return B200_SUCCESS;     // not registered by this solver: nothing to do
    });
}

b200_status b200_get_result(b200_solver* s, double* x)
{
    if (!s || !x) { g_last_error = "null argument"; return B200_UNKNOWN_ERROR; }
    return guarded([&]() -> b200_status {
        if (!s->analysed) throw std::runtime_error("get_result before any solve");
        CUDA_OK(cudaSetDevice(s->device));
        s->maybe_register(s->reg_x, s->reg_x_bytes, s->cand_x, x, sizeof(double) * s->N);     // pageable D2H of 24 B per row costs 2 ms on C3
        CUDA_OK(cudaMemcpyAsync(x, s->d_xnat.p, sizeof(double) * s->N, cudaMemcpyDeviceToHost, s->stream));
        CUDA_OK(cudaStreamSynchronize(s->stream));
        return B200_SUCCESS;
    });
}

// ---- wells -----------------------------------------------------------------------------------------

b200_wells* b200_wells_create(const char* mode, int use_well_conn)
{
    std::string m = mode ? mode : "";
    // WellContributions.cpp:31-49
    if (m == "b200" || m == "cusparse" || m == "opencl" || m == "fpga") return new b200_wells();
    if (m == "amgcl") {
        if (!use_well_conn) { g_last_error = "Error amgcl requires --matrix-add-well-contributions=true"; return nullptr; }
        return new b200_wells();
    }
    g_last_error = "Invalid accelerator mode";
    return nullptr;
}
void b200_wells_destroy(b200_wells* w) { delete w; }

b200_status b200_wells_set_block_size(b200_wells* w, unsigned int dim, unsigned int dim_wells)
{
    return guarded([&]() -> b200_status {
        if (!w) throw std::runtime_error("null wells");
        w->dim = dim; w->dim_wells = dim_wells;
        if (dim != 3 || dim_wells != 4)
            throw std::runtime_error("WellContributions::setBlockSize error: dim and dim_wells must be equal to 3 and 4");
        return B200_SUCCESS;
    });
}
b200_status b200_wells_add_num_blocks(b200_wells* w, unsigned int n)
{
    return guarded([&]() -> b200_status {
        if (!w) throw std::runtime_error("null wells");
        if (w->allocated) throw std::runtime_error("Error cannot add more sizes after allocated in WellContributions");
        w->num_blocks += n; w->num_std_wells++;
        return B200_SUCCESS;
    });
}
b200_status b200_wells_alloc(b200_wells* w)
{
    return guarded([&]() -> b200_status {
        if (!w) throw std::runtime_error("null wells");
        if (w->num_std_wells > 0) {
            if (w->dim != 3 || w->dim_wells != 4) throw std::runtime_error("setBlockSize(3, 4) must precede alloc");
            w->val_pointers.assign(w->num_std_wells + 1, 0);
            w->Ccols.resize(w->num_blocks); w->Bcols.resize(w->num_blocks);
            w->Cnnzs.resize((size_t) w->num_blocks * 12); w->Bnnzs.resize((size_t) w->num_blocks * 12);
            w->Dnnzs.resize((size_t) w->num_std_wells * 16);
            w->allocated = true;
        }
        return B200_SUCCESS;
    });
}
b200_status b200_wells_add_matrix(b200_wells* w, b200_well_matrix type, const int* colIndices, const double* values, unsigned int val_size)
{
    return guarded([&]() -> b200_status {
        if (!w) throw std::runtime_error("null wells");
        if (!w->allocated) throw std::runtime_error("Error cannot add wellcontribution before allocating memory in WellContributions");
        if (w->num_std_wells_so_far >= w->num_std_wells) throw std::runtime_error("more wells added than announced with addNumBlocks");
        const unsigned off = w->num_blocks_so_far;
        switch (type) {
            case B200_WELL_C:
            case B200_WELL_B:
                if (off + val_size > w->num_blocks) throw std::runtime_error("more blocks added than announced with addNumBlocks");
                if (!colIndices || !values) throw std::runtime_error("null colIndices / values");
                if (type == B200_WELL_C) {
                    memcpy(w->Cnnzs.data() + (size_t) off * 12, values, sizeof(double) * val_size * 12);
                    memcpy(w->Ccols.data() + off, colIndices, sizeof(int) * val_size);
                } else {
                    memcpy(w->Bnnzs.data() + (size_t) off * 12, values, sizeof(double) * val_size * 12);
                    memcpy(w->Bcols.data() + off, colIndices, sizeof(int) * val_size);
                    w->val_pointers[w->num_std_wells_so_far] = off;
                    w->num_blocks_so_far += val_size;          // WellContributions.cpp:205-208
                    w->num_std_wells_so_far++;
                    w->val_pointers[w->num_std_wells_so_far] = w->num_blocks_so_far;
                }
                break;
            case B200_WELL_D:
                if (!values) throw std::runtime_error("null values");
                memcpy(w->Dnnzs.data() + (size_t) w->num_std_wells_so_far * 16, values, sizeof(double) * 16);
                break;
            default:
                throw std::runtime_error("Error unsupported matrix ID for WellContributions::addMatrix()");
        }
        return B200_SUCCESS;
    });
}
unsigned int b200_wells_get_num_wells(const b200_wells* w) { return w ? w->num_std_wells + (unsigned) w->ms.size() : 0; }

b200_status b200_wells_get_multisegment_inverse(const b200_wells* w, unsigned int index, double* Dinv_out)
{
    return guarded([&]() -> b200_status {
        if (!w || !Dinv_out) throw std::runtime_error("null argument");
        if (index >= w->ms.size()) throw std::runtime_error("multisegment well index out of range");
        memcpy(Dinv_out, w->ms[index].Dinv.data(), sizeof(double) * w->ms[index].Dinv.size());
        return B200_SUCCESS;
    });
}

b200_status b200_wells_add_multisegment(b200_wells* w, unsigned int dim, unsigned int dim_wells, unsigned int Mb,
                                        const double* Bvalues, const unsigned int* BcolIndices, const unsigned int* BrowPointers,
                                        unsigned int DnumBlocks, const double* Dvalues, const int* DcolPointers,
                                        const int* DrowIndices, const double* Cvalues)
{
    return guarded([&]() -> b200_status {
        if (!w) throw std::runtime_error("null wells");
        if (dim != 3 || dim_wells != 4)
            throw std::runtime_error("WellContributions::addMultisegmentWellContribution error: dim and dim_wells must be equal to 3 and 4");
        if (Mb == 0 || !Bvalues || !BcolIndices || !BrowPointers || !Dvalues || !DcolPointers || !DrowIndices || !Cvalues)
            throw std::runtime_error("addMultisegmentWellContribution: null argument or no segments");
        const unsigned M = Mb * dim_wells;
        if (BrowPointers[0] != 0) throw std::runtime_error("addMultisegmentWellContribution: BrowPointers must start at 0");
        for (unsigned r = 0; r < Mb; ++r)
            if (BrowPointers[r + 1] < BrowPointers[r]) throw std::runtime_error("addMultisegmentWellContribution: BrowPointers must ascend");
        if ((size_t) DcolPointers[M] != (size_t) DnumBlocks * dim_wells * dim_wells)
            throw std::runtime_error("addMultisegmentWellContribution: DcolPointers[M] must equal DnumBlocks * dim_wells^2");
        b200::Wells::MsWell m;
        m.Mb = Mb;
        const unsigned nB = BrowPointers[Mb];
        m.Brows.assign(BrowPointers, BrowPointers + Mb + 1);
        m.Bcols.assign(BcolIndices, BcolIndices + nB);
        m.B.assign(Bvalues, Bvalues + (size_t) nB * 12);
        m.C.assign(Cvalues, Cvalues + (size_t) nB * 12);
        b200::invert_csc((int) M, DcolPointers, DrowIndices, Dvalues, m.Dinv);
        w->ms.push_back(std::move(m));
        return B200_SUCCESS;
    });
}

// ---- multi-GPU: one process (and one b200_solver) per GPU ---------------------------------------------

b200_status b200_dist_unique_id(unsigned char* id128)
{
    return guarded([&]() -> b200_status {
        if (!id128) throw std::runtime_error("null argument");
        g_nccl.load();
        NcclId id;
        NCCL_OK(g_nccl.GetUniqueId(&id));
        memcpy(id128, id.internal, 128);
        return B200_SUCCESS;
    });
}

b200_status b200_dist_init(b200_solver* s, int rank, int world, const unsigned char* id128)
{
    return guarded([&]() -> b200_status {
        if (!s || world < 1 || rank < 0 || rank >= world) throw std::runtime_error("bad arguments");
        if (s->analysed) throw std::runtime_error("b200_dist_init must precede the first solve");
        CUDA_OK(cudaSetDevice(s->device));
        s->dist.enabled = true; s->dist.rank = rank; s->dist.world = world;
        if (world > 1) {
            if (!id128) throw std::runtime_error("null NCCL id");
            g_nccl.load();
            NcclId id;
            memcpy(id.internal, id128, 128);
            NCCL_OK(g_nccl.CommInitRank(&s->dist.comm, world, id, rank));
        }
        return B200_SUCCESS;
    });
}

b200_status b200_dist_set_halo(b200_solver* s, int n_ghost, int n_neigh, const int* neigh_rank, const int* send_ptr,
                               const int* send_rows, const int* recv_ptr, unsigned char* ipc_handle64)
{
    return guarded([&]() -> b200_status {
        if (!s || n_ghost < 0 || n_neigh < 0 || n_neigh > 64) throw std::runtime_error("bad arguments");
        if (!s->dist.enabled) throw std::runtime_error("b200_dist_init first");
        if (s->analysed) throw std::runtime_error("b200_dist_set_halo must precede the first solve");
        if (n_neigh > 0 && (!neigh_rank || !send_ptr || !send_rows || !recv_ptr)) throw std::runtime_error("null halo arrays");
        CUDA_OK(cudaSetDevice(s->device));
        Dist& D = s->dist;
        D.n_ghost = n_ghost; D.nneigh = n_neigh;
        D.neigh_rank.assign(neigh_rank, neigh_rank + n_neigh);
        D.send_ptr.assign(1, 0); D.recv_ptr.assign(1, 0);
        if (n_neigh > 0) {
            D.send_ptr.assign(send_ptr, send_ptr + n_neigh + 1);
            D.recv_ptr.assign(recv_ptr, recv_ptr + n_neigh + 1);
            D.send_rows.assign(send_rows, send_rows + send_ptr[n_neigh]);
            if (D.recv_ptr[n_neigh] != n_ghost) throw std::runtime_error("recv_ptr does not cover the ghost range");
        } else if (n_ghost != 0) throw std::runtime_error("ghost cells without neighbours");
        // receive block: flags + two parity buffers; zeroed so that epoch 0 never matches
        const size_t bytes = kHaloRecvOffset + 2 * 3 * (size_t) std::max(n_ghost, 1) * sizeof(double);
        s->d_halo.alloc(bytes);
        CUDA_OK(cudaMemset(s->d_halo.p, 0, bytes));
        s->d_push_tickets.alloc(64);
        CUDA_OK(cudaMemset(s->d_push_tickets.p, 0, 64 * sizeof(unsigned)));
        s->d_dist_ctr.alloc(4);
        CUDA_OK(cudaMemset(s->d_dist_ctr.p, 0, 4 * sizeof(unsigned)));
        D.rank_base.assign(D.world, nullptr);
        D.rank_base[D.rank] = s->d_halo.p;
        D.peers.assign(n_neigh, HaloPeerD{});
        D.have_halo = true;
        D.peers_ready = (n_neigh == 0);
        D.mail_ready = false;
        if (ipc_handle64) {
            cudaIpcMemHandle_t h;
            static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
            CUDA_OK(cudaIpcGetMemHandle(&h, s->d_halo.p));
            memcpy(ipc_handle64, &h, 64);
        }
        return B200_SUCCESS;
    });
}

b200_status b200_dist_map_rank(b200_solver* s, int rank, const unsigned char* ipc_handle64)
{
    return guarded([&]() -> b200_status {
        if (!s) throw std::runtime_error("null argument");
        Dist& D = s->dist;
        if (!D.have_halo) throw std::runtime_error("b200_dist_set_halo first");
        if (rank < 0 || rank >= D.world || D.world > 64) throw std::runtime_error("bad rank");
        CUDA_OK(cudaSetDevice(s->device));
        if (rank != D.rank && !D.rank_base[rank]) {
            if (!ipc_handle64) throw std::runtime_error("null IPC handle");
            cudaIpcMemHandle_t h;
            memcpy(&h, ipc_handle64, 64);
            void* base = nullptr;
            CUDA_OK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
            D.rank_base[rank] = base;
        }
        bool all = true;
        for (void* p : D.rank_base) all = all && p != nullptr;
        if (all) {
            for (int r = 0; r < D.world; ++r) {
                unsigned char* b = reinterpret_cast<unsigned char*>(D.rank_base[r]);
                D.mail.flags[r] = reinterpret_cast<unsigned*>(b + 256);
                D.mail.vals[r] = reinterpret_cast<double*>(b + 768);
            }
            s->d_mail.alloc(1);
            CUDA_OK(cudaMemcpy(s->d_mail.p, &D.mail, sizeof(MailD), cudaMemcpyHostToDevice));
            D.mail_ready = true;
        }
        return B200_SUCCESS;
    });
}

b200_status b200_dist_connect_peer(b200_solver* s, int neigh_index, int peer_n_ghost, int peer_recv_offset, int peer_slot)
{
    return guarded([&]() -> b200_status {
        if (!s) throw std::runtime_error("null argument");
        Dist& D = s->dist;
        if (!D.have_halo || neigh_index < 0 || neigh_index >= D.nneigh) throw std::runtime_error("bad neighbour index");
        if (peer_slot < 0 || peer_slot >= 64 || peer_recv_offset < 0 || peer_recv_offset > peer_n_ghost) throw std::runtime_error("bad peer layout");
        void* base = D.rank_base[D.neigh_rank[neigh_index]];
        if (!base) throw std::runtime_error("neighbour rank not mapped (b200_dist_map_rank)");
        CUDA_OK(cudaSetDevice(s->device));
        HaloPeerD& P = D.peers[neigh_index];
        P.flag = reinterpret_cast<unsigned*>(base) + peer_slot;
        P.recv = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(base) + kHaloRecvOffset) + 3 * (size_t) peer_recv_offset;
        P.parity_stride = 3ll * peer_n_ghost;
        P.send_begin = D.send_ptr[neigh_index];
        P.send_end = D.send_ptr[neigh_index + 1];
        bool all = true;
        for (const HaloPeerD& q : D.peers) all = all && q.recv != nullptr;
        if (all) {
            s->d_peers.alloc(D.nneigh);
            CUDA_OK(cudaMemcpy(s->d_peers.p, D.peers.data(), sizeof(HaloPeerD) * D.nneigh, cudaMemcpyHostToDevice));
            D.peers_ready = true;
        }
        return B200_SUCCESS;
    });
}

// y_owned = (A [x_owned; x_ghost])_owned with the halo exchanged over peer memory (collective: every rank calls it)
b200_status b200_dist_spmv(b200_solver* s, const double* x, double* y)
{
    return guarded([&]() -> b200_status {
        if (!s || !x || !y) throw std::runtime_error("null argument");
        CUDA_OK(cudaSetDevice(s->device));
        if (!s->dist.enabled) throw std::runtime_error("not a multi-GPU solver");
        if (!s->have_system) throw std::runtime_error("no system uploaded");
        if (!s->have_factor) s->permute_values();
        s->to_device_p(x, s->d_tmp2.p);
        s->halo_push(s->d_tmp2.p, false);
        s->spmv<0>(s->d_tmp2.p, s->d_t.p, nullptr);
        s->halo_join();
        s->spmv_ghost<0>(s->d_t.p, nullptr, false);
        s->to_host_nat(s->d_t.p, y);
        return B200_SUCCESS;
    });
}

int b200_dist_rank(const b200_solver* s) { return s ? s->dist.rank : 0; }
int b200_dist_world(const b200_solver* s) { return s ? s->dist.world : 1; }

// ---- kernel-level entry points --------------------------------------------------------------------

b200_status b200_spmv(b200_solver* s, const double* x, double* y)
{
    return guarded([&]() -> b200_status {
        if (!s || !x || !y) throw std::runtime_error("null argument");
        CUDA_OK(cudaSetDevice(s->device));
        if (!s->have_system) throw std::runtime_error("no system uploaded");
        if (!s->have_factor) s->permute_values();
        s->to_device_p(x, s->d_tmp2.p);
        s->spmv<0>(s->d_tmp2.p, s->d_t.p, nullptr);
        s->to_host_nat(s->d_t.p, y);
        return B200_SUCCESS;
    });
}

b200_status b200_well_apply(b200_solver* s, const double* x, double* y)
{
    return guarded([&]() -> b200_status {
        if (!s || !x || !y) throw std::runtime_error("null argument");
        CUDA_OK(cudaSetDevice(s->device));
        if (!s->have_system) throw std::runtime_error("no system uploaded");
        s->to_device_p(x, s->d_tmp2.p);
        s->to_device_p(y, s->d_t.p);
        s->wells_apply<0>(s->d_tmp2.p, s->d_t.p, nullptr);
        s->to_host_nat(s->d_t.p, y);
        return B200_SUCCESS;
    });
}

b200_status b200_ilu0_factorize(b200_solver* s)
{
    return guarded([&]() -> b200_status {
        if (!s) throw std::runtime_error("null argument");
        CUDA_OK(cudaSetDevice(s->device));
        if (!s->have_system) throw std::runtime_error("no system uploaded");
        s->permute_values();
        s->factorize();
        CUDA_OK(cudaMemcpyAsync(s->h_S, s->d_S.p, sizeof(Scalars), cudaMemcpyDeviceToHost, s->stream));
        CUDA_OK(cudaStreamSynchronize(s->stream));
        CUDA_OK(cudaGetLastError());
        if (s->h_S->singular) {
            g_last_error = "ILU0: singular or non-finite pivot block";
            return B200_CREATE_PRECONDITIONER_FAILED;
        }
        return B200_SUCCESS;
    });
}

static double host_sentinel()
{
    double v;
    unsigned long long u = kSentinel;
    memcpy(&v, &u, sizeof v);
    return v;
}

b200_status b200_ilu0_apply(b200_solver* s, const double* d, double* v)
{
    return guarded([&]() -> b200_status {
        if (!s || !d || !v) throw std::runtime_error("null argument");
        CUDA_OK(cudaSetDevice(s->device));
        s->ensure_factor();
        s->to_device_p(d, s->d_tmp2.p);
        s->fill(s->d_w.p, host_sentinel());
        s->fill(s->d_y.p, host_sentinel());
        s->trsv_lower(s->d_tmp2.p, s->d_w.p, false);
        s->trsv_upper(s->d_w.p, s->d_y.p, nullptr, false);
        s->to_host_nat(s->d_y.p, v);
        return B200_SUCCESS;
    });
}

b200_status b200_get_ilu0(b200_solver* s, double* lu)
{
    return guarded([&]() -> b200_status {
        if (!s || !lu) throw std::runtime_error("null argument");
        CUDA_OK(cudaSetDevice(s->device));
        s->ensure_factor();
        std::vector<double> tmp((size_t) s->nnz);     // owned x owned blocks; ghost blocks of lu are left untouched
        CUDA_OK(cudaMemcpyAsync(tmp.data(), s->d_LU.p, sizeof(double) * s->nnz, cudaMemcpyDeviceToHost, s->stream));
        CUDA_OK(cudaStreamSynchronize(s->stream));
        for (long long q = 0; q < s->nnzb; ++q)
            memcpy(lu + (size_t) s->an.srcblk[q] * 9, tmp.data() + (size_t) q * 9, 9 * sizeof(double));
        return B200_SUCCESS;
    });
}

static void export_schedule(const LevelSchedule& S, int Nb, int* to, int* from, int* rpl, int* nlev)
{
    if (to) memcpy(to, S.toOrder.data(), sizeof(int) * Nb);
    if (from) memcpy(from, S.fromOrder.data(), sizeof(int) * Nb);
    if (rpl) for (int l = 0; l < S.nlev; ++l) rpl[l] = S.levelPtr[l + 1] - S.levelPtr[l];
    if (nlev) *nlev = S.nlev;
}

b200_status b200_get_level_schedule(b200_solver* s, int* to, int* from, int* rpl, int* nlev)
{
    return guarded([&]() -> b200_status {
        if (!s || !s->analysed) throw std::runtime_error("no analysed system");
        export_schedule(s->an.sched, s->Nb, to, from, rpl, nlev);
        return B200_SUCCESS;
    });
}

b200_status b200_level_schedule_host(int Nb, const int* rows, const int* cols, int* to, int* from, int* rpl, int* nlev)
{
    return guarded([&]() -> b200_status {
        if (Nb <= 0 || !rows || !cols) throw std::runtime_error("bad arguments");
        LevelSchedule S = level_schedule(Nb, rows, cols);
        export_schedule(S, Nb, to, from, rpl, nlev);
        return B200_SUCCESS;
    }, B200_ANALYSIS_FAILED);
}

// Host-only verification of the sweep schedule (no device): analyse the pattern, fill L/U with random
// values, run the packed streams through the chunk-by-chunk emulator and compare with the sequential
// natural-order substitution.  stats: [nparts, nlines, nstrips, stagesL, stagesU, chunksL, windowDepsL, globalDepsL,
// maxMetaInts, maxValsDoubles, maxRhsRows, levels].
// pseudo-random factor in the permuted pattern (small off-diagonal blocks, well-conditioned "inverse pivots"), a right-hand side
// and the sequential natural-order solves y = L^-1 rhs, x = U^-1 y on p-space storage: the reference the host emulations of the
// sweep kernels are compared with
static void sweep_check_reference(const Analysis& A, unsigned seed, std::vector<double>& LU, std::vector<double>& rhs, std::vector<double>& yr,
                                  std::vector<double>& xr)
{
    const int Nb = A.Nb;
    LU.assign((size_t) A.nnzb * 9, 0.0);
    unsigned long long st = seed * 2654435761ull + 88172645463325252ull;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double) (st >> 11) / 9007199254740992.0 - 0.5; };
    for (int q = 0; q < Nb; ++q)
        for (int k = A.prow[q]; k < A.prow[q + 1]; ++k)
            for (int e = 0; e < 9; ++e)
                LU[(size_t) k * 9 + e] = (k == A.pdiag[q] ? (e % 4 == 0 ? 1.0 : 0.0) : 0.0) + 0.3 * rnd();
    rhs.assign((size_t) 3 * Nb + 8, 0.0); yr.assign((size_t) 3 * Nb, 0.0); xr.assign((size_t) 3 * Nb, 0.0);
    for (int i = 0; i < 3 * Nb; ++i) rhs[i] = rnd();
    for (int r = 0; r < Nb; ++r) {
        const int q = A.iperm[r];
        double acc[3] = {rhs[3 * q], rhs[3 * q + 1], rhs[3 * q + 2]};
        for (int k = A.prow[q]; k < A.pdiag[q]; ++k)
            for (int c = 0; c < 3; ++c)
                for (int e = 0; e < 3; ++e) acc[c] -= LU[(size_t) k * 9 + c * 3 + e] * yr[3 * (size_t) A.pcol[k] + e];
        for (int c = 0; c < 3; ++c) yr[3 * (size_t) q + c] = acc[c];
    }
    for (int r = Nb - 1; r >= 0; --r) {
        const int q = A.iperm[r];
        double acc[3] = {yr[3 * q], yr[3 * q + 1], yr[3 * q + 2]};
        for (int k = A.pdiag[q] + 1; k < A.prow[q + 1]; ++k)
            for (int c = 0; c < 3; ++c)
                for (int e = 0; e < 3; ++e) acc[c] -= LU[(size_t) k * 9 + c * 3 + e] * xr[3 * (size_t) A.pcol[k] + e];
        const double* d = LU.data() + (size_t) A.pdiag[q] * 9;
        for (int c = 0; c < 3; ++c) xr[3 * (size_t) q + c] = d[c * 3] * acc[0] + d[c * 3 + 1] * acc[1] + d[c * 3 + 2] * acc[2];
    }
}

b200_status b200_sweep_schedule_check_host(int Nb, const int* rows, const int* cols, int parts, int stage_bytes, int window,
                                           unsigned seed, double* max_rel_err, long long* stats)
{
    return guarded([&]() -> b200_status {
        if (Nb <= 0 || !rows || !cols) throw std::runtime_error("bad arguments");
        AnalysisOptions opt;
        if (parts > 0) opt.parts = parts;
        if (stage_bytes > 0) opt.stageBytes = stage_bytes;
        if (window > 0) opt.window = window;
        Analysis A = analyse(Nb, rows, cols, opt);
        std::vector<double> LU, rhs, yr, xr, y((size_t) 3 * Nb + 8), x((size_t) 3 * Nb + 8);
        sweep_check_reference(A, seed, LU, rhs, yr, xr);
        std::vector<double> vL, vU;
        fill_stream_host(A.L, true, LU.data(), 1.0, vL);
        fill_stream_host(A.U, false, LU.data(), 1.0, vU);
        if (!emulate_sweep(A, A.L, true, vL, rhs.data(), y.data())) throw std::runtime_error("lower sweep schedule deadlocks");
        if (!emulate_sweep(A, A.U, false, vU, y.data(), x.data())) throw std::runtime_error("upper sweep schedule deadlocks");
        double num = 0.0, den = 0.0;
        for (int i = 0; i < 3 * Nb; ++i) {
            num = std::max(num, std::fabs(x[i] - xr[i]) + std::fabs(y[i] - yr[i]));
            den = std::max(den, std::fabs(xr[i]));
        }
        if (max_rel_err) *max_rel_err = den > 0.0 ? num / den : num;
        if (stats) {
            const long long v[12] = {A.nparts, A.nlines, A.nstrips, (long long) A.L.stages.size(), (long long) A.U.stages.size(), A.L.nchunks,
                                     A.L.nWindow, A.L.nExternal, std::max(A.L.maxMetaInts, A.U.maxMetaInts),
                                     std::max(A.L.maxValsDoubles, A.U.maxValsDoubles), std::max(A.L.maxRhsRows, A.U.maxRhsRows), A.nlev};
            memcpy(stats, v, sizeof v);
        }
        return B200_SUCCESS;
    }, B200_ANALYSIS_FAILED);
}

// The same check for the round-2 schedule (sweep2.hpp): builds the record streams for `parts` parts, `consumer_warps` consumer
// warps (groups x group width per part; 0 = the library's defaults), `helpers` helper warps, fills them from a pseudo-random
// factor and runs the host emulation of k_sweep2 against the sequential natural-order solves.
// stats (12): parts, lines, strips, records L, records U, empty records L, window deps L, external deps L, external rows L,
// helper blocks L, max groups, max group width.
// Opt-in colour ordering: the colouring alone, on the host (no device needed)
b200_status b200_graph_coloring_host(int Nb, const int* rows, const int* cols, unsigned seed, int* to_order, int* from_order,
                                     int* rows_per_color, int* ncolors)
{
    return guarded([&]() -> b200_status {
        if (Nb <= 0 || !rows || !cols || !to_order || !from_order || !ncolors) throw std::runtime_error("null argument");
        const b200::ColorOrder o = b200::graph_coloring(Nb, rows, cols, seed);
        std::copy(o.toOrder.begin(), o.toOrder.end(), to_order);
        std::copy(o.fromOrder.begin(), o.fromOrder.end(), from_order);
        if (rows_per_color) std::copy(o.rowsPerColor.begin(), o.rowsPerColor.end(), rows_per_color);
        *ncolors = o.ncolors;
        return B200_SUCCESS;
    });
}

// the ordering a solver with option reorder = 1 works in (after its first solve): to_order / from_order as above
b200_status b200_get_reorder(b200_solver* s, int* to_order, int* from_order, int* ncolors)
{
    return guarded([&]() -> b200_status {
        if (!s || !ncolors) throw std::runtime_error("null argument");
        if (!s->analysed || !s->level_sweeps) throw std::runtime_error("no colour ordering: set option reorder = 1 before the first solve");
        if (to_order) std::copy(s->color.toOrder.begin(), s->color.toOrder.end(), to_order);
        if (from_order) std::copy(s->color.fromOrder.begin(), s->color.fromOrder.end(), from_order);
        *ncolors = s->color.ncolors;
        return B200_SUCCESS;
    });
}

b200_status b200_sweep2_schedule_check_host(int Nb, const int* rows, const int* cols, int parts, int window, int ext_window,
                                            int consumer_warps, int helpers, int groups, int wg, unsigned seed, double relax,
                                            double* max_rel_err, long long* stats)
{
    return guarded([&]() -> b200_status {
        if (Nb <= 0 || !rows || !cols) throw std::runtime_error("bad arguments");
        AnalysisOptions opt;
        if (parts > 0) opt.parts = parts;
        if (window > 0) opt.window = window;
        (void) ext_window;
        opt.buildStreams = false;
        Analysis A = analyse(Nb, rows, cols, opt);
        Sweep2Options o2;
        if (consumer_warps > 0) o2.consumerWarps = consumer_warps;
        (void) helpers;
        (void) groups; (void) wg;
        Sweep2Plan L, U;
        build_sweep2_plans(A, rows, cols, o2, L, U);
        std::vector<double> LU, rhs, yr, xr, y((size_t) 3 * Nb + 8), x((size_t) 3 * Nb + 8);
        sweep_check_reference(A, seed, LU, rhs, yr, xr);
        std::vector<double> vL, vU;
        fill_stream2_host(L, true, LU.data(), 1.0, vL);
        fill_stream2_host(U, false, LU.data(), relax, vU);
        if (!emulate_sweep2(A, L, true, vL, rhs.data(), y.data())) throw std::runtime_error("lower sweep schedule deadlocks");
        if (!emulate_sweep2(A, U, false, vU, y.data(), x.data())) throw std::runtime_error("upper sweep schedule deadlocks");
        double num = 0.0, den = 0.0;
        for (int i = 0; i < 3 * Nb; ++i) {
            num = std::max(num, std::fabs(x[i] - relax * xr[i]) + std::fabs(y[i] - yr[i]));
            den = std::max(den, std::fabs(xr[i]));
        }
        if (max_rel_err) *max_rel_err = den > 0.0 ? num / den : num;
        if (stats) {
            // part pairs that read each other's rows in the same sweep (a ping-pong: every level pays the hand-over latency)
            std::sort(L.partEdges.begin(), L.partEdges.end());
            L.partEdges.erase(std::unique(L.partEdges.begin(), L.partEdges.end()), L.partEdges.end());
            long long mutual = 0;
            for (auto& e : L.partEdges) if (e.first < e.second && std::binary_search(L.partEdges.begin(), L.partEdges.end(), std::make_pair(e.second, e.first))) ++mutual;
            const long long v[12] = {A.nparts, A.nlines, A.nstrips, L.nrecords, U.nrecords, L.nmulti, L.nWindow, L.nExternal, L.nMultiLaneRecords,
                                     mutual, L.maxChunks, L.nLanes};
            memcpy(stats, v, sizeof v);
        }
        return B200_SUCCESS;
    }, B200_ANALYSIS_FAILED);
}

// Debugging aid: stage timeline of the last traced sweep (option "sweep_trace" = 1): per part (148) and stage
// (first 1024) four SM-clock stamps {consumer starts waiting, data landed, stage done, producer issued}.
b200_status b200_get_sweep_trace(b200_solver* s, long long* out, long long count)
{
    return guarded([&]() -> b200_status {
        if (!s || !out || !s->d_trace.p) throw std::runtime_error("no trace (set option sweep_trace first)");
        CUDA_OK(cudaSetDevice(s->device));
        CUDA_OK(cudaStreamSynchronize(s->stream));
        CUDA_OK(cudaMemcpy(out, s->d_trace.p, sizeof(long long) * std::min<size_t>((size_t) count, s->d_trace.n), cudaMemcpyDeviceToHost));
        return B200_SUCCESS;
    });
}

// Host-only replay of the factorisation plan (no device): analyse the pattern, run the elimination plan of
// k_ilu_factor_plan row by row in level order with the same arithmetic, return LU in the caller's pattern
// (inverse pivot in the diagonal slot).  Lets the CPU test-suite check the plan against the oracle's ILU0.
b200_status b200_factor_plan_check_host(int Nb, const int* rows, const int* cols, const double* vals, double* lu_out, int* max_row,
                                        int* max_ops)
{
    return guarded([&]() -> b200_status {
        if (Nb <= 0 || !rows || !cols || !vals || !lu_out) throw std::runtime_error("bad arguments");
        Analysis A = analyse(Nb, rows, cols, AnalysisOptions());
        const long long nnzb = rows[Nb];
        std::vector<double> LU((size_t) nnzb * 9);
        for (long long q = 0; q < nnzb; ++q) memcpy(LU.data() + q * 9, vals + (size_t) A.srcblk[q] * 9, 72);
        auto inv3h = [](const double* m, double* inv) {
            double t4 = m[0] * m[4], t6 = m[0] * m[5], t8 = m[1] * m[3], t10 = m[2] * m[3], t12 = m[1] * m[6], t14 = m[2] * m[6];
            double det = t4 * m[8] - t6 * m[7] - t8 * m[8] + t10 * m[7] + t12 * m[5] - t14 * m[4];
            double t17 = 1.0 / det;
            inv[0] = (m[4] * m[8] - m[5] * m[7]) * t17; inv[1] = -(m[1] * m[8] - m[2] * m[7]) * t17; inv[2] = (m[1] * m[5] - m[2] * m[4]) * t17;
            inv[3] = -(m[3] * m[8] - m[5] * m[6]) * t17; inv[4] = (m[0] * m[8] - t14) * t17; inv[5] = -(t6 - t10) * t17;
            inv[6] = (m[3] * m[7] - m[4] * m[6]) * t17; inv[7] = -(m[0] * m[7] - t12) * t17; inv[8] = (t4 - t8) * t17;
            return det != 0.0 && std::isfinite(det);
        };
        for (int l = 0; l < A.nflev; ++l)
            for (int t = A.flevPtr[l]; t < A.flevPtr[l + 1]; ++t) {
                const int i = A.flevRows[t];
                double* row = LU.data() + (size_t) A.prow[i] * 9;
                int o = A.facPtr[i];
                const int oe = A.facPtr[i + 1];
                while (o < oe) {
                    const int code = A.facOps[2 * o + 1], toff = code & 255, nupd = code >> 8;
                    const double* d = LU.data() + (size_t) A.facOps[2 * o] * 9;
                    double lij[9];
                    for (int r = 0; r < 3; ++r)
                        for (int c = 0; c < 3; ++c) lij[3 * r + c] = row[toff * 9 + 3 * r] * d[c] + row[toff * 9 + 3 * r + 1] * d[3 + c] + row[toff * 9 + 3 * r + 2] * d[6 + c];
                    memcpy(row + toff * 9, lij, 72);
                    for (int u = 0; u < nupd; ++u) {
                        const double* uu = LU.data() + (size_t) A.facOps[2 * (o + 1 + u)] * 9;
                        const int tgt = -(A.facOps[2 * (o + 1 + u) + 1] + 1);
                        for (int r = 0; r < 3; ++r)
                            for (int c = 0; c < 3; ++c) row[tgt * 9 + 3 * r + c] -= lij[3 * r] * uu[c] + lij[3 * r + 1] * uu[3 + c] + lij[3 * r + 2] * uu[6 + c];
                    }
                    o += 1 + nupd;
                }
                double inv[9];
                double* dg = LU.data() + (size_t) A.pdiag[i] * 9;
                if (!inv3h(dg, inv)) { g_last_error = "ILU0: singular or non-finite pivot block"; return B200_CREATE_PRECONDITIONER_FAILED; }
                memcpy(dg, inv, 72);
            }
        for (long long q = 0; q < nnzb; ++q) memcpy(lu_out + (size_t) A.srcblk[q] * 9, LU.data() + q * 9, 72);
        if (max_row) *max_row = A.facMaxRow;
        if (max_ops) *max_ops = A.facMaxOps;
        return B200_SUCCESS;
    }, B200_ANALYSIS_FAILED);
}

static int kind_of(const std::string& k)
{
    if (k == "ilu_apply") return -2;
    for (int i = 0; i < K_COUNT; ++i) if (k == kKindNames[i]) return i;
    return -1;
}

b200_status b200_time_kernel(b200_solver* s, const char* which, int reps, int flush_l2, double* ms_out, double* bytes_out)
{
    return guarded([&]() -> b200_status {
        if (!s || !which || reps <= 0) throw std::runtime_error("bad arguments");
        CUDA_OK(cudaSetDevice(s->device));
        const int kind = kind_of(which);
        if (kind == -1) throw std::runtime_error(std::string("unknown kernel '") + which + "'");
        s->ensure_factor();
        const int N = s->N;
        const long long flush_n = 64ll << 20;                    // 512 MB > 126 MB L2
        if (flush_l2) s->d_flush.alloc(flush_n);
        // operands: any finite data will do for timing; r as input vector
        CUDA_OK(cudaMemcpyAsync(s->d_tmp2.p, s->d_bstage.p, sizeof(double) * N, cudaMemcpyDeviceToDevice, s->stream));
        Scalars hs; memset(&hs, 0, sizeof hs);
        hs.rho = hs.rho_new = hs.alpha = hs.omega = hs.h = hs.tr = hs.tt = 1.0; hs.norm0 = 1.0; hs.tol = 0.0; hs.max_half = 1 << 30;
        const bool saved_profile = s->profile;
        s->profile = false;
        double total_ms = 0.0;
        for (int it = -2; it < reps; ++it) {                      // two warm-up launches
            CUDA_OK(cudaMemcpyAsync(s->d_S.p, &hs, sizeof hs, cudaMemcpyHostToDevice, s->stream));
            if (kind == K_LOWER || kind == -2) s->fill(s->d_w.p, host_sentinel());
            if (kind == K_UPPER || kind == -2 || kind == K_UPPER_SPMV) s->fill(s->d_y.p, host_sentinel());
            if (kind == K_UPPER_SPMV && !s->fused_now()) throw std::runtime_error("the fused upper sweep + SpMV is not active for this system");
            if (kind == K_UPPER || kind == K_UPPER_SPMV) CUDA_OK(cudaMemcpyAsync(s->d_w.p, s->d_tmp2.p, sizeof(double) * N, cudaMemcpyDeviceToDevice, s->stream));
            if (kind == K_VEC_XR1 || kind == K_VEC_XR2 || kind == K_WELL || kind == K_SPMV)
                CUDA_OK(cudaMemcpyAsync(s->d_y.p, s->d_tmp2.p, sizeof(double) * N, cudaMemcpyDeviceToDevice, s->stream));
            if (flush_l2) {
                k_flush_l2<<<s->num_sms * 8, 256, 0, s->stream>>>(s->d_flush.p, flush_n, (double) it);
                s->launch_count++;
            }
            CUDA_OK(cudaEventRecord(s->ev_c, s->stream));
            switch (kind) {
                case K_SPMV: s->spmv<0>(s->d_y.p, s->d_t.p, nullptr); break;
                case K_LOWER: s->trsv_lower(s->d_tmp2.p, s->d_w.p, false); break;
                case K_UPPER: s->trsv_upper(s->d_w.p, s->d_y.p, nullptr, false); break;
                case K_UPPER_SPMV: s->trsv_upper_spmv<1>(s->d_w.p, s->d_y.p, s->d_w.p, s->d_t.p, s->d_tmp2.p); break;
                case -2: s->trsv_lower(s->d_tmp2.p, s->d_w.p, false); s->trsv_upper(s->d_w.p, s->d_y.p, nullptr, false); break;
                case K_FACTOR: s->factorize(); break;
                case K_PERMUTE: s->permute_values(); break;
                case K_VEC_P: s->stats[K_VEC_P].launches++; s->launch_count++;
                    k_vec_p<<<s->vec_blocks, kVecThreads, 0, s->stream>>>(s->d_tmp2.p, s->d_p.p, s->d_v.p, N, s->d_S.p); break;
                case K_VEC_XR1: s->stats[K_VEC_XR1].launches++; s->launch_count++;
                    k_vec_xr1<<<s->vec_blocks, kVecThreads, 0, s->stream>>>(s->d_x.p, s->d_y.p, s->d_r.p, s->d_v.p, N, s->d_S.p, s->d_partials.p, s->d_ticket.p, 0, 0, DistRedD{}); break;
                case K_VEC_XR2: s->stats[K_VEC_XR2].launches++; s->launch_count++;
                    k_vec_xr2<<<s->vec_blocks, kVecThreads, 0, s->stream>>>(s->d_x.p, s->d_y.p, s->d_r.p, s->d_t.p, s->d_rt.p, N, s->d_S.p, s->d_partials.p, s->d_ticket.p, 0, 0, DistRedD{}); break;
                case K_WELL: s->wells_apply<0>(s->d_y.p, s->d_t.p, nullptr); break;
                default: throw std::runtime_error(std::string("kernel '") + which + "' cannot be timed in isolation");
            }
            CUDA_OK(cudaEventRecord(s->ev_d, s->stream));
            CUDA_OK(cudaEventSynchronize(s->ev_d));
            float ms = 0.f;
            CUDA_OK(cudaEventElapsedTime(&ms, s->ev_c, s->ev_d));
            if (it >= 0) total_ms += ms;
        }
        CUDA_OK(cudaGetLastError());
        s->profile = saved_profile;
        s->have_factor = (kind == K_PERMUTE) ? false : s->have_factor;   // A refreshed, LU stale only if values changed (they did not)
        s->have_factor = true;
        if (ms_out) *ms_out = total_ms / reps;
        if (bytes_out) *bytes_out = kind == -2 ? s->alg_bytes(K_LOWER) + s->alg_bytes(K_UPPER) : s->alg_bytes(kind);
        return B200_SUCCESS;
    });
}

b200_status b200_kernel_stats(b200_solver* s, const char* which, long long* launches, double* total_ms, double* bytes)
{
    return guarded([&]() -> b200_status {
        if (!s || !which) throw std::runtime_error("bad arguments");
        const int kind = kind_of(which);
        if (kind < 0) throw std::runtime_error(std::string("unknown kernel '") + which + "'");
        if (launches) *launches = s->stats[kind].launches;
        if (total_ms) *total_ms = s->stats[kind].ms;
        if (bytes) *bytes = s->alg_bytes(kind);
        return B200_SUCCESS;
    });
}

void b200_reset_stats(b200_solver* s)
{
    if (!s) return;
    for (auto& k : s->stats) k = KStat();
    s->launch_count = 0;
}

long long b200_launch_count(b200_solver* s) { return s ? s->launch_count : 0; }

b200_status b200_timer_start(b200_solver* s)
{
    return guarded([&]() -> b200_status {
        if (!s) throw std::runtime_error("null solver");
        CUDA_OK(cudaSetDevice(s->device));
        if (!s->ev_t0) { CUDA_OK(cudaEventCreate(&s->ev_t0)); CUDA_OK(cudaEventCreate(&s->ev_t1)); }
        CUDA_OK(cudaStreamSynchronize(s->stream));
        CUDA_OK(cudaEventRecord(s->ev_t0, s->stream));
        return B200_SUCCESS;
    });
}

b200_status b200_timer_stop(b200_solver* s, double* ms)
{
    return guarded([&]() -> b200_status {
        if (!s || !ms || !s->ev_t0) throw std::runtime_error("timer not started");
        CUDA_OK(cudaEventRecord(s->ev_t1, s->stream));
        CUDA_OK(cudaEventSynchronize(s->ev_t1));
        float f = 0.f;
        CUDA_OK(cudaEventElapsedTime(&f, s->ev_t0, s->ev_t1));
        *ms = f;
        return B200_SUCCESS;
    });
}

}  // extern "C"
