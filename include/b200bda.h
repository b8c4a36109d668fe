/*
 * b200bda.h -- C ABI of libb200bda.so, the B200-native (sm_100a) ILU0-BiCGSTAB backend for
 * OPM Flow's per-Newton-step linear solve (3x3-block BSR, fp64) plus the standard- and multisegment-well apply.
 *
 * Drop-in boundary.  The entry points below are exactly what a `bda::BdaSolver<3>` subclass and
 * an `Opm::WellContributions`-compatible container need; each one names the reference interface
 * it replaces (paths relative to the reference tree, opm/simulators/linalg/bda/).  The C++ shim
 * that binds them (b200SolverBackend<block_size>) is in opm-autodiff_b200/hostcpp/ and the two
 * string-chain patches a maintainer adds to the reference are shown in INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes, no CUDA / torch types.  All `*_host` pointers are host
 * memory owned by the caller and not retained after the call returns (page-locking of `vals`/`b`/`x`
 * happens only on request: b200_host_register, or the "pin_host" option).  Functions return a
 * b200_status; on anything but B200_SUCCESS b200_last_error() describes the failure.  There is
 * no CPU fallback anywhere in the library: without a usable sm_100 device every entry point that
 * computes fails loudly.
 */
#ifndef B200BDA_H
#define B200BDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* bda::SolverStatus, BdaSolver.hpp:32-37 (same order, same meaning). */
typedef enum {
    B200_SUCCESS = 0,                       /* BDA_SOLVER_SUCCESS */
    B200_ANALYSIS_FAILED = 1,               /* BDA_SOLVER_ANALYSIS_FAILED */
    B200_CREATE_PRECONDITIONER_FAILED = 2,  /* BDA_SOLVER_CREATE_PRECONDITIONER_FAILED */
    B200_UNKNOWN_ERROR = 3                  /* BDA_SOLVER_UNKNOWN_ERROR (also: bad arguments, CUDA errors) */
} b200_status;

/* bda::BdaResult, BdaResult.hpp:28-40 (first five fields, same meaning), followed by extras. */
typedef struct {
    int    iterations;   /* (int) min(it, maxit), as cusparseSolverBackend.cu:172 */
    double reduction;    /* norm / norm0 */
    int    converged;    /* norm < tolerance * norm0 reached within maxit */
    double conv_rate;    /* reduction^(1/it) */
    double elapsed;      /* seconds, whole solve_system call */
    /* extras (not in BdaResult) */
    double it;           /* Dune/cusparse half-step counter at exit: 0.5, 1.0, 1.5, ... */
    double norm0;
    double norm;
    int    breakdown;    /* 1 if a Dune SolverAbort guard (|rho|,|omega|,|h| tiny) fired */
    int    num_levels;   /* level sets of the ILU0 dependency DAG */
    double t_analysis;   /* seconds: sparsity analysis (first call only, else 0) */
    double t_copy;       /* seconds: host->device upload (GPU time) */
    double t_factor;     /* seconds: permutation + ILU0 factorisation (GPU time) */
    double t_krylov;     /* seconds: BiCGSTAB loop (GPU time) */
} b200_result;

typedef struct b200_solver b200_solver;   /* one per BdaSolver<3> object */
typedef struct b200_wells  b200_wells;    /* one per Opm::WellContributions object */

/* ---- solver: replaces cusparseSolverBackend<3> behind BdaSolver<3> ------------------------ */

/* BdaSolver ctor (BdaSolver.hpp:78): verbosity, maxit, tolerance, deviceID.  Selects the device
 * (cusparseSolverBackend.cu:201) and creates the stream.  NULL on failure. */
b200_solver* b200_create(int verbosity, int maxit, double tolerance, unsigned int device_id);

/* ~cusparseSolverBackend / finalize (cusparseSolverBackend.cu:54-57,249-279). */
void b200_destroy(b200_solver* s);

/* Options the reference fixes at compile time or takes from the FlexibleSolver JSON tree:
 *   "relaxation"  ILU0 relaxation w (setupPropertyTree.cpp:175-188; cusparse uses 1.0)  default 1.0
 *   "tolerance", "maxit", "verbosity"   override the constructor values
 *   "pin_host"    1: cudaHostRegister the caller's vals / b / x once they are seen on two consecutive calls.
 *                 Contract: a buffer handed in stays allocated until another one replaces it, b200_host_unregister releases
 *                 it, or the solver is destroyed (Flow's matrix, rhs and solution storage live as long as the simulator).
 *                 Off by default -- the library does not own that memory; see also b200_host_register.            default 0
 *   "wells_flat"  1: standard wells that fit one CTA (<= 1024 perforations, <= 128 wells) use the apply kernel with
 *                 host-resolved index chains (k_wells_flat); 0: always the general kernel                 default 1
 *   "wells_cluster"  1: the flat well apply runs on a thread-block cluster of 8 CTAs (k_wells_cluster)        default 1
 *   "use_graph"   1: replay the factorisation launches and the BiCGSTAB iteration body as CUDA graphs  default 1
 *   "lookahead"   iterations enqueued per convergence read-back (the device stops by itself)      default 2
 *   "profile"     1: time every kernel with CUDA events (no graphs), see b200_kernel_stats         default 0
 *   "sweep_parts", "sweep_warps", "sweep_groups", "sweep_helpers", "sweep_slots", "sweep_stage_bytes", "sweep_window",
 *   "sweep_ext_window"   schedule of the triangular sweeps: parts (CTAs, all resident), consumer warps per CTA and their
 *                 level groups, helper warps, ring slots and bytes per stage (0 = auto: a quarter of a part's factor bytes,
 *                 16 KB .. 80 KB), rows of the shared-memory window (0 = auto)
 *                 and of the external-row ring; set before the first solve
 *   "sweep_helper_sleep"  ns a helper warp sleeps between two polls of external rows (measured: 0 is 2.5 % faster than 60
 *                 on 44 k and 110 k rows and no slower on 1 M)                                       default 0
 *   "sweep_trace" 1: record the stage timeline of the sweeps (b200_get_sweep_trace; debugging)     default 0
 *   "spmv_blocks" upper bound of the SpMV grid                                                    default 16 per SM
 *   Size-dependent features, set before the first solve: 0 = off, 1 = automatic (on from 100 000 block rows), 2 = on     default 1
 *   "spmv_sell"   SpMV from a sliced-ELL copy of A (a lane per block row, coalesced value loads) instead of the BSR kernel
 *   "fuse_spmv"   the upper sweep's CTAs run the SpMV that follows it as their parts finish (needs spmv_sell)
 *   "defer_x"     x += alpha y / omega y are applied by idle CTAs of the next lower sweep instead of the vector kernels
 *                 (automatic = on at every size)
 *   "sweep_early" sweep helper warps fetch a stage's external rows one stage ahead
 *   "p2p_allreduce"  multi-GPU: 1 peer-memory mailboxes, 0 NCCL + finish kernel                    default 1
 *   "fuse_allreduce" multi-GPU: the mailbox exchange runs inside the kernel that completes the local sums      default 1
 *   "halo_side"   multi-GPU: the halo push runs on a side stream beside the well apply (joined before k_spmv_ghost)  default 1
 *   "reorder"     0: natural order, the level sets of the reference's level scheduling (default); 1: graph colouring (see
 *                 b200_graph_coloring_host / b200_get_reorder), opt-in, set before the first solve.  "reorder_seed": its seed  default 0, 1
 *   Round-2 sweeps and factorisation (set before the first solve):
 *   "sweep_v2"    1: k_sweep2 (a lane per block row, operands from global memory into registers); 0: the round-1 kernel  default 1
 *   "sweep_autotune"  1: below the size where every SM gets a part, the analysis times a lower + upper sweep for a few part
 *                 counts around the rule's value and keeps the fastest (once per pattern); 0: the rule's value -- part
 *                 counts differ in rounding only, so pin it for bit-reproducible runs across processes             default 1
 *   "s2_cw", "s2_poll_lead", "s2_prefetch"   consumer warps per part (<= 15), steps before its own at which a warp starts
 *                 to poll rows of other parts, records ahead pulled into L2                              default 15, 15, 2
 *   "tail_all_sms"  sweeps with a tail (SpMV, x update) launch one CTA per SM even when the schedule has fewer parts; the CTAs
 *                 beyond the parts only work on the tail.  0 never, 1 when the parts take at most half of the SMs, 2 always  default 1
 *   "fuse_ring_warps", "fuse_chunk"  ring-fed warps and slots per chunk of the SpMV tail                       default 8, 4
 *   "fac_rows3"   1: factorisation with three rows per warp (k_ilu_factor_plan3), 0: one row per warp         default 1
 *   "fac_warps"   warps per CTA of k_ilu_factor_plan3                                                      default 4
 *   "fac_pdl"     level launches of the factorisation chained by programmatic dependent launch               default 1
 * Unknown keys return B200_UNKNOWN_ERROR. */
b200_status b200_set_option(b200_solver* s, const char* key, double value);

/* BdaSolver<3>::solve_system (BdaSolver.hpp:86-88; cusparseSolverBackend.cu:480-499).
 *   N    scalar rows (= 3*Nb), nnz scalar nonzeros (= 9*nnzb), dim = 3
 *   vals_host  nnz doubles, row-major 3x3 blocks in BSR order (BdaBridge.cpp:231-232)
 *   rows_host  Nb+1 ints, cols_host nnzb ints (0-based, ascending per row, diagonal present;
 *              BdaBridge.cpp:167-189); read on the first call only (sparsity is fixed after it,
 *              cusparseSolverBackend.cu:312)
 *   b_host     N doubles; the initial guess is 0 (cusparseSolverBackend.cu:301)
 *   wells      may be NULL or hold zero wells
 * Synchronous: returns after the solve finished on the device (cusparseSolverBackend.cu:456).
 * Non-convergence is not an error: status SUCCESS with res->converged == 0. */
b200_status b200_solve_system(b200_solver* s, int N, int nnz, int dim,
                              const double* vals_host, const int* rows_host, const int* cols_host,
                              const double* b_host, b200_wells* wells, b200_result* res);

/* BdaSolver<3>::get_result (BdaSolver.hpp:90; cusparseSolverBackend.cu:464-475): N doubles D2H.
 * Always safe to call after a solve (the reference tests call it unconditionally). */
b200_status b200_get_result(b200_solver* s, double* x_host);

/* Page-lock (cudaHostRegister) a caller-owned host buffer -- the matrix values, the right-hand side, the solution vector --
 * so that the copies of b200_solve_system / b200_get_result run at PCIe speed (55 GB/s instead of ~10 GB/s pageable).  The
 * caller, who knows the buffer's lifetime, must call b200_host_unregister before freeing it (b200_destroy releases what is
 * left).  No reference counterpart: the reference copies from pageable memory (cusparseSolverBackend.cu:286-299). */
b200_status b200_host_register(b200_solver* s, void* ptr, size_t bytes);
b200_status b200_host_unregister(b200_solver* s, void* ptr);

/* Same solve with the system already resident in HBM (uploaded by the last b200_solve_system or
 * b200_upload_system): repeats permutation + ILU0 + BiCGSTAB without any host<->device copy of
 * the matrix.  This is the device-resident leg the benchmark's `value` is measured on. */
b200_status b200_upload_system(b200_solver* s, int N, int nnz, int dim,
                               const double* vals_host, const int* rows_host, const int* cols_host,
                               const double* b_host, b200_wells* wells);
b200_status b200_solve_resident(b200_solver* s, b200_result* res);

const char* b200_last_error(void);

/* ---- wells (standard and multisegment): replaces Opm::WellContributions (WellContributions.hpp:60-214) -------- */

typedef enum { B200_WELL_C = 0, B200_WELL_D = 1, B200_WELL_B = 2 } b200_well_matrix; /* MatrixType, :69-73 */

/* WellContributions ctor (WellContributions.cpp:31-49): accepts "b200" (and the reference's own
 * mode strings); anything else fails with "Invalid accelerator mode". */
b200_wells* b200_wells_create(const char* accelerator_mode, int use_well_conn);
void        b200_wells_destroy(b200_wells* w);
/* setBlockSize (WellContributions.cpp:215-225): fails unless dim == 3 and dim_wells == 4. */
b200_status b200_wells_set_block_size(b200_wells* w, unsigned int dim, unsigned int dim_wells);
/* addNumBlocks (:227-234): fails after alloc. */
b200_status b200_wells_add_num_blocks(b200_wells* w, unsigned int num_blocks);
/* alloc (:236-259). */
b200_status b200_wells_alloc(b200_wells* w);
/* addMatrix (:152-213): per well C, then D, then B; fails before alloc.  col_indices ignored for D. */
b200_status b200_wells_add_matrix(b200_wells* w, b200_well_matrix type, const int* col_indices,
                                  const double* values, unsigned int val_size);
unsigned int b200_wells_get_num_wells(const b200_wells* w);   /* getNumWells: standard + multisegment (WellContributions.hpp:164-166) */

/* WellContributions::addMultisegmentWellContribution (WellContributions.hpp:195-213, .cpp:261-271) and the constructor of
 * MultisegmentWellContribution (MultisegmentWellContribution.cpp:32-58): B and C in blocked CSR (Mb block rows = segments,
 * dim_wells x dim blocks, row-major [well eq][cell eq], C on B's pattern), D as the scalar CSC matrix handed to UMFPACK
 * (DcolPointers[dim_wells * Mb + 1], DrowIndices / Dvalues [DnumBlocks * dim_wells^2]).  May be called at any time before
 * the solve, independently of the standard wells' three phases.  Where the reference factorises D with UMFPACK on the host
 * and then solves on the host inside every operator apply (x D2H, y H2D: WellContributions.cu:167-187), this library
 * inverts D here (dense, partial pivoting) and applies y -= C^T (D^-1 (B x)) on the device.  Fails unless dim == 3 and
 * dim_wells == 4, or if D is singular.  Not supported on several ranks yet. */
b200_status b200_wells_add_multisegment(b200_wells* w, unsigned int dim, unsigned int dim_wells, unsigned int Mb,
                                        const double* Bvalues, const unsigned int* BcolIndices, const unsigned int* BrowPointers,
                                        unsigned int DnumBlocks, const double* Dvalues, const int* DcolPointers,
                                        const int* DrowIndices, const double* Cvalues);
/* Test hook (host only): the dense row-major (4 Mb)^2 inverse of D held for multisegment well `index`. */
b200_status b200_wells_get_multisegment_inverse(const b200_wells* w, unsigned int index, double* Dinv_out);

/* ---- multi-GPU: row slabs, halo exchange over NVLink peer memory, NCCL all-reduce ------------ */
/*
 * One process and one b200_solver per GPU (rank).  The reference has no multi-GPU accelerator path (it
 * disables the bridge under MPI, ISTLSolverEbos.hpp:136-141); the semantics implemented here are those of
 * its MPI CPU path, which is what a lifted ban would have to reproduce:
 *   - every rank holds the rows it owns, local column numbering with the ghost cells LAST
 *     (--owner-cells-first, ISTLSolverEbos.hpp:171-180; findOverlapRowsAndColumns.hpp:119-139);
 *   - operator: owned rows times the (owned + ghost) vector after copyOwnerToAll
 *     (WellModelGhostLastMatrixAdapter::apply, WellOperators.hpp:200-214);
 *   - preconditioner: ILU0 of the owned x owned block, no communication (block Jacobi:
 *     PreconditionerFactory.hpp:237-252, ParallelOverlappingILU0.hpp:440-494,857-895);
 *   - scalar products: owned entries only, summed over the ranks (Dune::OwnerOverlapCopyCommunication).
 * In this mode b200_solve_system takes N = 3 * owned rows, a pattern whose column indices run over
 * owned + n_ghost cells, and b / x of the owned rows only.  Standard wells must lie inside one rank.
 * Call order: b200_create, b200_dist_init, b200_dist_set_halo, (exchange the IPC handles), b200_dist_map_rank per rank,
 * b200_dist_connect_peer per neighbour, then b200_solve_system on every rank collectively.
 */

/* ncclGetUniqueId: 128 bytes, produced on rank 0 and sent to the other ranks by the host. */
b200_status b200_dist_unique_id(unsigned char* id128);
/* Marks the solver as one rank of `world` and joins the NCCL communicator (world == 1: no NCCL needed,
 * id128 may be NULL). */
b200_status b200_dist_init(b200_solver* s, int rank, int world, const unsigned char* id128);
/* Halo plan of this rank.  neigh_rank[n]: ranks exchanged with; send_rows[send_ptr[n] .. send_ptr[n+1]):
 * owned local rows whose x entries neighbour n needs, in the order n stores them as ghosts;
 * ghosts [recv_ptr[n], recv_ptr[n+1]) (ghost index = local column - owned rows) are filled by neighbour n.
 * Allocates the receive block and returns its CUDA IPC handle (64 bytes) for the neighbours. */
b200_status b200_dist_set_halo(b200_solver* s, int n_ghost, int n_neigh, const int* neigh_rank, const int* send_ptr,
                               const int* send_rows, const int* recv_ptr, unsigned char* ipc_handle64);
/* Maps the receive block of `rank` (its IPC handle from b200_dist_set_halo) into this process; the own rank needs no
 * handle.  Once every rank is mapped the dot products are all-reduced through peer-memory mailboxes (one NVLink round
 * trip per reduction, option "p2p_allreduce" = 1, default); neighbours must be mapped before b200_dist_connect_peer. */
b200_status b200_dist_map_rank(b200_solver* s, int rank, const unsigned char* ipc_handle64);
/* Wires the halo push to neighbour neigh_index.  peer_n_ghost: that rank's ghost count; peer_recv_offset: first ghost
 * index of MY section there (its recv_ptr[slot]); peer_slot: my index in ITS neighbour list. */
b200_status b200_dist_connect_peer(b200_solver* s, int neigh_index, int peer_n_ghost, int peer_recv_offset, int peer_slot);
/* Collective y_owned = (A [x_owned; x_ghost])_owned on the uploaded system (parity tests). */
b200_status b200_dist_spmv(b200_solver* s, const double* x_owned_host, double* y_owned_host);
int b200_dist_rank(const b200_solver* s);
int b200_dist_world(const b200_solver* s);

/* ---- kernel-level entry points (parity tests, roofline measurement) ------------------------ */

/* All operate on the system uploaded by the last solve/upload call; vectors are host arrays of
 * N doubles in the caller's natural ordering (the library permutes to its level ordering). */
b200_status b200_spmv(b200_solver* s, const double* x_host, double* y_host);            /* y = A x */
b200_status b200_well_apply(b200_solver* s, const double* x_host, double* y_inout_host); /* y -= C^T D^-1 B x */
b200_status b200_ilu0_factorize(b200_solver* s);                                        /* LU <- ILU0(A) */
b200_status b200_ilu0_apply(b200_solver* s, const double* d_host, double* v_host);      /* v = w (LU)^-1 d */
/* LU in the caller's pattern (nnz doubles): L blocks below the diagonal, U above, the diagonal
 * block holding the INVERSE of the pivot, as ParallelOverlappingILU0.hpp:440-494 leaves it. */
b200_status b200_get_ilu0(b200_solver* s, double* lu_vals_host);
/* Level scheduling of the pattern (bda/Reorder.cpp:266-318): to_order/from_order hold Nb ints,
 * rows_per_level up to Nb ints; returns the level count through num_levels. */
b200_status b200_get_level_schedule(b200_solver* s, int* to_order, int* from_order,
                                    int* rows_per_level, int* num_levels);
/* Host-only analysis (no device needed): same outputs from a raw pattern. */
b200_status b200_level_schedule_host(int Nb, const int* rows, const int* cols, int* to_order,
                                     int* from_order, int* rows_per_level, int* num_levels);

/* Host-only check of the triangular-sweep schedule this library would run for a pattern (no device
 * needed): emulates the packed pencil streams chunk by chunk with random factor values and compares
 * with the sequential natural-order substitution (ParallelOverlappingILU0.hpp:867-895).  parts /
 * stage_bytes / window <= 0 select the defaults.  stats (12 values, may be NULL): parts, lines, strips,
 * stages L, stages U, chunks L, shared-memory deps L, global deps L, max meta ints, max value doubles,
 * max rhs rows, reference levels.  Fails with B200_ANALYSIS_FAILED if the schedule could deadlock. */
b200_status b200_sweep_schedule_check_host(int Nb, const int* rows, const int* cols, int parts, int stage_bytes,
                                           int window, unsigned int seed, double* max_rel_err, long long* stats);

/* The same check for the round-2 sweep schedule (groups of consumer warps that stream their own records; csrc/sweep2.hpp,
 * k_sweep2): host emulation of the kernel with random factor values against the sequential natural-order substitution, the
 * upper sweep with relaxation `relax` (x = relax * U^-1 L^-1 rhs, ParallelOverlappingILU0.hpp:897-901).  Arguments <= 0 select
 * the defaults.  stats (12 values, may be NULL): parts, lines, strips, records L, records U, empty records L, shared-memory
 * deps L, external deps L, external rows L, helper blocks L, max groups, max warps per group. */
b200_status b200_sweep2_schedule_check_host(int Nb, const int* rows, const int* cols, int parts, int window, int ext_window,
                                            int consumer_warps, int helpers, int groups, int wg, unsigned int seed, double relax,
                                            double* max_rel_err, long long* stats);

/* Opt-in multi-colour ordering (the reference's --opencl-ilu-reorder=graph_coloring: Reorder.cpp:58-172 colorBlockedNodes,
 * :209-222 colorsToReordering, called from BILU0.cpp:86-91 with maxRowsPerColor = maxColsPerColor = Nb).  Host only: colours
 * the block rows so that connected rows differ, returns the colour-major order (to_order[natural row] = position,
 * from_order[position] = natural row), the rows per colour (room for 256, may be NULL) and the number of colours.  The
 * reference seeds its random weights from std::random_device (Reorder.cpp:35-43); here `seed` makes the result reproducible. */
b200_status b200_graph_coloring_host(int Nb, const int* rows, const int* cols, unsigned int seed, int* to_order, int* from_order,
                                     int* rows_per_color, int* ncolors);

/* The ordering a solver created with option "reorder" = 1 works in (valid after its first solve).  Such a solver takes ILU0 of
 * the colour-permuted matrix -- another, weaker preconditioner (+30-50 % iterations on grid problems), as the reference's OpenCL
 * backend does with graph_coloring -- and sweeps level by level (k_trsv_level, one launch per colour); vals / b / x stay in the
 * caller's order.  Not available on several ranks. */
b200_status b200_get_reorder(b200_solver* s, int* to_order, int* from_order, int* ncolors);

/* Host-only replay of the ILU0 elimination plan the device kernel executes (no device needed): LU in the caller's pattern
 * with the inverse pivot in the diagonal slot, as ParallelOverlappingILU0.hpp:440-494 leaves it.  max_row / max_ops (may be
 * NULL) return the longest block row and the longest plan; the device uses the planned kernel when they fit its buffers. */
b200_status b200_factor_plan_check_host(int Nb, const int* rows, const int* cols, const double* vals, double* lu_out,
                                        int* max_row, int* max_ops);

/* Time `reps` back-to-back launches of one kernel with CUDA events on the solver's stream.
 * which: "spmv", "ilu_apply", "ilu_lower", "ilu_upper", "ilu_factor", "vec_p", "vec_xr1", "vec_xr2",
 * "well_apply", "permute" (b200_kernel_stats also knows "halo_push", "spmv_ghost", "allreduce", "finish").  Returns the mean milliseconds per launch and the ALGORITHMIC bytes one
 * launch moves.  flush_l2 != 0 writes a >L2 scratch buffer before every launch. */
b200_status b200_time_kernel(b200_solver* s, const char* which, int reps, int flush_l2,
                             double* ms_per_launch, double* algorithmic_bytes);

/* Debugging aid: with option "sweep_trace" = 1 the triangular sweeps record, per part (148 x) and stage (first 1024),
 * four SM-clock stamps {consumer starts waiting, data landed, stage done, producer issued}; this copies the trace of the
 * last sweep (148 * 1024 * 4 values) to the host. */
b200_status b200_get_sweep_trace(b200_solver* s, long long* out, long long count);

/* Per-kernel totals accumulated over the solves run with option "profile" = 1.
 * Returns B200_UNKNOWN_ERROR for an unknown name. */
b200_status b200_kernel_stats(b200_solver* s, const char* which, long long* launches,
                              double* total_ms, double* algorithmic_bytes_per_launch);
void        b200_reset_stats(b200_solver* s);
/* Number of kernels this library launched since creation / since b200_reset_stats. */
long long   b200_launch_count(b200_solver* s);

/* Device-side stopwatch on the solver's own stream (CUDA events): start records an event, stop
 * records a second one, waits for it and returns the milliseconds in between -- the timed region of
 * the benchmark, immune to host jitter. */
b200_status b200_timer_start(b200_solver* s);
b200_status b200_timer_stop(b200_solver* s, double* ms);

/* 1 if a CUDA device with compute capability 10.x is visible, else 0 (never throws). */
int b200_device_available(void);
/* Library version string. */
const char* b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200BDA_H */
