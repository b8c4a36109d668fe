// analysis.hpp -- host-side sparsity analysis for the B200 ILU0-BiCGSTAB backend.
//
// Replaces, for this backend, what the reference does in BILU0::init (bda/BILU0.cpp:50-158) and
// cusparseSolverBackend::analyse_matrix (bda/cusparseSolverBackend.cu:348-422).  Done once per
// sparsity pattern, O(nnzb log) work.  It produces
//   (1) the reference's level sets of the lower-triangular dependency DAG, exactly as
//       bda::findLevelScheduling returns them (bda/Reorder.cpp:266-318) -- exported for parity and used
//       (on the symmetrised pattern) as the global topological key of everything below;
//   (2) the PENCIL schedule the triangular sweeps run.  Inter-SM signalling through L2 costs
//       0.25-0.46 us per hop on B200 (profiles/r1_pingpong_latency.txt): a sweep that crosses the chip
//       once per level set (298 of them for 100^3 cells) is latency bound at 1-3 us per level, 20x off
//       the HBM roofline.  So the rows are cut into P parts (P = resident CTAs, one per SM), each a
//       bundle of grid LINES (maximal runs r, r+1, ... of mutually dependent rows): a pencil.  Inside a
//       pencil consecutive level sets hand values over through SHARED memory (one named barrier per
//       level, ~0.1 us); only dependencies that leave the pencil travel through L2, and any dependency
//       path crosses few pencils, so the exposed hops drop from #levels to ~#pencils on a diagonal.
//       Mathematically nothing changes: any topological order of the DAG gives the same ILU0 factors
//       and the same triangular solves as the sequential reference (natural ordering).
//   (3) "p-space": rows renumbered part by part, inside a part by (global level, natural index) -- the
//       processing order, so every CTA streams its factor slice, its rhs and its output linearly --
//       the BSR pattern of the row-permuted matrix, and the packed lane-major streams of both sweeps.
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

namespace b200 {

constexpr int kRowsPerWarp = 10;   // 3 lanes per block row -> 30 active lanes per warp
constexpr int kPadCol = INT_MIN;   // padded dependency slot of a chunk (factor value 0)

struct LevelSchedule {
    int nlev = 0;
    std::vector<int> level;        // level of each natural row
    std::vector<int> toOrder;      // reference-identical (discovery order) natural -> ordered
    std::vector<int> fromOrder;    // ordered -> natural
    std::vector<int> levelPtr;     // nlev + 1
};

// Level sets exactly as bda::findLevelScheduling produces them for the structurally symmetric
// patterns it is written for: level 0 = rows without lower-triangular dependencies (ascending),
// each further level = rows whose dependencies are all done, discovered by scanning the previous
// level in order and the rows depending on each of its rows in ascending order (Reorder.cpp:288-311).
// Differences: (1) O(nnzb) (the reference never clears its candidate list, O(Nb * levels));
// (2) level 0 is found through the CSR lower entries, not through the CSC pattern
// (Reorder.cpp:276-282), so a structurally NON-symmetric pattern still gets a valid schedule.
inline LevelSchedule level_schedule(int Nb, const int* rows, const int* cols)
{
    LevelSchedule S;
    const int64_t nnzb = rows[Nb];
    S.level.assign(Nb, -1);
    S.toOrder.assign(Nb, -1);
    S.fromOrder.assign(Nb, -1);
    std::vector<int> cptr(Nb + 1, 0), crow(std::max<int64_t>(nnzb, 1));
    for (int64_t k = 0; k < nnzb; ++k) {
        if (cols[k] < 0 || cols[k] >= Nb) throw std::runtime_error("column index out of range");
        cptr[cols[k] + 1]++;
    }
    for (int c = 0; c < Nb; ++c) cptr[c + 1] += cptr[c];
    {
        std::vector<int> fill(cptr.begin(), cptr.end() - 1);
        for (int r = 0; r < Nb; ++r)
            for (int k = rows[r]; k < rows[r + 1]; ++k) crow[fill[cols[k]]++] = r;
    }
    std::vector<int> pending(Nb, 0);
    for (int r = 0; r < Nb; ++r) {
        int n = 0;
        for (int k = rows[r]; k < rows[r + 1] && cols[k] < r; ++k) ++n;
        pending[r] = n;
    }
    int next = 0;
    S.levelPtr.push_back(0);
    for (int r = 0; r < Nb; ++r)
        if (pending[r] == 0) { S.fromOrder[next] = r; S.toOrder[r] = next; S.level[r] = 0; ++next; }
    S.levelPtr.push_back(next);
    int active = 0;
    // A row joins a level when all of its lower dependencies are done; the reference appends it the
    // first time it is visited in that state while scanning the previous level, which is its first
    // visit of the sweep because readiness does not change during a sweep (doneRows is updated only
    // after the scan, Reorder.cpp:303-311).
    std::vector<int> seen(Nb, -1), hits(Nb, 0), cand;
    while (next < Nb) {
        const int lev = (int) S.levelPtr.size() - 1;
        const int start = next;
        cand.clear();
        for (; active < start; ++active) {
            const int p = S.fromOrder[active];
            for (int k = cptr[p]; k < cptr[p + 1]; ++k) {
                const int r = crow[k];
                if (r <= p || S.level[r] >= 0) continue;
                if (seen[r] != lev) { seen[r] = lev; hits[r] = 0; cand.push_back(r); }
                ++hits[r];
            }
        }
        for (int r : cand)
            if (hits[r] == pending[r]) { S.fromOrder[next] = r; S.toOrder[r] = next; S.level[r] = lev; ++next; }
        for (int r : cand) {
            if (S.level[r] < 0) pending[r] -= hits[r];
            hits[r] = 0;
        }
        if (next == start) throw std::runtime_error("level scheduling made no progress (cyclic dependency?)");
        S.levelPtr.push_back(next);
    }
    S.nlev = (int) S.levelPtr.size() - 1;
    return S;
}

// ---- packed sweep streams -------------------------------------------------------------------------
//
// A sweep (forward over L, backward over U) is cut, per part, into STAGES: tens of KB of consecutive
// work the producer warp of the CTA fetches with three bulk copies (meta ints, factor values, rhs
// rows) into one slot of a shared-memory ring.  The unit of work is a RECORD: <= 10 rows of one level
// (3 lanes per row, lane = 3 q + comp) with up to 3 dependencies each; rows with more dependencies get
// continuation records (the partial sums stay in registers).  Record c of a level belongs to consumer
// warp c % W; every warp walks its own static WORK LIST and passes every level barrier of the stage.
// Everything a lane needs is laid out so that it costs one load with an immediate offset.
//
// Shared-memory value space of a part ("xwin", rows of 3 doubles): [0, window) the most recent rows of
// the part (position in processing order & (window - 1)), [window, window + extWindow) a ring of rows
// owned by other parts, parked there by helper warps, and one all-zero row at window + extWindow that
// padded dependency slots point to.
//
//   meta blob (ints, 16-byte multiple):
//     header  [0] ngroups [1] nrecords [2] g_lo (even-aligned first p-row of the rhs copy) [3] rhs rows copied
//             [4] next (external rows of the stage) [5] off_ext [6] off_wl [7] off_items
//             [8] ext_base (external rows of the part before this stage) [9] level barriers in the stage [10] off_codes
//     groups  : at 12, ngroups cumulative ends: the external rows [0, end) of the stage list are needed by the
//               levels up to the group's; a helper warp delivers them group by group
//     ext     : at off_ext, next p-rows (rows of other parts, or of this part beyond the window), in group order;
//               row k of the list is parked at xwin row window + ((ext_base + k) & (extWindow - 1))
//     wl      : at off_wl, W + 1 record offsets (per warp) then W trailing barrier counts
//     items   : at off_items (16-byte aligned), one int4 per record, grouped by warp in processing order:
//               {24 (g0 - g_lo), first | last << 1 | count << 4 | level barriers to pass first << 16, ext_need, 3 g0}
//               row q is p-row g0 + q (lower sweep) or g0 - q (upper sweep); external rows [0, ext_need) must be parked
//     codes   : at off_codes (16-byte aligned), per record 32 x int4 (one per lane):
//               {32 * xwin row of dependency 0, 1, 2, byte offset of the lane's result in xwin or -1 (no store)}
//   vals blob (doubles): per record NF x 32 doubles, field f of lane l at sweep_vidx(f, l).
//     lower sweep, NF = 9 : field 3 j + v = L[block j of row q][comp][v]
//     upper sweep, NF = 12: field 3 j + v = (D^-1 U)[block j of row q][comp][v], field 9 + v = (w D^-1)[comp][v]
//     (the inverse pivot and the relaxation factor w are folded into the stream when it is filled, so that the
//     dependent part of a row is the same 9 fma for both sweeps: x = (w D^-1) y - sum (D^-1 U) x, with x = w U^-1 y)
// position of field f of lane l inside a record of the value stream: doubles are paired so that a lane fetches two with
// one 16-byte load (lower: 4 pairs + the ninth value alone; upper: 6 pairs)
inline int sweep_vidx(bool lower, int f, int l) { return (lower && f == 8) ? 256 + l : (f >> 1) * 64 + 2 * l + (f & 1); }
constexpr int kXwinStride = 4;     // doubles per xwin row (3 used): rows are 32-byte aligned, read as 16 + 8 bytes (48-byte rows would
                                   // remove the 2-way bank conflicts of the 8-byte reads; measured: no gain, 50 % more shared memory)

struct StageRef {
    long long meta_off;    // ints into SweepPlan::meta
    long long vals_off;    // doubles into the sweep's value stream (even)
    int meta_ints, vals_doubles;
    int g_lo, g_rows;      // rhs copy: rows [g_lo, g_lo + g_rows), both even
};
struct PartRef { int stage_begin, stage_end, row0, nrows; };
struct BuildRef {          // one per record: where its factor values come from
    long long vals_off;    // doubles into the value stream
    int src_off;           // into SweepPlan::src: 3 x count dependency blocks then count pivot blocks (p-space, -1: none)
    int count, first;
};

struct SweepPlan {
    std::vector<int> meta;
    std::vector<StageRef> stages;
    std::vector<PartRef> parts;
    std::vector<BuildRef> build;
    std::vector<int> src;
    long long nvals = 0;
    int nfields = 9;
    int maxMetaInts = 4, maxValsDoubles = 2, maxRhsRows = 2, maxExtRows = 2;
    long long nchunks = 0, nentries = 0, nExternal = 0, nWindow = 0, nExtRows = 0;
};

struct Analysis {
    int Nb = 0;
    int64_t nnzb = 0;
    int nlev = 0;                     // reference level sets (lower-triangular DAG)
    LevelSchedule sched;              // reference-identical schedule (exported for parity)
    std::vector<int> perm;            // p-space row -> natural row
    std::vector<int> iperm;           // natural row -> p-space row
    std::vector<int> prow, pcol;      // pattern of the permuted matrix; entries of a row keep their NATURAL column order
    std::vector<int> pdiag;           // index of the diagonal block of each p-space row
    std::vector<int> srcblk;          // p-space block -> natural block index
    // factorisation schedule: p-space rows grouped by global level of the symmetrised pattern
    int nflev = 0;
    std::vector<int> flevPtr, flevRows;
    // elimination plan of the factorisation (k_ilu_factor_plan): per p-space row the upstream blocks it needs, in the
    // order the left-looking elimination uses them.  op = {p-space block, code}: code >= 0: pivot of lower entry
    // `code & 255` of the row (offset from the row start) followed by `code >> 8` updates; code < 0: update of the row's
    // block -(code + 1) with the current L block times this upstream U block.
    std::vector<int> facPtr;          // Nb + 1, in ops
    std::vector<int> facOps;          // 2 ints per op
    int facMaxRow = 0, facMaxOps = 0;
    // triangular sweeps
    int nparts = 0, nlines = 0, window = 0, warps = 8, groups = 1, extWindow = 512;
    int nstrips = 0;
    std::vector<int> partPtr;         // nparts + 1, p-space rows
    std::vector<int> partMaxStep;     // rows in the largest level step of each part
    SweepPlan L, U;
    int64_t nnzL = 0;                 // strictly lower blocks
    int maxRowLen = 0;
};

// Sliced-ELL copy of the p-space matrix for the SpMV: slices of 32 consecutive rows, one LANE per row, slot k of a slice holds
// the k-th block of each of its rows as 9 x 32 doubles ([value][lane], a warp reads 256 contiguous bytes per load) plus 32
// column indices.  The slice width is its longest row, capped at 1.5 x its mean row length + 1: the blocks beyond the cap stay
// in the BSR arrays and are added by the owning lane (rows much longer than their neighbours: wells folded into the matrix).
struct SellPlan {
    int nslices = 0;
    std::vector<int> ptr;             // nslices + 1, in slots
    std::vector<int> over;            // per slice: 1 if some row has blocks beyond the slice width
    std::vector<int> col;             // 32 per slot; padding points at a valid row with zero values
    std::vector<int> src;             // 32 per slot: source block of the caller's array (as srcblk), -1 = padding
};

inline SellPlan build_sell(int Nb, const std::vector<int>& prow, const std::vector<int>& pcol, const std::vector<int>& srcblk)
{
    SellPlan P;
    P.nslices = (Nb + 31) / 32;
    P.ptr.assign(P.nslices + 1, 0);
    P.over.assign(P.nslices, 0);
    for (int s = 0; s < P.nslices; ++s) {
        const int r0 = 32 * s, r1 = std::min(Nb, r0 + 32);
        int longest = 0;
        for (int r = r0; r < r1; ++r) longest = std::max(longest, prow[r + 1] - prow[r]);
        const int cap = (int) ((3LL * (prow[r1] - prow[r0]) + 2 * (r1 - r0) - 1) / (2 * (r1 - r0))) + 1;
        const int width = std::min(longest, cap);
        P.over[s] = longest > width;
        P.ptr[s + 1] = P.ptr[s] + width;
    }
    P.col.assign((size_t) P.ptr[P.nslices] * 32, 0);
    P.src.assign((size_t) P.ptr[P.nslices] * 32, -1);
    for (int s = 0; s < P.nslices; ++s)
        for (int lane = 0; lane < 32; ++lane) {
            const int r = 32 * s + lane, width = P.ptr[s + 1] - P.ptr[s];
            for (int k = 0; k < width; ++k) {
                const size_t o = (size_t) (P.ptr[s] + k) * 32 + lane;
                if (r < Nb && prow[r] + k < prow[r + 1]) { P.col[o] = pcol[prow[r] + k]; P.src[o] = srcblk[prow[r] + k]; }
                else P.col[o] = std::min(r, Nb - 1);
            }
        }
    return P;
}

// Work units of the SpMV that runs inside the upper sweep (k_sweep<..., SPMV>): a CTA whose part is finished claims units
// of `slicesPerUnit` SELL slices from a global counter.  A unit may start when every part that owns one of its rows or
// columns has finished; units are listed in the order in which that is expected to happen (the upper sweep finishes the
// parts with the HIGHEST lowest-level first).
struct FusedPlan {
    std::vector<int> units;           // 2 per unit: first slice, slices
    std::vector<int> needPtr, need;   // parts a unit reads from (its own rows included)
};

inline FusedPlan build_fused(int Nb, const std::vector<int>& prow, const std::vector<int>& pcol, const std::vector<int>& partPtr,
                             const std::vector<int>& flevPtr, const std::vector<int>& flevRows, int slicesPerUnit)
{
    FusedPlan F;
    const int nparts = (int) partPtr.size() - 1, nslices = (Nb + 31) / 32;
    std::vector<int> partOfRow(Nb), minlev(nparts, INT32_MAX);
    for (int p = 0; p < nparts; ++p)
        for (int r = partPtr[p]; r < partPtr[p + 1]; ++r) partOfRow[r] = p;
    for (int l = 0; l + 1 < (int) flevPtr.size(); ++l)
        for (int k = flevPtr[l]; k < flevPtr[l + 1]; ++k) minlev[partOfRow[flevRows[k]]] = std::min(minlev[partOfRow[flevRows[k]]], l);
    const int nunits = (nslices + slicesPerUnit - 1) / slicesPerUnit;
    std::vector<std::vector<int>> needOf(nunits);
    std::vector<int> key(nunits), order(nunits), mark(nparts, -1);
    for (int u = 0; u < nunits; ++u) {
        const int r0 = 32 * u * slicesPerUnit, r1 = std::min(Nb, r0 + 32 * slicesPerUnit);
        int k = INT32_MAX;
        for (int r = r0; r < r1; ++r) {
            auto touch = [&](int p) { if (mark[p] != u) { mark[p] = u; needOf[u].push_back(p); k = std::min(k, minlev[p]); } };
            // the vector entries are gathered through L1: every row that shares a 128-byte line with a gathered one (rows
            // are 24 bytes) must be final too
            auto touch_row = [&](int c) { touch(partOfRow[std::max(0, c - 6)]); touch(partOfRow[c]); touch(partOfRow[std::min(Nb - 1, c + 6)]); };
            touch(partOfRow[r]);
            for (int e = prow[r]; e < prow[r + 1]; ++e) touch_row(pcol[e]);
        }
        key[u] = k; order[u] = u;
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] > key[b]; });
    F.needPtr.push_back(0);
    for (int u : order) {
        F.units.push_back(u * slicesPerUnit);
        F.units.push_back(std::min(slicesPerUnit, nslices - u * slicesPerUnit));
        for (int p : needOf[u]) F.need.push_back(p);
        F.needPtr.push_back((int) F.need.size());
    }
    return F;
}

struct AnalysisOptions {
    int parts = 148;            // resident CTAs of the sweep kernels
    int stageBytes = 16384;     // meta + values + rhs of one ring slot
    int window = 0;             // rows of the part kept in the shared-memory window (power of two); 0: four times the largest
                                // level step of any part, at least 256
    int warps = 8;              // consumer warps of a sweep CTA (the static work lists are cut for this many)
    int groups = 1;             // consecutive levels go to different warp groups (warps / groups warps each): while one group
                                // runs the dependent part of level l, the next ones already hold the operands of l+1, l+2
    int extWindow = 512;        // rows of the external-row ring (power of two); bounds ring slots x external rows per stage
    bool buildStreams = true;   // false: the round-2 schedule (sweep2.hpp) is built instead of the packed round-1 streams
};

namespace detail {

inline void build_sweep(const Analysis& A, const int* rows, const int* cols, const std::vector<int>& glev,
                        const std::vector<int>& partOf, bool lower, const AnalysisOptions& opt, SweepPlan& S)
{
    const int W = A.window, EW = A.extWindow, NWc = A.warps, NG = A.groups, NWG = A.warps / A.groups;
    const int zrow = W + EW;
    const int NF = lower ? 9 : 12;
    S.nfields = NF;
    S.parts.resize(A.nparts);
    // cols of a TmpRec: >= 0 xwin row of the window, kPadCol padding, otherwise -(p-row + 1) of an external row
    struct TmpRec { int g0, count, level, ps0, first, last; int cols[3][kRowsPerWarp], src[3][kRowsPerWarp], piv[kRowsPerWarp]; };
    for (int p = 0; p < A.nparts; ++p) {
        const int row0 = A.partPtr[p], nrows = A.partPtr[p + 1] - row0;
        const int slack = W - A.partMaxStep[p];
        S.parts[p].stage_begin = (int) S.stages.size();
        S.parts[p].row0 = row0;
        S.parts[p].nrows = nrows;
        std::vector<TmpRec> st;         // pending stage
        int st_bytes = 0, st_glo = 0, st_ghi = 0;
        int prev_level = -1;            // level of the last record emitted in this part
        int level_index = 0;            // levels of the part so far: level l belongs to warp group l % groups
        long long ext_base = 0;         // external rows of the part before the pending stage
        auto flush = [&]() {
            if (st.empty()) return;
            const int nrec = (int) st.size();
            // external rows of the stage, listed once, grouped by the first level that needs them
            std::vector<int> ext, gend, need(nrec, 0), order, bars;
            std::vector<std::vector<int>> wl(NWc), wbar(NWc);
            std::vector<int> pend(NWc, 0);
            int nbar_total = 0;
            {
                std::vector<std::pair<int, int>> seen;      // (p-row, index), kept sorted
                for (int c = 0; c < nrec;) {
                    int e = c;
                    while (e < nrec && st[e].level == st[c].level) ++e;
                    if (prev_level >= 0 && st[c].level != prev_level) { for (int w = 0; w < NWc; ++w) pend[w]++; nbar_total++; level_index++; }
                    prev_level = st[c].level;
                    int chunk = -1;                          // continuation records follow their first record to the same warp
                    for (int k = c; k < e; ++k) {
                        bool any = false;
                        for (int j = 0; j < 3; ++j)
                            for (int q = 0; q < st[k].count; ++q) {
                                int& code = st[k].cols[j][q];
                                if (code >= 0 || code == kPadCol) continue;
                                const int gd = -(code + 1);
                                auto it = std::lower_bound(seen.begin(), seen.end(), std::make_pair(gd, -1));
                                int idx;
                                if (it != seen.end() && it->first == gd) idx = it->second;
                                else { idx = (int) ext.size(); ext.push_back(gd); seen.insert(it, std::make_pair(gd, idx)); }
                                code = W + (int) ((ext_base + idx) & (EW - 1));
                                any = true;
                            }
                        need[k] = any ? 1 : 0;
                        if (st[k].first) ++chunk;
                        const int w = (level_index % NG) * NWG + chunk % NWG;
                        wl[w].push_back(k);
                        wbar[w].push_back(pend[w]);
                        pend[w] = 0;
                    }
                    if (!ext.empty() && (gend.empty() || (int) ext.size() != gend.back())) gend.push_back((int) ext.size());
                    for (int k = c; k < e; ++k) if (need[k]) need[k] = (int) ext.size();
                    c = e;
                }
            }
            const int next = (int) ext.size(), ngroups = (int) gend.size();
            if (next > EW || nbar_total >= 32768) throw std::runtime_error("sweep stage too large");
            const int off_ext = 12 + ngroups;
            const int off_wl = off_ext + next;
            const int off_items = (off_wl + 2 * NWc + 1 + 3) & ~3;
            const int off_codes = off_items + 4 * nrec;
            const int meta_ints = off_codes + 128 * nrec;
            StageRef R{};
            R.meta_off = (long long) S.meta.size();
            R.meta_ints = meta_ints;
            R.vals_off = S.nvals;
            const int glo_al = st_glo & ~1, ghi_al = (st_ghi + 1) & ~1;
            R.g_lo = glo_al;
            R.g_rows = ghi_al - glo_al;
            S.meta.resize(S.meta.size() + meta_ints, 0);
            int* m = S.meta.data() + R.meta_off;
            m[0] = ngroups; m[1] = nrec; m[2] = glo_al; m[3] = R.g_rows; m[4] = next; m[5] = off_ext; m[6] = off_wl; m[7] = off_items;
            m[8] = (int) (ext_base & (EW - 1)); m[9] = nbar_total; m[10] = off_codes;
            std::copy(gend.begin(), gend.end(), m + 12);
            std::copy(ext.begin(), ext.end(), m + off_ext);
            int* w0 = m + off_wl;
            int o = 0;
            for (int w = 0; w < NWc; ++w) {
                w0[w] = o;
                for (size_t t = 0; t < wl[w].size(); ++t, ++o) {
                    const TmpRec& rc = st[wl[w][t]];
                    int* it = m + off_items + 4 * o;
                    it[0] = 24 * (rc.g0 - glo_al);
                    it[1] = rc.first | (rc.last << 1) | (rc.count << 4) | (wbar[w][t] << 16);
                    it[2] = need[wl[w][t]];
                    it[3] = 3 * rc.g0;
                    int* cd = m + off_codes + 128 * o;
                    for (int l = 0; l < 32; ++l) {
                        const int q = l / 3, comp = l - 3 * q;
                        const bool act = q < rc.count;
                        for (int j = 0; j < 3; ++j) {
                            const int code = act ? rc.cols[j][q] : kPadCol;
                            cd[4 * l + j] = 8 * kXwinStride * (code == kPadCol ? zrow : code);
                        }
                        cd[4 * l + 3] = (act && rc.last) ? 8 * kXwinStride * ((rc.ps0 + q) & (W - 1)) + 8 * comp : -1;
                    }
                    BuildRef B{};
                    B.vals_off = R.vals_off + (long long) o * NF * 32;
                    B.src_off = (int) S.src.size();
                    B.count = rc.count;
                    B.first = rc.first;
                    for (int j = 0; j < 3; ++j) for (int q = 0; q < rc.count; ++q) S.src.push_back(rc.src[j][q]);
                    for (int q = 0; q < rc.count; ++q) S.src.push_back(lower ? -1 : rc.piv[q]);
                    S.build.push_back(B);
                }
                w0[NWc + 1 + w] = pend[w];
            }
            w0[NWc] = o;
            const long long vo = (long long) nrec * NF * 32;
            if (vo > INT_MAX || S.src.size() > (size_t) INT_MAX) throw std::runtime_error("sweep stage too large");
            R.vals_doubles = (int) vo;
            S.nvals += vo;
            S.maxMetaInts = std::max(S.maxMetaInts, meta_ints);
            S.maxValsDoubles = std::max(S.maxValsDoubles, R.vals_doubles);
            S.maxRhsRows = std::max(S.maxRhsRows, R.g_rows);
            S.maxExtRows = std::max(S.maxExtRows, next);
            S.nExtRows += next;
            S.nchunks += nrec;
            S.nentries += nbar_total;
            S.stages.push_back(R);
            ext_base += next;
            st.clear();
            st_bytes = 0;
        };
        // walk the part in processing order: positions 0..nrows-1 (lower) or reversed (upper)
        auto g_of = [&](int ps) { return lower ? row0 + ps : row0 + nrows - 1 - ps; };
        int pos = 0;
        while (pos < nrows) {
            const int lev = glev[A.perm[g_of(pos)]];
            int end = pos;
            while (end < nrows && glev[A.perm[g_of(end)]] == lev) ++end;
            for (int s = pos; s < end; s += kRowsPerWarp) {
                const int count = std::min(kRowsPerWarp, end - s);
                // dependencies of the rows of this chunk
                std::vector<int> dcode[kRowsPerWarp], dsrc[kRowsPerWarp];
                int nd = 0, piv[kRowsPerWarp];
                for (int q = 0; q < count; ++q) {
                    const int ps = s + q, g = g_of(ps), r = A.perm[g];
                    for (int k = rows[r]; k < rows[r + 1]; ++k) {
                        const int c = cols[k];
                        if (lower ? c < r : c > r) {
                            const int gd = A.iperm[c];
                            int code = -(gd + 1);
                            if (partOf[c] == p) {
                                const int pd = lower ? gd - row0 : row0 + nrows - 1 - gd;      // processing position of the dependency
                                if (pd >= ps) throw std::runtime_error("internal: dependency not earlier in processing order");
                                if (ps - pd <= slack) { code = pd & (W - 1); S.nWindow++; }
                            }
                            if (code < 0) S.nExternal++;
                            dcode[q].push_back(code);
                            dsrc[q].push_back(A.prow[g] + (k - rows[r]));
                        }
                    }
                    piv[q] = A.pdiag[g];
                    nd = std::max(nd, (int) dcode[q].size());
                }
                const int npass = std::max(1, (nd + 2) / 3);
                int bytes = 0;
                std::vector<TmpRec> recs(npass);
                for (int r = 0; r < npass; ++r) {
                    TmpRec& t = recs[r];
                    t.g0 = g_of(s); t.count = count; t.level = lev; t.ps0 = s; t.first = r == 0; t.last = r == npass - 1;
                    for (int q = 0; q < kRowsPerWarp; ++q) {
                        t.piv[q] = q < count ? piv[q] : -1;
                        for (int j = 0; j < 3; ++j) {
                            const int d = 3 * r + j;
                            const bool have = q < count && d < (int) dcode[q].size();
                            t.cols[j][q] = have ? dcode[q][d] : kPadCol;
                            t.src[j][q] = have ? dsrc[q][d] : -1;
                            if (have && dcode[q][d] < 0) bytes += 4;
                        }
                    }
                    bytes += 16 + 512 + NF * 256;
                }
                bytes += 24 * count;
                if (!st.empty() && st_bytes + bytes > opt.stageBytes) flush();
                if (st.empty()) { st_glo = INT_MAX; st_ghi = 0; st_bytes = 192; }
                const int g0 = g_of(s);
                const int glo = lower ? g0 : g0 - count + 1, ghi = lower ? g0 + count : g0 + 1;
                st_glo = std::min(st_glo, glo); st_ghi = std::max(st_ghi, ghi);
                st_bytes += bytes;
                for (auto& t : recs) st.push_back(t);
            }
            pos = end;
        }
        flush();
        S.parts[p].stage_end = (int) S.stages.size();
    }
}

}  // namespace detail

inline Analysis analyse(int Nb, const int* rows, const int* cols, const AnalysisOptions& opt = AnalysisOptions())
{
    Analysis A;
    A.Nb = Nb;
    A.nnzb = rows[Nb];
    if (rows[0] != 0) throw std::runtime_error("rows[0] must be 0");
    if (Nb >= (1 << 30)) throw std::runtime_error("Nb too large");
    if (opt.window != 0 && (opt.window < 64 || (opt.window & (opt.window - 1)))) throw std::runtime_error("window must be a power of two >= 64");
    A.window = opt.window;
    if (opt.warps < 1 || opt.warps > 26) throw std::runtime_error("consumer warps must be in 1..26");
    if (opt.groups < 1 || opt.warps % opt.groups) throw std::runtime_error("consumer warps must be a multiple of the warp groups");
    A.warps = opt.warps;
    A.groups = opt.groups;
    if (opt.extWindow < 64 || (opt.extWindow & (opt.extWindow - 1))) throw std::runtime_error("extWindow must be a power of two >= 64");
    A.extWindow = opt.extWindow;
    for (int r = 0; r < Nb; ++r) {
        bool diag = false;
        for (int k = rows[r]; k < rows[r + 1]; ++k) {
            if (k > rows[r] && cols[k] <= cols[k - 1]) throw std::runtime_error("columns must be strictly ascending in every row");
            diag |= (cols[k] == r);
            A.nnzL += cols[k] < r;
        }
        if (!diag) throw std::runtime_error("diagonal block missing in block row " + std::to_string(r));
        A.maxRowLen = std::max(A.maxRowLen, rows[r + 1] - rows[r]);
    }
    A.sched = level_schedule(Nb, rows, cols);
    A.nlev = A.sched.nlev;

    // symmetrised lower adjacency: for row i all j < i with A_ij != 0 or A_ji != 0.  Levels taken on it
    // are valid for BOTH sweeps (forward over L ascending, backward over U descending) even if the
    // pattern is not structurally symmetric; for symmetric patterns they are the reference's level sets.
    std::vector<int> sptr(Nb + 1, 0), sadj;
    {
        std::vector<int> cnt(Nb, 0);
        for (int r = 0; r < Nb; ++r)
            for (int k = rows[r]; k < rows[r + 1]; ++k) {
                const int c = cols[k];
                if (c < r) cnt[r]++;
                else if (c > r) cnt[c]++;        // transpose entry (c, r) with r < c
            }
        for (int r = 0; r < Nb; ++r) sptr[r + 1] = sptr[r] + cnt[r];
        sadj.resize(std::max(sptr[Nb], 1));
        std::vector<int> fill(sptr.begin(), sptr.end() - 1);
        for (int r = 0; r < Nb; ++r)
            for (int k = rows[r]; k < rows[r + 1]; ++k) {
                const int c = cols[k];
                if (c < r) sadj[fill[r]++] = c;
                else if (c > r) sadj[fill[c]++] = r;
            }
        // duplicates (entry present in both triangles) are harmless for max-level computations
    }
    std::vector<int> glev(Nb, 0);
    int ng = 0;
    for (int r = 0; r < Nb; ++r) {
        int l = 0;
        for (int k = sptr[r]; k < sptr[r + 1]; ++k) l = std::max(l, glev[sadj[k]] + 1);
        glev[r] = l;
        ng = std::max(ng, l + 1);
    }
    A.nflev = ng;

    // ---- lines: maximal runs of consecutive rows each depending on its predecessor ---------------------
    const int Preq = std::max(1, opt.parts);
    const int lineCap = std::max(16, (Nb + Preq - 1) / Preq);
    std::vector<int> lineOf(Nb), lineStart;
    {
        int len = 0;
        for (int r = 0; r < Nb; ++r) {
            bool chained = false;
            if (r > 0 && len < lineCap)
                for (int k = sptr[r]; k < sptr[r + 1]; ++k) chained |= (sadj[k] == r - 1);
            if (!chained) { lineStart.push_back(r); len = 0; }
            lineOf[r] = (int) lineStart.size() - 1;
            ++len;
        }
    }
    const int nlines = (int) lineStart.size();
    lineStart.push_back(Nb);
    A.nlines = nlines;
    std::vector<double> lweight(nlines, 0.0);
    std::vector<int> lfar(nlines, 0);       // largest line-index distance to a lower neighbour line
    for (int r = 0; r < Nb; ++r) {
        const int l = lineOf[r];
        lweight[l] += (double) (rows[r + 1] - rows[r]) + 1.0;
        for (int k = sptr[r]; k < sptr[r + 1]; ++k) lfar[l] = std::max(lfar[l], l - lineOf[sadj[k]]);
    }
    // lines per "plane": the typical far offset (median over the lines that have one)
    int perPlane = 1;
    {
        std::vector<int> far;
        for (int l = 0; l < nlines; ++l) if (lfar[l] > 1) far.push_back(lfar[l]);
        if ((int64_t) far.size() * 10 >= nlines && !far.empty()) {
            std::nth_element(far.begin(), far.begin() + far.size() / 2, far.end());
            perPlane = far[far.size() / 2];
        }
    }
    const int P = std::min(Preq, nlines);
    const int nplanes = std::max(1, nlines / std::max(1, perPlane));
    const double T = (double) nlines / P;                       // lines per part
    int tk = (int) std::lround(std::sqrt(T));
    tk = std::max(1, std::min(tk, nplanes));
    int G = (int) std::lround((double) nplanes / tk);
    G = std::max(1, std::min(G, P));
    A.nstrips = G;
    // strips: consecutive lines in natural order, equal weight
    double totalw = 0.0;
    for (double w : lweight) totalw += w;
    std::vector<int> stripOf(nlines);
    std::vector<double> stripw(G, 0.0);
    {
        double cum = 0.0;
        for (int l = 0; l < nlines; ++l) {
            int s = (int) ((cum + 0.5 * lweight[l]) * G / totalw);
            s = std::max(0, std::min(G - 1, s));
            stripOf[l] = s;
            stripw[s] += lweight[l];
            cum += lweight[l];
        }
    }
    // bands per strip: P shared out in proportion to strip weight (largest remainder), at least 1 where lines exist
    std::vector<int> bands(G, 0);
    {
        std::vector<std::pair<double, int>> rem;
        int used = 0;
        for (int s = 0; s < G; ++s) {
            const double share = stripw[s] / totalw * P;
            bands[s] = (int) share;
            if (bands[s] == 0 && stripw[s] > 0.0) bands[s] = 1;
            used += bands[s];
            rem.emplace_back(share - (int) share, s);
        }
        std::sort(rem.begin(), rem.end(), [](auto& a, auto& b) { return a.first > b.first; });
        for (size_t i = 0; used < P && !rem.empty(); ++i, ++used) bands[rem[i % rem.size()].second]++;
        while (used > P) {                                 // the >= 1 floor may overshoot
            int big = (int) (std::max_element(bands.begin(), bands.end()) - bands.begin());
            if (bands[big] <= 1) break;
            bands[big]--; used--;
        }
    }
    // inside a strip: lines ordered by the level of their first row (a diagonal band), cut by weight
    std::vector<int> partOfLine(nlines, -1);
    int nparts = 0;
    {
        std::vector<std::vector<int>> linesOf(G);
        for (int l = 0; l < nlines; ++l) linesOf[stripOf[l]].push_back(l);
        for (int s = 0; s < G; ++s) {
            auto& v = linesOf[s];
            if (v.empty()) continue;
            std::stable_sort(v.begin(), v.end(), [&](int a, int b) { return glev[lineStart[a]] < glev[lineStart[b]]; });
            const int B = std::max(1, std::min<int>(bands[s], (int) v.size()));
            double cum = 0.0;
            int lastBand = -1;
            const int base = nparts;
            for (int l : v) {
                int b = (int) ((cum + 0.5 * lweight[l]) * B / stripw[s]);
                b = std::max(0, std::min(B - 1, b));
                if (b > lastBand + 1) b = lastBand + 1;     // no empty bands
                if (b > lastBand) lastBand = b;
                partOfLine[l] = base + b;
                cum += lweight[l];
            }
            nparts = base + lastBand + 1;
        }
    }
    A.nparts = nparts;
    std::vector<int> partOf(Nb);
    for (int r = 0; r < Nb; ++r) partOf[r] = partOfLine[lineOf[r]];

    // ---- p-space: part-major, then (global level, natural row) -----------------------------------------
    A.partPtr.assign(nparts + 1, 0);
    for (int r = 0; r < Nb; ++r) A.partPtr[partOf[r] + 1]++;
    for (int p = 0; p < nparts; ++p) A.partPtr[p + 1] += A.partPtr[p];
    A.perm.resize(Nb);
    A.iperm.resize(Nb);
    {
        std::vector<int> fill(A.partPtr.begin(), A.partPtr.end() - 1);
        for (int r = 0; r < Nb; ++r) A.perm[fill[partOf[r]]++] = r;
        for (int p = 0; p < nparts; ++p)
            std::stable_sort(A.perm.begin() + A.partPtr[p], A.perm.begin() + A.partPtr[p + 1],
                             [&](int a, int b) { return glev[a] < glev[b]; });
        for (int q = 0; q < Nb; ++q) A.iperm[A.perm[q]] = q;
    }
    A.partMaxStep.assign(nparts, 0);
    int maxStep = 1;
    for (int p = 0; p < nparts; ++p) {
        int run = 0;
        for (int q = A.partPtr[p]; q < A.partPtr[p + 1]; ++q) {
            run = (q > A.partPtr[p] && glev[A.perm[q]] == glev[A.perm[q - 1]]) ? run + 1 : 1;
            A.partMaxStep[p] = std::max(A.partMaxStep[p], run);
        }
        maxStep = std::max(maxStep, A.partMaxStep[p]);
    }
    if (A.window == 0) {
        A.window = 256;
        while (A.window < 4 * maxStep && A.window < 4096) A.window *= 2;
    }

    // permuted BSR pattern: rows permuted, every row keeps its entries in natural column order (so the
    // left-looking elimination order and the L/U split are exactly those of the natural-order reference)
    A.prow.resize((size_t) Nb + 1);
    A.pcol.resize(A.nnzb);
    A.pdiag.resize(Nb);
    A.srcblk.resize(A.nnzb);
    A.prow[0] = 0;
    for (int q = 0; q < Nb; ++q) {
        const int r = A.perm[q];
        A.prow[q + 1] = A.prow[q] + (rows[r + 1] - rows[r]);
    }
    for (int q = 0; q < Nb; ++q) {
        const int r = A.perm[q];
        int o = A.prow[q];
        for (int k = rows[r]; k < rows[r + 1]; ++k, ++o) {
            A.pcol[o] = A.iperm[cols[k]];
            A.srcblk[o] = k;
            if (cols[k] == r) A.pdiag[q] = o;
        }
    }
    // factorisation schedule in p-space rows
    A.flevPtr.assign(ng + 1, 0);
    for (int r = 0; r < Nb; ++r) A.flevPtr[glev[r] + 1]++;
    for (int l = 0; l < ng; ++l) A.flevPtr[l + 1] += A.flevPtr[l];
    A.flevRows.resize(Nb);
    {
        std::vector<int> fill(A.flevPtr.begin(), A.flevPtr.end() - 1);
        for (int q = 0; q < Nb; ++q) A.flevRows[fill[glev[A.perm[q]]]++] = q;
    }
    // elimination plan: rows are final once their level has run, so every upstream block can be prefetched up front
    A.facPtr.assign((size_t) Nb + 1, 0);
    A.facOps.reserve((size_t) 4 * A.nnzb);
    for (int q = 0; q < Nb; ++q) {
        const int rs = A.prow[q], re = A.prow[q + 1], di = A.pdiag[q];
        A.facMaxRow = std::max(A.facMaxRow, re - rs);
        for (int kj = rs; kj < di; ++kj) {
            const int j = A.pcol[kj];
            const size_t head = A.facOps.size();
            A.facOps.push_back(A.pdiag[j]);
            A.facOps.push_back(0);
            int nupd = 0;
            for (int jk = A.pdiag[j] + 1; jk < A.prow[j + 1]; ++jk) {
                const int colk = A.pcol[jk];
                for (int ik = kj + 1; ik < re; ++ik)
                    if (A.pcol[ik] == colk) { A.facOps.push_back(jk); A.facOps.push_back(-(ik - rs + 1)); ++nupd; break; }
            }
            A.facOps[head + 1] = (kj - rs) | (nupd << 8);
        }
        A.facPtr[q + 1] = (int) (A.facOps.size() / 2);
        A.facMaxOps = std::max(A.facMaxOps, A.facPtr[q + 1] - A.facPtr[q]);
    }
    if (opt.buildStreams) {
        detail::build_sweep(A, rows, cols, glev, partOf, true, opt, A.L);
        detail::build_sweep(A, rows, cols, glev, partOf, false, opt, A.U);
    }
    return A;
}

// ---- host emulation of the sweep kernels (schedule verification without a GPU) ---------------------
//
// Interprets the packed streams exactly as k_sweep does, one chunk at a time, round-robin over the
// parts; a chunk whose out-of-window dependency has not been produced yet makes its part yield.
// Returns false if a full round makes no progress (the schedule would deadlock on the device).
inline void fill_stream_host(const SweepPlan& S, bool lower, const double* LU, double relax, std::vector<double>& vals)
{
    vals.assign((size_t) std::max<long long>(S.nvals, 1), 0.0);
    for (const BuildRef& B : S.build)
        for (int l = 0; l < 32; ++l) {
            const int q = l / 3, comp = l - 3 * q;
            if (q >= B.count) continue;
            double inv[3] = {0.0, 0.0, 0.0};
            if (!lower) {
                const int kp = S.src[B.src_off + 3 * B.count + q];
                for (int e = 0; e < 3; ++e) inv[e] = LU[(size_t) kp * 9 + comp * 3 + e];
                if (B.first) for (int v = 0; v < 3; ++v) vals[B.vals_off + sweep_vidx(false, 9 + v, l)] = relax * inv[v];
            }
            for (int j = 0; j < 3; ++j) {
                const int k = S.src[B.src_off + j * B.count + q];
                for (int v = 0; v < 3; ++v) {
                    double x = 0.0;
                    if (k >= 0) {
                        if (lower) x = LU[(size_t) k * 9 + comp * 3 + v];
                        else x = inv[0] * LU[(size_t) k * 9 + v] + inv[1] * LU[(size_t) k * 9 + 3 + v] + inv[2] * LU[(size_t) k * 9 + 6 + v];
                    }
                    vals[B.vals_off + sweep_vidx(lower, 3 * j + v, l)] = x;
                }
            }
        }
}

inline bool emulate_sweep(const Analysis& A, const SweepPlan& S, bool lower, const std::vector<double>& vals, const double* rhs,
                          double* out)
{
    const int W = A.window, EW = A.extWindow, NW = A.warps, zrow = W + EW, NF = S.nfields;
    const double NaN = std::nan("");
    for (int i = 0; i < 3 * A.Nb; ++i) out[i] = NaN;
    // One cursor per consumer warp, exactly the control flow of k_sweep: walk the work list, pass `nbar` level
    // barriers (all W warps must arrive), wait for the parked external rows, compute, finally the trailing barriers.
    struct Warp { int t, bars; bool loaded, in_tail, done, at_bar; double carry[32]; };
    struct Part { int stage; std::vector<Warp> w; int arrivals; std::vector<double> xwin; bool started; };
    std::vector<Part> parts(A.nparts);
    for (int p = 0; p < A.nparts; ++p) {
        parts[p].stage = S.parts[p].stage_begin;
        parts[p].w.assign(NW, Warp{});
        parts[p].arrivals = 0;
        parts[p].xwin.assign((size_t) kXwinStride * (zrow + 1), NaN);
        for (int e = 0; e < 3; ++e) parts[p].xwin[(size_t) kXwinStride * zrow + e] = 0.0;
        parts[p].started = false;
    }
    int remaining = A.nparts;
    std::vector<char> finished(A.nparts, 0);
    while (remaining > 0) {
        bool progress = false;
        for (int p = 0; p < A.nparts; ++p) {
            if (finished[p]) continue;
            Part& P = parts[p];
            const PartRef& PR = S.parts[p];
            while (true) {
                if (P.stage >= PR.stage_end) { finished[p] = 1; --remaining; progress = true; break; }
                const StageRef& R = S.stages[P.stage];
                const int* m = S.meta.data() + R.meta_off;
                const int* extl = m + m[5];
                const int* wl = m + m[6];
                const int* items = m + m[7];
                const int* codes = m + m[10];
                const int ext_base = m[8];
                if (!P.started) {
                    int nitems = 0;
                    for (int w = 0; w < NW; ++w) {
                        Warp& c = P.w[w];
                        c.t = wl[w]; c.bars = 0; c.loaded = c.in_tail = c.done = c.at_bar = false;
                        int bars = wl[NW + 1 + w];
                        for (int t = wl[w]; t < wl[w + 1]; ++t, ++nitems) bars += (items[4 * t + 1] >> 16) & 0xffff;
                        if (bars != m[9]) throw std::runtime_error("emulate: warps disagree on the barrier count of a stage");
                    }
                    if (nitems != m[1]) throw std::runtime_error("emulate: work lists do not cover the records of the stage");
                    if (m[4] > EW) throw std::runtime_error("emulate: external rows of a stage exceed the ring");
                    if ((long long) m[1] * NF * 32 != R.vals_doubles) throw std::runtime_error("emulate: value blob size mismatch");
                    P.arrivals = 0;
                    P.started = true;
                }
                bool local = false;
                int ndone = 0;
                for (int w = 0; w < NW; ++w) {
                    Warp& c = P.w[w];
                    if (c.done) { ++ndone; continue; }
                    if (c.at_bar) continue;
                    if (!c.in_tail && !c.loaded) {
                        if (c.t == wl[w + 1]) { c.in_tail = true; c.bars = wl[NW + 1 + w]; }
                        else { c.bars = (items[4 * c.t + 1] >> 16) & 0xffff; c.loaded = true; }
                    }
                    if (c.bars > 0) { c.at_bar = true; P.arrivals++; local = true; continue; }
                    if (c.in_tail) { c.done = true; ++ndone; local = true; continue; }
                    const int* it = items + 4 * c.t;
                    const int count = (it[1] >> 4) & 15, first = it[1] & 1, last = (it[1] >> 1) & 1, need = it[2], g0 = it[3] / 3;
                    if (it[0] != 24 * (g0 - R.g_lo) || it[3] != 3 * g0) throw std::runtime_error("emulate: bad record item");
                    bool ready = true;
                    for (int x = 0; x < need && ready; ++x) ready = !std::isnan(out[3 * (size_t) extl[x]]);
                    if (!ready) continue;                     // the helper warp has not delivered this group yet
                    const double* v = vals.data() + R.vals_off + (size_t) c.t * NF * 32;
                    const int* cd = codes + 128 * c.t;
                    for (int l = 0; l < 3 * count; ++l) {
                        const int q = l / 3, comp = l - 3 * q;
                        const int g = lower ? g0 + q : g0 - q;
                        if (g < R.g_lo || g >= R.g_lo + R.g_rows) throw std::runtime_error("emulate: row outside the stage's rhs window");
                        double acc;
                        if (!first) acc = c.carry[l];
                        else if (lower) acc = rhs[3 * (size_t) g + comp];
                        else acc = v[sweep_vidx(false, 9, l)] * rhs[3 * (size_t) g] + v[sweep_vidx(false, 10, l)] * rhs[3 * (size_t) g + 1] + v[sweep_vidx(false, 11, l)] * rhs[3 * (size_t) g + 2];
                        for (int j = 0; j < 3; ++j) {
                            if (cd[4 * l + j] % (8 * kXwinStride)) throw std::runtime_error("emulate: bad dependency code");
                            const int code = cd[4 * l + j] / (8 * kXwinStride);
                            if (code < 0 || code > zrow) throw std::runtime_error("emulate: bad dependency code");
                            const double* x;
                            if (code >= W && code < zrow) {      // parked external row: map the ring slot back to the list
                                const int k = (code - W - ext_base) & (EW - 1);
                                if (k >= need) throw std::runtime_error("emulate: external row beyond ext_need");
                                x = out + 3 * (size_t) extl[k];
                            } else x = P.xwin.data() + kXwinStride * (size_t) code;
                            for (int e = 0; e < 3; ++e) {
                                if (std::isnan(x[e])) throw std::runtime_error("emulate: read of a value that was not produced yet");
                                acc -= v[sweep_vidx(lower, 3 * j + e, l)] * x[e];
                            }
                        }
                        c.carry[l] = acc;
                    }
                    for (int l = 0; l < 32; ++l) {
                        const int wout = cd[4 * l + 3];
                        if ((wout >= 0) != (last && l < 3 * count)) throw std::runtime_error("emulate: store flag mismatch");
                        if (wout < 0) continue;
                        const int q = l / 3, comp = l - 3 * q;
                        const int g = lower ? g0 + q : g0 - q;
                        P.xwin[(size_t) wout / 8] = c.carry[l];
                        out[3 * (size_t) g + comp] = c.carry[l];
                    }
                    c.t++; c.loaded = false;
                    local = true;
                }
                if (P.arrivals == NW) {
                    for (int w = 0; w < NW; ++w) { P.w[w].at_bar = false; P.w[w].bars--; }
                    P.arrivals = 0;
                    local = true;
                }
                if (ndone == NW) { P.stage++; P.started = false; progress = true; continue; }
                if (!local) break;          // every warp waits for another part: yield
                progress = true;
            }
        }
        if (!progress) return false;
    }
    return true;
}

}  // namespace b200
