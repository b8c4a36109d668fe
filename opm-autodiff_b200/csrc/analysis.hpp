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
// A sweep (forward over L, backward over U) is cut, per part, into STAGES: a few KB of consecutive
// work the producer warp of the CTA fetches with three bulk copies (meta ints, factor values, rhs
// rows) into one slot of a shared-memory ring.  A stage holds CHUNKS of <= 10 rows of one level
// (3 lanes per row).  Chunk c of a level belongs to consumer warp c % W; every warp walks its own
// static WORK LIST and passes every level barrier of the stage.
//
// Shared-memory value space of a part ("xwin", rows of 3 doubles): [0, window) the most recent rows of
// the part (position in processing order & (window - 1)), [window, window + extWindow) a ring of rows
// owned by other parts, parked there by helper warps, and one all-zero row at window + extWindow that
// padded dependency slots point to.  A dependency code is simply the xwin row to read.
//
//   meta blob (ints, 16-byte multiple):
//     header  [0] ngroups [1] nchunks [2] g_lo (even-aligned first p-row of the rhs copy) [3] rhs rows copied
//             [4] next (external rows of the stage) [5] off_ext [6] off_wl [7] off_items
//             [8] ext_base (external rows of the part before this stage) [9] level barriers in the stage
//     groups  : at 12, ngroups cumulative ends: the external rows [0, end) of the stage list are needed by the
//               levels up to the group's; a helper warp delivers them group by group
//     cols    : per chunk nd x count ints, xwin rows (see above)
//     ext     : at off_ext, next p-rows (rows of other parts, or of this part beyond the window), in group order;
//               row k of the list is parked at xwin row window + ((ext_base + k) & (extWindow - 1))
//     wl      : at off_wl, W + 1 item offsets (per warp, in items) then W trailing barrier counts
//     items   : at off_items (16-byte aligned), one int4 per chunk, grouped by warp in processing order:
//               {g0, count | nd << 4 | barriers to pass first << 16, cols_off | vals_off << 16, wpos0 | ext_need << 16}
//               row q of the chunk is p-row g0 + q (lower sweep) or g0 - q (upper sweep) and is written to
//               xwin row (wpos0 + q) & (window - 1); ext_need: external rows [0, ext_need) must be parked
//   vals blob (doubles, 16-byte multiple): per chunk (nd [+1 inverse pivot for U]) x 3 x (3 count) doubles,
//     value ((j*3 + v) * 3 count + 3 q + comp) = LU[block j of row q][comp][v]  (lane-major: conflict-free, coalesced)
struct StageRef {
    long long meta_off;    // ints into SweepPlan::meta
    long long vals_off;    // doubles into the sweep's value stream (even)
    int meta_ints, vals_doubles;
    int g_lo, g_rows;      // rhs copy: rows [g_lo, g_lo + g_rows), both even
};
struct PartRef { int stage_begin, stage_end, row0, nrows; };
struct BuildRef {          // one per chunk: where the factor values of the chunk come from
    long long vals_off;    // doubles into the value stream
    int src_off;           // into SweepPlan::src: nd_eff x count p-space block indices (-1: padding)
    int count, nd_eff, pad;
};

struct SweepPlan {
    std::vector<int> meta;
    std::vector<StageRef> stages;
    std::vector<PartRef> parts;
    std::vector<BuildRef> build;
    std::vector<int> src;
    long long nvals = 0;
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
    // triangular sweeps
    int nparts = 0, nlines = 0, window = 0, warps = 8, extWindow = 1024;
    int nstrips = 0;
    std::vector<int> partPtr;         // nparts + 1, p-space rows
    std::vector<int> partMaxStep;     // rows in the largest level step of each part
    SweepPlan L, U;
    int64_t nnzL = 0;                 // strictly lower blocks
    int maxRowLen = 0;
};

struct AnalysisOptions {
    int parts = 148;            // resident CTAs of the sweep kernels
    int stageBytes = 16384;     // meta + values + rhs of one ring slot
    int window = 2048;          // rows of the part kept in the shared-memory window (power of two)
    int warps = 8;              // consumer warps of a sweep CTA (the static work lists are cut for this many)
    int extWindow = 1024;       // rows of the external-row ring (power of two); bounds ring slots x external rows per stage
};

namespace detail {

inline void build_sweep(const Analysis& A, const int* rows, const int* cols, const std::vector<int>& glev,
                        const std::vector<int>& partOf, bool lower, const AnalysisOptions& opt, SweepPlan& S)
{
    const int W = A.window, EW = A.extWindow, NWc = A.warps;
    const int zrow = W + EW;
    S.parts.resize(A.nparts);
    // cols of a TmpChunk: >= 0 xwin row of the window, kPadCol padding, otherwise -(p-row + 1) of an external row
    struct TmpChunk { int g0, count, nd, level, ps0; std::vector<int> cols, src; };
    for (int p = 0; p < A.nparts; ++p) {
        const int row0 = A.partPtr[p], nrows = A.partPtr[p + 1] - row0;
        const int slack = W - A.partMaxStep[p];
        S.parts[p].stage_begin = (int) S.stages.size();
        S.parts[p].row0 = row0;
        S.parts[p].nrows = nrows;
        std::vector<TmpChunk> st;       // pending stage
        int st_bytes = 0, st_glo = 0, st_ghi = 0;
        int prev_level = -1;            // level of the last chunk emitted in this part
        long long ext_base = 0;         // external rows of the part before the pending stage
        auto flush = [&]() {
            if (st.empty()) return;
            const int nch = (int) st.size();
            // external rows of the stage, listed once, grouped by the first level that needs them
            std::vector<int> ext, gend, need(nch, 0);
            std::vector<std::vector<int>> wl(NWc);          // chunk indices per warp
            std::vector<std::vector<int>> wbar(NWc);        // barriers before each of them
            std::vector<int> pend(NWc, 0);
            int nbar_total = 0;
            {
                std::vector<std::pair<int, int>> seen;      // (p-row, index), kept sorted
                for (int c = 0; c < nch;) {
                    int e = c;
                    while (e < nch && st[e].level == st[c].level) ++e;
                    if (prev_level >= 0 && st[c].level != prev_level) { for (int w = 0; w < NWc; ++w) pend[w]++; nbar_total++; }
                    prev_level = st[c].level;
                    for (int k = c; k < e; ++k) {
                        bool any = false;
                        for (int& code : st[k].cols) {
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
                        const int w = (k - c) % NWc;
                        wl[w].push_back(k);
                        wbar[w].push_back(pend[w]);
                        pend[w] = 0;
                    }
                    if (gend.empty() || (int) ext.size() != gend.back()) { if (!ext.empty()) gend.push_back((int) ext.size()); }
                    for (int k = c; k < e; ++k) if (need[k]) need[k] = (int) ext.size();
                    c = e;
                }
            }
            for (auto& c : st) for (int& code : c.cols) if (code == kPadCol) code = zrow;
            const int next = (int) ext.size(), ngroups = (int) gend.size();
            if (next >= 65536 || nch >= 65536 || nbar_total >= 32768 || next > EW) throw std::runtime_error("sweep stage too large");
            const int off_cols = 12 + ngroups;
            int ncols = 0;
            for (auto& c : st) ncols += c.nd * c.count;
            const int off_ext = off_cols + ncols;
            const int off_wl = off_ext + next;
            const int off_items = (off_wl + 2 * NWc + 1 + 3) & ~3;
            const int meta_ints = off_items + 4 * nch;
            StageRef R{};
            R.meta_off = (long long) S.meta.size();
            R.meta_ints = meta_ints;
            R.vals_off = S.nvals;
            const int glo_al = st_glo & ~1, ghi_al = (st_ghi + 1) & ~1;
            R.g_lo = glo_al;
            R.g_rows = ghi_al - glo_al;
            S.meta.resize(S.meta.size() + meta_ints, 0);
            int* m = S.meta.data() + R.meta_off;
            m[0] = ngroups; m[1] = nch; m[2] = glo_al; m[3] = R.g_rows; m[4] = next; m[5] = off_ext; m[6] = off_wl; m[7] = off_items;
            m[8] = (int) (ext_base & (EW - 1)); m[9] = nbar_total;
            std::copy(gend.begin(), gend.end(), m + 12);
            std::vector<int> cols_off(nch), vals_off(nch);
            int co = off_cols;
            long long vo = 0;
            for (int c = 0; c < nch; ++c) {
                const TmpChunk& t = st[c];
                const int nd_eff = t.nd + (lower ? 0 : 1);
                if (co >= 65536 || vo >= 65536) throw std::runtime_error("sweep stage too large for the packed chunk offsets");
                cols_off[c] = co; vals_off[c] = (int) vo;
                std::copy(t.cols.begin(), t.cols.end(), m + co);
                co += t.nd * t.count;
                BuildRef B{};
                B.vals_off = R.vals_off + vo;
                B.src_off = (int) S.src.size();
                B.count = t.count;
                B.nd_eff = nd_eff;
                S.src.insert(S.src.end(), t.src.begin(), t.src.end());
                S.build.push_back(B);
                vo += (long long) nd_eff * 9 * t.count;
            }
            std::copy(ext.begin(), ext.end(), m + off_ext);
            {
                int* w0 = m + off_wl;
                int o = 0;
                for (int w = 0; w < NWc; ++w) {
                    w0[w] = o;
                    for (size_t t = 0; t < wl[w].size(); ++t, ++o) {
                        const int c = wl[w][t];
                        const TmpChunk& ch = st[c];
                        int* it = m + off_items + 4 * o;
                        it[0] = ch.g0;
                        it[1] = ch.count | (ch.nd << 4) | (wbar[w][t] << 16);
                        it[2] = cols_off[c] | (vals_off[c] << 16);
                        it[3] = (ch.ps0 & (W - 1)) | (need[c] << 16);
                    }
                    w0[NWc + 1 + w] = pend[w];
                }
                w0[NWc] = o;
            }
            vo = (vo + 1) & ~1LL;
            if (vo > INT_MAX || S.src.size() > (size_t) INT_MAX) throw std::runtime_error("sweep stage too large");
            R.vals_doubles = (int) vo;
            S.nvals += vo;
            S.maxMetaInts = std::max(S.maxMetaInts, meta_ints);
            S.maxValsDoubles = std::max(S.maxValsDoubles, R.vals_doubles);
            S.maxRhsRows = std::max(S.maxRhsRows, R.g_rows);
            S.maxExtRows = std::max(S.maxExtRows, next);
            S.nExtRows += next;
            S.nchunks += nch;
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
                TmpChunk t;
                t.count = std::min(kRowsPerWarp, end - s);
                t.g0 = g_of(s);
                t.ps0 = s;
                t.level = lev;
                t.nd = 0;
                for (int q = 0; q < t.count; ++q) {
                    const int r = A.perm[g_of(s + q)];
                    int n = 0;
                    for (int k = rows[r]; k < rows[r + 1]; ++k) n += lower ? cols[k] < r : cols[k] > r;
                    t.nd = std::max(t.nd, n);
                }
                if (t.nd > 0xfff) throw std::runtime_error("block row too long for the sweep chunk descriptor");
                const int nd_eff = t.nd + (lower ? 0 : 1);
                t.cols.assign((size_t) t.nd * t.count, kPadCol);
                t.src.assign((size_t) nd_eff * t.count, -1);
                int nextc = 0;
                for (int q = 0; q < t.count; ++q) {
                    const int ps = s + q, g = g_of(ps), r = A.perm[g];
                    int j = 0;
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
                            if (code < 0) { S.nExternal++; nextc++; }
                            t.cols[(size_t) j * t.count + q] = code;
                            t.src[(size_t) j * t.count + q] = A.prow[g] + (k - rows[r]);
                            ++j;
                        }
                    }
                    if (!lower) t.src[(size_t) t.nd * t.count + q] = A.pdiag[g];
                }
                const int bytes = 16 + 4 * t.nd * t.count + 72 * nd_eff * t.count + 24 * t.count + 8 + 4 * nextc + 4;
                if (!st.empty() && st_bytes + bytes > opt.stageBytes) flush();
                if (st.empty()) { st_glo = INT_MAX; st_ghi = 0; st_bytes = 160; }
                const int glo = lower ? t.g0 : t.g0 - t.count + 1, ghi = lower ? t.g0 + t.count : t.g0 + 1;
                st_glo = std::min(st_glo, glo); st_ghi = std::max(st_ghi, ghi);
                st_bytes += bytes;
                st.push_back(std::move(t));
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
    if (opt.window < 64 || (opt.window & (opt.window - 1))) throw std::runtime_error("window must be a power of two >= 64");
    A.window = opt.window;
    if (opt.warps < 1 || opt.warps > 14) throw std::runtime_error("consumer warps must be in 1..14");
    A.warps = opt.warps;
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
    for (int p = 0; p < nparts; ++p) {
        int run = 0;
        for (int q = A.partPtr[p]; q < A.partPtr[p + 1]; ++q) {
            run = (q > A.partPtr[p] && glev[A.perm[q]] == glev[A.perm[q - 1]]) ? run + 1 : 1;
            A.partMaxStep[p] = std::max(A.partMaxStep[p], run);
        }
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
    detail::build_sweep(A, rows, cols, glev, partOf, true, opt, A.L);
    detail::build_sweep(A, rows, cols, glev, partOf, false, opt, A.U);
    return A;
}

// ---- host emulation of the sweep kernels (schedule verification without a GPU) ---------------------
//
// Interprets the packed streams exactly as k_sweep does, one chunk at a time, round-robin over the
// parts; a chunk whose out-of-window dependency has not been produced yet makes its part yield.
// Returns false if a full round makes no progress (the schedule would deadlock on the device).
inline void fill_stream_host(const SweepPlan& S, const double* LU, std::vector<double>& vals)
{
    vals.assign((size_t) std::max<long long>(S.nvals, 1), 0.0);
    for (const BuildRef& B : S.build)
        for (int j = 0; j < B.nd_eff; ++j)
            for (int q = 0; q < B.count; ++q) {
                const int k = S.src[B.src_off + j * B.count + q];
                for (int comp = 0; comp < 3; ++comp)
                    for (int v = 0; v < 3; ++v)
                        vals[B.vals_off + (size_t) (j * 3 + v) * 3 * B.count + 3 * q + comp] = k >= 0 ? LU[(size_t) k * 9 + comp * 3 + v] : 0.0;
            }
}

inline bool emulate_sweep(const Analysis& A, const SweepPlan& S, bool lower, const std::vector<double>& vals, const double* rhs,
                          double* out, double relax)
{
    const int W = A.window, EW = A.extWindow, NW = A.warps, zrow = W + EW;
    const double NaN = std::nan("");
    for (int i = 0; i < 3 * A.Nb; ++i) out[i] = NaN;
    // One cursor per consumer warp, exactly the control flow of k_sweep: walk the work list, pass `nbar` level
    // barriers (all W warps must arrive), wait for the parked external rows, compute, finally the trailing barriers.
    struct Warp { int t, bars, tail_left; bool loaded, in_tail, done, at_bar; };
    struct Part { int stage; std::vector<Warp> w; int arrivals; std::vector<double> xwin; bool started; };
    std::vector<Part> parts(A.nparts);
    for (int p = 0; p < A.nparts; ++p) {
        parts[p].stage = S.parts[p].stage_begin;
        parts[p].w.assign(NW, Warp{0, 0, 0, false, false, false, false});
        parts[p].arrivals = 0;
        parts[p].xwin.assign((size_t) 3 * (zrow + 1), NaN);
        for (int e = 0; e < 3; ++e) parts[p].xwin[(size_t) 3 * zrow + e] = 0.0;
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
                const int ext_base = m[8];
                if (!P.started) {
                    int nitems = 0;
                    for (int w = 0; w < NW; ++w) {
                        P.w[w] = Warp{wl[w], 0, 0, false, false, false, false};
                        int bars = wl[NW + 1 + w];
                        for (int t = wl[w]; t < wl[w + 1]; ++t, ++nitems) bars += (items[4 * t + 1] >> 16) & 0xffff;
                        if (bars != m[9]) throw std::runtime_error("emulate: warps disagree on the barrier count of a stage");
                    }
                    if (nitems != m[1]) throw std::runtime_error("emulate: work lists do not cover the chunks of the stage");
                    if (m[4] > EW) throw std::runtime_error("emulate: external rows of a stage exceed the ring");
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
                    const int g0 = it[0], count = it[1] & 15, nd = (it[1] >> 4) & 0xfff;
                    const int co = it[2] & 0xffff, vo = (int) ((unsigned) it[2] >> 16);
                    const int wpos0 = it[3] & 0xffff, need = (int) ((unsigned) it[3] >> 16);
                    bool ready = true;
                    for (int x = 0; x < need && ready; ++x) ready = !std::isnan(out[3 * (size_t) extl[x]]);
                    if (!ready) continue;                     // the helper warp has not delivered this group yet
                    const double* v = vals.data() + R.vals_off + vo;
                    double res[kRowsPerWarp][3];
                    for (int q = 0; q < count; ++q) {
                        const int g = lower ? g0 + q : g0 - q;
                        if (g < R.g_lo || g >= R.g_lo + R.g_rows) throw std::runtime_error("emulate: row outside the stage's rhs window");
                        double acc[3];
                        for (int comp = 0; comp < 3; ++comp) {
                            double a = rhs[3 * (size_t) g + comp];
                            for (int j = 0; j < nd; ++j) {
                                const int code = m[co + j * count + q];
                                if (code < 0 || code > zrow) throw std::runtime_error("emulate: bad dependency code");
                                const double* x;
                                if (code >= W && code < zrow) {      // parked external row: map the ring slot back to the list
                                    const int k = (code - W - ext_base) & (EW - 1);
                                    if (k >= need) throw std::runtime_error("emulate: external row beyond ext_need");
                                    x = out + 3 * (size_t) extl[k];
                                } else x = P.xwin.data() + 3 * (size_t) code;
                                for (int e = 0; e < 3; ++e) {
                                    if (std::isnan(x[e])) throw std::runtime_error("emulate: read of a value that was not produced yet");
                                    a -= v[(size_t) (j * 3 + e) * 3 * count + 3 * q + comp] * x[e];
                                }
                            }
                            acc[comp] = a;
                        }
                        for (int comp = 0; comp < 3; ++comp) {
                            double r = acc[comp];
                            if (!lower) {
                                r = 0.0;
                                for (int e = 0; e < 3; ++e) r += v[(size_t) (nd * 3 + e) * 3 * count + 3 * q + comp] * acc[e];
                                r *= relax;
                            }
                            res[q][comp] = r;
                        }
                    }
                    for (int q = 0; q < count; ++q) {
                        const int g = lower ? g0 + q : g0 - q;
                        for (int comp = 0; comp < 3; ++comp) {
                            P.xwin[3 * (size_t) ((wpos0 + q) & (W - 1)) + comp] = res[q][comp];
                            out[3 * (size_t) g + comp] = res[q][comp];
                        }
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
