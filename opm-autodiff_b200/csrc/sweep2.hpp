// sweep2.hpp -- host-side schedule of the round-2 triangular sweeps (k_sweep2, kernels.cuh) + a host emulation of the kernel.
//
// What changed against the round-1 sweep (analysis.hpp build_sweep / kernels.cuh k_sweep) and why (measured on B200 with
// tools/microbench/sweep2_proto.cu, fp64lat.cu, dsmem_pingpong.cu, pollrate.cu and the traces of tools/s2_trace.py; numbers in
// DESIGN.md):
//   * a lone warp issues ~1 instruction per 4-5 cycles, so the time of a level was the INSTRUCTION COUNT of a record (operand
//     fetch + dependent part + stores, ~500 cycles for a 64-row level whatever the layout).  Now the 32-row chunks of consecutive
//     levels ("steps") of a part go ROUND ROBIN to the consumer warps: while the warps of step l run its dependent part (barrier, 6
//     shared loads, 27 fma, 2 shared stores), the others store their results and fetch the operands of their next chunk, several
//     steps ahead.  Step-to-step hand-over is one named barrier: bar.arrive by the warps that finished step l - 1, bar.sync by
//     the warps that run step l.
//   * a lane owns a block ROW (27 factor values in registers, no redundant reads of the dependencies);
//   * the factor values no longer pass through a shared-memory ring (TMA write + LDS read = twice through the 128 B/cycle port):
//     each warp streams ITS OWN records straight from global memory into registers right after it has finished a chunk -- (warps
//     / chunks per step) step times ahead of their use, which covers the HBM latency -- so the stream of a part is one
//     independent sequential stream per consumer warp;
//   * records carry exactly the rows they hold (no lane padding): [pair k][row] 16-byte value pairs, 16 bytes of codes per row,
//     one 16-byte header per record in a separate array (fetched 32 records at a time, one per lane);
//   * NO helper warps and no external-row ring.  Rows owned by other parts still travel through L2 under the NaN-sentinel
//     protocol of round 1 (the value is its own ready flag), but the lane that needs a row fetches it ITSELF, together with its
//     operands, several steps ahead, and folds its contribution into the start value of the row BEFORE it waits for the previous
//     step -- so an external dependency costs nothing inside the dependent chain, and a row that is not there yet is simply
//     polled again by the lane that needs it (a helper warp that polled, parked and published for everybody needed 1200-2000
//     cycles per round, and every round sat between the producer and the consumer of a face row).
// The mathematics, the parts (pencils), p-space and the shared-memory window of recent rows are those of round 1 (analysis.hpp).
// Reference semantics: ParallelOverlappingILU0.hpp:867-901.
#pragma once
#include "analysis.hpp"

namespace b200 {

constexpr int kS2Barriers = 15;        // named barriers 1..15: barrier 1 + l % 15 separates step l from step l + 1
constexpr int kS2MaxWarps = 15;        // consumer warps of a CTA (<= kS2Barriers: a barrier is reused only after every warp has passed it)

// record header (one int2):
//   x  first p-space row of the record (rows x + q, lower sweep, or x - q, upper sweep), 29 bits | lead << 29 (the warp's first
//      record of the step, chunk 0: it counts the part's steps) | low two bits of `gap` << 30 (gap = steps since the warp's
//      previous record, at most 15: the warp tracks the step of its record with it)
//   y  rows (6 bits) | flags << 6 | barrier to wait on << 12 | barrier to arrive at << 16 | warps on the barrier waited on << 20 |
//      warps on the barrier arrived at << 25 | high two bits of `gap` << 30
// codes (one int4 per LANE): {d0 | d1 << 16, d2 | out << 16, extA, extB}
//   d_j = 8 x window slot of dependency j (8 x window = the all-zero row: no dependency, or an external one), out = 8 x window
//   slot of the result; extA / extB = p-space row of an EXTERNAL dependency (a row of another part, or of this part beyond the
//   window) whose factor block sits in slot 2 / slot 1 of the lane, -1 = none;
//   the low three bits of d0, d1, d2 (always zero in a slot code) carry meta bits 0-2, 3-5, 6-8:
//   meta = row of the lane within the record (5 bits) | position of the lane within its row << 5 | further lanes of the row << 7.
// A block row with more than three dependencies (or more than two external ones) takes SEVERAL LANES of the record (at most
// kS2RowLanes): every lane subtracts its own three blocks from its own start value (the rhs for the first lane, zero for the
// others) and the partial sums are added with two warp shuffles before the store.  (Continuation records -- the first version
// -- put the second record's operand fetch, a full memory round trip, on the chain of every level that had such a row: 1.2 us
// instead of 0.25 us per level on the Norne-size system with its fault connections.)  Rows with more than 3 x kS2RowLanes
// dependencies still take continuation records (every lane keeps its partial sum in registers between them).
constexpr int kS2RowLanes = 4;
enum : int { S2_FIRST = 1, S2_LAST = 2, S2_SYNC = 4, S2_ARRIVE = 8, S2_EXT = 16, S2_MULTI = 32 };
// S2_EXT: some lane of the record has an external dependency; S2_MULTI: some row of the record takes several lanes

struct S2Part { int ncw, nsteps, row0, nrows, stream0, pad0, pad1, pad2; };
struct S2Stream { long long vals_off;      // doubles into the sweep's value stream
                  long long code_off;      // int4 units into the code array
                  int hdr_off, nrec; };
struct S2Build { long long vals_off; int src_off, cnt, first, pad; };

struct Sweep2Plan {
    std::vector<S2Part> parts;
    std::vector<S2Stream> streams;
    std::vector<int> hdrs;            // 2 per record
    std::vector<int> codes;           // 4 per lane of a record
    std::vector<S2Build> build;       // one per record
    std::vector<int> src;             // per record: 3 x cnt dependency blocks (p-space block index, -1 none), then cnt pivot blocks (upper); cnt = lanes
    std::vector<int> stepChunks;      // host only (emulation, statistics): per part nsteps entries, chunks (<= 32 lanes) of every step
    std::vector<int> stepPtr;         // nparts + 1 offsets into stepChunks
    long long nvals = 0;
    int npairs = 14;                  // value pairs per row: 14 (27 values + pad) lower, 18 (27 + 9) upper
    long long nrecords = 0, nmulti = 0, nExternal = 0, nWindow = 0, nOwnExternal = 0, nMultiLaneRecords = 0, nLanes = 0;
    std::vector<std::pair<int, int>> partEdges;      // (owner part, reading part) of every external dependency that crosses parts (host only, statistics)
    int maxChunks = 0;
};

struct Sweep2Options {
    int consumerWarps = kS2MaxWarps;  // warps that take records (round robin over the chunks of consecutive steps)
};

inline int s2_pack(int cnt, int flags, int sync_id, int arrive_id, int sync_warps, int arrive_warps)
{
    return cnt | (flags << 6) | (sync_id << 12) | (arrive_id << 16) | (sync_warps << 20) | (arrive_warps << 25);
}
inline int s2_cnt(int y) { return y & 63; }
inline int s2_g0(int x) { return x & 0x1fffffff; }
inline int s2_lead(int x) { return (x >> 29) & 1; }
inline int s2_gap(int x, int y) { return (int) (((unsigned) x >> 30) | (((unsigned) y >> 30) << 2)); }
inline int s2_flags(int y) { return (y >> 6) & 63; }
inline int s2_meta(const int* cd) { return (cd[0] & 7) | (((cd[0] >> 16) & 7) << 3) | ((cd[1] & 7) << 6); }
inline int s2_ord(int meta) { return meta & 31; }
inline int s2_k(int meta) { return (meta >> 5) & 3; }
inline int s2_nsec(int meta) { return (meta >> 7) & 3; }

namespace detail {

// The steps (level sets) of a part are cut into CHUNKS of <= 32 lanes (a lane per row and three dependencies); the chunks of
// consecutive steps go round robin to the consumer warps, so a warp that has finished its chunk of step l fetches the operands
// of its next chunk -- of step l + (warps / chunks per step) -- at once, that many step times before they are needed.  A warp
// has two records of the same step only when a step has more chunks than there are warps, or for rows with more than
// 3 x kS2RowLanes dependencies (continuation records; the partial sums stay in registers).
inline void build_sweep2(const Analysis& A, const int* rows, const int* cols, const std::vector<int>& glev, const std::vector<int>& partOf,
                         bool lower, const Sweep2Options& opt, Sweep2Plan& S)
{
    const int W = A.window, zcode = 8 * W;
    if (8LL * (W + 1) > 65535) throw std::runtime_error("sweep window exceeds the 16-bit slot codes");
    if (A.Nb >= (1 << 29)) throw std::runtime_error("too many block rows for the sweep record headers");
    const int NCW = std::max(1, std::min(opt.consumerWarps, kS2MaxWarps));
    S.npairs = lower ? 14 : 18;
    S.parts.resize(A.nparts);
    S.stepPtr.assign(1, 0);
    for (int p = 0; p < A.nparts; ++p) {
        const int row0 = A.partPtr[p], nrows = A.partPtr[p + 1] - row0;
        const int slack = W - A.partMaxStep[p];
        auto g_of = [&](int ps) { return lower ? row0 + ps : row0 + nrows - 1 - ps; };
        // steps: runs of equal global level in processing order
        std::vector<int> stepPtr(1, 0);
        for (int pos = 1; pos <= nrows; ++pos)
            if (pos == nrows || glev[A.perm[g_of(pos)]] != glev[A.perm[g_of(pos - 1)]]) stepPtr.push_back(pos);
        const int nsteps = (int) stepPtr.size() - 1;
        // pass 0: the dependencies of every row, grouped three to a lane (an external dependency sits in slot 2 of its lane; the order
        // in which the contributions of a row are subtracted is free), and the chunks (<= 32 lanes, whole rows) of every step
        struct Group { int code[3], src[3], ext[2]; };
        struct RowPlan { std::vector<Group> groups; int lanes, passes; };
        std::vector<RowPlan> rplan(nrows);
        for (int ps = 0; ps < nrows; ++ps) {
            const int gq = g_of(ps), r = A.perm[gq];
            std::vector<std::pair<int, int>> win, ext;          // (slot code | p-row, source block)
            for (int k = rows[r]; k < rows[r + 1]; ++k) {
                const int c = cols[k];
                if (!(lower ? c < r : c > r)) continue;
                const int gd = A.iperm[c], sb = A.prow[gq] + (k - rows[r]);
                int slot = -1;
                if (partOf[c] == p) {
                    const int pd = lower ? gd - row0 : row0 + nrows - 1 - gd;
                    if (pd >= ps) throw std::runtime_error("internal: dependency not earlier in processing order");
                    if (ps - pd <= slack) slot = pd & (W - 1);
                }
                if (slot >= 0) { win.emplace_back(8 * slot, sb); S.nWindow++; }
                else { ext.emplace_back(gd, sb); S.nExternal++; if (partOf[c] == p) S.nOwnExternal++; else S.partEdges.emplace_back(partOf[c], p); }
            }
            RowPlan& R = rplan[ps];
            size_t iw = 0, ie = 0;
            do {
                Group g;
                for (int j = 0; j < 3; ++j) { g.code[j] = zcode; g.src[j] = -1; }
                g.ext[0] = g.ext[1] = -1;
                int free_slots = 3;
                if (ie < ext.size()) { g.ext[0] = ext[ie].first; g.src[2] = ext[ie].second; ++ie; free_slots = 2; }
                if (ie < ext.size()) { g.ext[1] = ext[ie].first; g.src[1] = ext[ie].second; ++ie; free_slots = 1; }
                for (int j = 0; j < free_slots && iw < win.size(); ++j, ++iw) { g.code[j] = win[iw].first; g.src[j] = win[iw].second; }
                R.groups.push_back(g);
            } while (iw < win.size() || ie < ext.size());
            R.lanes = std::min(kS2RowLanes, (int) R.groups.size());
            R.passes = ((int) R.groups.size() + R.lanes - 1) / R.lanes;
        }
        std::vector<std::vector<int>> stepChunkStart(nsteps);      // first row (position) of every chunk of the step
        for (int st = 0; st < nsteps; ++st) {
            int lanes = 0;
            for (int ps = stepPtr[st]; ps < stepPtr[st + 1]; ++ps) {
                if (stepChunkStart[st].empty() || lanes + rplan[ps].lanes > 32) { stepChunkStart[st].push_back(ps); lanes = 0; }
                lanes += rplan[ps].lanes;
            }
        }
        auto chunks_of = [&](int st) { return (int) stepChunkStart[st].size(); };
        auto warps_of = [&](int st) { return std::min(chunks_of(st), NCW); };
        S2Part& P = S.parts[p];
        P.ncw = NCW; P.nsteps = nsteps; P.row0 = row0; P.nrows = nrows; P.pad0 = P.pad1 = P.pad2 = 0;
        P.stream0 = (int) S.streams.size();
        // per warp: headers, codes, build refs collected separately, appended stream by stream at the end
        struct WarpStream { std::vector<int> hdr, codes; std::vector<S2Build> build; std::vector<std::vector<int>> src; long long vals = 0; };
        std::vector<WarpStream> ws((size_t) NCW);
        int base = 0;                                   // warp of chunk 0 of the current step
        std::vector<int> lastStep(NCW, -1);             // step of the warp's previous record
        for (int st = 0; st < nsteps; ++st) {
            const int nchunks = chunks_of(st);
            S.stepChunks.push_back(nchunks);
            S.maxChunks = std::max(S.maxChunks, nchunks);
            const int sync_id = 1 + (st + kS2Barriers - 1) % kS2Barriers, arrive_id = 1 + st % kS2Barriers;
            const int sync_warps = st > 0 ? warps_of(st - 1) + warps_of(st) : 0;
            const int arrive_warps = st < nsteps - 1 ? warps_of(st) + warps_of(st + 1) : 0;
            std::vector<std::vector<std::pair<int, int>>> recs(NCW);     // per warp: (chunk, pass)
            std::vector<int> npass(nchunks, 1), chunkEnd(nchunks);
            for (int c = 0; c < nchunks; ++c) {
                chunkEnd[c] = c + 1 < nchunks ? stepChunkStart[st][c + 1] : stepPtr[st + 1];
                for (int ps = stepChunkStart[st][c]; ps < chunkEnd[c]; ++ps) npass[c] = std::max(npass[c], rplan[ps].passes);
                for (int pz = 0; pz < npass[c]; ++pz) recs[(base + c) % NCW].emplace_back(c, pz);
            }
            for (int wv = 0; wv < NCW; ++wv) {
                WarpStream& wsr = ws[wv];
                const int nr = (int) recs[wv].size();
                if (nr > 1) S.nmulti++;
                for (int t = 0; t < nr; ++t) {
                    const int c = recs[wv][t].first, pz = recs[wv][t].second;
                    const int ps0 = stepChunkStart[st][c], ps1 = chunkEnd[c];
                    int cnt = 0;
                    for (int ps = ps0; ps < ps1; ++ps) cnt += rplan[ps].lanes;
                    int flags = 0;
                    if (t == 0 && st > 0) flags |= S2_SYNC;
                    if (t == nr - 1 && st < nsteps - 1) flags |= S2_ARRIVE;
                    if (pz == 0) flags |= S2_FIRST;
                    if (pz == npass[c] - 1) flags |= S2_LAST;
                    const int g0 = g_of(ps0);
                    S2Build B{};
                    B.vals_off = wsr.vals; B.cnt = cnt; B.first = pz == 0; B.src_off = 0;
                    std::vector<int> src((size_t) (lower ? 3 : 4) * cnt, -1);
                    for (int ps = ps0; ps < ps1; ++ps) if (rplan[ps].lanes > 1) flags |= S2_MULTI;
                    int lane = 0;
                    for (int ps = ps0; ps < ps1; ++ps) {
                        const RowPlan& R = rplan[ps];
                        for (int k = 0; k < R.lanes; ++k, ++lane) {
                            Group gz;
                            for (int j = 0; j < 3; ++j) { gz.code[j] = zcode; gz.src[j] = -1; }
                            gz.ext[0] = gz.ext[1] = -1;
                            const int gi = pz * R.lanes + k;
                            const Group& Gr = gi < (int) R.groups.size() ? R.groups[gi] : gz;
                            for (int j = 0; j < 3; ++j) src[(size_t) j * cnt + lane] = Gr.src[j];
                            if (Gr.ext[0] >= 0) flags |= S2_EXT;
                            const int outc = k == 0 ? 8 * (ps & (W - 1)) : zcode;
                            const int meta = (flags & S2_MULTI) ? (ps - ps0) | (k << 5) | ((R.lanes - 1) << 7) : 0;      // (a lane per row otherwise: nothing to say)
                            wsr.codes.push_back((Gr.code[0] | (meta & 7)) | ((Gr.code[1] | ((meta >> 3) & 7)) << 16));
                            wsr.codes.push_back((Gr.code[2] | ((meta >> 6) & 7)) | (outc << 16));
                            wsr.codes.push_back(Gr.ext[0]);
                            wsr.codes.push_back(Gr.ext[1]);
                            if (!lower) src[(size_t) 3 * cnt + lane] = A.pdiag[g_of(ps)];
                        }
                    }
                    if (flags & S2_MULTI) S.nMultiLaneRecords++;
                    S.nLanes += cnt;
                    wsr.build.push_back(B);
                    wsr.src.push_back(std::move(src));
                    wsr.vals += 2LL * S.npairs * cnt;
                    const int gap = std::min(15, st - lastStep[wv]);
                    lastStep[wv] = st;
                    wsr.hdr.push_back((int) ((unsigned) g0 | ((unsigned) (t == 0 && c == 0) << 29) | ((unsigned) (gap & 3) << 30)));
                    wsr.hdr.push_back((int) ((unsigned) s2_pack(cnt, flags, sync_id, arrive_id, sync_warps, arrive_warps) | ((unsigned) (gap >> 2) << 30)));
                    S.nrecords++;
                }
            }
            base = (base + nchunks) % NCW;
        }
        S.stepPtr.push_back((int) S.stepChunks.size());
        // append the warp streams
        for (size_t w = 0; w < ws.size(); ++w) {
            S2Stream R{};
            R.vals_off = S.nvals; R.code_off = (long long) (S.codes.size() / 4); R.hdr_off = (int) (S.hdrs.size() / 2);
            R.nrec = (int) (ws[w].hdr.size() / 2);
            S.hdrs.insert(S.hdrs.end(), ws[w].hdr.begin(), ws[w].hdr.end());
            S.codes.insert(S.codes.end(), ws[w].codes.begin(), ws[w].codes.end());
            for (size_t i = 0; i < ws[w].build.size(); ++i) {
                S2Build B = ws[w].build[i];
                B.vals_off += S.nvals;
                B.src_off = (int) S.src.size();
                S.src.insert(S.src.end(), ws[w].src[i].begin(), ws[w].src[i].end());
                S.build.push_back(B);
            }
            S.nvals += ws[w].vals;
            if (S.hdrs.size() / 2 > (size_t) INT_MAX || S.src.size() > (size_t) INT_MAX) throw std::runtime_error("sweep schedule too large");
            S.streams.push_back(R);
        }
    }
}

}  // namespace detail

inline void build_sweep2_plans(const Analysis& A, const int* rows, const int* cols, const Sweep2Options& opt, Sweep2Plan& L, Sweep2Plan& U)
{
    // global levels of the symmetrised pattern and the part of every natural row, as analyse() computed them
    std::vector<int> glev(A.Nb, 0), partOf(A.Nb, 0);
    for (int l = 0; l < A.nflev; ++l)
        for (int k = A.flevPtr[l]; k < A.flevPtr[l + 1]; ++k) glev[A.perm[A.flevRows[k]]] = l;
    for (int p = 0; p < A.nparts; ++p)
        for (int q = A.partPtr[p]; q < A.partPtr[p + 1]; ++q) partOf[A.perm[q]] = p;
    detail::build_sweep2(A, rows, cols, glev, partOf, true, opt, L);
    detail::build_sweep2(A, rows, cols, glev, partOf, false, opt, U);
}

// value f of a row of a record with cnt rows: pair f / 2 of row q sits at doubles 2 * ((f / 2) * cnt + q) + (f & 1)
//   f = 9 j + 3 c + e : L_j[c][e] (lower) or (D^-1 U_j)[c][e] (upper) of dependency j;  upper only: f = 27 + 3 c + e : (w D^-1)[c][e]
inline size_t s2_vidx(int f, int q, int cnt) { return 2 * ((size_t) (f >> 1) * cnt + q) + (f & 1); }

inline void fill_stream2_host(const Sweep2Plan& S, bool lower, const double* LU, double relax, std::vector<double>& vals)
{
    vals.assign((size_t) std::max<long long>(S.nvals, 1), 0.0);
    for (const S2Build& B : S.build)
        for (int q = 0; q < B.cnt; ++q) {
            double* out = vals.data() + B.vals_off;
            double inv[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (!lower) {
                const int kp = S.src[B.src_off + 3 * B.cnt + q];
                for (int e = 0; e < 9; ++e) inv[e] = LU[(size_t) kp * 9 + e];
                for (int e = 0; e < 9; ++e) out[s2_vidx(27 + e, q, B.cnt)] = B.first ? relax * inv[e] : 0.0;
            }
            for (int j = 0; j < 3; ++j) {
                const int k = S.src[B.src_off + j * B.cnt + q];
                for (int c = 0; c < 3; ++c)
                    for (int e = 0; e < 3; ++e) {
                        double x = 0.0;
                        if (k >= 0) {
                            const double* u = LU + (size_t) k * 9;
                            x = lower ? u[3 * c + e] : inv[3 * c] * u[e] + inv[3 * c + 1] * u[3 + e] + inv[3 * c + 2] * u[6 + e];
                        }
                        out[s2_vidx(9 * j + 3 * c + e, q, B.cnt)] = x;
                    }
            }
        }
}

// Host emulation of k_sweep2: interprets the packed streams exactly as the kernel does (per part: steps in order, the warps of a
// step walk their records), round robin over the parts; a step one of whose external rows has not been produced yet makes its
// part yield.  Checks the invariants the kernel relies on (barrier ids and counts, window slots never read before written, ...).
// Returns false on deadlock.
inline bool emulate_sweep2(const Analysis& A, const Sweep2Plan& S, bool lower, const std::vector<double>& vals, const double* rhs, double* out)
{
    const int W = A.window, NP = S.npairs;
    const double NaN = std::nan("");
    for (int i = 0; i < 3 * A.Nb; ++i) out[i] = NaN;
    struct WarpCur { int rec = 0; long long vals = 0, code = 0; double carry[32][3]; };
    struct PartState { int step = 0, base = 0; std::vector<WarpCur> w; std::vector<double> xs; bool done = false; };
    std::vector<PartState> ps(A.nparts);
    for (int p = 0; p < A.nparts; ++p) {
        const S2Part& P = S.parts[p];
        ps[p].w.resize((size_t) P.ncw);
        for (size_t w = 0; w < ps[p].w.size(); ++w) { ps[p].w[w].vals = S.streams[P.stream0 + w].vals_off; ps[p].w[w].code = S.streams[P.stream0 + w].code_off; }
        ps[p].xs.assign((size_t) 3 * (W + 1), NaN);
        for (int e = 0; e < 3; ++e) ps[p].xs[(size_t) 3 * W + e] = 0.0;
        ps[p].done = P.nsteps == 0;
    }
    int remaining = 0;
    for (auto& s : ps) remaining += !s.done;
    while (remaining > 0) {
        bool progress = false;
        for (int p = 0; p < A.nparts; ++p) {
            PartState& T = ps[p];
            if (T.done) continue;
            const S2Part& P = S.parts[p];
            const int* chunks = S.stepChunks.data() + S.stepPtr[p];
            while (!T.done) {
                const int st = T.step, nch = chunks[st], nw = std::min(nch, P.ncw);
                // are the external rows of the step there?  (walk the step's records without executing them)
                bool ready = true;
                for (int k = 0; k < nw && ready; ++k) {
                    const int wv = (T.base + k) % P.ncw;
                    const S2Stream& R = S.streams[P.stream0 + wv];
                    const WarpCur& c = T.w[wv];
                    if (c.rec >= R.nrec) throw std::runtime_error("emulate2: warp stream ended before the part's last step");
                    {
                        const int* h = S.hdrs.data() + 2 * (size_t) (R.hdr_off + c.rec);
                        const int flags = s2_flags(h[1]);
                        if ((flags & S2_SYNC) != (st > 0 ? S2_SYNC : 0)) throw std::runtime_error("emulate2: SYNC flag mismatch");
                        if (st > 0) {
                            if (((h[1] >> 12) & 15) != 1 + (st - 1) % kS2Barriers) throw std::runtime_error("emulate2: wrong barrier to wait on");
                            if (((h[1] >> 20) & 31) != std::min(chunks[st - 1], P.ncw) + nw) throw std::runtime_error("emulate2: wrong warp count on the barrier waited on");
                        }
                    }
                    long long code = c.code;
                    for (int rec = c.rec; rec < R.nrec && ready; ++rec) {
                        const int* h = S.hdrs.data() + 2 * (size_t) (R.hdr_off + rec);
                        const int cnt = s2_cnt(h[1]), flags = s2_flags(h[1]);
                        bool any = false;
                        for (int q = 0; q < cnt && ready; ++q)
                            for (int j = 0; j < 2; ++j) {
                                const int er = S.codes[4 * (size_t) (code + q) + 2 + j];
                                if (er < -1 || er >= A.Nb) throw std::runtime_error("emulate2: bad external row");
                                if (er >= 0) { any = true; if (std::isnan(out[3 * (size_t) er])) ready = false; }
                            }
                        if (any != ((flags & S2_EXT) != 0)) throw std::runtime_error("emulate2: EXT flag mismatch");
                        code += cnt;
                        if ((flags & S2_ARRIVE) && st < P.nsteps - 1) break;
                    }
                }
                if (!ready) break;      // yield
                // run the step: every warp of the step, its records up to and including the one flagged ARRIVE (last step: all that is left)
                for (int k = 0; k < nw; ++k) {
                    const int wv = (T.base + k) % P.ncw;
                    const S2Stream& R = S.streams[P.stream0 + wv];
                    WarpCur& c = T.w[wv];
                    bool first_rec = true;
                    while (true) {
                        if (c.rec >= R.nrec) {
                            if (st == P.nsteps - 1 && !first_rec) break;
                            throw std::runtime_error("emulate2: warp stream ended inside a step");
                        }
                        const int* h = S.hdrs.data() + 2 * (size_t) (R.hdr_off + c.rec);
                        const int g0 = s2_g0(h[0]), cnt = s2_cnt(h[1]), flags = s2_flags(h[1]);
                        if (!first_rec && (flags & S2_SYNC)) throw std::runtime_error("emulate2: SYNC flag in the middle of a step");
                        if (cnt < 1 || cnt > 32) throw std::runtime_error("emulate2: bad row count");
                        first_rec = false;
                        const double* v = vals.data() + c.vals;
                        const int* cd = S.codes.data() + 4 * (size_t) c.code;
                        bool multi = false;
                        for (int q = 0; q < cnt; ++q) {
                            const int meta = s2_meta(cd + 4 * q), ord = (flags & S2_MULTI) ? s2_ord(meta) : q, k = s2_k(meta), nsec = s2_nsec(meta);
                            if (!(flags & S2_MULTI) && meta) throw std::runtime_error("emulate2: lane meta in a record without multi-lane rows");
                            if (k > nsec || nsec >= kS2RowLanes || (k > 0 && (q == 0 || s2_meta(cd + 4 * (q - 1)) != ((meta & ~(3 << 5)) | ((k - 1) << 5)))))
                                throw std::runtime_error("emulate2: bad lane meta");
                            if (nsec > 0) multi = true;
                            const int gq = lower ? g0 + ord : g0 - ord;
                            if (gq < P.row0 || gq >= P.row0 + P.nrows) throw std::runtime_error("emulate2: row outside the part");
                            double acc[3];
                            if (!(flags & S2_FIRST)) for (int e = 0; e < 3; ++e) acc[e] = c.carry[q][e];
                            else if (k > 0) for (int e = 0; e < 3; ++e) acc[e] = 0.0;
                            else if (lower) for (int e = 0; e < 3; ++e) acc[e] = rhs[3 * (size_t) gq + e];
                            else for (int cc = 0; cc < 3; ++cc) {
                                acc[cc] = 0.0;
                                for (int e = 0; e < 3; ++e) acc[cc] += v[s2_vidx(27 + 3 * cc + e, q, cnt)] * rhs[3 * (size_t) gq + e];
                            }
                            // external dependencies: slot 2 (extA), slot 1 (extB), straight from the result vector
                            const int d[3] = {cd[4 * q] & 0xfff8, (cd[4 * q] >> 16) & 0xfff8, cd[4 * q + 1] & 0xfff8};
                            for (int j = 0; j < 2; ++j) {
                                const int er = cd[4 * q + 2 + j], slot = 2 - j;
                                if (er < 0) continue;
                                if (d[slot] != 8 * W) throw std::runtime_error("emulate2: slot of an external dependency does not point at the zero row");
                                for (int cc = 0; cc < 3; ++cc)
                                    for (int e = 0; e < 3; ++e) acc[cc] -= v[s2_vidx(9 * slot + 3 * cc + e, q, cnt)] * out[3 * (size_t) er + e];
                            }
                            for (int j = 0; j < 3; ++j) {
                                if (d[j] % 8 || d[j] / 8 > W) throw std::runtime_error("emulate2: bad dependency code");
                                const double* x = T.xs.data() + 3 * (size_t) (d[j] / 8);
                                for (int cc = 0; cc < 3; ++cc)
                                    for (int e = 0; e < 3; ++e) {
                                        if (std::isnan(x[e])) throw std::runtime_error("emulate2: read of a value that was not produced yet");
                                        acc[cc] -= v[s2_vidx(9 * j + 3 * cc + e, q, cnt)] * x[e];
                                    }
                            }
                            for (int e = 0; e < 3; ++e) c.carry[q][e] = acc[e];
                        }
                        if (multi != ((flags & S2_MULTI) != 0)) throw std::runtime_error("emulate2: MULTI flag mismatch");
                        // stores happen after every lane of the record has read its dependencies (the kernel: all lanes in lock step);
                        // the lanes of a row add their partial sums as the kernel's two shuffle rounds do
                        if (flags & S2_LAST) {
                            double fin[32][3];
                            for (int q = 0; q < cnt; ++q) for (int e = 0; e < 3; ++e) fin[q][e] = c.carry[q][e];
                            for (int q = 0; q < cnt; ++q) {            // round 1: (0, 1), (2, 3)
                                const int meta = s2_meta(cd + 4 * q), k = s2_k(meta), nsec = s2_nsec(meta);
                                if (!(k & 1) && k + 1 <= nsec) for (int e = 0; e < 3; ++e) fin[q][e] += c.carry[q + 1][e];
                            }
                            for (int q = 0; q < cnt; ++q) {            // round 2: 0 += 2
                                const int meta = s2_meta(cd + 4 * q), k = s2_k(meta), nsec = s2_nsec(meta);
                                if (k == 0 && nsec >= 2) {
                                    const int m2 = s2_meta(cd + 4 * (q + 2)), k2 = s2_k(m2), n2 = s2_nsec(m2);
                                    for (int e = 0; e < 3; ++e) fin[q][e] += c.carry[q + 2][e] + ((!(k2 & 1) && k2 + 1 <= n2) ? c.carry[q + 3][e] : 0.0);
                                }
                            }
                            for (int q = 0; q < cnt; ++q) {
                                const int meta = s2_meta(cd + 4 * q), ord = (flags & S2_MULTI) ? s2_ord(meta) : q, k = s2_k(meta);
                                if (k != 0) continue;
                                const int gq = lower ? g0 + ord : g0 - ord;
                                const int oc = (cd[4 * q + 1] >> 16) & 0xffff;
                                if (oc % 8 || oc / 8 >= W) throw std::runtime_error("emulate2: bad result slot");
                                for (int e = 0; e < 3; ++e) { T.xs[(size_t) 3 * (oc / 8) + e] = fin[q][e]; out[3 * (size_t) gq + e] = fin[q][e]; }
                            }
                        }
                        c.vals += 2LL * NP * cnt; c.code += cnt; c.rec++;
                        if (st == P.nsteps - 1) {
                            if (flags & S2_ARRIVE) throw std::runtime_error("emulate2: ARRIVE flag on the last step");
                            continue;
                        }
                        if (flags & S2_ARRIVE) {
                            if (((h[1] >> 16) & 15) != 1 + st % kS2Barriers) throw std::runtime_error("emulate2: wrong barrier to arrive at");
                            if (((h[1] >> 25) & 31) != nw + std::min(chunks[st + 1], P.ncw)) throw std::runtime_error("emulate2: wrong warp count on the barrier arrived at");
                            break;
                        }
                    }
                }
                T.base = (T.base + nch) % P.ncw;
                T.step++;
                progress = true;
                if (T.step == P.nsteps) {
                    for (size_t w = 0; w < T.w.size(); ++w)
                        if (T.w[w].rec != S.streams[P.stream0 + w].nrec) throw std::runtime_error("emulate2: records left over at the end of a part");
                    T.done = true; --remaining;
                }
            }
        }
        if (!progress) return false;
    }
    return true;
}

}  // namespace b200
