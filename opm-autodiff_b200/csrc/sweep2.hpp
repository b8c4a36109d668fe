// sweep2.hpp -- host-side schedule of the round-2 triangular sweeps (k_sweep2, sweep2.cuh) + a host emulation of the kernel.
//
// What changed against the round-1 sweep (analysis.hpp build_sweep / kernels.cuh k_sweep) and why (measured on B200 with
// tools/microbench/sweep2_proto.cu, fp64lat.cu, dsmem_pingpong.cu; numbers in DESIGN.md):
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
//   * records carry exactly the rows they hold (no lane padding): [pair k][row] 16-byte value pairs, 8 bytes of codes per row,
//     one 16-byte header per record in a separate array (fetched 32 records at a time, one per lane).
// The mathematics, the parts (pencils), p-space, the window / external-row value space and the NaN-sentinel dataflow between
// parts are those of round 1 (analysis.hpp).  Reference semantics: ParallelOverlappingILU0.hpp:867-901.
#pragma once
#include "analysis.hpp"

namespace b200 {

constexpr int kS2Barriers = 15;        // named barriers 1..15: barrier 1 + l % 15 separates step l from step l + 1
constexpr int kS2MaxWarps = 15;        // consumer warps of a CTA (<= kS2Barriers: a barrier is reused only after every warp has passed it)
constexpr int kS2Reuse = 2;            // an external row parked for step m may be read by the steps m .. m + kS2Reuse - 1 (listed once for them)

// record header (one int4):
//   x  first p-space row of the record (rows x + q, lower sweep, or x - q, upper sweep)
//   y  rows | flags << 8 | barrier to wait on << 16 | barrier to arrive at << 20
//   z  external rows needed (prefix of the part's list; 0 = nothing new to wait for) | warps on the barrier waited on << 24
//   w  external rows no step >= this one reads any more (prefix: ring slots free for reuse) | warps on the barrier arrived at << 24
enum : int { S2_FIRST = 1, S2_LAST = 2, S2_SYNC = 4, S2_ARRIVE = 8, S2_LEAD = 16 };

struct S2Part { int ncw, nsteps, row0, nrows, stream0, ext0, next, pad; };
struct S2Stream { long long vals_off;      // doubles into the sweep's value stream
                  long long code_off;      // int2 units into the code array
                  int hdr_off, nrec; };
struct S2Build { long long vals_off; int src_off, cnt, first, pad; };

struct Sweep2Plan {
    std::vector<S2Part> parts;
    std::vector<S2Stream> streams;
    std::vector<int> hdrs;            // 4 per record
    std::vector<int> codes;           // 2 per row of a record: {d0 | d1 << 16, d2 | out << 16}, each = 8 x slot of the value space
    std::vector<int> ext;             // p-rows, per part in order of need
    std::vector<S2Build> build;       // one per record
    std::vector<int> src;             // per record: 3 x cnt dependency blocks (p-space block index, -1 none), then cnt pivot blocks (upper)
    std::vector<int> stepChunks;      // host only (emulation, statistics): per part nsteps entries, 32-row chunks of every step
    std::vector<int> stepPtr;         // nparts + 1 offsets into stepChunks
    long long nvals = 0;
    int npairs = 14;                  // value pairs per row: 14 (27 values + pad) lower, 18 (27 + 9) upper
    long long nrecords = 0, nmulti = 0, nExternal = 0, nWindow = 0, nExtRows = 0;
    int maxChunks = 0;
};

struct Sweep2Options {
    int consumerWarps = kS2MaxWarps;  // warps that take records (round robin over the chunks of consecutive steps)
};

namespace detail {

// The steps (level sets) of a part are cut into CHUNKS of <= 32 rows (a lane per row); the chunks of consecutive steps go round
// robin to the consumer warps, so a warp that has finished its chunk of step l fetches the operands of its next chunk -- of step
// l + (warps / chunks per step) -- at once, that many step times before they are needed.  A warp has two records of the same
// step only when a step has more chunks than there are warps, or for rows with more than three dependencies (continuation
// records; the partial sums stay in registers).
inline void build_sweep2(const Analysis& A, const int* rows, const int* cols, const std::vector<int>& glev, const std::vector<int>& partOf,
                         bool lower, const Sweep2Options& opt, Sweep2Plan& S)
{
    const int W = A.window, EW = A.extWindow, zslot = W + EW;
    if (8LL * (zslot + 1) > 65535) throw std::runtime_error("sweep window + external ring exceed the 16-bit slot codes");
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
        auto chunks_of = [&](int st) { return (stepPtr[st + 1] - stepPtr[st] + 31) / 32; };
        auto warps_of = [&](int st) { return std::min(chunks_of(st), NCW); };
        S2Part& P = S.parts[p];
        P.ncw = NCW; P.nsteps = nsteps; P.row0 = row0; P.nrows = nrows; P.pad = 0;
        P.stream0 = (int) S.streams.size(); P.ext0 = (int) S.ext.size();
        // per warp: headers, codes, build refs collected separately, appended stream by stream at the end
        struct WarpStream { std::vector<int> hdr, codes; std::vector<S2Build> build; std::vector<std::vector<int>> src; long long vals = 0; };
        std::vector<WarpStream> ws((size_t) NCW);
        // external rows: listed in the order the steps need them; a row already listed for one of the last kS2Reuse steps is reused
        std::vector<std::pair<int, int>> seen;          // (p-row, index in the part's list) of the rows that may be reused, sorted by p-row
        std::vector<int> listedBefore(1, 0);            // rows listed before step m
        int next_total = 0;                             // external rows of the part so far
        int base = 0;                                   // warp of chunk 0 of the current step
        for (int st = 0; st < nsteps; ++st) {
            const int pos0 = stepPtr[st], n = stepPtr[st + 1] - pos0;
            const int nchunks = chunks_of(st);
            S.stepChunks.push_back(nchunks);
            S.maxChunks = std::max(S.maxChunks, nchunks);
            {   // rows listed before step st - kS2Reuse + 1 may not be reused any more
                const int keep_from = listedBefore[std::max(0, st - kS2Reuse + 1)];
                size_t o = 0;
                for (size_t i = 0; i < seen.size(); ++i) if (seen[i].second >= keep_from) seen[o++] = seen[i];
                seen.resize(o);
            }
            const int freed = listedBefore[std::max(0, st - kS2Reuse + 1)];
            // pass 1: dependencies of every row of the step; external rows get their list index
            struct RowDeps { std::vector<int> code, src; int piv; };
            std::vector<RowDeps> rd(n);
            for (int q = 0; q < n; ++q) {
                const int ps = pos0 + q, gq = g_of(ps), r = A.perm[gq];
                for (int k = rows[r]; k < rows[r + 1]; ++k) {
                    const int c = cols[k];
                    if (!(lower ? c < r : c > r)) continue;
                    const int gd = A.iperm[c];
                    int slot = -1;
                    if (partOf[c] == p) {
                        const int pd = lower ? gd - row0 : row0 + nrows - 1 - gd;
                        if (pd >= ps) throw std::runtime_error("internal: dependency not earlier in processing order");
                        if (ps - pd <= slack) { slot = pd & (W - 1); S.nWindow++; }
                    }
                    if (slot < 0) {
                        S.nExternal++;
                        auto it = std::lower_bound(seen.begin(), seen.end(), std::make_pair(gd, -1));
                        int idx;
                        if (it != seen.end() && it->first == gd) idx = it->second;
                        else { idx = next_total++; S.ext.push_back(gd); seen.insert(it, std::make_pair(gd, idx)); }
                        slot = W + (idx & (EW - 1));
                    }
                    rd[q].code.push_back(8 * slot);
                    rd[q].src.push_back(A.prow[gq] + (k - rows[r]));
                }
                rd[q].piv = A.pdiag[gq];
            }
            // (a step that adds no external row of its own has nothing new to wait for)
            const int need = next_total > listedBefore[st] ? next_total : 0;
            if (next_total - freed > EW) throw std::runtime_error("external-row ring of the triangular sweeps too small for the rows of " + std::to_string(kS2Reuse) + " steps");
            if (next_total >= (1 << 24)) throw std::runtime_error("too many external rows in one part of the triangular sweeps");
            listedBefore.push_back(next_total);
            // pass 2: records.  Chunk c of the step belongs to warp (base + c) % NCW.
            const int sync_id = 1 + (st + kS2Barriers - 1) % kS2Barriers, arrive_id = 1 + st % kS2Barriers;
            const int sync_warps = st > 0 ? warps_of(st - 1) + warps_of(st) : 0;
            const int arrive_warps = st < nsteps - 1 ? warps_of(st) + warps_of(st + 1) : 0;
            std::vector<std::vector<std::pair<int, int>>> recs(NCW);     // per warp: (chunk, pass)
            std::vector<int> npass(nchunks, 1);
            for (int c = 0; c < nchunks; ++c) {
                const int cnt = std::min(32, n - 32 * c);
                int nd = 0;
                for (int q = 0; q < cnt; ++q) nd = std::max(nd, (int) rd[32 * c + q].code.size());
                npass[c] = std::max(1, (nd + 2) / 3);
                for (int ps = 0; ps < npass[c]; ++ps) recs[(base + c) % NCW].emplace_back(c, ps);
            }
            for (int wv = 0; wv < NCW; ++wv) {
                WarpStream& wsr = ws[wv];
                const int nr = (int) recs[wv].size();
                if (nr > 1) S.nmulti++;
                for (int t = 0; t < nr; ++t) {
                    const int c = recs[wv][t].first, ps = recs[wv][t].second;
                    const int cnt = std::min(32, n - 32 * c);
                    int flags = 0;
                    if (t == 0 && st > 0) flags |= S2_SYNC;
                    if (t == nr - 1 && st < nsteps - 1) flags |= S2_ARRIVE;
                    if (t == 0 && c == 0) flags |= S2_LEAD;
                    if (ps == 0) flags |= S2_FIRST;
                    if (ps == npass[c] - 1) flags |= S2_LAST;
                    const int g0 = g_of(pos0 + 32 * c);
                    S2Build B{};
                    B.vals_off = wsr.vals; B.cnt = cnt; B.first = ps == 0; B.src_off = 0;
                    std::vector<int> src((size_t) (lower ? 3 : 4) * cnt, -1);
                    for (int q = 0; q < cnt; ++q) {
                        const RowDeps& R = rd[32 * c + q];
                        int code[3] = {8 * zslot, 8 * zslot, 8 * zslot};
                        for (int j = 0; j < 3; ++j) {
                            const int d = 3 * ps + j;
                            if (d < (int) R.code.size()) { code[j] = R.code[d]; src[(size_t) j * cnt + q] = R.src[d]; }
                        }
                        const int outc = 8 * ((pos0 + 32 * c + q) & (W - 1));
                        wsr.codes.push_back(code[0] | (code[1] << 16));
                        wsr.codes.push_back(code[2] | (outc << 16));
                        if (!lower) src[(size_t) 3 * cnt + q] = R.piv;
                    }
                    wsr.build.push_back(B);
                    wsr.src.push_back(std::move(src));
                    wsr.vals += 2LL * S.npairs * cnt;
                    wsr.hdr.push_back(g0);
                    wsr.hdr.push_back(cnt | (flags << 8) | (sync_id << 16) | (arrive_id << 20));
                    wsr.hdr.push_back((t == 0 ? need : 0) | (sync_warps << 24));      // the warp's first record of the step waits for the step's external rows
                    wsr.hdr.push_back(freed | (arrive_warps << 24));
                    S.nrecords++;
                }
            }
            base = (base + nchunks) % NCW;
        }
        P.next = next_total;
        S.nExtRows += next_total;
        S.stepPtr.push_back((int) S.stepChunks.size());
        // append the warp streams
        for (size_t w = 0; w < ws.size(); ++w) {
            S2Stream R{};
            R.vals_off = S.nvals; R.code_off = (long long) (S.codes.size() / 2); R.hdr_off = (int) (S.hdrs.size() / 4);
            R.nrec = (int) (ws[w].hdr.size() / 4);
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
            if (S.hdrs.size() / 4 > (size_t) INT_MAX || S.src.size() > (size_t) INT_MAX) throw std::runtime_error("sweep schedule too large");
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
// step walk their records; the helper parks external rows in list order under the ring's flow control), round robin over the
// parts; a step whose external rows have not been produced yet makes its part yield.  Checks the invariants the kernel relies on
// (barrier ids and counts, ring slots, window slots never read before written, ...).  Returns false on deadlock.
inline bool emulate_sweep2(const Analysis& A, const Sweep2Plan& S, bool lower, const std::vector<double>& vals, const double* rhs, double* out)
{
    const int W = A.window, EW = A.extWindow, zslot = W + EW, NP = S.npairs;
    const double NaN = std::nan("");
    for (int i = 0; i < 3 * A.Nb; ++i) out[i] = NaN;
    struct WarpCur { int rec = 0; long long vals = 0, code = 0; double carry[32][3]; };
    struct PartState { int step = 0, base = 0; std::vector<WarpCur> w; std::vector<double> xs; int parked = 0; bool done = false; };
    std::vector<PartState> ps(A.nparts);
    for (int p = 0; p < A.nparts; ++p) {
        const S2Part& P = S.parts[p];
        ps[p].w.resize((size_t) P.ncw);
        for (size_t w = 0; w < ps[p].w.size(); ++w) { ps[p].w[w].vals = S.streams[P.stream0 + w].vals_off; ps[p].w[w].code = S.streams[P.stream0 + w].code_off; }
        ps[p].xs.assign((size_t) 3 * (zslot + 1), NaN);
        for (int e = 0; e < 3; ++e) ps[p].xs[(size_t) 3 * zslot + e] = 0.0;
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
                // which external rows does this step need?  (the first record of every warp of the step)
                int need = 0, freed = -1;
                for (int k = 0; k < nw; ++k) {
                    const int wv = (T.base + k) % P.ncw;
                    const S2Stream& R = S.streams[P.stream0 + wv];
                    const WarpCur& c = T.w[wv];
                    if (c.rec >= R.nrec) throw std::runtime_error("emulate2: warp stream ended before the part's last step");
                    const int* h = S.hdrs.data() + 4 * (size_t) (R.hdr_off + c.rec);
                    const int flags = (h[1] >> 8) & 255;
                    if ((flags & S2_SYNC) != (st > 0 ? S2_SYNC : 0)) throw std::runtime_error("emulate2: SYNC flag mismatch");
                    if (((flags & S2_LEAD) != 0) != (k == 0)) throw std::runtime_error("emulate2: LEAD flag mismatch");
                    if (st > 0) {
                        if (((h[1] >> 16) & 15) != 1 + (st - 1) % kS2Barriers) throw std::runtime_error("emulate2: wrong barrier to wait on");
                        if (((unsigned) h[2] >> 24) != (unsigned) (std::min(chunks[st - 1], P.ncw) + nw)) throw std::runtime_error("emulate2: wrong warp count on the barrier waited on");
                    }
                    const int nd = h[2] & 0xffffff, fr = h[3] & 0xffffff;
                    if (nd) { if (need && nd != need) throw std::runtime_error("emulate2: warps of a step disagree on the external rows"); need = nd; }
                    if (freed >= 0 && fr != freed) throw std::runtime_error("emulate2: warps of a step disagree on the freed prefix");
                    freed = fr;
                }
                if (need) {
                    if (need > P.next) throw std::runtime_error("emulate2: external need beyond the part's list");
                    if (need > freed + EW) throw std::runtime_error("emulate2: ring flow control would deadlock");
                    // park (the helper's job): in list order, as far as the rows have been produced
                    bool ready = true;
                    while (T.parked < need) {
                        const int k = T.parked, gd = S.ext[P.ext0 + k];
                        if (std::isnan(out[3 * (size_t) gd])) { ready = false; break; }
                        for (int e = 0; e < 3; ++e) T.xs[(size_t) 3 * (W + (k & (EW - 1))) + e] = out[3 * (size_t) gd + e];
                        T.parked++;
                    }
                    if (!ready) break;      // yield
                }
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
                        const int* h = S.hdrs.data() + 4 * (size_t) (R.hdr_off + c.rec);
                        const int g0 = h[0], cnt = h[1] & 255, flags = (h[1] >> 8) & 255;
                        if (!first_rec && (flags & (S2_SYNC | S2_LEAD))) throw std::runtime_error("emulate2: SYNC / LEAD flag in the middle of a step");
                        if (!first_rec && (h[2] & 0xffffff)) throw std::runtime_error("emulate2: external wait in the middle of a step");
                        if (cnt < 1 || cnt > 32) throw std::runtime_error("emulate2: bad row count");
                        first_rec = false;
                        const double* v = vals.data() + c.vals;
                        const int* cd = S.codes.data() + 2 * (size_t) c.code;
                        for (int q = 0; q < cnt; ++q) {
                            const int gq = lower ? g0 + q : g0 - q;
                            if (gq < P.row0 || gq >= P.row0 + P.nrows) throw std::runtime_error("emulate2: row outside the part");
                            double acc[3];
                            if (!(flags & S2_FIRST)) for (int e = 0; e < 3; ++e) acc[e] = c.carry[q][e];
                            else if (lower) for (int e = 0; e < 3; ++e) acc[e] = rhs[3 * (size_t) gq + e];
                            else for (int cc = 0; cc < 3; ++cc) {
                                acc[cc] = 0.0;
                                for (int e = 0; e < 3; ++e) acc[cc] += v[s2_vidx(27 + 3 * cc + e, q, cnt)] * rhs[3 * (size_t) gq + e];
                            }
                            const int d[3] = {cd[2 * q] & 0xffff, (cd[2 * q] >> 16) & 0xffff, cd[2 * q + 1] & 0xffff};
                            for (int j = 0; j < 3; ++j) {
                                if (d[j] % 8 || d[j] / 8 > zslot) throw std::runtime_error("emulate2: bad dependency code");
                                const int slot = d[j] / 8;
                                if (slot >= W && slot < zslot) {         // parked external row: must be one of the live rows of the ring
                                    bool live = false;
                                    for (int kk = std::max(freed, T.parked - EW); kk < T.parked && !live; ++kk) live = (kk & (EW - 1)) == slot - W;
                                    if (!live) throw std::runtime_error("emulate2: read of a ring slot that holds no live row");
                                }
                                const double* x = T.xs.data() + 3 * (size_t) slot;
                                for (int cc = 0; cc < 3; ++cc)
                                    for (int e = 0; e < 3; ++e) {
                                        if (std::isnan(x[e])) throw std::runtime_error("emulate2: read of a value that was not produced yet");
                                        acc[cc] -= v[s2_vidx(9 * j + 3 * cc + e, q, cnt)] * x[e];
                                    }
                            }
                            for (int e = 0; e < 3; ++e) c.carry[q][e] = acc[e];
                        }
                        // stores happen after every row of the record has read its dependencies (the kernel: all lanes in lock step)
                        if (flags & S2_LAST)
                            for (int q = 0; q < cnt; ++q) {
                                const int gq = lower ? g0 + q : g0 - q;
                                const int oc = (cd[2 * q + 1] >> 16) & 0xffff;
                                if (oc % 8 || oc / 8 >= W) throw std::runtime_error("emulate2: bad result slot");
                                for (int e = 0; e < 3; ++e) { T.xs[(size_t) 3 * (oc / 8) + e] = c.carry[q][e]; out[3 * (size_t) gq + e] = c.carry[q][e]; }
                            }
                        c.vals += 2LL * NP * cnt; c.code += cnt; c.rec++;
                        if (st == P.nsteps - 1) {
                            if (flags & S2_ARRIVE) throw std::runtime_error("emulate2: ARRIVE flag on the last step");
                            continue;
                        }
                        if (flags & S2_ARRIVE) {
                            if (((h[1] >> 20) & 15) != 1 + st % kS2Barriers) throw std::runtime_error("emulate2: wrong barrier to arrive at");
                            if (((unsigned) h[3] >> 24) != (unsigned) (nw + std::min(chunks[st + 1], P.ncw))) throw std::runtime_error("emulate2: wrong warp count on the barrier arrived at");
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
