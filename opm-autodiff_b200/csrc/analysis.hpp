// analysis.hpp -- host-side sparsity analysis for the B200 ILU0-BiCGSTAB backend.
//
// Replaces, for this backend, what the reference does in BILU0::init (bda/BILU0.cpp:50-158) and
// cusparseSolverBackend::analyse_matrix (bda/cusparseSolverBackend.cu:348-422): it derives the
// level sets of the lower-triangular dependency DAG (bda/Reorder.cpp:266-318), a symmetric
// permutation P into level order, the permuted BSR pattern of P A P^T and the work-chunk table
// of the triangular-solve kernels.  Done once per sparsity pattern, O(nnzb) work.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace b200 {

constexpr int kRowsPerWarp = 10;   // 3 lanes per block row -> 30 active lanes per warp

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
    // CSC pattern (rows that hold an entry in each column, ascending)
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
    // remaining lower dependencies per row
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
    // A row joins level L when its LAST dependency sits in level L-1.  The reference appends it at
    // the first scanned row for which canBeStarted() holds, i.e. when all dependencies are in
    // levels < L -- it is first *seen ready* while scanning its earliest dependent predecessor of
    // the previous level.  Counting down `pending` over all previous-level predecessors and
    // appending at the first scanned one reproduces that order: every predecessor in the previous
    // level is scanned in this sweep, and predecessors in older levels were counted earlier.
    std::vector<int> seen(Nb, -1);          // level sweep in which the row was first met
    std::vector<int> cand;                  // rows met in this sweep, in first-met order
    std::vector<int> hits(Nb, 0);
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
        // The reference tests canBeStarted(r) at EVERY visit; r becomes eligible in this sweep iff
        // all its dependencies are done, and then it is appended at its first visit.
        for (int r : cand) {
            if (hits[r] == pending[r]) {
                S.fromOrder[next] = r; S.toOrder[r] = next; S.level[r] = lev; ++next;
            }
        }
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

struct Analysis {
    int Nb = 0;
    int64_t nnzb = 0;
    int nlev = 0;
    LevelSchedule sched;              // reference-identical schedule (exported for parity)
    std::vector<int> perm;            // p-space row -> natural row (levels ascending, rows sorted inside a level)
    std::vector<int> iperm;           // natural row -> p-space row
    std::vector<int> levelPtr;        // nlev + 1, in p-space rows
    std::vector<int> prow, pcol;      // pattern of P A P^T (columns ascending)
    std::vector<int> pdiag;           // index of the diagonal block of each p-space row
    std::vector<int> srcblk;          // p-space block -> natural block index
    std::vector<int> chunks;          // trisolve work chunks: start * 16 + count, never crossing a level
    int maxRowLen = 0;
};

inline Analysis analyse(int Nb, const int* rows, const int* cols)
{
    Analysis A;
    A.Nb = Nb;
    A.nnzb = rows[Nb];
    if (rows[0] != 0) throw std::runtime_error("rows[0] must be 0");
    for (int r = 0; r < Nb; ++r) {
        bool diag = false;
        for (int k = rows[r]; k < rows[r + 1]; ++k) {
            if (k > rows[r] && cols[k] <= cols[k - 1]) throw std::runtime_error("columns must be strictly ascending in every row");
            diag |= (cols[k] == r);
        }
        if (!diag) throw std::runtime_error("diagonal block missing in block row " + std::to_string(r));
    }
    A.sched = level_schedule(Nb, rows, cols);
    A.nlev = A.sched.nlev;
    A.levelPtr = A.sched.levelPtr;
    // own ordering: same level sets, rows ascending inside a level (gather locality)
    A.perm.resize(Nb);
    A.iperm.resize(Nb);
    {
        std::vector<int> fill(A.levelPtr.begin(), A.levelPtr.end() - 1);
        for (int r = 0; r < Nb; ++r) A.perm[fill[A.sched.level[r]]++] = r;
        for (int q = 0; q < Nb; ++q) A.iperm[A.perm[q]] = q;
    }
    A.prow.resize((size_t) Nb + 1);
    A.pcol.resize(A.nnzb);
    A.srcblk.resize(A.nnzb);
    A.pdiag.resize(Nb);
    A.prow[0] = 0;
    for (int q = 0; q < Nb; ++q) {
        const int r = A.perm[q];
        A.prow[q + 1] = A.prow[q] + (rows[r + 1] - rows[r]);
        A.maxRowLen = std::max(A.maxRowLen, rows[r + 1] - rows[r]);
    }
    std::vector<std::pair<int, int>> tmp;
    for (int q = 0; q < Nb; ++q) {
        const int r = A.perm[q];
        tmp.clear();
        for (int k = rows[r]; k < rows[r + 1]; ++k) tmp.emplace_back(A.iperm[cols[k]], k);
        std::sort(tmp.begin(), tmp.end());
        int o = A.prow[q];
        for (auto& e : tmp) {
            A.pcol[o] = e.first;
            A.srcblk[o] = e.second;
            if (e.first == q) A.pdiag[q] = o;
            ++o;
        }
    }
    for (int l = 0; l < A.nlev; ++l)
        for (int s = A.levelPtr[l]; s < A.levelPtr[l + 1]; s += kRowsPerWarp)
            A.chunks.push_back(s * 16 + std::min(kRowsPerWarp, A.levelPtr[l + 1] - s));
    if (Nb >= (1 << 27)) throw std::runtime_error("Nb too large for the chunk encoding");
    return A;
}

}  // namespace b200
