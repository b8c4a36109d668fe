// coloring.hpp -- opt-in multi-colour ordering of the block rows (host side, once per pattern).
//
// The reference offers it for its OpenCL ILU0 (`--opencl-ilu-reorder=graph_coloring`, BdaBridge.cpp:72-80, BILU0.cpp:86-91):
// rows are coloured so that no two rows that are connected (A_ij != 0 or A_ji != 0) share a colour, the matrix is permuted
// colour by colour (P A P^T, Reorder.cpp:179-222) and ILU0 is taken of the PERMUTED matrix -- a different, weaker
// preconditioner (tests/test_ref_reorder.py: +30-50 % iterations on the grids here) whose level sets are the colours (~15
// instead of nx + ny + nz).  Never the default here: it is outside BASELINE.json's +-10 % iteration bound.
//
// Restated from Reorder.cpp:58-172 (colorBlockedNodes: Jones-Plassmann with sequential sweeps, one colour per sweep, random
// weights; a row joins the colour of the sweep if none of its neighbours already has that colour and every uncoloured
// neighbour has a strictly smaller weight) with the limits of its call site (BILU0.cpp:89: maxRowsPerColor = maxColsPerColor
// = Nb, counted in scalar rows / newly touched scalar columns, the column marks never reset), and colorsToReordering
// (Reorder.cpp:209-222: colour-major, natural order inside a colour).  The reference seeds its generator from
// std::random_device (Reorder.cpp:35-43), so its colouring differs from run to run; here the seed is an argument and the
// result is reproducible.  Checked on the CPU against the compiled reference for validity and colour counts
// (tests/test_ref_reorder.py).
#pragma once
#include <algorithm>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

namespace b200 {

constexpr int kMaxColors = 256;      // Reorder.hpp:30 MAX_COLORS

struct ColorOrder {
    int ncolors = 0;
    std::vector<int> toOrder;        // natural row -> position in the colour-major order
    std::vector<int> fromOrder;      // position -> natural row
    std::vector<int> rowsPerColor;
};

inline ColorOrder graph_coloring(int Nb, const int* rows, const int* cols, unsigned seed)
{
    // neighbours of a row in either direction (the reference walks the CSR row and the CSC column of the node)
    std::vector<int> nptr((size_t) Nb + 1, 0), nadj;
    {
        std::vector<int> cnt(Nb, 0);
        for (int r = 0; r < Nb; ++r)
            for (int k = rows[r]; k < rows[r + 1]; ++k)
                if (cols[k] != r) { cnt[r]++; cnt[cols[k]]++; }
        for (int r = 0; r < Nb; ++r) nptr[r + 1] = nptr[r] + cnt[r];
        nadj.resize((size_t) std::max(nptr[Nb], 1));
        std::vector<int> fill(nptr.begin(), nptr.end() - 1);
        for (int r = 0; r < Nb; ++r)
            for (int k = rows[r]; k < rows[r + 1]; ++k)
                if (cols[k] != r) { nadj[fill[r]++] = cols[k]; nadj[fill[cols[k]]++] = r; }
    }
    const unsigned bs = 3;
    const unsigned maxRows = (unsigned) Nb, maxCols = (unsigned) Nb;
    std::mt19937 gen(seed);
    std::vector<int> weight(Nb), color(Nb);
    std::vector<char> touched(Nb, 0);
    int left = Nb;
    for (int attempt = 0; attempt < 100; ++attempt) {
        std::uniform_int_distribution<int> uniform{};
        for (int& w : weight) w = uniform(gen);
        std::fill(color.begin(), color.end(), -1);
        for (int c = 0; c < kMaxColors; ++c) {
            unsigned rowsIn = 0, colsIn = 0;
            for (int i = 0; i < Nb; ++i) {
                if (color[i] != -1) continue;
                bool top = true;
                for (int k = nptr[i]; k < nptr[i + 1] && top; ++k) {
                    const int j = nadj[k], jc = color[j];
                    if (jc != -1 && jc != c) continue;                // coloured in an earlier sweep
                    if (jc == c || weight[i] <= weight[j]) top = false;
                }
                if (!top) continue;
                unsigned fresh = 0;
                for (int k = rows[i]; k < rows[i + 1]; ++k)
                    if (!touched[cols[k]]) { touched[cols[k]] = 1; fresh += bs; }
                if (colsIn + fresh > maxCols) break;
                colsIn += fresh;
                color[i] = c;
                rowsIn += bs;
                if (rowsIn + bs - 1 >= maxRows) break;
            }
            left = (int) std::count(color.begin(), color.end(), -1);
            if (left == 0) {
                ColorOrder o;
                o.ncolors = c + 1;
                o.toOrder.resize(Nb); o.fromOrder.resize(Nb); o.rowsPerColor.assign(o.ncolors, 0);
                int pos = 0;
                for (int cc = 0; cc < o.ncolors; ++cc)
                    for (int i = 0; i < Nb; ++i)
                        if (color[i] == cc) { o.rowsPerColor[cc]++; o.toOrder[i] = pos; o.fromOrder[pos] = i; ++pos; }
                return o;
            }
        }
    }
    throw std::runtime_error("graph colouring: no colouring with " + std::to_string(kMaxColors) + " colours after 100 tries (" +
                             std::to_string(left) + " rows left)");
}

// P A P^T of the PATTERN (Reorder.cpp:179-208): row p of the result is row fromOrder[p], its columns renumbered with toOrder
// and sorted ascending; src[k'] = block of the caller's arrays that lands at k' (the values follow on the device).
inline void permute_pattern(int Nb, const int* rows, const int* cols, const ColorOrder& o, std::vector<int>& prows,
                            std::vector<int>& pcols, std::vector<int>& src)
{
    prows.assign((size_t) Nb + 1, 0);
    pcols.resize((size_t) rows[Nb]); src.resize((size_t) rows[Nb]);
    std::vector<std::pair<int, int>> row;
    int out = 0;
    for (int p = 0; p < Nb; ++p) {
        const int r = o.fromOrder[p];
        row.clear();
        for (int k = rows[r]; k < rows[r + 1]; ++k) row.emplace_back(o.toOrder[cols[k]], k);
        std::sort(row.begin(), row.end());
        for (const auto& e : row) { pcols[out] = e.first; src[out] = e.second; ++out; }
        prows[p + 1] = out;
    }
}

}  // namespace b200
