// istl_mm.hpp -- reader/writer for the blocked MatrixMarket files Flow dumps when the linear-solver
// verbosity exceeds 10 (ISTLSolverEbos.hpp:245-252 -> WriteSystemMatrixHelper.hpp:31-84; header line
// "% ISTL_STRUCT blocked <rows> <cols>", MatrixMarketSpecializations.hpp:26-59) and the reference's own
// fixtures use (tests/matr33.txt:1-3, tests/rhs3.txt:1-3).  Produces the raw BSR arrays a BdaBridge hands
// to its backend (BdaBridge.cpp:167-189,231-232): columns ascending per row, row-major 3x3 blocks.
#pragma once
#include <algorithm>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace b200mm {

struct Bsr {
    int Nb = 0, bs = 3;
    std::vector<int> rows, cols;
    std::vector<double> vals;
};

inline Bsr read_matrix(const std::string& path)
{
    std::ifstream f(path);
    if (!f) throw std::runtime_error("Could not read matrix file " + path);
    std::string line;
    std::getline(f, line);
    if (line.rfind("%%MatrixMarket matrix coordinate real general", 0) != 0) throw std::runtime_error("not a coordinate MatrixMarket file: " + path);
    int br = 1, bc = 1;
    std::streampos pos = f.tellg();
    while (std::getline(f, line) && !line.empty() && line[0] == '%') {
        std::istringstream is(line);
        std::string pct, tag, kind;
        is >> pct >> tag >> kind;
        if (tag == "ISTL_STRUCT" && kind == "blocked") is >> br >> bc;
        pos = f.tellg();
    }
    if (br != bc) throw std::runtime_error("only square blocks are supported");
    std::istringstream hs(line);
    long n = 0, m = 0, nnz = 0;
    hs >> n >> m >> nnz;
    if (n <= 0 || n != m || n % br) throw std::runtime_error("bad matrix size line in " + path);
    (void) pos;
    std::map<std::pair<int, int>, std::vector<double>> blocks;
    for (long q = 0; q < nnz; ++q) {
        long i, j; double v;
        if (!(f >> i >> j >> v)) throw std::runtime_error("truncated MatrixMarket file " + path);
        --i; --j;
        auto& blk = blocks[{(int) (i / br), (int) (j / bc)}];
        if (blk.empty()) blk.assign((size_t) br * bc, 0.0);
        blk[(size_t) (i % br) * bc + (j % bc)] = v;
    }
    Bsr A;
    A.bs = br; A.Nb = (int) (n / br);
    A.rows.assign((size_t) A.Nb + 1, 0);
    for (auto& kv : blocks) A.rows[(size_t) kv.first.first + 1]++;
    for (int r = 0; r < A.Nb; ++r) A.rows[(size_t) r + 1] += A.rows[r];
    for (auto& kv : blocks) {            // std::map iterates (row, col) ascending
        A.cols.push_back(kv.first.second);
        A.vals.insert(A.vals.end(), kv.second.begin(), kv.second.end());
    }
    return A;
}

inline std::vector<double> read_vector(const std::string& path)
{
    std::ifstream f(path);
    if (!f) throw std::runtime_error("Could not read rhs file " + path);
    std::string line;
    std::getline(f, line);
    if (line.rfind("%%MatrixMarket matrix array real general", 0) != 0) throw std::runtime_error("not an array MatrixMarket file: " + path);
    while (std::getline(f, line) && !line.empty() && line[0] == '%') {}
    std::istringstream hs(line);
    long n = 0, m = 0;
    hs >> n >> m;
    std::vector<double> v((size_t) n * std::max(m, 1L));
    for (auto& x : v) if (!(f >> x)) throw std::runtime_error("truncated vector file " + path);
    return v;
}

inline void write_vector(const std::string& path, const std::vector<double>& v, int bs)
{
    std::ofstream f(path);
    f.precision(17);
    f << "%%MatrixMarket matrix array real general\n% ISTL_STRUCT blocked " << bs << " 1\n" << v.size() << " 1\n";
    for (double x : v) f << x << "\n";
}

}  // namespace b200mm
