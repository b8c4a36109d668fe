// ref_glue.cpp -- TEST-ONLY extern "C" entry points into the UNMODIFIED reference
// functions of opm/simulators/linalg/bda/Reorder.cpp and BlockedMatrix.cpp, which
// are compiled where they lie under /root/reference (see oracle/Makefile).  Used
// to pin the oracle's and the product's level sets / permutations.  Never shipped.
#include <vector>
#include <cstring>
#include <opm/simulators/linalg/bda/Reorder.hpp>
#include <opm/simulators/linalg/bda/BlockedMatrix.hpp>

extern "C" {

// bda::csrPatternToCsc (Reorder.cpp:333-366) + bda::findLevelScheduling (:266-318).
// rowsPerColor must hold Nb ints.  Returns numColors.
int ref_level_schedule(int Nb, int* rows, int* cols, int* toOrder, int* fromOrder, int* rowsPerColor)
{
    std::vector<int> cscRows(rows[Nb]), cscPtr(Nb + 1), rpc;
    bda::csrPatternToCsc(cols, rows, cscRows.data(), cscPtr.data(), Nb);
    int numColors = 0;
    bda::findLevelScheduling(cols, rows, cscRows.data(), cscPtr.data(), Nb, &numColors, toOrder, fromOrder, rpc);
    for (int i = 0; i < numColors; ++i) rowsPerColor[i] = rpc[i];
    return numColors;
}

// bda::findGraphColoring<3> (Reorder.cpp:322-330); randomly seeded in the reference.
int ref_graph_coloring(int Nb, int* rows, int* cols, int* toOrder, int* fromOrder, int* rowsPerColor)
{
    std::vector<int> cscRows(rows[Nb]), cscPtr(Nb + 1), rpc;
    bda::csrPatternToCsc(cols, rows, cscRows.data(), cscPtr.data(), Nb);
    int numColors = 0;
    bda::findGraphColoring<3>(cols, rows, cscRows.data(), cscPtr.data(), Nb, Nb, Nb, &numColors, toOrder, fromOrder, rpc);
    for (int i = 0; i < numColors; ++i) rowsPerColor[i] = rpc[i];
    return numColors;
}

// bda::reorderBlockedMatrixByPattern<3> (Reorder.cpp:179-208): P A P^T with per-row sort.
void ref_reorder_matrix(int Nb, int nnzb, double* vals, int* cols, int* rows, int* toOrder, int* fromOrder,
                        double* rvals, int* rcols, int* rrows)
{
    bda::BlockedMatrix<3> mat(Nb, nnzb, vals, cols, rows);
    bda::BlockedMatrix<3> rmat(Nb, nnzb, rvals, rcols, rrows);
    bda::reorderBlockedMatrixByPattern<3>(&mat, toOrder, fromOrder, &rmat);
}

// bda::blockMult / blockMultSub (BlockedMatrix.cpp:69-100), 3x3 row-major.
void ref_block_mult(double* a, double* b, double* c) { bda::blockMult<3>(a, b, c); }
void ref_block_mult_sub(double* a, double* b, double* c) { bda::blockMultSub<3>(a, b, c); }

}
