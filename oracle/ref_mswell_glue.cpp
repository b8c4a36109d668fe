// ref_mswell_glue.cpp -- TEST-ONLY extern "C" entry point into the UNMODIFIED reference class
// Opm::MultisegmentWellContribution (opm/simulators/linalg/bda/MultisegmentWellContribution.cpp, compiled where it lies
// under /root/reference; see oracle/Makefile, target _ref/libref_mswell.so).  The class calls UMFPACK, which is not in this
// image: the five umfpack_di_* functions it uses are provided below as a dense LU with partial pivoting (an exact solve of
// the same matrix), so that the reference's own B x / C^T z loops and data layout are what the oracle's restatement
// (orc_ms_apply) is pinned against.  Never shipped.
#include <config.h>
#include <cmath>
#include <cstring>
#include <vector>

#include <umfpack.h>
#include <opm/simulators/linalg/bda/MultisegmentWellContribution.hpp>

namespace {
struct Sym { int n; };
struct Num { int n; std::vector<double> LU; std::vector<int> piv; };
}

extern "C" {

int umfpack_di_symbolic(int n_row, int n_col, const int[], const int[], const double[], void** Symbolic, const double[], double[])
{
    if (n_row != n_col) return 1;
    *Symbolic = new Sym{n_row};
    return UMFPACK_OK;
}

int umfpack_di_numeric(const int Ap[], const int Ai[], const double Ax[], void* Symbolic, void** Numeric, const double[], double[])
{
    const int n = static_cast<Sym*>(Symbolic)->n;
    Num* N = new Num{n, std::vector<double>((size_t) n * n, 0.0), std::vector<int>(n)};
    for (int c = 0; c < n; ++c)
        for (int q = Ap[c]; q < Ap[c + 1]; ++q) N->LU[(size_t) Ai[q] * n + c] += Ax[q];
    for (int k = 0; k < n; ++k) {
        int pr = k;
        double best = std::fabs(N->LU[(size_t) k * n + k]);
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(N->LU[(size_t) i * n + k]) > best) { best = std::fabs(N->LU[(size_t) i * n + k]); pr = i; }
        N->piv[k] = pr;
        if (best == 0.0) { *Numeric = N; return 1; }
        if (pr != k)
            for (int c = 0; c < n; ++c) std::swap(N->LU[(size_t) k * n + c], N->LU[(size_t) pr * n + c]);
        for (int i = k + 1; i < n; ++i) {
            const double l = N->LU[(size_t) i * n + k] / N->LU[(size_t) k * n + k];
            N->LU[(size_t) i * n + k] = l;
            if (l != 0.0)
                for (int c = k + 1; c < n; ++c) N->LU[(size_t) i * n + c] -= l * N->LU[(size_t) k * n + c];
        }
    }
    *Numeric = N;
    return UMFPACK_OK;
}

int umfpack_di_solve(int, const int[], const int[], const double[], double X[], const double B[], void* Numeric, const double[], double[])
{
    const Num* N = static_cast<Num*>(Numeric);
    const int n = N->n;
    for (int i = 0; i < n; ++i) X[i] = B[i];
    for (int k = 0; k < n; ++k) if (N->piv[k] != k) std::swap(X[k], X[N->piv[k]]);
    for (int i = 1; i < n; ++i) { double s = X[i]; for (int c = 0; c < i; ++c) s -= N->LU[(size_t) i * n + c] * X[c]; X[i] = s; }
    for (int i = n - 1; i >= 0; --i) {
        double s = X[i];
        for (int c = i + 1; c < n; ++c) s -= N->LU[(size_t) i * n + c] * X[c];
        X[i] = s / N->LU[(size_t) i * n + i];
    }
    return UMFPACK_OK;
}

void umfpack_di_free_symbolic(void** Symbolic) { delete static_cast<Sym*>(*Symbolic); *Symbolic = nullptr; }
void umfpack_di_free_numeric(void** Numeric) { delete static_cast<Num*>(*Numeric); *Numeric = nullptr; }

// Constructs Opm::MultisegmentWellContribution exactly as WellContributions::addMultisegmentWellContribution does
// (WellContributions.cpp:261-271) and calls its apply(h_x, h_y) once (MultisegmentWellContribution.cpp:70-110): y is
// updated in place.
void ref_mswell_apply(unsigned dim, unsigned dim_wells, unsigned Mb, const double* Bvalues, const unsigned* BcolIndices,
                      const unsigned* BrowPointers, unsigned DnumBlocks, double* Dvalues, int* DcolPointers, int* DrowIndices,
                      const double* Cvalues, double* x, double* y)
{
    const unsigned nB = BrowPointers[Mb];
    std::vector<double> Bv(Bvalues, Bvalues + (size_t) nB * dim * dim_wells), Cv(Cvalues, Cvalues + (size_t) nB * dim * dim_wells);
    std::vector<unsigned int> Bc(BcolIndices, BcolIndices + nB), Br(BrowPointers, BrowPointers + Mb + 1);
    Opm::MultisegmentWellContribution well(dim, dim_wells, Mb, Bv, Bc, Br, DnumBlocks, Dvalues, DcolPointers, DrowIndices, Cv);
    well.apply(x, y);
}

}
