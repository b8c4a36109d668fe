// ref_cusparse_glue.cu -- TEST/BENCH-ONLY extern "C" entry point into the UNMODIFIED reference GPU backend
// (opm/simulators/linalg/bda/cusparseSolverBackend.cu + WellContributions.cu/.cpp), compiled where the sources lie under
// /root/reference against the shim headers of oracle/ref_shims (see oracle/Makefile, target _ref/libref_cusparse.so).
// It exists to time the incumbent GPU backend on the same B200 and the same system as the product and as a
// weak cross-check of the solution.  Never linked, imported or shipped by the product.
#include <config.h>
#include <chrono>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <opm/simulators/linalg/bda/cusparseSolverBackend.hpp>
#include <opm/simulators/linalg/bda/WellContributions.hpp>
#include <opm/simulators/linalg/bda/MultisegmentWellContribution.hpp>

// MultisegmentWellContribution.cpp needs UMFPACK (absent); the class is never instantiated here (standard wells only), the
// linker only needs the symbols WellContributions.cpp refers to.
namespace Opm {
MultisegmentWellContribution::MultisegmentWellContribution(unsigned int, unsigned int, unsigned int, std::vector<double>&,
                                                           std::vector<unsigned int>&, std::vector<unsigned int>&, unsigned int,
                                                           double*, UMFPackIndex*, UMFPackIndex*, std::vector<double>&)
{
    throw std::logic_error("MultisegmentWellContribution is not available in the reference shim build (no UMFPACK)");
}
MultisegmentWellContribution::~MultisegmentWellContribution() {}
void MultisegmentWellContribution::apply(double*, double*) {}
void MultisegmentWellContribution::setCudaStream(cudaStream_t) {}
void MultisegmentWellContribution::setReordering(int*, bool) {}
}

static std::string g_err;

extern "C" {

const char* ref_cusparse_last_error() { return g_err.c_str(); }

// Runs `nsolves` calls of cusparseSolverBackend<3>::solve_system + get_result on ONE backend object, each with a freshly
// built WellContributions (as ISTLSolverEbos.hpp:265-272 does per solve).  The first call includes initialize / analyse
// (cusparseSolverBackend.cu:480-499).  wall_s[i]: wall-clock seconds of call i (solve_system + get_result);
// iters[i], reduction[i], converged[i] from BdaResult.  Wells in the CSR-over-wells layout (nwells = 0: none).
// Returns 0, or 1 with ref_cusparse_last_error() set.
int ref_cusparse_solve(int N, int nnz, double* vals, int* rows, int* cols, double* b, int nwells, const unsigned* wptr,
                       int* Bcols, int* Ccols, double* B, double* C, double* Dinv, double tol, int maxit, int nsolves, double* x,
                       double* wall_s, int* iters, double* reduction, int* converged)
{
    try {
        bda::cusparseSolverBackend<3> backend(0, maxit, tol, 0);
        for (int s = 0; s < nsolves; ++s) {
            auto t0 = std::chrono::steady_clock::now();
            Opm::WellContributions wc("cusparse", false);
            if (nwells > 0) {
                wc.setBlockSize(3, 4);
                for (int w = 0; w < nwells; ++w) wc.addNumBlocks(wptr[w + 1] - wptr[w]);
                wc.alloc();
                for (int w = 0; w < nwells; ++w) {
                    const unsigned s0 = wptr[w], n = wptr[w + 1] - wptr[w];
                    wc.addMatrix(Opm::WellContributions::MatrixType::C, Ccols + s0, C + (size_t) s0 * 12, n);
                    wc.addMatrix(Opm::WellContributions::MatrixType::D, Bcols + s0, Dinv + (size_t) w * 16, 1);
                    wc.addMatrix(Opm::WellContributions::MatrixType::B, Bcols + s0, B + (size_t) s0 * 12, n);
                }
            }
            bda::BdaResult res;
            bda::SolverStatus st = backend.solve_system(N, nnz, 3, vals, rows, cols, b, wc, res);
            if (st != bda::SolverStatus::BDA_SOLVER_SUCCESS) throw std::runtime_error("solve_system status " + std::to_string((int) st));
            backend.get_result(x);
            wall_s[s] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            iters[s] = res.iterations; reduction[s] = res.reduction; converged[s] = res.converged ? 1 : 0;
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

}
