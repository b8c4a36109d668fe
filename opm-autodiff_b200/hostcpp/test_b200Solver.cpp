// test_b200Solver.cpp -- the reference's boundary test (tests/test_cusparseSolver.cpp:49-112) restated for
// the B200 backend without Boost/Dune: read a blocked MatrixMarket system, construct the backend through
// the BdaSolver<3> interface with (verbosity, maxiter, tol), pass an EMPTY WellContributions, call
// solve_system + get_result unconditionally and print x (one value per line, 17 digits) for the pytest
// wrapper to compare against the golden vector.  Also applies BdaBridge's zero-diagonal fix-up
// (BdaBridge.cpp:125-161) because the bridge does so before every backend call.
//   usage: test_b200Solver matrix.mm rhs.mm [tol] [maxiter] [verbosity]
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "b200SolverBackend.hpp"
#include "istl_mm.hpp"

static int checkZeroDiagonal(b200mm::Bsr& A)
{
    int numZeros = 0;
    for (int r = 0; r < A.Nb; ++r)
        for (int k = A.rows[r]; k < A.rows[r + 1]; ++k)
            if (A.cols[k] == r)
                for (int rr = 0; rr < 3; ++rr) {
                    double& v = A.vals[(size_t) k * 9 + rr * 3 + rr];
                    if (v == 0.0) { v = 1e-15; ++numZeros; }
                }
    return numZeros;
}

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: %s matrix.mm rhs.mm [tol] [maxiter] [verbosity]\n", argv[0]); return 2; }
    const double tolerance = argc > 3 ? std::atof(argv[3]) : 0.5;      // tests/options_flexiblesolver.json:2
    const int maxit = argc > 4 ? std::atoi(argv[4]) : 20;              // :3
    const int verbosity = argc > 5 ? std::atoi(argv[5]) : 0;
    try {
        b200mm::Bsr A = b200mm::read_matrix(argv[1]);
        std::vector<double> rhs = b200mm::read_vector(argv[2]);
        if (A.bs != 3) { std::fprintf(stderr, "BdaSolver only accepts blocksize = 3\n"); return 3; }
        checkZeroDiagonal(A);
        std::unique_ptr<bda::BdaSolver<3>> backend;
        try {
            backend = std::make_unique<bda::b200SolverBackend<3>>(verbosity, maxit, tolerance, 0u);
        } catch (const std::logic_error& e) {
            std::fprintf(stderr, "Problem with initializing a device: %s\n", e.what());   // the reference SKIPS here
            return 77;
        }
        Opm::WellContributions wellContribs("b200", false);
        if (std::getenv("B200_TEST_MSWELL")) {
            // one multisegment well with a single segment perforating cell 0, B = 0 (so it changes nothing): exercises
            // WellContributions::addMultisegmentWellContribution (WellContributions.hpp:195-213) through the C++ surface
            std::vector<double> Bv(12, 0.0), Cv(12, 1.0), Dv = {2, 0, 0, 0, 0, 2, 0, 0, 0, 0, 2, 0, 0, 0, 0, 2};
            std::vector<unsigned int> Bc = {0u}, Br = {0u, 1u};
            std::vector<int> Dcol = {0, 4, 8, 12, 16}, Drow = {0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3};
            wellContribs.addMultisegmentWellContribution(3, 4, 1, Bv, Bc, Br, 1, Dv.data(), Dcol.data(), Drow.data(), Cv);
            if (wellContribs.getNumWells() != 1) { std::fprintf(stderr, "getNumWells() != 1\n"); return 4; }
        }
        bda::BdaResult result;
        const int N = A.Nb * 3, nnz = (int) A.cols.size() * 9;
        std::vector<double> x((size_t) N, 0.0);
        // the glue code page-locks the buffers that live as long as the simulator, once
        auto* b200 = static_cast<bda::b200SolverBackend<3>*>(backend.get());
        b200->registerHostBuffer(A.vals.data(), sizeof(double) * (size_t) nnz);
        b200->registerHostBuffer(rhs.data(), sizeof(double) * (size_t) N);
        b200->registerHostBuffer(x.data(), sizeof(double) * (size_t) N);
        bda::SolverStatus st = backend->solve_system(N, nnz, 3, A.vals.data(), A.rows.data(), A.cols.data(), rhs.data(), wellContribs, result);
        backend->get_result(x.data());
        b200->unregisterHostBuffer(x.data());
        b200->unregisterHostBuffer(rhs.data());
        b200->unregisterHostBuffer(A.vals.data());
        std::printf("status %d converged %d iterations %d reduction %.6e\n", (int) st, (int) result.converged, result.iterations, result.reduction);
        for (double v : x) std::printf("%.17g\n", v);
        return st == bda::SolverStatus::BDA_SOLVER_SUCCESS ? 0 : 1;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
