// bda_compat.hpp -- stand-ins for the three reference headers a BdaSolver backend is written against,
// so that b200SolverBackend.hpp compiles OUTSIDE the OPM tree (the originals pull in Dune, UMFPACK and
// OpenCL headers that are not in this image).  Inside opm-simulators define B200_IN_OPM_TREE and the
// real headers are used instead; names, signatures and member meaning are identical:
//   bda::BdaResult, bda::SolverStatus, bda::BdaSolver<block_size>   bda/BdaResult.hpp:28-40, bda/BdaSolver.hpp:32-90
//   Opm::WellContributions (standard wells and addMultisegmentWellContribution)  bda/WellContributions.hpp:60-214
#pragma once

#ifdef B200_IN_OPM_TREE
#include <opm/simulators/linalg/bda/BdaResult.hpp>
#include <opm/simulators/linalg/bda/BdaSolver.hpp>
#include <opm/simulators/linalg/bda/WellContributions.hpp>
#else

#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200bda.h"

namespace bda {

struct BdaResult {
    int iterations = 0;
    double reduction = 0.0;
    bool converged = false;
    double conv_rate = 0.0;
    double elapsed = 0.0;
};

enum class SolverStatus { BDA_SOLVER_SUCCESS, BDA_SOLVER_ANALYSIS_FAILED, BDA_SOLVER_CREATE_PRECONDITIONER_FAILED, BDA_SOLVER_UNKNOWN_ERROR };

}  // namespace bda

namespace Opm {

// Host container with the reference's three-phase protocol; the data goes straight into the C ABI's
// b200_wells object, which the backend uploads in one piece at solve time.
class WellContributions {
public:
    enum class MatrixType { C, D, B };

    WellContributions(std::string accelerator_mode, bool useWellConn)
        : h_(b200_wells_create(accelerator_mode.c_str(), useWellConn ? 1 : 0))
    {
        if (!h_) throw std::logic_error(b200_last_error());
    }
    ~WellContributions() { b200_wells_destroy(h_); }
    WellContributions(const WellContributions&) = delete;
    WellContributions& operator=(const WellContributions&) = delete;

    unsigned int getNumWells() { return b200_wells_get_num_wells(h_); }
    void setBlockSize(unsigned int dim, unsigned int dim_wells) { check(b200_wells_set_block_size(h_, dim, dim_wells)); }
    void addNumBlocks(unsigned int numBlocks) { check(b200_wells_add_num_blocks(h_, numBlocks)); }
    void alloc() { check(b200_wells_alloc(h_)); }
    void addMatrix(MatrixType type, int* colIndices, double* values, unsigned int val_size)
    {
        check(b200_wells_add_matrix(h_, static_cast<b200_well_matrix>(static_cast<int>(type)), colIndices, values, val_size));
    }
    // WellContributions.hpp:195-213 (UMFPackIndex is int for Dune >= 2.7)
    void addMultisegmentWellContribution(unsigned int dim, unsigned int dim_wells, unsigned int Mb, std::vector<double>& Bvalues,
                                         std::vector<unsigned int>& BcolIndices, std::vector<unsigned int>& BrowPointers,
                                         unsigned int DnumBlocks, double* Dvalues, int* DcolPointers, int* DrowIndices,
                                         std::vector<double>& Cvalues)
    {
        check(b200_wells_add_multisegment(h_, dim, dim_wells, Mb, Bvalues.data(), BcolIndices.data(), BrowPointers.data(),
                                          DnumBlocks, Dvalues, DcolPointers, DrowIndices, Cvalues.data()));
    }
    b200_wells* handle() { return h_; }

private:
    static void check(b200_status st)
    {
        if (st != B200_SUCCESS) throw std::logic_error(b200_last_error());
    }
    b200_wells* h_;
};

}  // namespace Opm

namespace bda {

using Opm::WellContributions;

template <unsigned int block_size>
class BdaSolver {
protected:
    int verbosity = 0;
    int maxit = 200;
    double tolerance = 1e-2;
    std::string bitstream = "";
    int N = 0, Nb = 0, nnz = 0, nnzb = 0;
    unsigned int platformID = 0, deviceID = 0;
    bool initialized = false;

public:
    BdaSolver(int linear_solver_verbosity, int max_it, double tolerance_, unsigned int deviceID_)
        : verbosity(linear_solver_verbosity), maxit(max_it), tolerance(tolerance_), deviceID(deviceID_) {}
    virtual ~BdaSolver() {}
    virtual SolverStatus solve_system(int N, int nnz, int dim, double* vals, int* rows, int* cols, double* b,
                                      WellContributions& wellContribs, BdaResult& res) = 0;
    virtual void get_result(double* x) = 0;
};

}  // namespace bda
#endif
