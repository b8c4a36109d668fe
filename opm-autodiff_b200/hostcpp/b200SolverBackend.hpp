// b200SolverBackend.hpp -- the bda::BdaSolver<block_size> subclass that puts the B200 library behind
// the reference's accelerator plugin surface (cf. bda/cusparseSolverBackend.hpp:38-146).  Header-only;
// link with -lb200bda.  Flow selects it with --accelerator-mode=b200 once the two string-chain
// branches of INTEGRATION.md are in BdaBridge.cpp:65-120 and WellContributions.cpp:31-49.
#pragma once
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <string>

#include "bda_compat.hpp"

namespace bda {

template <unsigned int block_size>
class b200SolverBackend : public BdaSolver<block_size> {
    using Base = BdaSolver<block_size>;
    b200_solver* handle_ = nullptr;
    b200_result last_{};

public:
    /// Same arguments as cusparseSolverBackend (cusparseSolverBackend.hpp:121-126); device errors throw
    /// std::logic_error like cudaCheckLastError/OPM_THROW (cuda_header.hpp:34-44).
    b200SolverBackend(int linear_solver_verbosity, int maxit, double tolerance, unsigned int deviceID)
        : Base(linear_solver_verbosity, maxit, tolerance, deviceID)
    {
        handle_ = b200_create(linear_solver_verbosity, maxit, tolerance, deviceID);
        if (!handle_) throw std::logic_error(std::string("b200SolverBackend: ") + b200_last_error());
        this->initialized = true;
    }
    ~b200SolverBackend() override { b200_destroy(handle_); }
    b200SolverBackend(const b200SolverBackend&) = delete;
    b200SolverBackend& operator=(const b200SolverBackend&) = delete;

    /// ILU relaxation and the other knobs of include/b200bda.h (b200_set_option).
    void setOption(const std::string& key, double value)
    {
        if (b200_set_option(handle_, key.c_str(), value) != B200_SUCCESS) throw std::logic_error(b200_last_error());
    }

    SolverStatus solve_system(int N, int nnz, int dim, double* vals, int* rows, int* cols, double* b,
                              WellContributions& wellContribs, BdaResult& res) override
    {
        this->N = N; this->nnz = nnz; this->Nb = N / (int) block_size; this->nnzb = nnz / (int) (block_size * block_size);
        b200_wells* w = wells_of(wellContribs);
        const b200_status st = b200_solve_system(handle_, N, nnz, dim, vals, rows, cols, b, w, &last_);
        res.iterations = last_.iterations;
        res.reduction = last_.reduction;
        res.converged = last_.converged != 0;
        res.conv_rate = last_.conv_rate;
        res.elapsed = last_.elapsed;
        switch (st) {
            case B200_SUCCESS: return SolverStatus::BDA_SOLVER_SUCCESS;
            case B200_ANALYSIS_FAILED: return SolverStatus::BDA_SOLVER_ANALYSIS_FAILED;
            case B200_CREATE_PRECONDITIONER_FAILED: return SolverStatus::BDA_SOLVER_CREATE_PRECONDITIONER_FAILED;
            default:
                // runtime/device errors throw in the reference backends (OPM_THROW(std::logic_error, ...))
                throw std::logic_error(std::string("b200SolverBackend::solve_system: ") + b200_last_error());
        }
    }

    void get_result(double* x) override
    {
        if (b200_get_result(handle_, x) != B200_SUCCESS) throw std::logic_error(b200_last_error());
    }

    /// Page-lock a caller-owned buffer that outlives the solves -- Flow's matrix values, right-hand side and solution vector --
    /// so that the copies run at PCIe speed (b200_host_register).  The library never pins memory on its own:
    /// the glue code that knows the buffers' lifetime (BdaBridge) calls this once and unregisters before they are freed or resized.
    void registerHostBuffer(void* p, std::size_t bytes)
    {
        if (b200_host_register(handle_, p, bytes) != B200_SUCCESS) throw std::logic_error(b200_last_error());
    }
    void unregisterHostBuffer(void* p)
    {
        if (b200_host_unregister(handle_, p) != B200_SUCCESS) throw std::logic_error(b200_last_error());
    }

    const b200_result& lastResult() const { return last_; }
    b200_solver* handle() { return handle_; }

private:
#ifdef B200_IN_OPM_TREE
    static b200_wells* wells_of(WellContributions& wc) { return wc.getNumWells() > 0 ? wc.b200Handle() : nullptr; }
#else
    static b200_wells* wells_of(WellContributions& wc) { return wc.getNumWells() > 0 ? wc.handle() : nullptr; }
#endif
};

}  // namespace bda
