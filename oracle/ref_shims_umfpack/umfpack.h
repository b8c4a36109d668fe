/* Shim: the five UMFPACK (SuiteSparse) entry points MultisegmentWellContribution.cpp calls, declared as the real
 * umfpack.h declares them (int indices: UMFPackIndex is int for Dune < 2.7, MultisegmentWellContribution.hpp:80-84).
 * SuiteSparse is not in this image; oracle/ref_mswell_glue.cpp implements them with a dense LU with partial pivoting, so
 * that the reference's own constructor and apply() run UNMODIFIED in a TEST-ONLY checker (oracle/_ref/libref_mswell.so). */
#ifndef B200_REF_SHIM_UMFPACK_H
#define B200_REF_SHIM_UMFPACK_H
#ifdef __cplusplus
extern "C" {
#endif
#define UMFPACK_A 0
#define UMFPACK_OK 0
int umfpack_di_symbolic(int n_row, int n_col, const int Ap[], const int Ai[], const double Ax[], void** Symbolic,
                        const double Control[], double Info[]);
int umfpack_di_numeric(const int Ap[], const int Ai[], const double Ax[], void* Symbolic, void** Numeric,
                       const double Control[], double Info[]);
int umfpack_di_solve(int sys, const int Ap[], const int Ai[], const double Ax[], double X[], const double B[],
                     void* Numeric, const double Control[], double Info[]);
void umfpack_di_free_symbolic(void** Symbolic);
void umfpack_di_free_numeric(void** Numeric);
#ifdef __cplusplus
}
#endif
#endif
