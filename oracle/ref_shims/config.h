/* Shim for building pieces of the reference as a TEST-ONLY checker (oracle/_ref). */
#ifndef B200_REF_SHIM_CONFIG_H
#define B200_REF_SHIM_CONFIG_H
#define HAVE_FPGA 0
#define HAVE_OPENCL 0
#ifndef B200_REF_UMFPACK_SHIM            /* libref_mswell.so passes -DHAVE_SUITESPARSE_UMFPACK=1 and the umfpack.h shim */
#define HAVE_SUITESPARSE_UMFPACK 0
#endif
#include <cassert>
#include <cstring>
#endif
