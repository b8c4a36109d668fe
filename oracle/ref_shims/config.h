/* Shim for building pieces of the reference as a TEST-ONLY checker (oracle/_ref). */
#ifndef B200_REF_SHIM_CONFIG_H
#define B200_REF_SHIM_CONFIG_H
#define HAVE_FPGA 0
#define HAVE_OPENCL 0
#define HAVE_SUITESPARSE_UMFPACK 0
#include <cassert>
#include <cstring>
#endif
