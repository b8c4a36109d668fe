/* Shim: version macros (dune-common is not in this image). */
#ifndef B200_REF_SHIM_DUNE_VERSION_HH
#define B200_REF_SHIM_DUNE_VERSION_HH
#define DUNE_VERSION_NEWER(module, major, minor) 0
#endif
