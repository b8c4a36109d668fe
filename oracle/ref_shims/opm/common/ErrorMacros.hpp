/* Shim: OPM_THROW from opm-common (not in this image). */
#ifndef B200_REF_SHIM_ERRORMACROS_HPP
#define B200_REF_SHIM_ERRORMACROS_HPP
#include <sstream>
#include <stdexcept>
#include <cassert>
#include <cstring>
#define OPM_THROW(Exception, message)                         \
    do {                                                      \
        std::ostringstream oss__;                             \
        oss__ << "[" << __FILE__ << ":" << __LINE__ << "] " << message; \
        throw Exception(oss__.str());                         \
    } while (false)
#endif
