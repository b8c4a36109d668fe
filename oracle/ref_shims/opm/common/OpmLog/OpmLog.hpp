/* Shim: the reference logs through opm-common's OpmLog (not in this image). */
#ifndef B200_REF_SHIM_OPMLOG_HPP
#define B200_REF_SHIM_OPMLOG_HPP
#include <iostream>
#include <string>
namespace Opm {
struct OpmLog {
    static void info(const std::string& m)    { if (verbose()) std::cerr << "[ref info] " << m << "\n"; }
    static void warning(const std::string& m) { std::cerr << "[ref warning] " << m << "\n"; }
    static void error(const std::string& m)   { std::cerr << "[ref error] " << m << "\n"; }
    static void debug(const std::string&)     {}
    static bool& verbose() { static bool v = false; return v; }
};
}
#endif
