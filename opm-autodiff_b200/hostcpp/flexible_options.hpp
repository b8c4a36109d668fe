// flexible_options.hpp -- the part of a FlexibleSolver property tree the B200 backend honours, read from the JSON files
// Flow takes with --linear-solver-configuration-json-file (keys and defaults: FlexibleSolver_impl.hpp:147-150,
// setupPropertyTree.cpp:175-188; example tests/options_flexiblesolver.json).  A deliberately small reader: objects,
// strings and numbers only, values may be quoted numbers as in the reference's files.  No Boost.
#pragma once
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>

namespace b200opt {

struct Options {
    double tol = 1e-2;
    int maxiter = 200;
    int verbosity = 0;
    double relaxation = 1.0;
    std::string solver = "bicgstab", preconditioner = "ParOverILU0";
    int ilulevel = 0;
};

// flat map "a.b.c" -> value text
inline void parse_object(const std::string& s, size_t& i, const std::string& prefix, std::map<std::string, std::string>& out)
{
    auto ws = [&]() { while (i < s.size() && std::isspace((unsigned char) s[i])) ++i; };
    auto str = [&]() {
        if (s[i] != '"') throw std::runtime_error("JSON: string expected");
        size_t j = s.find('"', i + 1);
        if (j == std::string::npos) throw std::runtime_error("JSON: unterminated string");
        std::string r = s.substr(i + 1, j - i - 1);
        i = j + 1;
        return r;
    };
    ws();
    if (i >= s.size() || s[i] != '{') throw std::runtime_error("JSON: object expected");
    ++i;
    while (true) {
        ws();
        if (i < s.size() && s[i] == '}') { ++i; return; }
        const std::string key = str();
        ws();
        if (i >= s.size() || s[i] != ':') throw std::runtime_error("JSON: ':' expected");
        ++i; ws();
        const std::string name = prefix.empty() ? key : prefix + "." + key;
        if (s[i] == '{') parse_object(s, i, name, out);
        else if (s[i] == '"') out[name] = str();
        else { size_t j = i; while (j < s.size() && s[j] != ',' && s[j] != '}' && !std::isspace((unsigned char) s[j])) ++j; out[name] = s.substr(i, j - i); i = j; }
        ws();
        if (i < s.size() && s[i] == ',') ++i;
    }
}

inline Options from_json_text(const std::string& text, bool strict = true)
{
    std::map<std::string, std::string> m;
    size_t i = 0;
    parse_object(text, i, "", m);
    auto get = [&](const char* k, const std::string& d) { auto it = m.find(k); return it == m.end() ? d : it->second; };
    Options o;
    o.tol = std::atof(get("tol", "1e-2").c_str());
    o.maxiter = std::atoi(get("maxiter", "200").c_str());
    o.verbosity = std::atoi(get("verbosity", "0").c_str());
    o.solver = get("solver", "bicgstab");
    o.preconditioner = get("preconditioner.type", "ParOverILU0");
    o.ilulevel = std::atoi(get("preconditioner.ilulevel", "0").c_str());
    std::string pl = o.preconditioner;
    for (auto& c : pl) c = (char) std::tolower((unsigned char) c);
    const bool ilu0 = pl == "ilu0" || pl == "paroverilu0";
    if (ilu0) o.relaxation = std::atof(get("preconditioner.relaxation", "1.0").c_str());
    if (strict) {
        if (o.solver != "bicgstab") throw std::invalid_argument("the b200 backend implements solver 'bicgstab' only, got '" + o.solver + "'");
        if (!ilu0) throw std::invalid_argument("the b200 backend implements preconditioner ILU0 / ParOverILU0 only, got '" + o.preconditioner + "'");
        if (o.ilulevel != 0) throw std::invalid_argument("the b200 backend implements fill level 0 only");
    }
    return o;
}

inline Options from_json_file(const std::string& path, bool strict = true)
{
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot read " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return from_json_text(ss.str(), strict);
}

}  // namespace b200opt
