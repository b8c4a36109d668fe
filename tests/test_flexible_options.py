"""FlexibleSolver property-tree options as the backend reads them (Python mirror and C++ header)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# tests/options_flexiblesolver.json of the reference, verbatim structure: CPR with an ILU0 fine smoother
REFERENCE_OPTIONS = {
    "tol": "0.5", "maxiter": "20",
    "preconditioner": {"type": "cpr", "finesmoother": {"type": "ILU0", "relaxation": "1.0"},
                       "coarsesolver": {"tol": "0.5", "maxiter": "20", "preconditioner": {"type": "amg", "maxlevel": "5"},
                                        "verbosity": "0", "solver": "bicgstab"},
                       "verbosity": "11", "weights_filename": "weight_cpr.txt"},
    "verbosity": "10", "solver": "bicgstab"}
ILU_OPTIONS = {"tol": 1e-2, "maxiter": 200, "verbosity": 0, "solver": "bicgstab",
               "preconditioner": {"type": "ParOverILU0", "relaxation": 0.9, "ilulevel": 0}}       # setupPropertyTree.cpp:175-188


def test_python_mirror(built):
    from opm_autodiff_b200 import bridge
    o = bridge.FlexibleSolverOptions.from_tree(ILU_OPTIONS)
    assert (o.tol, o.maxiter, o.verbosity, o.relaxation) == (1e-2, 200, 0, 0.9)
    # the reference's own GPU tests read only tol / maxiter / verbosity from the CPR file (test_cusparseSolver.cpp:61-66)
    o = bridge.FlexibleSolverOptions.from_tree(REFERENCE_OPTIONS, strict=False)
    assert (o.tol, o.maxiter, o.verbosity, o.relaxation) == (0.5, 20, 10, 1.0)
    with pytest.raises(ValueError):
        bridge.FlexibleSolverOptions.from_tree(REFERENCE_OPTIONS)            # CPR is not this backend's preconditioner
    with pytest.raises(ValueError):
        bridge.FlexibleSolverOptions.from_tree(dict(ILU_OPTIONS, solver="gmres"))
    with pytest.raises(ValueError):
        bridge.FlexibleSolverOptions.from_tree(dict(ILU_OPTIONS, preconditioner={"type": "ILU0", "ilulevel": 1}))


def test_cpp_header(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text('#include "flexible_options.hpp"\n#include <cstdio>\nint main(int c, char** v) { try { auto o = b200opt::from_json_file(v[1], v[2][0] == \'1\');'
                   ' std::printf("%g %d %d %g\\n", o.tol, o.maxiter, o.verbosity, o.relaxation); return 0; } catch (const std::exception& e) { std::printf("ERR %s\\n", e.what()); return 3; } }\n')
    exe = str(tmp_path / "t")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "opm-autodiff_b200", "hostcpp"), "-o", exe, str(src)])
    a, b = tmp_path / "ilu.json", tmp_path / "cpr.json"
    a.write_text(json.dumps(ILU_OPTIONS)); b.write_text(json.dumps(REFERENCE_OPTIONS, indent=4))
    assert subprocess.run([exe, str(a), "1"], capture_output=True, text=True).stdout.split() == ["0.01", "200", "0", "0.9"]
    assert subprocess.run([exe, str(b), "0"], capture_output=True, text=True).stdout.split() == ["0.5", "20", "10", "1"]
    r = subprocess.run([exe, str(b), "1"], capture_output=True, text=True)
    assert r.returncode == 3 and "ILU0" in r.stdout
