"""opm-simulators.patch (the caller-side glue of INTEGRATION.md) must apply to the reference tree it was written for."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
FILES = ["opm/simulators/linalg/ISTLSolverEbos.hpp", "opm/simulators/linalg/bda/BdaBridge.cpp",
         "opm/simulators/linalg/bda/WellContributions.cpp", "opm/simulators/linalg/bda/WellContributions.hpp"]


@pytest.mark.skipif(not os.path.isdir(REF) or shutil.which("patch") is None, reason="needs the reference tree and patch(1)")
def test_patch_applies_to_the_reference_tree(tmp_path):
    for f in FILES:
        os.makedirs(os.path.dirname(tmp_path / f), exist_ok=True)
        shutil.copy(os.path.join(REF, f), tmp_path / f)
    with open(os.path.join(ROOT, "opm-simulators.patch")) as fh:
        r = subprocess.run(["patch", "-p1", "--batch"], cwd=tmp_path, stdin=fh, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    txt = (tmp_path / FILES[1]).read_text()
    assert 'accelerator_mode.compare("b200") == 0' in txt and "b200SolverBackend<block_size>" in txt
    ebos = (tmp_path / FILES[0]).read_text()
    assert ebos.count("cpuSolverStale_") >= 5 and '(accelerator_mode != "b200")' in ebos


def test_patch_touches_only_the_caller_side():
    with open(os.path.join(ROOT, "opm-simulators.patch")) as fh:
        names = [l.split()[1][2:] for l in fh if l.startswith("+++ ")]
    assert sorted(names) == sorted(FILES)
