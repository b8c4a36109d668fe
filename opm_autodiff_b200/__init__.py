"""Import shim: the package directory is named ``opm-autodiff_b200`` (not a Python identifier);
this module makes it importable as ``opm_autodiff_b200``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "opm-autodiff_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
