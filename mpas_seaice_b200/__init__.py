"""Import shim: the product package lives in the directory ``mpas-seaice_b200/`` (the name the
build contract fixes); a hyphen is not importable, so ``import mpas_seaice_b200`` resolves here and
this module re-points its search path at the real directory."""
import os as _os

_real = _os.path.normpath(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "mpas-seaice_b200"))
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
