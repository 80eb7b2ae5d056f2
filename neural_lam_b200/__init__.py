"""Importable alias for the package directory ``neural-lam_b200/`` (a hyphen is not a valid
Python identifier, so ``import neural_lam_b200`` resolves here and executes the real package
``__init__`` with ``__path__`` pointing at ``neural-lam_b200/``)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "neural-lam_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__, encoding="utf-8") as _f:
    exec(compile(_f.read(), __file__, "exec"), globals())
del _f, _os
