"""Import alias for the ``hyena-dna_b200`` package directory.

The product package lives in ``hyena-dna_b200/`` (a hyphen cannot appear in a Python
identifier), so ``import hyena_dna_b200`` resolves here and re-exports it: this module's
``__path__`` points at the hyphenated directory and its ``__init__`` is executed in place.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "hyena-dna_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
