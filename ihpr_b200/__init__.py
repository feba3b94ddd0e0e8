"""Importable alias of the product package.

The package directory is named after the reference repository
(``integral-human-pose-regression-for-3d-human-pose-estimation_b200/``), which is not a valid
Python identifier; ``import ihpr_b200`` resolves to that directory.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "integral-human-pose-regression-for-3d-human-pose-estimation_b200")
__path__ = [_REAL]
__file__ = _os.path.join(_REAL, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f
