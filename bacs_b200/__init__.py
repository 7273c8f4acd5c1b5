"""Importable alias for the package directory ``bacs-continual-semantic-segmentation_b200/``.

The contract names the package directory with hyphens, which Python cannot import by
name; this shim makes ``import bacs_b200`` (and ``bacs_b200.loss`` etc.) resolve to it."""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "bacs-continual-semantic-segmentation_b200")
__path__ = [_REAL]
__file__ = _os.path.join(_REAL, "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"), globals())
