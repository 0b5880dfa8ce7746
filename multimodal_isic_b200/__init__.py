"""Importable alias of the ``multimodal-isic_b200/`` package directory (a hyphen cannot appear
in a Python identifier).  All code lives in ``multimodal-isic_b200/``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "multimodal-isic_b200")
__path__[:] = [_real]
with open(_os.path.join(_real, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _fh
