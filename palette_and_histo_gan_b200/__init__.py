"""Importable alias of the package directory `palette-and-histo-gan_b200/` (a hyphen is not a valid
Python identifier).  `import palette_and_histo_gan_b200` executes that directory's `__init__.py` with
this module's `__path__` pointing at it, so submodules resolve there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "palette-and-histo-gan_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
