"""Import shim: the package sources live in ``clip-mixer_b200/`` (the layout the build contract
names; a hyphen is not importable), this module makes them importable as ``clip_mixer_b200``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "clip-mixer_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
