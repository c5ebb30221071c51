"""Import shim: exposes the package directory `diffusion-tts_b200/` under the importable name
`diffusion_tts_b200` (a hyphen is not legal in a Python module name)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), 'diffusion-tts_b200')]
with open(_os.path.join(__path__[0], '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], '__init__.py'), 'exec'))
del _os, _f
