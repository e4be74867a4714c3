"""Importable alias of the `e2-tts-pytorch_b200/` package directory (a hyphen is not a valid module name).
`import e2_tts_pytorch_b200` executes e2-tts-pytorch_b200/__init__.py with this module as the package."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'e2-tts-pytorch_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
