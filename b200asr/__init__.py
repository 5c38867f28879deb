"""Importable alias for the product package.

The product lives in ``automatic-speech-recognition-asr-onnx_b200/`` (a directory
name Python cannot import directly because of the hyphens).  ``import b200asr``
executes that package's ``__init__`` with ``__path__`` pointing at it, so
``b200asr.engine`` etc. resolve to the files in the hyphenated directory.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "automatic-speech-recognition-asr-onnx_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
