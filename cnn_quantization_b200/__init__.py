"""Import shim: the product package lives in the directory ``cnn-quantization_b200/`` (the name the project
layout prescribes); a hyphen is not importable, so ``import cnn_quantization_b200`` resolves here and this
module re-points its package path at the real directory and executes its ``__init__``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "cnn-quantization_b200")
__path__[:] = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f
