"""ctypes access to oracle/_build/liboracle_leaf.so (plain-C leaf restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os

import numpy as np

from . import build_oracle

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(build_oracle.BUILD, "liboracle_leaf.so")
        if not os.path.exists(path):
            build_oracle.build_leaf()
        l = ctypes.CDLL(path)
        fp, i64, i32, f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float
        l.oracle_float2gemmlowp.restype = i32
        l.oracle_float2gemmlowp.argtypes = [fp, fp, i64, f32, f32, i32, i32, i32, fp]
        l.oracle_quantize1_rows.restype = None
        l.oracle_quantize1_rows.argtypes = [fp, fp, fp, i64, i64, fp, fp, fp, i32, i32]
        _lib = l
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def float2gemmlowp(x, range_, offset, num_bits, int_exp=False, enforce_true_zero=True, noise=None):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    nz = None if noise is None else np.ascontiguousarray(noise, dtype=np.float32)
    same = lib().oracle_float2gemmlowp(_ptr(x), _ptr(out), x.size, float(range_), float(offset), int(num_bits),
                                       int(int_exp), int(enforce_true_zero), _ptr(nz))
    return x if same else out


def quantize1_rows(x, delta, offset, num_bits, bits=None, want_grid=False):
    x = np.ascontiguousarray(x, dtype=np.float32)
    rows, cols = x.shape
    delta = np.ascontiguousarray(np.atleast_1d(delta), dtype=np.float32)
    offset = np.ascontiguousarray(np.atleast_1d(offset), dtype=np.float32)
    per_row = int(delta.size > 1)
    b = None if bits is None else np.ascontiguousarray(bits, dtype=np.float32)
    y = np.empty_like(x)
    grid = np.empty_like(x) if want_grid else None
    lib().oracle_quantize1_rows(_ptr(x), _ptr(y), _ptr(grid), rows, cols, _ptr(delta), _ptr(offset), _ptr(b), per_row,
                                int(num_bits))
    return (y, grid) if want_grid else y
