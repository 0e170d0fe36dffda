"""Drop-in for the reference's compiled extension module ``int_quantization``
(kernels/int_quantization.cpp:6-12, kernels/gemmlowp.cu:30-45): same function name, same seven positional
arguments, same semantics, backed by ``fqb200_float2gemmlowp`` in libfqb200.so.

Differences from the reference kernel, none of them numerical: launches on PyTorch's *current* stream under
the tensor's device (the reference uses the legacy default stream without a device guard), validates dtype /
device, covers the whole GPU instead of 32 CTAs, and accepts ``noise=None`` for the all-zero noise tensor the
reference always passes.  ``range <= 0`` returns the input tensor itself, like the reference.
"""
import torch

from . import ops

__all__ = ["float2gemmlowp"]


def float2gemmlowp(input, range, offset, num_bits, int_exp, enforce_true_zero, noise=None):
    """``float2gemmlowp(in, range, offset, num_bits, int_exp, enforce_true_zero, noise) -> Tensor``.

    ``range`` / ``offset`` may be python floats or 0-d tensors; tensors are converted with ``float()`` exactly
    as pybind does for the reference (that conversion synchronises with the device)."""
    range_f = float(range)
    offset_f = float(offset)
    if range_f <= 0:
        return input
    if noise is not None and not isinstance(noise, torch.Tensor):
        raise TypeError("noise must be a tensor or None")
    return ops.float2gemmlowp(input, range_f, offset_f, int(num_bits), bool(int_exp), bool(enforce_true_zero), noise)
