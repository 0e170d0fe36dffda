"""The plain-C leaf restatement (oracle/fq_leaf.c) against the numpy/torch oracle and the reference fixtures."""
import numpy as np
import pytest
import torch

from golden_inputs import regen
from oracle import c_leaf, fq_oracle as O


@pytest.mark.parametrize("tz,rng,off", [(True, 7.3, -3.1), (False, 5.0, 0.0), (False, 6.0, 0.5), (False, 4.0, -1.0),
                                        (True, 1e-3, -4e-4)])
@pytest.mark.parametrize("bits", [2, 4, 8])
def test_compiled_leaf_c_equals_numpy(tz, rng, off, bits):
    rs = np.random.RandomState(bits)
    x = (rs.standard_normal(20011) * 2 + 0.3).astype(np.float32)
    x[:6] = [0.0, np.inf, -np.inf, np.nan, 1e38, -1e38]
    noise = rs.uniform(-0.5, 0.5, x.size).astype(np.float32)
    for nz in (None, noise):
        a = c_leaf.float2gemmlowp(x, rng, off, bits, False, tz, nz)
        b = O.float2gemmlowp(x, rng, off, bits, False, tz, nz)
        assert np.array_equal(a, b, equal_nan=True)
    assert c_leaf.float2gemmlowp(x, 0.0, off, bits) is not None
    assert O.float2gemmlowp(x, -1.0, off, bits) is x


def test_torch_leaf_c_equals_reference_fixture(golden):
    arrays, meta = golden
    for name in ("leaf_rows_bits", "leaf_rows_nobits"):
        x = regen(meta[name]["input"])
        bits = arrays[name + ".bits"] if name + ".bits" in arrays else None
        y = c_leaf.quantize1_rows(x, arrays[name + ".delta"], arrays[name + ".offset"], 4, bits)
        assert np.array_equal(y, arrays[name + ".y"])
    name = "leaf_cfg1_int4"
    x = regen(meta[name]["input"]).reshape(1, -1)
    y = c_leaf.quantize1_rows(x, arrays[name + ".delta"], arrays[name + ".offset"], 4)
    assert np.array_equal(y.reshape(-1), arrays[name + ".y"].reshape(-1))
