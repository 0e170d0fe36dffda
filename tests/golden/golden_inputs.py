"""Deterministic inputs for the golden fixtures (numpy RandomState: stable across numpy versions).

Shared by the fixture generators (which need /root/reference or a GPU) and by the tests (which do not).
"""
import numpy as np


def make_input(seed, shape, dist="normal", relu=False, chan_scale=True):
    rs = np.random.RandomState(seed)
    if dist == "normal":
        x = rs.standard_normal(shape)
    elif dist == "laplace":
        x = rs.laplace(size=shape)
    else:
        raise ValueError(dist)
    if chan_scale and len(shape) >= 2:
        cdim = 1 if len(shape) == 4 else 0
        sc = rs.uniform(0.1, 3.0, size=shape[cdim])
        shp = [1] * len(shape)
        shp[cdim] = shape[cdim]
        x = x * sc.reshape(shp) + 0.25 * sc.reshape(shp)
    if relu:
        x = np.maximum(x, 0)
    return x.astype(np.float32)


def regen(spec):
    """Rebuild a fixture input from the ``input`` entry of ref_cpu_meta.json."""
    x = make_input(spec["seed"], tuple(spec["shape"]), spec.get("dist", "normal"), spec.get("relu", False),
                   spec.get("chan_scale", True))
    if "mult" in spec:
        x = (x * spec["mult"]).astype(np.float32)
    if "const_row" in spec:
        r, v = spec["const_row"]
        x[r, :] = v
    return x
