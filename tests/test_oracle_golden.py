"""The oracle (oracle/fq_oracle.py) against fixtures produced by the REAL reference on CPU.

The reference has no tests of its own; tests/golden/make_golden.py ran its int_quantizer.py hot path on
seeded inputs and stored the outputs plus the (delta, offset, bit_alloc) its leaf received.  Same torch
CPU build -> the restatement is expected to be bit-identical; the asserts allow 1e-6 on parameters and a
vanishing fraction of one-step flips so a different host CPU (other SIMD reduction order) cannot break
the suite.
"""
import numpy as np
import pytest
import torch

from golden_inputs import regen
from conftest import fq_mismatch
from oracle import fq_oracle as O

torch.set_num_threads(1)


def _params_close(a, b, name):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    assert a.shape == b.shape, name
    assert np.allclose(a, b, rtol=2e-6, atol=1e-9), (name, np.abs(a - b).max())


def _case_names(kind_prefixes):
    import json, os
    with open(os.path.join(os.path.dirname(__file__), "golden", "ref_cpu_meta.json")) as f:
        meta = json.load(f)
    return sorted(n for n in meta if n.startswith(kind_prefixes))


@pytest.mark.parametrize("name", _case_names(("leaf_",)))
def test_leaf_given_params_bit_exact(golden, name):
    arrays, meta = golden
    info = meta[name]
    x = torch.from_numpy(regen(info["input"]))
    delta = torch.from_numpy(arrays[name + ".delta"])
    offset = torch.from_numpy(arrays[name + ".offset"])
    bits = torch.from_numpy(arrays[name + ".bits"]) if name + ".bits" in arrays else None
    if info["kind"] == "leaf":
        delta, offset = delta.reshape(()), offset.reshape(())
    y = O.gemmlowp_quantize1(x, delta, offset, info["num_bits"], bit_alloc=bits)
    assert np.array_equal(y.numpy(), arrays[name + ".y"])  # integer grid and dequant bit-exact
    # unique-level invariant
    if info["kind"] == "leaf":
        assert np.unique(y.numpy()).size <= 2 ** info["num_bits"]


def _run_act(info, x):
    p = info["params"]
    positive = bool(info.get("half_range") or info.get("force_positive"))
    nb = info["num_bits"]
    if p.get("mtd_quant"):
        tgt = p.get("bit_alloc_target_act") or nb
        return O.mid_tread_activation(x, tgt, p.get("pcq_act", False), positive), None
    kw = dict(bit_alloc_act=p.get("bit_alloc_act", False), bit_alloc_prior=p.get("bit_alloc_prior", "gaus"),
              bit_alloc_target=p.get("bit_alloc_target_act"), bit_alloc_round=p.get("bit_alloc_rmode", "round") == "round")
    if p.get("clipping", "no") != "no":
        return O.clipping_quantize(x, nb, p["clipping"], p.get("pcq_act", False), positive, return_parts=True, **kw)
    return O.quantize_activation_per_channel(x, nb, positive, return_parts=True, **kw)


@pytest.mark.parametrize("name", _case_names(("act_", "mt_act")))
def test_activation_paths(golden, name):
    arrays, meta = golden
    info = meta[name]
    x = torch.from_numpy(regen(info["input"]))
    y, parts = _run_act(info, x)
    if parts is not None:
        _params_close(parts["delta"], arrays[name + ".delta"], name + ".delta")
        _params_close(parts["offset"], arrays[name + ".offset"], name + ".offset")
        if name + ".bits" in arrays:
            assert parts["bits"] is not None
            assert np.array_equal(parts["bits"].numpy(), arrays[name + ".bits"]), name
    frac, _ = fq_mismatch(y.numpy(), arrays[name + ".y"])
    assert frac <= 1e-4, (name, frac)


@pytest.mark.parametrize("name", _case_names(("w_", "mt_w")))
def test_weight_paths(golden, name):
    arrays, meta = golden
    info = meta[name]
    p = info["params"]
    w = torch.from_numpy(regen(info["input"]))
    nb = info.get("override_num_bits") or info["num_bits"]
    if p.get("mtd_quant"):
        y = O.mid_tread_weights_per_channel(w, p["bit_alloc_target_weight"])
    else:
        y, parts = O.quantize_weights_per_channel(w, nb, p.get("bit_alloc_weight", False),
                                                  p.get("bit_alloc_target_weight"), True, return_parts=True)
        _params_close(parts["delta"], arrays[name + ".delta"], name)
        if name + ".bits" in arrays:
            assert np.array_equal(parts["bits"].numpy(), arrays[name + ".bits"])
    frac, _ = fq_mismatch(y.numpy(), arrays[name + ".y"])
    assert frac <= 1e-4, (name, frac)


def test_bit_allocation_known_answers(golden):
    arrays, _ = golden
    sig = torch.from_numpy(arrays["bits.sigma"])
    for key, tgt, rnd in (("t4_round", 4, True), ("t4_ceil", 4, False), ("t5p3_round", 5.3, True),
                          ("t3_round", 3, True), ("t2_round", 2, True)):
        b = O.get_bits_alloc_fixed_target(sig, tgt, rnd)
        assert np.array_equal(b.numpy(), arrays["bits." + key]), key
        assert b.min() >= 0 and b.max() <= 8
    # the fixed-target loop lands near the target (SURVEY section 4 anchor: within ~0.01 for round mode)
    assert abs(float(O.get_bits_alloc_fixed_target(sig, 4, True).mean()) - 4) < 0.02
    assert abs(float(O.get_bits_alloc_fixed_target(sig, 5.3, True).mean()) - 5.3) < 0.02


def test_alpha_tables_and_multipliers(golden):
    arrays, _ = golden
    om, al = O.omega_alpha_tables()
    assert np.allclose(om, arrays["tables.omega"], rtol=0, atol=0)
    assert np.allclose(al, arrays["tables.alpha"], rtol=1e-9, atol=1e-12)
    # known answers: optimum of the Laplace MSE at omega = 2^M
    want = {1: 1.859, 2: 2.829, 3: 3.897, 4: 5.015, 5: 6.203, 6: 7.413, 7: 8.618, 8: 9.890}
    got = O.get_alpha_mult(torch.tensor([2.0 ** m for m in want]), sym=True)
    for g, (m, w) in zip(got, want.items()):
        assert abs(g - w) < 2e-3, (m, g, w)
    x = torch.from_numpy(arrays["tables.mult_in"])
    assert np.allclose(O.get_alpha_mult(x.clone(), True), arrays["tables.mult_sym"], rtol=1e-9)
    assert np.allclose(O.get_alpha_mult(x.clone()[:6], False), arrays["tables.mult_pos"], rtol=1e-9)
    # the hard-coded ACIQ factors are the 2-decimal roundings of the same optimum (int_quantizer.py:12-19, 84)
    for m in range(2, 9):
        assert abs(O.ALPHA_LAPLACE[m] - want[m]) < 0.03


def test_statistics(golden):
    arrays, _ = golden
    x = torch.from_numpy(regen(dict(seed=4000, shape=(6, 10, 5, 7))))
    names = ["min", "max", "mean", "b", "std"]
    for pre, st in (("tensor", O.act_stats(x, names)), ("sampleavg", O.act_stats(x, names, True)),
                    ("pc", O.act_stats_perchannel(x, names)), ("pcavg", O.act_stats_perchannel(x, names, True))):
        for k in names:
            _params_close(st[k].numpy(), arrays["stats.%s.%s" % (pre, k)], pre + k)


def test_weight_correction_properties():
    rs = np.random.RandomState(5)
    w = torch.from_numpy((rs.standard_normal((8, 4, 3, 3)) * 0.1).astype(np.float32))
    wq = O.quantize_weights_per_channel(w, 4)
    c = O.weight_correction(w, wq, bias_corr=True, var_corr=False)
    assert np.allclose(c.view(8, -1).mean(-1).numpy(), w.view(8, -1).mean(-1).numpy(), atol=1e-7)
    c2 = O.weight_correction(w, wq, bias_corr=True, var_corr=True)
    assert np.allclose(c2.view(8, -1).std(-1).numpy(), w.view(8, -1).std(-1).numpy(), rtol=1e-4)
    assert O.weight_correction(w, wq, False, False) is wq
