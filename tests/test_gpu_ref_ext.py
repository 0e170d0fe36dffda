"""a1 against the REFERENCE's own CUDA extension, executed here on the B200.

oracle/build_oracle.py compiles kernels/int_quantization.cpp + gemmlowp.cu unmodified (sm_100a) into oracle/_ref/ in the
build container; the .so travels to the GPU box with the snapshot.  This is the strongest pin for the compiled leaf:
same inputs, the reference's kernel vs ours, bit for bit (including the FFMA contraction of the non-true-zero form)."""
import glob
import importlib.util
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref_ext():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    found = glob.glob(os.path.join(ROOT, "oracle", "_ref", "int_quantization*.so"))
    if not found:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    spec = importlib.util.spec_from_file_location("int_quantization", found[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("n", [1, 5, 1024, 100003, 1 << 22])
@pytest.mark.parametrize("bits", [2, 4, 8])
def test_float2gemmlowp_bit_exact_vs_reference_kernel(ref_ext, n, bits):
    import cnn_quantization_b200 as fq
    g = torch.Generator(device="cuda").manual_seed(n + bits)
    x = torch.randn(n, device="cuda", generator=g) * 2 + 0.3
    zeros = torch.zeros_like(x)
    noise = torch.rand(n, device="cuda", generator=g) - 0.5
    for tz, rng, off in ((True, 7.3, -3.1), (False, 5.0, 0.0), (False, 6.0, 0.5), (True, 9.0, -1e-4), (False, 4.0, -1.0),
                         (True, 1e-3, -4e-4), (True, 300.0, -150.0)):
        for nz in (zeros, noise):
            want = ref_ext.float2gemmlowp(x, rng, off, bits, False, tz, nz)
            got = fq.int_quantization.float2gemmlowp(x, rng, off, bits, False, tz, None if nz is zeros else nz)
            torch.cuda.synchronize()
            assert torch.equal(got, want), (tz, rng, off, nz is noise)
    want = ref_ext.float2gemmlowp(x, 7.3, -3.1, bits, True, True, zeros)  # int_exp
    assert torch.equal(fq.int_quantization.float2gemmlowp(x, 7.3, -3.1, bits, True, True, None), want)
    assert ref_ext.float2gemmlowp(x, 0.0, 0.0, bits, False, True, zeros) is not None  # range <= 0: input returned


def test_minmax_path_vs_reference_kernel_with_reference_host_logic(ref_ext):
    """gemmlowpMinMaxQuantize (int_quantizer.py:361-379, 605-614) with the reference's own kernel as the leaf vs our fused
    launch: the range comes from torch reductions there and from our statistics phase here."""
    import cnn_quantization_b200 as fq
    from test_gpu_parity import params
    torch.manual_seed(5)
    for shape, tag, positive in (((8, 16, 14, 14), "activation", False), ((8, 16, 14, 14), "activation", True),
                                 ((4, 1000), "activation_classifier", False), ((64, 32, 3, 3), "weight", False)):
        x = torch.randn(*shape, device="cuda")
        if positive:
            x = torch.relu(x)
        avg = "activation" in tag and "classifier" not in tag
        t = x.view(x.shape[0], -1) if avg else x.view(-1)
        mn = t.min(-1)[0].mean() if avg else t.min()
        mx = t.max(-1)[0].mean() if avg else t.max()
        if positive:
            mn = 0
        delta = mx - mn
        preserve_zero = bool((mn + delta) > 0 and mn < 0)
        want = ref_ext.float2gemmlowp(x.contiguous(), float(delta), float(mn), 8, False, preserve_zero, torch.zeros_like(x))
        q = fq.int_quantizer("int8", params())
        q.half_range = positive
        got = q(x, "id", tag)
        step = float(delta) / 255
        bad = (got - want).abs() > 1e-5 * want.abs() + 1e-9
        assert float(bad.float().mean()) <= 2e-4
        assert float((got - want).abs().max()) <= step * 1.01
