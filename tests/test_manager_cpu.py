"""The hook-based manager (cnn_quantization_b200.manager) against the REAL reference manager.

tests/golden/make_census.py drove the reference's class-swap manager + quantize_model + one forward on CPU and
recorded every quantize_instant call and the logits.  Here the very same call sites run through this package's
manager with the CPU oracle injected as quantizer factory (the CUDA quantizer cannot run without a GPU; the GPU
variant of this test is in test_gpu_pipeline.py).  Call list must match exactly, logits to fp32 round-off.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import fq_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def census():
    with open(os.path.join(GOLD, "ref_census.json")) as f:
        return json.load(f), np.load(os.path.join(GOLD, "ref_pipeline.npz"))


def run_ours(name, info, factory):
    from cnn_quantization_b200 import pipeline
    flags = dict(arch=info["arch"], **info["flags"])
    import cnn_quantization_b200.manager as M
    args = M.make_args(**flags)
    qm = M.QuantizationManagerInference(args, M.get_params(args), quantizer_factory=factory)
    qm.record = True
    import torchvision.models as models
    qm.enable()
    try:
        torch.manual_seed(12345)
        model = models.__dict__[args.arch](weights=None)
    finally:
        qm.stop_stamping()
    M.set_node_names(model)
    if "resnet" in args.arch:
        M.resnet_mark_before_relu(model)
        M.search_absorbe_bn(model)
        qm.bn_folding = True
    model.eval()
    qm.quantize_model(model)
    n_w = len(qm.calls)
    qm.attach(model)
    rs = np.random.RandomState(12345)
    x = torch.from_numpy(rs.standard_normal((info["batch"], 3, info["hw"], info["hw"])).astype(np.float32))
    with torch.no_grad():
        y = model(x)
    qm.detach()
    calls = [[c[0], c[1], c[2], list(c[3])] for c in qm.calls]
    return calls[:n_w], calls[n_w:], y.numpy()


@pytest.mark.parametrize("name", ["resnet18_w4a4", "resnet50_w4a4", "resnet50_w8a8", "vgg16_w4a4"])
def test_call_sites_and_logits_match_reference(census, name):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    meta, logits = census
    info = meta[name]
    w_calls, a_calls, y = run_ours(name, info, O.oracle_int_quantizer)
    assert w_calls == info["weight_calls"]
    assert a_calls == info["act_calls"]
    ref = logits[name]
    assert y.shape == ref.shape
    assert np.allclose(y, ref, rtol=1e-4, atol=1e-5 * float(np.abs(ref).max())), float(np.abs(y - ref).max())


def test_census_counts_match_survey(census):
    """55 hooked activation tensors / 54 weight tensors for ResNet-50; 21 / 16 for VGG-16."""
    meta, _ = census
    assert len(meta["resnet50_w4a4"]["act_calls"]) == 55 and len(meta["resnet50_w4a4"]["weight_calls"]) == 54
    assert len(meta["vgg16_w4a4"]["act_calls"]) == 21 and len(meta["vgg16_w4a4"]["weight_calls"]) == 16
    half = sum(1 for c in meta["resnet50_w4a4"]["act_calls"] if c[2])
    assert half == 33


def test_tag_table_overrides():
    import cnn_quantization_b200.manager as M
    args = M.make_args(arch="vgg16", qtype="int4", qweight="int4", clipping="laplace", per_channel_quant_weights=True,
                       per_channel_quant_act=True, bit_alloc_act=True, bit_alloc_weight=True, bias_corr_weight=True)
    qm = M.QuantizationManagerInference(args, M.get_params(args), quantizer_factory=O.oracle_int_quantizer)
    q = qm.quantizers
    assert q["activation"].num_bits == 4 and q["activation"].force_positive and not q["activation"].pcq_w
    assert q["activation_linear"].pcq_a is False and q["activation_linear"].clipping == "laplace"
    assert q["activation_classifier"].num_bits == 8 and q["activation_classifier"].clipping == "no"
    assert q["activation_pooling"].num_bits == 8 and not q["activation_pooling"].pcq_a
    assert q["weight"].num_bits == 4 and q["weight"].pcq_w and q["weight"].clipping == "no"
    assert q["weight_classifier"].num_bits == 8 and q["weight_classifier"].pcq_w
    assert qm.get_quantizer("no-such-tag") is qm.quantizer_default
    assert qm.ignore_ids == ["conv0_activation"]
    with pytest.raises(NotImplementedError):  # offline statistics need this package's CUDA quantizers
        M.QuantizationManagerInference(M.make_args(qtype="int8", stats_mode="use"), {}, quantizer_factory=O.oracle_int_quantizer)
    with pytest.raises(FileNotFoundError):    # use mode without collected files
        a = M.make_args(qtype="int8", stats_mode="use", stats_base_dir="/nonexistent/fqb200")
        M.QuantizationManagerInference(a, M.get_params(a))


def test_bn_folding_is_exact_for_eval_bn():
    import cnn_quantization_b200.manager as M
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, bias=False), torch.nn.BatchNorm2d(8), torch.nn.ReLU())
    net[1].running_mean.normal_()
    net[1].running_var.uniform_(0.5, 2)
    net[1].weight.data.normal_()
    net[1].bias.data.normal_()
    net.eval()
    x = torch.randn(2, 3, 9, 9)
    want = net(x)
    M.search_absorbe_bn(net)
    assert getattr(net[1], "absorbed", False)
    got = net(x)  # BN now has mean 0 / var 1 / no affine
    assert torch.allclose(got, want, atol=1e-5)
