"""GPU pipeline: the hook manager with the CUDA quantizers on the same seeded models / inputs the real reference ran
(tests/golden/make_census.py), plus multi-batch validation bookkeeping."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def census():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    with open(os.path.join(GOLD, "ref_census.json")) as f:
        return json.load(f), np.load(os.path.join(GOLD, "ref_pipeline.npz"))


@pytest.mark.parametrize("name", ["resnet18_w4a4", "resnet50_w4a4", "resnet50_w8a8", "vgg16_w4a4"])
def test_cuda_pipeline_call_sites_and_logits(census, name):
    from cnn_quantization_b200 import pipeline
    meta, logits = census
    info = meta[name]
    torch.backends.cudnn.allow_tf32 = False  # compare against fp32 CPU convolutions
    torch.backends.cuda.matmul.allow_tf32 = False
    model, qm = pipeline.build_quantized_model(dict(arch=info["arch"], **info["flags"]), "cuda")
    qm.record = True
    rs = np.random.RandomState(12345)
    x = torch.from_numpy(rs.standard_normal((info["batch"], 3, info["hw"], info["hw"])).astype(np.float32)).cuda()
    with torch.no_grad():
        y = model(x).cpu().numpy()
    qm.detach()
    calls = [[c[0], c[1], c[2], list(c[3])] for c in qm.calls]
    assert calls == info["act_calls"]
    ref = logits[name]
    # 4-bit grids amplify last-ulp differences of cuDNN vs CPU convolutions into occasional one-step flips that then
    # propagate, so logits are compared as a whole: same direction, same scale
    cos = float((y * ref).sum() / (np.linalg.norm(y) * np.linalg.norm(ref)))
    assert cos > 0.95, cos
    assert abs(np.linalg.norm(y) / np.linalg.norm(ref) - 1) < 0.1


def test_weights_match_reference_quantize_model(census):
    """quantize_model with the fused CUDA weight launch (quantize + bias correction) == the oracle manager on CPU."""
    from cnn_quantization_b200 import pipeline
    from oracle import fq_oracle as O
    meta, _ = census
    info = meta["resnet18_w4a4"]
    flags = dict(arch=info["arch"], **info["flags"])
    m_gpu, q1 = pipeline.build_quantized_model(flags, "cuda")
    m_cpu, q2 = pipeline.build_quantized_model(flags, "cpu", quantizer_factory=O.oracle_int_quantizer)
    q1.detach()
    q2.detach()
    for (n1, p1), (n2, p2) in zip(m_gpu.named_parameters(), m_cpu.named_parameters()):
        assert n1 == n2
        a, b = p1.detach().cpu().numpy(), p2.detach().numpy()
        scale = float(np.abs(b).max()) + 1e-12
        bad = np.abs(a - b) > 1e-5 * scale
        assert bad.mean() <= 2e-3, (n1, float(bad.mean()))


def test_validate_accumulates_like_average_meters():
    from cnn_quantization_b200 import pipeline
    model, qm = pipeline.build_quantized_model("resnet18_w4a4", "cuda")
    batches = [pipeline.synthetic_batch(4, seed=s, hw=64) for s in (1, 2, 3)]
    total = pipeline.validate(model, batches, "cuda")
    loss, top1, top5, n = pipeline.reduce_metrics(total)
    qm.detach()
    assert n == 12 and 0 <= top1 <= top5 <= 100 and np.isfinite(loss)


@pytest.mark.parametrize("config", ["resnet50_w8a8", "resnet50_w4a4", "resnet101_w4a4", "vgg16_w4a4", "vgg16_w4a4_mtq", "resnet18_w4a4"])
def test_every_baseline_config_runs_and_quantizes(config):
    """Every BASELINE.json configuration (small batch / image): finite logits, activations really on a coarse grid."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cnn_quantization_b200 import ops, pipeline
    model, qm = pipeline.build_quantized_model(config, "cuda")
    qm.record = True
    x, t = pipeline.synthetic_batch(4, seed=3, hw=64)
    ops.profile_reset(enable=True)
    with torch.no_grad():
        y = model(x.cuda())
    prof = ops.profile_collect()
    ops.profile_reset(enable=False)
    qm.detach()
    assert torch.isfinite(y).all()
    arch = pipeline.CONFIGS[config]["arch"]
    expected = {"resnet50": 55, "resnet101": 106, "vgg16": 21, "resnet18": 22}[arch]
    blocks = {"resnet50": 16, "resnet101": 33, "vgg16": 0, "resnet18": 8}[arch]
    assert len(qm.calls) == expected
    quant = sum(v["launches"] for k, v in prof["modes"].items() if k not in ("E", "P"))
    assert quant == expected  # exactly one kernel launch per hooked tensor
    # + per residual block either one fused add+ReLU kernel ("E") or - where the launch of the block's last convolution
    # can take the shortcut as an operand (int8 per-sample min/max and per-channel Laplace, both on channels-last) -
    # nothing at all: that launch's mode carries an "r"
    fused = sum(v["launches"] for k, v in prof["modes"].items() if k.endswith("r"))
    assert prof["modes"].get("E", {"launches": 0})["launches"] + fused == blocks
    assert fused == 0   # NCHW input here; the channels-last census is test_channels_last_pipeline_matches_nchw


def test_bias_buffer_follows_the_module_and_detach_restores_it():
    """ADVICE (round 1, medium): the fused conv bias must survive .to() / state_dict() while attached."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cnn_quantization_b200 import pipeline
    model, qm = pipeline.build_quantized_model("resnet18_w4a4", "cuda")
    conv = model.layer1[0].conv1
    assert conv.bias is None and conv._fq_bias.is_cuda
    before = conv._fq_bias.clone()
    model.cpu()
    assert not conv._fq_bias.is_cuda           # the buffer moved with the module
    model.cuda()
    assert torch.equal(conv._fq_bias, before)
    qm.detach()
    assert conv.bias is not None and conv.bias.is_cuda and torch.equal(conv.bias.data, before)
    assert "_fq_bias" not in dict(conv.named_buffers())


def test_relu_is_not_skipped_after_an_inplace_modification():
    """ADVICE (round 1, low): the non-negativity tag dies with any in-place op on the tensor."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cnn_quantization_b200.manager import _relu_forward_skipping
    relu = torch.nn.ReLU()
    relu.forward = _relu_forward_skipping(relu)
    x = torch.rand(8, device="cuda")
    x._fq_nonneg = x._version
    assert relu(x) is x
    x -= 1.0
    y = relu(x)
    assert float(y.min()) >= 0.0


def test_empty_positive_range_passthrough_is_rectified_when_the_relu_is_fused():
    """ADVICE (round 1, low): compiled leaf, half range, every sample <= 0 -> the reference hands the input back and its
    ReLU zeroes it; with the ReLU skipped the kernel's pass-through branch must do that."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import cnn_quantization_b200 as fq
    from test_gpu_parity import params
    x = -torch.rand(4, 8, 6, 6, device="cuda") - 0.1
    q = fq.int_quantizer("int8", params())
    q.half_range = True
    y = q(x.clone(), "conv1_activation", "activation")                      # reference behaviour: input handed back
    assert torch.equal(y, x)
    y = q(x.clone(), "conv1_activation", "activation", relu_follows=True)   # fused ReLU: rectified, tagged
    assert float(y.abs().max()) == 0.0 and y._fq_nonneg == y._version


@pytest.mark.parametrize("config", ["resnet50_w4a4", "vgg16_w4a4", "resnet50_w8a8"])
def test_pipeline_extensions_do_not_change_results(config):
    """Conv-bias fusion, in-place activations, skipping the ReLU after a half-range quantization and the fused residual
    add + ReLU are exact rewrites: switching them all off gives bit-identical logits."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cnn_quantization_b200 import manager as M, pipeline
    import torchvision.models as models
    x, _ = pipeline.synthetic_batch(4, seed=11, hw=64)

    def run(native_extensions):
        flags = dict(pipeline.CONFIGS[config])
        args = M.make_args(**flags)
        qm = M.QuantizationManagerInference(args, M.get_params(args))
        if not native_extensions:
            qm.fuse_conv_bias = qm.skip_redundant_relu = qm.fuse_residual_relu = qm.fast_maxpool = False
            for q in list(qm.quantizers.values()) + [qm.quantizer_default]:
                if hasattr(q, "inplace"):
                    q.inplace = False
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
        model.eval().cuda()
        qm.quantize_model(model)
        qm.attach(model)
        with torch.no_grad():
            y = model(x.cuda().clone())
        qm.detach()
        return y

    a, b = run(True), run(False)
    assert torch.equal(a, b)


@pytest.mark.parametrize("config", ["vgg16_w4a4", "vgg16_w4a4_mtq"])
def test_vgg_pooling_runs_inside_the_quantization_launches(config):
    """VGG-16 W4A4 on channels-last memory: the five max poolings run inside the launches of the convolutions in front of
    them ("Dp"), no pooling kernel is left, and the logits agree with the run that keeps them separate."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cnn_quantization_b200 import ops, pipeline
    x, _ = pipeline.synthetic_batch(4, seed=13, hw=64)
    xin = x.cuda().contiguous(memory_format=torch.channels_last)
    outs = []
    for fuse in (True, False):
        model, qm = pipeline.build_quantized_model(config, "cuda", channels_last=True)
        if not fuse:
            for m in model.modules():
                m.__dict__.pop("_fq_pool_module", None)
        qm.record = True
        ops.profile_reset(enable=True)
        with torch.no_grad():
            outs.append(model(xin.clone()).float().cpu().numpy())
        prof = ops.profile_collect()
        ops.profile_reset(enable=False)
        qm.detach()
        assert len(qm.calls) == 21
        pooled = sum(v["launches"] for k, v in prof["modes"].items() if k.endswith("p"))
        assert pooled == (5 if fuse else 0)
        assert prof["modes"].get("P", {"launches": 0})["launches"] == (0 if fuse else 5)
    a, b = outs
    cos = float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b)))
    assert cos > 0.999, cos


def test_int8_channels_last_block_epilogue_is_exact():
    """configs[1] on channels-last memory: the 16 block epilogues run inside the per-sample min/max launches ("Br") and
    the logits are bit-identical to the run that keeps them as separate add + ReLU kernels."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cnn_quantization_b200 import ops, pipeline
    x, _ = pipeline.synthetic_batch(4, seed=9, hw=64)
    xin = x.cuda().contiguous(memory_format=torch.channels_last)
    from cnn_quantization_b200.int_quantizer import IntQuantizer
    outs = []
    # (block epilogue in the launch, shortcut of the 4 down-sampling blocks deferred, the launch refuses the operand)
    for fuse, defer, refuse in ((True, True, False), (True, False, False), (False, False, False), (True, True, True)):
        model, qm = pipeline.build_quantized_model("resnet50_w8a8", "cuda", channels_last=True)
        qm.fuse_residual_into_quant, qm.defer_shortcut = fuse, defer
        keep = IntQuantizer._residual_kw
        if refuse:   # the fallback: a deferred shortcut is quantized after all (manager.finish_deferred), then add + ReLU
            IntQuantizer._residual_kw = lambda self, *a, **k: {}
        try:
            ops.profile_reset(enable=True)
            with torch.no_grad():
                outs.append(model(xin.clone()))
            prof = ops.profile_collect()
        finally:
            IntQuantizer._residual_kw = keep
            ops.profile_reset(enable=False)
            qm.detach()
        fused = sum(v["launches"] for k, v in prof["modes"].items() if k.endswith("r"))
        assert fused == (16 if fuse and not refuse else 0)
        assert prof["modes"].get("E", {"launches": 0})["launches"] == (0 if fuse and not refuse else 16)
        assert prof["modes"].get("S", {"launches": 0})["launches"] == (4 if defer else 0)
        assert sum(v["launches"] for k, v in prof["modes"].items() if k.endswith("p")) == 1 and "P" not in prof["modes"]   # the stem
    for o in outs[1:]:
        assert torch.equal(outs[0], o)


def test_channels_last_pipeline_matches_nchw():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cnn_quantization_b200 import ops, pipeline
    torch.backends.cudnn.allow_tf32 = False
    x, _ = pipeline.synthetic_batch(4, seed=5, hw=64)
    outs = []
    for cl in (False, True):
        model, qm = pipeline.build_quantized_model("resnet50_w4a4", "cuda", channels_last=cl)
        xin = x.cuda().contiguous(memory_format=torch.channels_last) if cl else x.cuda()
        ops.profile_reset(enable=True)
        with torch.no_grad():
            outs.append(model(xin).float().cpu().numpy())
        prof = ops.profile_collect()
        ops.profile_reset(enable=False)
        qm.detach()
        # NCHW: 55 hooked tensors + 16 fused residual add + ReLU kernels; channels-last: the 16 block epilogues run inside
        # the quantization launch of the block's last convolution and the stem's max pooling inside that of the first one
        assert prof["launches"] == (55 if cl else 55 + 16)
        if cl:
            assert sum(v["launches"] for k, v in prof["modes"].items() if k.endswith("p")) == 1 and "P" not in prof["modes"]
        if cl:
            assert sum(v["launches"] for k, v in prof["modes"].items() if k.endswith("r")) == 16
            assert prof["modes"]["S"]["launches"] == 4   # the shortcut convolutions of the 4 down-sampling blocks: statistics only
    a, b = outs
    cos = float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b)))
    assert cos > 0.97, cos
