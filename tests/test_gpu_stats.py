"""Offline statistics (`-sm collect` / `-sm use`) and activation bias correction (`-bca`, rank 2)
against fixtures produced by the REAL reference's statistics managers (tests/golden/make_stats_golden.py):
same on-disk formats in both directions, same use-mode results."""
import os
import pickle
import shutil

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
W4A4 = dict(qtype="int4", qweight="int4", clipping="laplace", per_channel_quant_weights=True, per_channel_quant_act=True,
            bit_alloc_act=True, bit_alloc_weight=True, bias_corr_weight=True)


def batches():
    rs = np.random.RandomState(2024)
    return [torch.from_numpy(rs.standard_normal((2, 3, 64, 64)).astype(np.float32)) for _ in range(2)]


def run(flags, base_dir):
    from cnn_quantization_b200 import pipeline
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = dict(arch="resnet18", stats_folder="resnet18", stats_base_dir=base_dir, **flags)
    model, qm = pipeline.build_quantized_model(cfg, "cuda")
    outs = []
    with torch.no_grad():
        for x in batches():
            outs.append(model(x.cuda()).cpu().numpy())
    qm.__exit__()
    return np.stack(outs)


@pytest.fixture(scope="module")
def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.mark.parametrize("name,flags", [("use_w4a4", W4A4), ("use_w4a4_bca", dict(bias_corr_act=True, **W4A4)),
                                        ("use_w8a8", dict(qtype="int8", qweight="int8"))])
def test_use_mode_with_reference_statistics(need_gpu, name, flags):
    """Statistics files written by the reference drive this package's use mode; logits match the reference's use mode."""
    ref = np.load(os.path.join(GOLD, "ref_stats_logits.npz"))[name]
    got = run(dict(stats_mode="use", **flags), os.path.join(GOLD, "ref_stats"))
    assert got.shape == ref.shape
    cos = float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref)))
    assert cos > 0.97, cos
    assert abs(np.linalg.norm(got) / np.linalg.norm(ref) - 1) < 0.08


def test_collect_mode_writes_the_reference_formats(need_gpu, tmp_path):
    import pandas as pd
    base = str(tmp_path)
    run(dict(stats_mode="collect", qtype="int4", qweight="int4"), base)
    run(dict(stats_mode="collect", qtype="int4", qweight="int4", per_channel_quant_act=True), base)
    ours = pd.read_csv(os.path.join(base, "statistics", "resnet18", "resnet18_summary.csv"), index_col=0)
    ref = pd.read_csv(os.path.join(GOLD, "ref_stats", "statistics", "resnet18", "resnet18_summary.csv"), index_col=0)
    assert list(ours.columns) == list(ref.columns)
    assert list(ours.index) == list(ref.index)
    assert list(ours["internal_name"]) == list(ref["internal_name"])
    for stat in ("min", "max", "mean", "std", "b", "mean_abs", "kurtosis", "dim"):
        for kind in ("min", "mean", "max"):
            col = "%s_%s" % (kind, stat)
            a, b = ours[col].to_numpy(dtype=np.float64), ref[col].to_numpy(dtype=np.float64)
            scale = np.abs(b).max() + 1e-12
            assert np.allclose(a, b, rtol=2e-3, atol=2e-4 * scale), (col, np.abs(a - b).max())
    with open(os.path.join(base, "statistics", "per_channel", "resnet18", "resnet18_statistics_perchannel_summary.pkl"), "rb") as f:
        mine = pickle.load(f)
    with open(os.path.join(GOLD, "ref_stats", "statistics", "per_channel", "resnet18", "resnet18_statistics_perchannel_summary.pkl"), "rb") as f:
        theirs = pickle.load(f)
    assert sorted(mine) == sorted(theirs)
    for layer in theirs:
        assert list(mine[layer].columns) == list(theirs[layer].columns)
        assert len(mine[layer]) == len(theirs[layer])
        for col in ("min_min", "max_max", "mean_mean", "mean_std", "mean_b", "mean_std_pos", "mean_max", "mean_min"):
            a, b = mine[layer][col].to_numpy(dtype=np.float64), theirs[layer][col].to_numpy(dtype=np.float64)
            scale = np.abs(b).max() + 1e-12
            assert np.allclose(a, b, rtol=2e-3, atol=5e-4 * scale), (layer, col, np.abs(a - b).max())
    # and the files written here are usable by this package's own use mode
    got = run(dict(stats_mode="use", **W4A4), base)
    ref_logits = np.load(os.path.join(GOLD, "ref_stats_logits.npz"))["use_w4a4"]
    cos = float((got * ref_logits).sum() / (np.linalg.norm(got) * np.linalg.norm(ref_logits)))
    assert cos > 0.97, cos


def test_use_mode_is_single_pass(need_gpu):
    """With offline statistics every activation is an apply-only launch (mode A, 8 B/element): no statistics phases."""
    from cnn_quantization_b200 import ops, pipeline
    cfg = dict(arch="resnet18", stats_folder="resnet18", stats_base_dir=os.path.join(GOLD, "ref_stats"), stats_mode="use", **W4A4)
    model, qm = pipeline.build_quantized_model(cfg, "cuda")
    x = batches()[0].cuda()
    with torch.no_grad():
        model(x)
        ops.profile_reset(enable=True)
        model(x)
    prof = ops.profile_collect()
    ops.profile_reset(enable=False)
    qm.detach()
    assert set(prof["modes"]) - {"E"} == {"A"}, prof["modes"].keys()   # "E": the fused residual add + ReLU of the 8 blocks
    assert prof["modes"]["A"]["launches"] == 22


def test_use_mode_on_channels_last_memory_finishes_the_blocks_in_the_launch(need_gpu):
    """`-sm use` on a channels-last model: the given-parameter launch of a block's last convolution also does the block's
    residual add + ReLU (FQB200_RANGE_GIVEN through fqb200_fused), and the stem's max pooling where the stem is 4-bit;
    logits equal the NCHW run's (pure elementwise rewrites: bit-identical up to cuDNN's NHWC / NCHW convolution order)."""
    from cnn_quantization_b200 import ops, pipeline
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = dict(arch="resnet18", stats_folder="resnet18", stats_base_dir=os.path.join(GOLD, "ref_stats"), stats_mode="use", **W4A4)
    x = batches()[0].cuda()
    outs = []
    for cl, fuse in ((False, True), (True, True), (True, False)):
        model, qm = pipeline.build_quantized_model(cfg, "cuda", channels_last=cl)
        qm.fuse_residual_into_quant = fuse
        xin = x.contiguous(memory_format=torch.channels_last) if cl else x
        ops.profile_reset(enable=True)
        with torch.no_grad():
            outs.append(model(xin).float())
        prof = ops.profile_collect()
        ops.profile_reset(enable=False)
        qm.detach()
        fused = sum(v["launches"] for k, v in prof["modes"].items() if k.endswith("r"))
        assert fused == (8 if cl and fuse else 0)
        assert prof["modes"].get("E", {"launches": 0})["launches"] == (0 if cl and fuse else 8)
        # the shortcut convolutions of the 3 down-sampling blocks launch nothing at all: quantized inside the consuming launch
        assert sum(v["launches"] for k, v in prof["modes"].items() if k[0] == "A") == (19 if cl and fuse else 22)
    nchw, cl_fused, cl_plain = outs
    assert torch.equal(cl_fused, cl_plain)   # given parameters: the epilogue in the launch is an exact rewrite
    a, b = nchw.cpu().numpy(), cl_fused.cpu().numpy()   # cuDNN's NHWC / NCHW kernels sum in different orders: 4-bit grids flip
    cos = float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b)))
    assert cos > 0.97, cos


def test_given_parameters_through_the_descriptor_entry_point(need_gpu):
    """FQB200_RANGE_GIVEN == fqb200_quantize1 (bit-equal), with bias, block epilogue and pooling on top."""
    import torch.nn.functional as F
    from cnn_quantization_b200 import _lib as L, ops
    g = torch.Generator(device="cuda").manual_seed(3)
    n, c, h, w = 6, 96, 12, 10
    x = (torch.randn(n, c, h, w, device="cuda", generator=g) * 1.4).contiguous(memory_format=torch.channels_last)
    r = torch.randn(n, c, h, w, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(c, device="cuda", generator=g) * 0.2
    delta = torch.rand(c, device="cuda", generator=g) * 3 + 0.5
    offset = -torch.rand(c, device="cuda", generator=g) * 1.5
    bits = torch.randint(1, 7, (c,), device="cuda", generator=g).float()
    lay = (n, c, h * w)
    for bt in (None, bits):
        base = ops.quantize1(x, delta, offset, 4, bits=bt, layout=lay, bias=bias)
        kw = dict(channels_last=True, range_mode=L.RANGE_GIVEN, num_bits=4, given=(delta, offset, bt), bias=bias)
        assert torch.equal(ops.fused(x, lay, **kw), base)
        assert torch.equal(ops.fused(x, lay, residual=r, residual_relu=True, **kw), torch.relu(base + r))
        assert torch.equal(ops.fused(x, lay, pool=(2, 2), **kw), F.max_pool2d(base, 2))
        assert torch.equal(ops.fused(x, lay, pool=(3, 3), **kw), F.max_pool2d(base, 3, 2, 1))
    y = x.clone()
    out = ops.fused(y, lay, out=y, channels_last=True, range_mode=L.RANGE_GIVEN, num_bits=4, given=(delta, offset, None), bias=bias)
    assert out.data_ptr() == y.data_ptr() and torch.equal(out, ops.quantize1(x, delta, offset, 4, layout=lay, bias=bias))
    with pytest.raises(Exception):
        ops.fused(x, lay, channels_last=True, range_mode=L.RANGE_GIVEN, num_bits=4)   # no parameters
