#!/usr/bin/env python
"""Offline-statistics fixtures from the REAL reference (build container only):

  1. `-sm collect` without and with -pcq_a on a seeded ResNet-18 (2 batches of 2 images, 64x64) -> the reference's own
     summary files, copied to tests/golden/ref_stats/ (CSV + pickle, a few KB);
  2. `-sm use` W4A4 (+ -bca) on the same model with those files -> logits in tests/golden/ref_stats_logits.npz.

The reference's statistics managers write under ~/mxt-sim; their module-level `base_dir` is pointed at a scratch
directory inside this repository instead.  The compiled leaf is routed to the CPU restatement exactly as in
make_census.py.
"""
import os
import shutil
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_census as mc  # noqa: E402  (sets up the reference import, stubs and the CPU leaf)

from pytorch_quantizer.quantization.inference import statistic_manager as sm_mod  # noqa: E402
from pytorch_quantizer.quantization.inference import statistic_manager_perchannel as smp_mod  # noqa: E402

SCRATCH = os.path.join(HERE, "_scratch_stats")
OUT = os.path.join(HERE, "ref_stats")
sm_mod.base_dir = SCRATCH
smp_mod.base_dir = SCRATCH

W4A4 = dict(qtype="int4", qweight="int4", clipping="laplace", per_channel_quant_weights=True, per_channel_quant_act=True,
            bit_alloc_act=True, bit_alloc_weight=True, bias_corr_weight=True)


def batches():
    rs = np.random.RandomState(2024)
    return [torch.from_numpy(rs.standard_normal((2, 3, 64, 64)).astype(np.float32)) for _ in range(2)]


def run(flags, xs):
    mc.Singleton._instances.clear()
    from itertools import count
    for cls in (mc.iqm.Conv2dWithId, mc.iqm.LinearWithId, mc.iqm.MaxPool2dWithId, mc.iqm.AvgPool2dWithId, mc.iqm.BatchNorm2dWithId):
        cls._id = count(0)
    args = mc.make_args(arch="resnet18", stats_folder="resnet18", **flags)
    outs = []
    with mc.iqm.QuantizationManagerInference(args, mc.qparams(args)) as qm:
        torch.manual_seed(12345)
        model = mc.models.resnet18(weights=None)
        mc.set_node_names(model)
        mc.resnet_mark_before_relu(model)
        mc.search_absorbe_bn(model)
        qm.bn_folding = True
        model.eval()
        qm.quantize_model(model)
        with torch.no_grad():
            for x in xs:
                outs.append(model(x).numpy())
    return np.stack(outs)


def main():
    torch.set_num_threads(8)
    shutil.rmtree(SCRATCH, ignore_errors=True)
    xs = batches()
    run(dict(stats_mode="collect", qtype="int4", qweight="int4"), xs)                              # per tensor
    run(dict(stats_mode="collect", qtype="int4", qweight="int4", per_channel_quant_act=True), xs)  # per channel
    logits = {"use_w4a4": run(dict(stats_mode="use", **W4A4), xs),
              "use_w4a4_bca": run(dict(stats_mode="use", bias_corr_act=True, **W4A4), xs),
              "use_w8a8": run(dict(stats_mode="use", qtype="int8", qweight="int8"), xs)}
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(os.path.join(OUT, "statistics", "resnet18"))
    os.makedirs(os.path.join(OUT, "statistics", "per_channel", "resnet18"))
    shutil.copy(os.path.join(SCRATCH, "statistics", "resnet18", "resnet18_summary.csv"), os.path.join(OUT, "statistics", "resnet18"))
    shutil.copy(os.path.join(SCRATCH, "statistics", "per_channel", "resnet18", "resnet18_statistics_perchannel_summary.pkl"),
                os.path.join(OUT, "statistics", "per_channel", "resnet18"))
    np.savez_compressed(os.path.join(HERE, "ref_stats_logits.npz"), **logits)
    shutil.rmtree(SCRATCH, ignore_errors=True)
    for k, v in logits.items():
        print(k, v.shape, float(np.abs(v).mean()))


if __name__ == "__main__":
    main()
