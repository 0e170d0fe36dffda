#!/usr/bin/env python
"""Drive the REAL reference manager (class swap + quantize_model + one forward) on CPU and record every
quantize_instant call (id, tag, half_range, shape) plus the logits -> tests/golden/ref_census.json / ref_pipeline.npz.

Build container only (needs /root/reference).  The compiled leaf cannot run without a GPU, so
``IntQuantizer.__gemmlowpQuantize__`` is routed to the CPU restatement of kernels/gemmlowp.cu (oracle a1) with the
reference's own preserve_zero rule; every other line executed is the reference's.
"""
import json
import os
import sys
import types
from argparse import Namespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from oracle import fq_oracle as O  # noqa: E402

stub = types.ModuleType("int_quantization")
stub.float2gemmlowp = lambda t, d, o, b, ie, tz, noise: O.float2gemmlowp(t, float(d), float(o), b, ie, tz, None)
sys.modules["int_quantization"] = stub
for name in ("mlflow", "tensorboardX", "bokeh"):
    sys.modules.setdefault(name, types.ModuleType(name))

import torchvision.models as models  # noqa: E402
import pytorch_quantizer.quantization.qtypes.int_quantizer  # noqa: E402,F401
from pytorch_quantizer.quantization.inference import inference_quantization_manager as iqm  # noqa: E402
from utils.absorb_bn import search_absorbe_bn  # noqa: E402
from utils.mark_relu import resnet_mark_before_relu  # noqa: E402
from utils.model_naming import set_node_names  # noqa: E402
from utils.misc import Singleton  # noqa: E402

iq_mod = sys.modules["pytorch_quantizer.quantization.qtypes.int_quantizer"]


def _leaf_cpu(self, tensor, delta, offset):
    preserve_zero = self.enforce_true_zero and (offset + delta) > 0 and offset < 0
    return stub.float2gemmlowp(tensor.contiguous(), delta, offset, self.num_bits, self.int_exp, bool(preserve_zero), None)


iq_mod.IntQuantizer.__gemmlowpQuantize__ = _leaf_cpu
torch.Tensor.cuda = lambda self, *a, **k: self  # utils/absorb_bn.py:19-20 hard-codes .cuda()

CONFIGS = {
    # name: (arch, input hw, batch, flags)
    "resnet18_w4a4": ("resnet18", 64, 2, dict(qtype="int4", qweight="int4", clipping="laplace", per_channel_quant_weights=True,
                                               per_channel_quant_act=True, bit_alloc_act=True, bit_alloc_weight=True,
                                               bias_corr_weight=True)),
    "resnet50_w4a4": ("resnet50", 64, 2, dict(qtype="int4", qweight="int4", clipping="laplace", per_channel_quant_weights=True,
                                               per_channel_quant_act=True, bit_alloc_act=True, bit_alloc_weight=True,
                                               bias_corr_weight=True)),
    "resnet50_w8a8": ("resnet50", 64, 2, dict(qtype="int8", qweight="int8")),
    "vgg16_w4a4": ("vgg16", 64, 2, dict(qtype="int4", qweight="int4", clipping="laplace", per_channel_quant_weights=True,
                                         per_channel_quant_act=True, bit_alloc_act=True, bit_alloc_weight=True,
                                         bias_corr_weight=True, bit_alloc_target_act=5.3, bit_alloc_target_weight=5.3)),
}


def make_args(**over):
    d = dict(arch="resnet18", qtype=None, qweight="int8", q_off=False, clipping="no", stats_mode="no", stats_kind="mean",
             stats_folder=None, stats_batch_avg=False, kld_threshold=False, measure_stats=False,
             per_channel_quant_weights=False, per_channel_quant_act=False, bit_alloc_act=False, bit_alloc_weight=False,
             bit_alloc_rmode="round", bit_alloc_prior="gaus", bit_alloc_target_act=None, bit_alloc_target_weight=None,
             bias_corr_act=False, bias_corr_weight=False, var_corr_weight=False, measure_entropy=False,
             mid_thread_quant=False, rho_act=None, rho_weight=None, preserve_zero=False)
    d.update(over)
    return Namespace(**d)


def qparams(a):
    return {"int": {"clipping": a.clipping, "stats_kind": a.stats_kind, "true_zero": a.preserve_zero, "kld": a.kld_threshold,
                    "pcq_weights": a.per_channel_quant_weights, "pcq_act": a.per_channel_quant_act,
                    "bit_alloc_act": a.bit_alloc_act, "bit_alloc_weight": a.bit_alloc_weight,
                    "bit_alloc_rmode": a.bit_alloc_rmode, "bit_alloc_prior": a.bit_alloc_prior,
                    "bit_alloc_target_act": a.bit_alloc_target_act, "bit_alloc_target_weight": a.bit_alloc_target_weight,
                    "bcorr_act": a.bias_corr_act, "bcorr_weight": a.bias_corr_weight, "vcorr_weight": a.var_corr_weight,
                    "logger": None, "measure_entropy": a.measure_entropy, "mtd_quant": a.mid_thread_quant},
            "qmanager": {"rho_act": a.rho_act, "rho_weight": a.rho_weight}}


def run(name):
    arch, hw, batch, flags = CONFIGS[name]
    Singleton._instances.clear()
    for cls in (iqm.Conv2dWithId, iqm.LinearWithId, iqm.MaxPool2dWithId, iqm.AvgPool2dWithId, iqm.BatchNorm2dWithId):
        from itertools import count
        cls._id = count(0)
    args = make_args(arch=arch, **flags)
    calls = []
    orig = iqm.TruncationOpManagerInference.quantize_instant

    def spy(self, tensor, id, tag="", stat_id=None, half_range=False, override_att=None, verbose=False):
        calls.append([id, tag, bool(half_range), list(tensor.shape)])
        return orig(self, tensor, id, tag, stat_id, half_range, override_att, False)

    iqm.TruncationOpManagerInference.quantize_instant = spy
    try:
        with iqm.QuantizationManagerInference(args, qparams(args)) as qm:
            torch.manual_seed(12345)
            model = models.__dict__[arch](weights=None)
            set_node_names(model)
            if "resnet" in arch:
                resnet_mark_before_relu(model)
                search_absorbe_bn(model)
                qm.bn_folding = True
            model.eval()
            qm.quantize_model(model)
            n_weight_calls = len(calls)
            rs = np.random.RandomState(12345)
            x = torch.from_numpy(rs.standard_normal((batch, 3, hw, hw)).astype(np.float32))
            with torch.no_grad():
                y = model(x)
    finally:
        iqm.TruncationOpManagerInference.quantize_instant = orig
    return dict(weight_calls=calls[:n_weight_calls], act_calls=calls[n_weight_calls:]), y.numpy()


def main():
    torch.set_num_threads(8)
    census, logits = {}, {}
    for name in CONFIGS:
        c, y = run(name)
        c.update(arch=CONFIGS[name][0], hw=CONFIGS[name][1], batch=CONFIGS[name][2], flags=CONFIGS[name][3])
        census[name] = c
        logits[name] = y
        print(name, len(c["weight_calls"]), "weight calls,", len(c["act_calls"]), "activation calls")
    with open(os.path.join(HERE, "ref_census.json"), "w") as f:
        json.dump(census, f)
    np.savez_compressed(os.path.join(HERE, "ref_pipeline.npz"), **logits)


if __name__ == "__main__":
    main()
