#!/usr/bin/env python
"""Run the REFERENCE's own inference manager (class swap, quantize_model, one forward) in a fresh process, either

  --mode reference   as shipped: its Python + its compiled ``int_quantization`` extension (oracle/_ref), or
  --mode dropin      after the zero-edit alias of INTEGRATION.md section 1 - ``sys.modules["int_quantization"]`` and
                     ``pytorch_quantizer.quantization.qtypes.int_quantizer`` point at this package BEFORE the reference
                     manager is imported - so every ``quantize_instant`` of the unmodified manager lands in libfqb200.so.

Writes logits and bookkeeping to --out (npz).  Test infrastructure (GPU box only; uses oracle/_ref).
"""
import argparse
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["reference", "dropin"], required=True)
    ap.add_argument("--config", default="resnet50_w4a4")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--hw", type=int, default=224)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()

    import numpy as np
    import torch
    from oracle import ref_live
    from oracle.host import host_threads
    torch.set_num_threads(host_threads())
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    launches = -1
    if a.mode == "dropin":
        # ---- INTEGRATION.md section 1, verbatim --------------------------------------------------------------
        import cnn_quantization_b200 as fq
        sys.modules["int_quantization"] = fq.int_quantization                       # (a) the compiled extension
        for name in ("mlflow", "tensorboardX", "bokeh"):
            sys.modules.setdefault(name, types.ModuleType(name))
        sys.path.insert(0, ref_live.PYREF)
        import pytorch_quantizer.quantization.qtypes as qtypes                       # (b) the quantizer
        import importlib
        fq_iq = importlib.import_module("cnn_quantization_b200.int_quantizer")   # the MODULE (fq.int_quantizer is the factory)
        sys.modules["pytorch_quantizer.quantization.qtypes.int_quantizer"] = fq_iq
        qtypes.int_quantizer = fq_iq.int_quantizer
        # ---------------------------------------------------------------------------------------------------------
        ns = ref_live.load(extension=fq.int_quantization)
        assert ns.iq is fq_iq and ns.qtypes.int_quantizer is fq_iq.int_quantizer
    else:
        ns = ref_live.load()
    from cnn_quantization_b200 import manager as M, pipeline   # argument namespace / synthetic batch helpers only
    flags = dict(pipeline.CONFIGS[a.config])
    args = M.make_args(**flags)
    model, qm = ref_live.build_reference_model(args, M.get_params(args), "cuda")
    quantizer_class = type(qm.op_manager.get_quantizer("activation")).__module__
    x, _ = pipeline.synthetic_batch(a.batch, seed=11, hw=a.hw)
    with torch.no_grad():
        y = model(x.cuda())
    torch.cuda.synchronize()
    qm.__exit__()
    if a.mode == "dropin":
        from cnn_quantization_b200 import ops
        launches = ops._prof["launches"]
    np.savez(a.out, logits=y.float().cpu().numpy(), launches=np.int64(launches), quantizer_module=np.array(quantizer_class))
    print("ok", a.mode, quantizer_class, launches)


if __name__ == "__main__":
    main()
