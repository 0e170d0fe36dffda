"""Inference pipeline around the fake-quantization hot path: the part of the reference's
``inference/inference_sim.py`` (``InferenceModel.__init__`` :131-229, ``validate`` :278-343) that the benchmark
needs, on synthetic ImageNet-shaped batches and random-init torchvision weights (no dataset / checkpoints offline).

One process per GPU: each rank builds the same seeded model, quantizes the same weights (deterministic, so no
broadcast), runs its shard of the batch, and the four validation counters are combined with ONE all-reduce at the
end (the reference uses single-process ``torch.nn.DataParallel``, inference_sim.py:196-200).  On-the-fly statistics
are therefore per shard - exactly what DataParallel replicas compute in the reference.
"""
import torch
import torch.nn as nn

from . import manager as M

__all__ = ["CONFIGS", "build_quantized_model", "synthetic_batch", "validate", "accuracy_counts", "reduce_metrics"]

# BASELINE.json configs -> reference CLI flags
_W4A4 = dict(qtype="int4", qweight="int4", clipping="laplace", per_channel_quant_weights=True, per_channel_quant_act=True,
             bit_alloc_act=True, bit_alloc_weight=True, bias_corr_weight=True)
CONFIGS = {
    "resnet50_w8a8": dict(arch="resnet50", qtype="int8", qweight="int8"),
    "resnet50_w4a4": dict(arch="resnet50", **_W4A4),
    "resnet101_w4a4": dict(arch="resnet101", **_W4A4),
    "vgg16_w4a4": dict(arch="vgg16", bit_alloc_target_act=5.3, bit_alloc_target_weight=5.3, **_W4A4),
    "vgg16_w4a4_mtq": dict(arch="vgg16", bit_alloc_target_act=5.3, bit_alloc_target_weight=5.3, mid_thread_quant=True, **_W4A4),
    "resnet18_w4a4": dict(arch="resnet18", **_W4A4),
}


def build_quantized_model(config, device, seed=12345, quantizer_factory=None, channels_last=False):
    """Model creation as in InferenceModel.__init__: seeded random-init torchvision model (the reference loads
    pretrained weights; none are available offline), node names, before-relu marks and BN folding for ResNets,
    ``.to(device)``, ``quantize_model``.  Returns (model, manager); the manager stays attached and enabled."""
    import torchvision.models as models
    flags = dict(CONFIGS[config]) if isinstance(config, str) else dict(config)
    args = M.make_args(**flags)
    qm = M.QuantizationManagerInference(args, M.get_params(args), quantizer_factory=quantizer_factory)
    qm.enable()
    try:
        torch.manual_seed(seed)  # inference_sim.py:127
        model = models.__dict__[args.arch](weights=None)
    finally:
        qm.stop_stamping()
    M.set_node_names(model)
    if "resnet" in args.arch:
        M.resnet_mark_before_relu(model)
    if "resnet" in args.arch or args.arch in ("vgg16_bn", "inception_v3"):
        M.search_absorbe_bn(model)
        qm.bn_folding = True
    model.eval()
    model.to(device)
    if channels_last:
        model.to(memory_format=torch.channels_last)
    qm.quantize_model(model)
    qm.attach(model)
    return model, qm


def synthetic_batch(batch, seed, device="cpu", hw=224, pin=False, channels_last=False):
    """ImageNet-shaped input batch + labels; N(0,1) per pixel is what a normalised image roughly looks like."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 3, hw, hw, generator=g)
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    t = torch.randint(0, 1000, (batch,), generator=g)
    if pin and torch.cuda.is_available():
        x, t = x.pin_memory(), t.pin_memory()
    if device != "cpu":
        x, t = x.to(device, non_blocking=True), t.to(device, non_blocking=True)
    return x, t


def accuracy_counts(output, target):
    """[loss_sum, correct@1, correct@5, count] as a device tensor (no host sync): the sums behind the reference's
    AverageMeters (inference_sim.py:319-325, utils/meters.py:81-95)."""
    loss_sum = nn.functional.cross_entropy(output, target, reduction="sum")
    _, pred = output.topk(5, 1, True, True)
    hit = pred.eq(target.view(-1, 1))
    c1 = hit[:, :1].sum()
    c5 = hit.sum()
    return torch.stack([loss_sum.float(), c1.float(), c5.float(), torch.tensor(float(target.numel()), device=output.device)])


def validate(model, batches, device):
    """``validate()`` of the reference on an iterable of (input, target) host batches: H2D copy, forward through the
    hooked model, metric accumulation on the device.  Returns the 4-vector of accuracy_counts summed over batches."""
    total = torch.zeros(4, device=device)
    with torch.no_grad():
        for x, t in batches:
            x = x.to(device, non_blocking=True)
            t = t.to(device, non_blocking=True)
            total += accuracy_counts(model(x), t)
    return total


def reduce_metrics(total):
    """The only collective of the path: all-reduce(SUM) of [loss_sum, correct@1, correct@5, count] over the ranks,
    then (loss, top1 %, top5 %) like the reference's AverageMeter averages."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    loss_sum, c1, c5, n = total.tolist()
    return loss_sum / n, 100.0 * c1 / n, 100.0 * c5 / n, int(n)
