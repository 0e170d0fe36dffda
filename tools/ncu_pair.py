"""One NCHW and one channels-last fused launch on the same tensor shape, for an `ncu --set full` A/B capture."""
import sys
import torch
sys.path.insert(0, ".")
import cnn_quantization_b200 as fq
from cnn_quantization_b200 import _lib as L
n, c, hw = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (512, 64, 56)))
for cl in (False, True):
    x = torch.randn(n, c, hw, hw, device="cuda")
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    fq.ops.fused(x, (n, c, hw * hw), range_mode=L.RANGE_LAPLACE, num_bits=4, bit_alloc=True, out=x, channels_last=cl)
    torch.cuda.synchronize()
