"""Mode-D timing over the census for the current FQB_* knobs + leader sub-stamps (development tool)."""
import ctypes, sys
import torch
sys.path.insert(0, ".")
import cnn_quantization_b200 as fq
from cnn_quantization_b200 import _lib as L
lib = L.load()
lib.fqb200_debug_timing.argtypes = [ctypes.c_void_p]
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
census = [(64, 112), (64, 56), (256, 56), (128, 28), (512, 28), (256, 14), (1024, 14), (512, 14)]
tot_ms = 0.0
w = {(64, 112): 1, (64, 56): 6, (256, 56): 4, (128, 56): 1, (128, 28): 7, (512, 28): 5, (256, 28): 1, (256, 14): 11, (1024, 14): 7, (512, 14): 1}
for (c, hw) in census:
    x = torch.randn(512, c, hw, hw, device="cuda")
    out = torch.empty_like(x)
    ts = []
    for rep in range(6):
        flush.zero_(); buf.zero_()
        lib.fqb200_debug_timing(buf.data_ptr())
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fq.ops.fused(x, (512, c, hw * hw), range_mode=L.RANGE_LAPLACE, num_bits=4, bit_alloc=True, out=out)
        e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    lib.fqb200_debug_timing(None)
    ts.sort(); ms = ts[len(ts) // 2]
    t = buf.cpu().tolist()
    r = lambda i: (t[i] - t[0]) / 1e3 if t[i] else float("nan")
    tot_ms += ms * w[(c, hw)]
    print("%4dx%3d  %.3f ms  frac %.3f | S1 %.0f L1 %.1f S2 %.0f L2 %.1f (reduce %.1f, b/std %.1f, bitalloc %.1f, params %.1f) A %.0f" % (
        c, hw, ms, x.numel() * 16 / ms / 1e6 / 6577.4, r(2), r(3) - r(2), r(6) - r(4), r(7) - r(6), r(10) - r(6), r(11) - r(10), r(12) - r(11),
        r(7) - r(12), ms * 1e3 - r(8)))
print("census-weighted total (no 7x7, no 128x56/256x28/512x14 singles' exact weights): %.2f ms" % tot_ms)
