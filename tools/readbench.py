"""Read-only streaming experiments (development tool): what bandwidth do the statistics phases reach, and what does a
plain reduction reach on the same tensor?"""
import sys
import torch
sys.path.insert(0, ".")
import cnn_quantization_b200 as fq
from cnn_quantization_b200 import _lib as L

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


for (n, c, hw) in [(512, 64, 112), (512, 256, 56), (512, 64, 56), (512, 256, 14)]:
    x = torch.randn(n, c, hw, hw, device="cuda")
    gb = x.numel() * 4 / 1e9
    lay = (n, c, hw * hw)
    t = timeit(lambda: fq.ops.fused(x, lay, stats_only=True))
    print("%s stats_only (S1+S2, 2 reads): %.3f ms  %.0f GB/s" % ((n, c, hw), t, 2 * gb / t * 1e3))
    t = timeit(lambda: fq.ops.fused(x, (1, 1, x.numel()), stats_only=True))
    print("%s stats_only per-tensor G=1 (flat): %.3f ms  %.0f GB/s" % ((n, c, hw), t, 2 * gb / t * 1e3))
    t = timeit(lambda: x.sum())
    print("%s torch sum (1 read): %.3f ms  %.0f GB/s" % ((n, c, hw), t, gb / t * 1e3))
    t = timeit(lambda: x.amax())
    print("%s torch amax (1 read): %.3f ms  %.0f GB/s" % ((n, c, hw), t, gb / t * 1e3))
    t = timeit(lambda: torch.relu_(x))
    print("%s torch relu_ (1R+1W): %.3f ms  %.0f GB/s" % ((n, c, hw), t, 2 * gb / t * 1e3))
