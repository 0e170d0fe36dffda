"""Phase-boundary timestamps of the fused kernel (development tool; %globaltimer, ns)."""
import ctypes, sys
import torch
sys.path.insert(0, ".")
import cnn_quantization_b200 as fq
from cnn_quantization_b200 import _lib as L
lib = L.load()
lib.fqb200_debug_timing.argtypes = [ctypes.c_void_p]
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
names = ["start", "b0 S1 done", "leader1 begin", "leader1 end", "b0 released", "b0 S2 done", "leader2 begin", "leader2 end",
         "b0 released", "b0 apply done"]
for (n, c, hw) in [(512, 64, 112), (512, 64, 56), (512, 256, 14), (512, 1024, 14), (512, 2048, 7), (512, 512, 7)]:
    x = torch.randn(n, c, hw, hw, device="cuda")
    out = torch.empty_like(x)
    lay = (n, c, hw * hw)
    for inplace in (False, True):
        for rep in range(3):
            flush.zero_()
            buf.zero_()
            lib.fqb200_debug_timing(buf.data_ptr())
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fq.ops.fused(x, lay, range_mode=L.RANGE_LAPLACE, num_bits=4, bit_alloc=True, out=(x if inplace else out))
            e.record()
            torch.cuda.synchronize()
        lib.fqb200_debug_timing(None)
        t = buf.cpu().tolist()
        rel = [(v - t[0]) / 1e3 if v else None for v in t[:10]]
        print((n, c, hw), "inplace" if inplace else "outofplace", "total %.1f us" % (s.elapsed_time(e) * 1e3),
              " | ".join("%s %.1f" % (nm, r) for nm, r in zip(names, rel) if r is not None))
