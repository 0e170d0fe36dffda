"""Phase-boundary timestamps of the fused kernels, NCHW next to channels-last (development tool; %globaltimer, us)."""
import ctypes, sys
import torch
sys.path.insert(0, ".")
import cnn_quantization_b200 as fq
from cnn_quantization_b200 import _lib as L
lib = L.load()
lib.fqb200_debug_timing.argtypes = [ctypes.c_void_p]
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
order = [(0, "start"), (13, "S1 streamed"), (1, "S1 combined"), (2, "L1 begin"), (3, "L1 end"), (4, "released"),
         (14, "S2 streamed"), (5, "S2 combined"), (6, "L2 begin"), (7, "L2 end"), (8, "released"), (9, "apply done")]
for (n, c, hw) in [(512, 64, 56), (512, 256, 14), (512, 128, 28), (512, 1024, 14), (512, 2048, 7)]:
    for cl in (False, True):
        x = torch.randn(n, c, hw, hw, device="cuda")
        if cl:
            x = x.contiguous(memory_format=torch.channels_last)
        lay = (n, c, hw * hw)
        for rep in range(3):
            flush.zero_()
            buf.zero_()
            lib.fqb200_debug_timing(buf.data_ptr())
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fq.ops.fused(x, lay, range_mode=L.RANGE_LAPLACE, num_bits=4, bit_alloc=True, out=x, channels_last=cl)
            e.record()
            torch.cuda.synchronize()
        lib.fqb200_debug_timing(None)
        t = buf.cpu().tolist()
        print((n, c, hw), "NHWC" if cl else "NCHW", "total %.1f us" % (s.elapsed_time(e) * 1e3),
              " | ".join("%s %.1f" % (nm, (t[i] - t[0]) / 1e3) for i, nm in order if t[i]))
