"""Per-CTA phase finish times (development tool)."""
import ctypes, sys
import numpy as np
import torch
sys.path.insert(0, ".")
import cnn_quantization_b200 as fq
from cnn_quantization_b200 import _lib as L
lib = L.load()
lib.fqb200_debug_timing.argtypes = [ctypes.c_void_p]
buf = torch.zeros(16 + 4 * 1024, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (n, c, hw) in [(512, 64, 112), (512, 256, 14)]:
    x = torch.randn(n, c, hw, hw, device="cuda")
    out = torch.empty_like(x)
    for rep in range(2):
        flush.zero_(); buf.zero_()
        lib.fqb200_debug_timing(buf.data_ptr())
        fq.ops.fused(x, (n, c, hw * hw), range_mode=L.RANGE_LAPLACE, num_bits=4, bit_alloc=True, out=out)
        torch.cuda.synchronize()
    lib.fqb200_debug_timing(None)
    t = buf.cpu().numpy()
    t0 = t[0]
    per = t[16:16 + 4 * 296].reshape(296, 4)
    s1 = (per[:, 0] - t0) / 1e3
    s2 = (per[:, 1] - t[4]) / 1e3
    ap = (per[:, 2] - t[8]) / 1e3
    sm = per[:, 3]
    for name, v in (("S1", s1), ("S2", s2), ("A", ap)):
        q = np.percentile(v, [0, 10, 25, 50, 75, 90, 100])
        print((n, c, hw), name, "finish us: min %.0f p10 %.0f p25 %.0f med %.0f p75 %.0f p90 %.0f max %.0f" % tuple(q))
    # correlation with SM id / CTA index
    order = np.argsort(s1)
    print("  fastest CTAs (idx, sm):", [(int(i), int(sm[i])) for i in order[:8]])
    print("  slowest CTAs (idx, sm):", [(int(i), int(sm[i])) for i in order[-8:]])
    # SMs by parity / die
    by_sm = {}
    for i in range(296):
        by_sm.setdefault(int(sm[i]), []).append(float(s1[i]))
    lone = [v for k, v in by_sm.items() if len(v) == 1]
    print("  SMs used:", len(by_sm), "SMs with 1 CTA:", len(lone), "with 2:", sum(1 for v in by_sm.values() if len(v) == 2), "3+:", sum(1 for v in by_sm.values() if len(v) > 2))
    smid = np.array(sorted(by_sm))
    mean_by = np.array([np.mean(by_sm[k]) for k in smid])
    half = len(smid) // 2
    print("  mean S1 finish, low-numbered SMs %.0f, high-numbered %.0f; even %.0f odd %.0f" % (mean_by[:half].mean(), mean_by[half:].mean(), mean_by[smid % 2 == 0].mean(), mean_by[smid % 2 == 1].mean()))
    print("  S1 finish vs CTA index quartiles:", [round(float(s1[i:i + 74].mean()), 0) for i in range(0, 296, 74)])
