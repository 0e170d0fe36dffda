import sys, torch
sys.path.insert(0, '.')
import cnn_quantization_b200 as fq
g = torch.Generator(device="cuda").manual_seed(1)
n = 1 << 24
for rep in range(8):
    a = torch.randn(n, device="cuda", generator=g) * (10.0 ** torch.randint(-6, 7, (n,), device="cuda", generator=g))
    b = torch.exp(torch.rand(n, device="cuda", generator=g) * 27.6 - 18.4)
    if rep % 2:
        bits = b.view(torch.int32)
        pat = torch.tensor([0x7FFFFF, 0x000000, 0x000001, 0x7FFFFE, 0x400000, 0x3FFFFF], device="cuda", dtype=torch.int32)
        sel = pat[torch.randint(0, 6, (n,), device="cuda", generator=g)]
        b = ((bits & ~0x7FFFFF) | sel).view(torch.float32)
    if rep == 7:
        a[:1000] = float("inf"); a[1000:2000] = float("nan"); a[2000:3000] = 3e38; a[3000:4000] = 1e-42
    fast, ieee = fq.ops._test_division(a.contiguous(), b.contiguous())
    bad = (fast != ieee) & ~(torch.isnan(fast) & torch.isnan(ieee))
    nb = int(bad.sum())
    if nb:
        qi = ieee[bad].abs()
        print(rep, "mismatch", nb, "ieee range", float(qi.min()), float(qi.max()), "a range", float(a[bad].abs().min()), float(a[bad].abs().max()))
        idx = bad.nonzero()[:5, 0]
        for i in idx.tolist():
            print("   a=%r b=%r fast=%r ieee=%r" % (float(a[i]), float(b[i]), float(fast[i]), float(ieee[i])))
    else:
        print(rep, "exact")
