"""Time of the ResNet stem convolution (512x3x224x224 -> 64, 7x7, stride 2) in cuDNN by input memory format / channel
padding (diagnostic; the convolution is third party, the question is which call the pipeline should make)."""
import torch
import torch.nn.functional as F

torch.backends.cudnn.benchmark = True
dev = "cuda"
x = torch.randn(512, 3, 224, 224, device=dev)
w = torch.randn(64, 3, 7, 7, device=dev) * 0.05


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


xc = x.contiguous(memory_format=torch.channels_last)
wc = w.contiguous(memory_format=torch.channels_last)
x4 = F.pad(x, (0, 0, 0, 0, 0, 1)).contiguous(memory_format=torch.channels_last)
w4 = F.pad(w, (0, 0, 0, 0, 0, 1)).contiguous(memory_format=torch.channels_last)
x8 = F.pad(x, (0, 0, 0, 0, 0, 5)).contiguous(memory_format=torch.channels_last)
w8 = F.pad(w, (0, 0, 0, 0, 0, 5)).contiguous(memory_format=torch.channels_last)
ref = F.conv2d(x, w, None, 2, 3)
print("NCHW in -> NCHW out            %.3f ms" % t(lambda: F.conv2d(x, w, None, 2, 3)))
print("NHWC C=3 -> NHWC out           %.3f ms" % t(lambda: F.conv2d(xc, wc, None, 2, 3)))
print("NHWC C=4 (zero pad) -> NHWC    %.3f ms" % t(lambda: F.conv2d(x4, w4, None, 2, 3)))
print("NHWC C=8 (zero pad) -> NHWC    %.3f ms" % t(lambda: F.conv2d(x8, w8, None, 2, 3)))
print("NCHW in + convert out to NHWC  %.3f ms" % t(lambda: F.conv2d(x, w, None, 2, 3).contiguous(memory_format=torch.channels_last)))
print("pad C 3->4 of the input alone  %.3f ms" % t(lambda: F.pad(xc, (0, 0, 0, 0, 0, 1)).contiguous(memory_format=torch.channels_last)))
for name, y in (("C=3 NHWC", F.conv2d(xc, wc, None, 2, 3)), ("C=4", F.conv2d(x4, w4, None, 2, 3)), ("C=8", F.conv2d(x8, w8, None, 2, 3))):
    print(name, "max |diff| vs NCHW result: %.3g (out is channels-last: %s)" % (float((y - ref).abs().max()), y.is_contiguous(memory_format=torch.channels_last)))
