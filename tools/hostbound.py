"""Host enqueue time against device time of one validation step (diagnostic): is a configuration launch-bound?
Usage: python tools/hostbound.py [config batch ...]"""
import sys
import time
import torch
sys.path.insert(0, ".")
from cnn_quantization_b200 import pipeline

args = sys.argv[1:] or ["resnet50_w4a4", "512", "resnet101_w4a4", "128", "resnet101_w4a4", "32"]
dev = torch.device("cuda:0")
for i in range(0, len(args), 2):
    config, batch = args[i], int(args[i + 1])
    model, qm = pipeline.build_quantized_model(config, dev, channels_last=True)
    x, t = pipeline.synthetic_batch(batch, seed=7, channels_last=True)
    x = x.to(dev).contiguous(memory_format=torch.channels_last)
    t = t.to(dev)
    with torch.no_grad():
        for _ in range(3):
            model(x)
        torch.cuda.synchronize()
        host, devt = [], []
        for _ in range(5):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            s.record()
            model(x)
            e.record()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            host.append((t1 - t0) * 1e3)
            devt.append(s.elapsed_time(e))
    print("%s batch %d: host enqueue %.2f ms, device %.2f ms per step" % (config, batch, min(host), min(devt)))
    qm.detach()
    del model, qm, x
    torch.cuda.empty_cache()
