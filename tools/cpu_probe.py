"""How many host threads are actually usable on this box, and how the CPU oracle pipeline scales with them."""
import os, sys, time
sys.path.insert(0, ".")
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try:
        print(f, open(f).read().strip())
    except OSError as e:
        print(f, "n/a")
print("loadavg", open("/proc/loadavg").read().strip())
import torch
from cnn_quantization_b200 import pipeline
from oracle import fq_oracle
model, qm = pipeline.build_quantized_model("resnet50_w4a4", "cpu", quantizer_factory=fq_oracle.oracle_int_quantizer)
for batch in (2, 16):
    x, t = pipeline.synthetic_batch(batch, seed=1)
    for th in (8, 16, 32, 64, 128):
        if th > (os.cpu_count() or 1):
            continue
        torch.set_num_threads(th)
        with torch.no_grad():
            t0 = time.perf_counter(); model(x); t1 = time.perf_counter()
            if t1 - t0 < 20:
                model(x)
            t2 = time.perf_counter()
        print("batch", batch, "threads", th, "first %.2f s second %.2f s" % (t1 - t0, t2 - t1), flush=True)
        if t1 - t0 > 30:
            break
