#!/usr/bin/env python
"""bench.py - ResNet-50 W4A4 (ACIQ Laplace + per-channel bit allocation + weight bias correction) inference
throughput through the fused sm_100a fake-quantization path.

    python bench.py --gpus N --steps K --warmup W            # this framework (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's pure-PyTorch algorithm on host CPU cores

A step = one forward of the hooked model over one synthetic 512x3x224x224 batch per GPU (55 hooked activation
tensors, 5.79 G elements; BASELINE.json config "ResNet-50 W4A4 -pcq_w -pcq_a -c laplace -baa -baw -bcw, batch 512").
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "resnet50_w4a4_images_per_s"
UNIT = "images/s"


# DRAM traffic of the dominant kernel, from profiles/ (bench.py cannot run under ncu itself): ncu dram__bytes_read.sum +
# dram__bytes_write.sum averaged over the 53 mode-D launches of one step of the named workload.
NCU_TRAFFIC_BYTES_PER_LAUNCH = {
    ("resnet50_w4a4", 512, True): (1527.45e6, "profiles/r01e_dram_bytes_fused_launches_step1.csv"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fqb200", choices=["fqb200", "reference"])
    ap.add_argument("--config", default="resnet50_w4a4")
    ap.add_argument("--batch", type=int, default=512, help="images per GPU per step")
    ap.add_argument("--cpu-batch", type=int, default=32, help="images per step of the CPU reference arm / cpu_baseline")
    ap.add_argument("--nchw", action="store_true",
                    help="keep the model's tensors in contiguous NCHW memory (fq_fused_kernel) instead of the default "
                         "torch.channels_last memory format (fq_fused_nhwc_kernel, no cuDNN layout conversions)")
    ap.add_argument("--channels-last", action="store_true", help="accepted for compatibility: this is the default")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short BASELINE configs[1] (W8A8) run")
    args = ap.parse_args()
    args.channels_last = not args.nchw
    return args


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """SM clock, power and throttle reasons sampled through NVML every 100 ms while the timed region runs."""

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.thread, self.err = index, [], threading.Event(), None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.rows.append((sm, pw, rs))
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                return
            time.sleep(0.1)

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        seen = set()
        for _, _, rs in self.rows:
            for k, bit in names.items():
                if rs & bit:
                    seen.add(k)
        busy = sorted(sm for sm, pw, _ in self.rows if pw > 250.0) or sorted(sm for sm, _, _ in self.rows)
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": self.max_sm, "samples": len(self.rows),
                "power_w_max": max((pw for _, pw, _ in self.rows), default=None), "reasons": sorted(seen)}


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's algorithm (oracle port) on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_pipeline_images_per_s(config, batch, steps, warmup):
    import torch
    from cnn_quantization_b200 import pipeline
    from oracle import fq_oracle
    from oracle.host import host_threads
    cores = host_threads()
    torch.set_num_threads(cores)
    model, qm = pipeline.build_quantized_model(config, "cpu", quantizer_factory=fq_oracle.oracle_int_quantizer)
    x, t = pipeline.synthetic_batch(batch, seed=1)
    with torch.no_grad():
        for _ in range(warmup):
            pipeline.accuracy_counts(model(x), t)
        t0 = time.perf_counter()
        for _ in range(steps):
            pipeline.accuracy_counts(model(x), t)
        dt = time.perf_counter() - t0
    qm.detach()
    return batch * steps / dt, dt / steps, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 10))  # ~3 s per 32-image step on 16 host threads
    warm = max(0, min(args.warmup, 2))
    ips, sec, cores = cpu_pipeline_images_per_s(args.config, args.cpu_batch, steps, warm)
    sample = "%d steps of %d images (of the %d-image batch) through the oracle port of int_quantizer.py, %d host threads" % (
        steps, args.cpu_batch, args.batch, cores)
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[2]: %s, batch %d/GPU, 224x224" % (args.config, args.batch),
                       "cpu_sample_batch": args.cpu_batch},
            "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# this framework
# ---------------------------------------------------------------------------------------------------
def run_fqb200(args):
    import torch
    import torch.distributed as dist
    import cnn_quantization_b200 as fq
    from cnn_quantization_b200 import ops, pipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the fake-quantization path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    fq._lib.load()
    torch.backends.cudnn.benchmark = True  # inference_sim.py:205

    model, qm = pipeline.build_quantized_model(args.config, dev, channels_last=args.channels_last)
    x_host, t_host = pipeline.synthetic_batch(args.batch, seed=1000 + rank, pin=True, channels_last=args.channels_last)
    x_dev, t_dev = x_host.to(dev), t_host.to(dev)
    total = torch.zeros(4, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        with torch.no_grad():
            total.add_(pipeline.accuracy_counts(model(x_dev), t_dev))

    host_metrics = torch.zeros(4).pin_memory()

    def step_e2e():
        with torch.no_grad():
            x = x_host.to(dev, non_blocking=True)
            t = t_host.to(dev, non_blocking=True)
            host_metrics.copy_(pipeline.accuracy_counts(model(x), t), non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller reads the step's metrics

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()

    # ---- timed region 1: inputs resident in HBM ---------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.profile_reset(enable=True)
    barrier()
    if os.environ.get("FQB_CUDA_PROFILER"):  # ncu --profile-from-start off: capture the timed region only
        torch.cuda.profiler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    barrier()
    if os.environ.get("FQB_CUDA_PROFILER"):
        torch.cuda.profiler.stop()
    ms = ev0.elapsed_time(ev1)
    prof = ops.profile_collect()
    ops.profile_reset(enable=False)
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 2: end to end from pinned host memory --------------------------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    t_ms = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t_ms.tolist()
    loss, top1, top5, n_img = pipeline.reduce_metrics(total)  # the path's one collective

    if rank == 0:
        peak, peak_src = peaks()
        images = world * args.batch * args.steps
        value = images / (ms / 1e3)
        e2e = images / (ms_e2e / 1e3)
        # dominant kernel: fq_fused_kernel in mode D (3 reads + 1 write = 16 B/element), per-launch CUDA events
        dom = prof["modes"].get("D", {"launches": 0, "elems": 0, "ms": 0.0, "bytes": 0})
        achieved = (dom["bytes"] / 1e9) / (dom["ms"] / 1e3) if dom["ms"] > 0 else 0.0
        traffic, traffic_src = NCU_TRAFFIC_BYTES_PER_LAUNCH.get((args.config, args.batch, args.channels_last), (None, None))
        quant_ms = sum(m["ms"] for m in prof["modes"].values())
        quant_elems = sum(m["elems"] for m in prof["modes"].values())
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[2]: %s (-pcq_w -pcq_a -c laplace -baa -baw -bcw), batch %d per GPU, "
                                   "3x224x224, random-init torchvision weights" % (args.config, args.batch),
                       "parallelism": "dp%d (batch sharded per rank, one all-reduce of 4 metrics)" % world,
                       "l2": "inputs larger than L2 (308 MB input, every hooked tensor 51 MB - 1.6 GB)",
                       "memory_format": "torch.channels_last (same logical NCHW tensors and results; --nchw selects contiguous NCHW)"
                                        if args.channels_last else "contiguous NCHW",
                       "conv": "cuDNN fp32 %s via torch (third party in the reference too)" % ("NHWC" if args.channels_last else "NCHW")},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": x_host.numel() * 4 + t_host.numel() * 8, "d2h_bytes_per_step": 16},
            "gpu_launches": prof["launches"],
            "roofline": {"bound": "hbm", "kernel": "%s mode D (stats, deviations, apply)" % (
                             "fq_fused_nhwc_kernel" if args.channels_last else "fq_fused_kernel<4>"),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "peak_source": peak_src, "traffic": traffic, "traffic_unit": "bytes per launch (ncu, DRAM read + write)",
                         "traffic_source": traffic_src, "launches": dom["launches"],
                         "algorithmic_bytes_per_launch": dom["bytes"] / max(dom["launches"], 1),
                         "algorithmic_bytes_per_elem": 16, "avg_launch_ms": dom["ms"] / max(dom["launches"], 1)},
            "quant": {"gelem_per_s": quant_elems / (quant_ms / 1e3) / 1e9 if quant_ms else None,
                      "ms_per_step": quant_ms / args.steps, "share_of_step": quant_ms / ms,
                      "modes": {k: {"launches": v["launches"], "ms": v["ms"], "GBps": (v["bytes"] / 1e9) / (v["ms"] / 1e3) if v["ms"] else None}
                                for k, v in prof["modes"].items()},
                      "by_layout": {k: {"launches": v["launches"], "ms_avg": v["ms"] / v["launches"],
                                        "frac": (v["bytes"] / 1e9) / (v["ms"] / 1e3) / peak if v["ms"] else None}
                                    for k, v in sorted(prof["shapes"].items(), key=lambda kv: -kv[1]["ms"])}},
            "clocks": clocks,
            "check": {"loss": loss, "top1": top1, "top5": top5, "images": n_img},
        }
        if world == 1 and not args.no_secondary and args.config == "resnet50_w4a4":
            line["config1_w8a8"] = secondary(args, dev, peak, "resnet50_w8a8", False, "B",
                                             "BASELINE configs[1]: resnet50_w8a8 (--qtype int8), per-sample min/max + apply")
            other = not args.channels_last
            line["channels_last_variant" if other else "nchw_variant"] = secondary(
                args, dev, peak, args.config, other, "D",
                "headline config with the model's tensors in %s" % (
                    "torch.channels_last memory (fq_fused_nhwc_kernel, cuDNN NHWC convs)" if other else
                    "contiguous NCHW memory (fq_fused_kernel, cuDNN converts layouts internally); select with --nchw"))
        if world == 1 and not args.no_cpu_baseline:
            ips, sec, cores = cpu_pipeline_images_per_s(args.config, args.cpu_batch, 3, 1)
            line["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "3 steps (after 1 warm-up) of %d images through the oracle port of int_quantizer.py "
                                              "+ torch CPU convs, %d host threads (cgroup quota)" % (args.cpu_batch, cores)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def secondary(args, dev, peak, config, channels_last, mode, note):
    """A short (3 timed steps, inputs resident) run of another configuration next to the headline one; reports the
    pipeline rate and the roofline of the dominant kernel mode (D: 16 B/element, B: 12 B/element)."""
    import torch
    from cnn_quantization_b200 import ops, pipeline
    model, qm = pipeline.build_quantized_model(config, dev, channels_last=channels_last)
    x, t = pipeline.synthetic_batch(args.batch, seed=7, channels_last=channels_last)
    x, t = x.to(dev), t.to(dev)
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        for _ in range(3):
            pipeline.accuracy_counts(model(x), t)
        torch.cuda.synchronize()
        ops.profile_reset(enable=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            pipeline.accuracy_counts(model(x), t)
        e1.record()
        torch.cuda.synchronize()
    prof = ops.profile_collect()
    ops.profile_reset(enable=False)
    qm.detach()
    del model, qm, x
    torch.cuda.empty_cache()
    ms = e0.elapsed_time(e1) / 3
    b = prof["modes"].get(mode, {"bytes": 0, "ms": 0.0, "launches": 0})
    gbs = (b["bytes"] / 1e9) / (b["ms"] / 1e3) if b["ms"] else None
    quant_ms = sum(m["ms"] for m in prof["modes"].values()) / 3
    return {"workload": "%s, batch %d" % (note, args.batch), "value": args.batch / (ms / 1e3),
            "unit": UNIT, "ms_per_step": ms, "steps": 3, "quant_ms_per_step": quant_ms,
            "roofline": {"kernel": "fused kernel mode %s" % mode, "algorithmic_bytes_per_elem": {"D": 16, "B": 12}[mode],
                         "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak if gbs else None,
                         "launches": b["launches"]}}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_fqb200(args)


if __name__ == "__main__":
    main()
