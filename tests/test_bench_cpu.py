"""bench.py contract checks that need no GPU: the reference arm (CPU oracle port) and the host-thread detection."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_threads_respects_affinity_and_quota():
    from oracle.host import host_threads
    n = host_threads()
    assert 1 <= n <= len(os.sched_getaffinity(0))
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
    except OSError:
        return
    if quota != "max":
        assert n <= -(-int(quota) // int(period))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-batch", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "resnet50_w4a4_images_per_s" and line["unit"] == "images/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1
    # the reference's own Python when build() staged it (oracle/_ref/pyref), else the oracle port
    from oracle import ref_live
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_live.python_available() else "port")
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["cpu_baseline"]["cores"] >= 1 and "images" in line["cpu_baseline"]["sample"]
    assert 0.0 < line["cpu_baseline"]["quant_share_of_step"] < 1.0
    assert "configs[2]" in line["config"]["workload"] and line["config"]["steps_requested"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_metric_and_workload_follow_the_config():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-batch", "1", "--config", "resnet18_w4a4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["metric"] == "resnet18_w4a4_images_per_s" and "resnet18_w4a4" in line["config"]["workload"]
    assert len(out.stdout.strip().splitlines()) == 1   # ONE line on stdout


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
