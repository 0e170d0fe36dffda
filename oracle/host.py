"""Test infrastructure (like everything under oracle/): how many host threads the CPU legs may use."""
import os


def host_threads():
    """Host threads this process can really run: the affinity mask capped by the cgroup CPU quota.  (os.cpu_count() is
    the machine's 128; the GPU boxes give a container 16 CPUs of quota, and 128 OpenMP threads under that quota spend
    their time throttled at barriers - a 500x slowdown that would flatter the GPU arm.)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:  # cgroup v2
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, -(-int(quota) // int(period))))
    except (OSError, ValueError):
        try:  # cgroup v1
            quota = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if quota > 0 and period > 0:
                n = min(n, max(1, -(-quota // period)))
        except (OSError, ValueError):
            pass
    return max(1, n)
