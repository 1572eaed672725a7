#!/usr/bin/env python3
"""How does the reference's CPU path scale with threads on this box?  (cgroup quota, SMT, noisy neighbours...)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    if os.path.exists(p):
        print(p, open(p).read().strip())
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
print(open("/proc/loadavg").read().strip())
print('usable_cpus', bench.usable_cpus())
for t in (1, 8, 16, 32, 64, 128):
    if t > (os.cpu_count() or 1):
        break
    r = bench.cpu_reference_throughput(16000, 160, n_inst=max(8 * t, 64), prep_calls=440, timed_calls=1500 if t >= 8 else 200, threads=t)
    print(f"threads {t:4d}: {r['value']:.4g} blocks/s  ({r['value'] / t:.4g} per thread, {r['seconds']:.2f} s)")
