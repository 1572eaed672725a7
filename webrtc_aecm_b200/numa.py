"""Bind the calling process to the CPUs (hence, by first touch, the memory) of the NUMA node a GPU hangs off.

Host planes that feed a GPU over PCIe should live on the socket that owns the GPU's root port: DMA from the
remote socket runs at a fraction of the link rate.  Pure host plumbing (sysfs + sched_setaffinity)."""
from __future__ import annotations

import os


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_local_cpus(device_index: int):
    """CPUs local to the GPU's PCIe root (empty set if unknown)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        # nvml reports an 8-hex-digit domain, sysfs uses 4
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/local_cpulist"
        with open(path) as f:
            return _parse_cpulist(f.read())
    except Exception:
        return set()


def bind_to_gpu(device_index: int) -> int:
    """Restrict the process to the GPU-local CPUs; returns how many CPUs remain (0 = left unchanged)."""
    cpus = gpu_local_cpus(device_index)
    allowed = os.sched_getaffinity(0)
    cpus &= allowed
    if not cpus:
        return 0
    os.sched_setaffinity(0, cpus)
    return len(cpus)
