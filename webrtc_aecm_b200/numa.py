"""Place the calling process -- its threads and, more importantly, the pinned host planes it is about to allocate --
next to the GPU it feeds over PCIe.

Host planes that feed a GPU should live in the memory of the socket that owns the GPU's PCIe root port: DMA from the
other socket crosses the inter-socket link and runs at a fraction of the link rate (measured here: 2.6 MB H2D in 226 us
remote vs 50-74 us local).  Two mechanisms, both plain Linux (sysfs + sched_setaffinity + set_mempolicy):

* CPU affinity to the GPU-local CPUs, so that first-touch places memory there -- possible only when the container may
  run on those CPUs at all;
* a *memory policy* (MPOL_PREFERRED) for the GPU's node, which works even when the container's cpuset confines every
  thread to one socket (then `nvidia-smi topo` shows the same "CPU affinity" for all GPUs and eight processes
  would otherwise stack all their pinned planes on that one node) -- provided the cpuset's mems allow the node and
  the container may call set_mempolicy.

Nothing here touches the GPU; `bind_to_gpu` returns a dict describing what it found and did (it goes into bench.py's
JSON line as ``host``)."""
from __future__ import annotations

import ctypes
import glob
import os

_SYS_SET_MEMPOLICY = 238  # x86_64
_MPOL_PREFERRED = 1


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


def _gpu_sysfs_dir(device_index: int):
    import pynvml

    pynvml.nvmlInit()
    # CUDA and NVML enumerate alike unless CUDA_VISIBLE_DEVICES reorders: map through it when it is a list of indices
    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    idx = device_index
    if vis and all(p.strip().isdigit() for p in vis.split(",")):
        idx = int(vis.split(",")[device_index])
    h = pynvml.nvmlDeviceGetHandleByIndex(idx)
    bus = pynvml.nvmlDeviceGetPciInfo(h).busId
    bus = bus.decode() if isinstance(bus, bytes) else bus
    dom, rest = bus.split(":", 1)  # nvml reports an 8-hex-digit domain, sysfs uses 4
    return f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}"


def gpu_local_cpus(device_index: int):
    """CPUs local to the GPU's PCIe root (empty set if unknown)."""
    try:
        with open(os.path.join(_gpu_sysfs_dir(device_index), "local_cpulist")) as f:
            return _parse_cpulist(f.read())
    except Exception:
        return set()


def gpu_numa_node(device_index: int) -> int:
    """NUMA node of the GPU's PCIe root as the kernel sees it (sysfs is not filtered by the container's cpuset); -1 if unknown."""
    try:
        with open(os.path.join(_gpu_sysfs_dir(device_index), "numa_node")) as f:
            return int(f.read().strip())
    except Exception:
        return -1


def numa_nodes():
    """{node: set of its CPUs} for the nodes the kernel has online."""
    nodes = {}
    for d in glob.glob("/sys/devices/system/node/node[0-9]*"):
        try:
            with open(os.path.join(d, "cpulist")) as f:
                nodes[int(os.path.basename(d)[4:])] = _parse_cpulist(f.read())
        except (OSError, ValueError):
            pass
    return nodes


def mems_allowed():
    """Nodes this process may allocate memory on (cpuset.mems)."""
    try:
        with open("/proc/self/status") as f:
            for line in f:
                if line.startswith("Mems_allowed_list:"):
                    return _parse_cpulist(line.split(":", 1)[1])
    except OSError:
        pass
    return set()


def prefer_memory_node(node: int) -> int:
    """set_mempolicy(MPOL_PREFERRED, {node}) for the calling thread (inherited by allocations it makes from now on,
    the CUDA driver's pinned pages included).  Returns 0 or the errno."""
    libc = ctypes.CDLL(None, use_errno=True)
    mask = (ctypes.c_ulong * 16)()
    mask[node // 64] = 1 << (node % 64)
    r = libc.syscall(_SYS_SET_MEMPOLICY, _MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(16 * 64 + 1))
    return 0 if r == 0 else ctypes.get_errno()


def bind_to_gpu(device_index: int, world: int = 1) -> dict:
    """Move the process next to GPU `device_index` (see the module text).  Never fails: whatever cannot be determined or
    is not permitted is reported in the returned dict and left alone."""
    allowed = os.sched_getaffinity(0)
    nodes = numa_nodes()
    mems = mems_allowed()
    node = gpu_numa_node(device_index)
    local = gpu_local_cpus(device_index)
    if not local and node in nodes:
        local = set(nodes[node])
    info = {"gpu": device_index, "gpu_numa_node": node, "numa_nodes_online": sorted(nodes), "mems_allowed": sorted(mems),
            "cpus_allowed": len(allowed), "cpu_binding": "unchanged", "memory_policy": "default (first touch)"}
    usable = local & allowed
    if usable and usable != allowed:
        os.sched_setaffinity(0, usable)
        info["cpu_binding"] = f"{len(usable)} GPU-local CPUs"
    elif usable:
        info["cpu_binding"] = f"unchanged: all {len(allowed)} allowed CPUs are GPU-local"
    elif local:
        info["cpu_binding"] = "unchanged: none of the allowed CPUs is GPU-local (cpuset confines the process to another socket)"
    else:
        info["cpu_binding"] = "unchanged: the GPU's local CPUs are unknown (no sysfs entry / no NVML)"
    info["numa_bound_cpus"] = len(os.sched_getaffinity(0))
    if node >= 0 and (not mems or node in mems) and len(nodes) > 1:
        err = prefer_memory_node(node)
        info["memory_policy"] = f"MPOL_PREFERRED node {node}" if err == 0 else f"default: set_mempolicy refused (errno {err})"
    elif node >= 0 and mems and node not in mems:
        info["memory_policy"] = f"default: the GPU's node {node} is outside cpuset.mems {sorted(mems)}"
    return info
