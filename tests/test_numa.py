"""webrtc_aecm_b200/numa.py is pure host plumbing (sysfs, sched_setaffinity, set_mempolicy): it must describe what it found
and never fail, with or without a GPU, NVML or NUMA."""
import os

from webrtc_aecm_b200 import numa


def test_cpulist_parser():
    assert numa._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert numa._parse_cpulist("") == set()


def test_bind_to_gpu_reports_and_never_raises():
    before = os.sched_getaffinity(0)
    try:
        info = numa.bind_to_gpu(0, 8)
        for key in ("gpu", "gpu_numa_node", "numa_nodes_online", "mems_allowed", "cpus_allowed", "cpu_binding", "memory_policy",
                    "numa_bound_cpus"):
            assert key in info
        assert info["numa_bound_cpus"] == len(os.sched_getaffinity(0)) >= 1
        assert os.sched_getaffinity(0) <= before  # it may only narrow the mask
    finally:
        os.sched_setaffinity(0, before)


def test_topology_readers():
    nodes = numa.numa_nodes()
    assert all(isinstance(n, int) and cpus for n, cpus in nodes.items())
    mems = numa.mems_allowed()
    assert not nodes or not mems or mems <= set(nodes) | mems
    # preferring a node the process may allocate on is either permitted (0) or refused with an errno; both are fine
    if mems:
        assert isinstance(numa.prefer_memory_node(min(mems)), int)
