#!/usr/bin/env python3
"""Host <-> device copy ceiling of the box, per GPU alone and with all GPUs at once, for every placement of the pinned
host buffer the process is allowed to choose (one torchrun rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/hostlink_probe.py

Explains bench.py's multi-GPU `e2e` numbers: a 10 ms tick of N instances moves N x 640 B host->device and N x 320 B back,
so the end-to-end rate of the whole box cannot exceed (aggregate copy bandwidth) / 960 B per instance-call.  Prints one
JSON line (rank 0)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from webrtc_aecm_b200 import numa  # noqa: E402

MB = 1 << 20


def copy_gbs(host, dev_buf, direction, reps, stream):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for _ in range(reps):
            if direction == "h2d":
                dev_buf.copy_(host, non_blocking=True)
            else:
                host.copy_(dev_buf, non_blocking=True)
    stream.synchronize()
    return host.numel() * reps / (time.perf_counter() - t0) / 1e9


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nodes = sorted(numa.mems_allowed() or numa.numa_nodes() or {0})
    info = {"gpu": local, "gpu_numa_node": numa.gpu_numa_node(local), "local_cpus": len(numa.gpu_local_cpus(local)),
            "cpus_allowed": len(os.sched_getaffinity(0)), "mems_allowed": nodes, "numa_nodes_online": sorted(numa.numa_nodes())}
    size = 64 * MB
    dbuf = torch.empty(size, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    results = {}
    placements = [("first_touch", None)] + [(f"node{n}", n) for n in nodes if len(nodes) > 1]
    for name, node in placements:
        err = numa.prefer_memory_node(node) if node is not None else 0
        if err:
            results[name] = {"error": f"set_mempolicy errno {err}"}
            continue
        host = torch.empty(size, dtype=torch.uint8).pin_memory()
        host.fill_(1)
        hout = torch.empty(size, dtype=torch.uint8).pin_memory()
        r = {}
        for mode in ("alone", "all"):
            for direction in ("h2d", "d2h"):
                if mode == "alone":  # ranks take turns
                    val = 0.0
                    for turn in range(world):
                        if world > 1:
                            dist.barrier()
                        if turn == rank:
                            val = copy_gbs(host if direction == "h2d" else hout, dbuf, direction, 20, s_in)
                    r[f"{direction}_alone"] = val
                else:
                    if world > 1:
                        dist.barrier()
                    r[f"{direction}_all"] = copy_gbs(host if direction == "h2d" else hout, dbuf, direction, 40, s_in)
            # both directions at once (what the pipelined e2e path does), all ranks together
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s_in):
            for _ in range(40):
                dbuf.copy_(host, non_blocking=True)
        d2 = torch.empty(size // 2, dtype=torch.uint8, device=dev)
        with torch.cuda.stream(s_out):
            for _ in range(40):
                hout[:size // 2].copy_(d2, non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()
        secs = time.perf_counter() - t0
        r["duplex_all_h2d"] = size * 40 / secs / 1e9
        r["duplex_all_d2h"] = size // 2 * 40 / secs / 1e9
        results[name] = r
        del host, hout
    info["gbs"] = results
    gathered = [info]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, info)
    if rank == 0:
        agg = {}
        for name, _ in placements:
            rows = [g["gbs"].get(name, {}) for g in gathered]
            if any("error" in r for r in rows):
                agg[name] = rows[0]
                continue
            agg[name] = {k: round(sum(r[k] for r in rows), 1) for k in rows[0]}
        best = max((v.get("duplex_all_h2d", 0) + v.get("duplex_all_d2h", 0), k) for k, v in agg.items() if "error" not in v)
        print(json.dumps({"probe": "hostlink", "n_gpus": world, "buffer_mb": size // MB, "sum_over_gpus_gbs": agg,
                          "e2e_ceiling_blocks_per_s": {k: round((v["duplex_all_h2d"] + v["duplex_all_d2h"]) * 1e9 / 960 * 2.5)
                                                       for k, v in agg.items() if "error" not in v},
                          "best_placement": best[1], "per_gpu": gathered}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
