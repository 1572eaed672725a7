#!/usr/bin/env python3
"""bench.py -- AECM blocks/s on B200 (our arm) and on the host cores (reference arm).

Workload = BASELINE.json configs[3]: 262 144 AecmCore instances in total, 16 kHz, STRONG-scaled over the GPUs of the
run (--gpus 1: all 262 144 on one GPU; --gpus 8: 32 768 per GPU).  ``--batch B`` overrides the per-GPU batch (then the
run is weak-scaled, B per GPU).  A *step* is one lockstep 10 ms API call -- WebRtcAecm_BufferFarend + WebRtcAecm_Process
for every instance of the batch -- i.e. one launch of the fused persistent kernel, 2 or 3 64-sample blocks per instance
(2.5 on average).  ``--calls-per-step K`` fuses K consecutive calls into one launch (offline mode; in ``config``).

  value      device-resident inputs, CUDA events per step on the launch stream, L2 flushed between steps
  e2e        AecmB200_SubmitStream/WaitStream with pinned HOST planes: H2D of far+near and D2H of out+status inside the timing
  parity     the timed steps (device-resident AND e2e) of >= 64 sampled instances are replayed on the CPU by the unmodified
             reference (oracle/_ref) from the first sample of the stream and compared bit for bit: ``parity_checked``
  roofline   K-amortised algorithmic bytes / mean kernel time vs the measured HBM peak; next to it the
             *measured* DRAM bytes per launch and the issue-slot fraction (profiles/traffic.json, from ncu captures)
  realtime   the same batch paced: one tick per 10 ms for --realtime-seconds, host planes in, host planes out, counting
             the ticks whose output was not back in host memory before the next tick was due
  cpu_baseline  the unmodified reference on the host cores, bounded sample, rank 0 at N=1 only

``--impl reference`` times the reference's own C implementation on all host cores for the same workload.
``--group`` drives all GPUs from ONE process through the C multi-GPU API (AecmB200_Group*), end to end only.
Under torchrun every rank owns one GPU and an independent shard of instances (no collective on the data path);
time is the max over ranks, value is the whole-job aggregate.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOTAL_BATCH = 262144          # BASELINE.json configs[3]
# per block and instance a one-block launch moves ~3231 B in and ~2709 B out = 5940 B, of which
# 384 B are stream I/O (far, near, out) and 262 B the far-history row written plus the row read; the remaining 5294 B
# are per-instance state, which a launch that keeps the record on-chip for K blocks moves once per K blocks.
ALGO_BYTES_PER_BLOCK_NOMINAL = 5940
ALGO_STREAM_BYTES = 384 + 262
ALGO_STATE_BYTES = ALGO_BYTES_PER_BLOCK_NOMINAL - ALGO_STREAM_BYTES
PREP_BLOCKS = 1100           # > 1024 blocks: startupState == 2, far-end VAD and channel store logic in steady state
SEED = 0xAEC2
METRIC = "AECM blocks/sec (64-sample, 16 kHz) across batch at 1/2/4/8 B200; bit-exact"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--total-batch", type=int, default=TOTAL_BATCH, help="instances over all GPUs (strong scaling)")
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (overrides --total-batch; weak scaling)")
    ap.add_argument("--fs", type=int, default=16000)
    ap.add_argument("--calls-per-step", type=int, default=1)
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-realtime", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--realtime-seconds", type=float, default=10.0)
    ap.add_argument("--parity-sample", type=int, default=64, help="instances replayed on the CPU reference (whole job)")
    ap.add_argument("--clock-sampler", default="nvml", choices=["nvml", "smi"], help="NVML thread (1 ms) or nvidia-smi -lms 20")
    ap.add_argument("--group", action="store_true", help="one process, all GPUs through AecmB200_Group* (e2e only)")
    ap.add_argument("--group-threads", type=int, default=0, help="--group: 1 = one submitting host thread per shard")
    return ap.parse_args()


def workload(args, world):
    """(per-GPU batch, scaling) of this run."""
    if args.batch:
        return args.batch, "weak"
    if args.total_batch % world:
        raise SystemExit(f"--total-batch {args.total_batch} does not divide over {world} GPUs")
    return args.total_batch // world, "strong"


def config_dict(args, world):
    """Identical for both arms (the driver compares them): a function of the command line only."""
    B, scaling = workload(args, world)
    fs, cps = args.fs, args.calls_per_step
    total = B * world
    return {"workload": f"batch={total} AecmCore instances ({'BASELINE configs[3], ' if total == TOTAL_BATCH and fs == 16000 else ''}"
                        f"{B} per GPU x {world} GPU), {fs} Hz, synthetic echo streams, cng=1 echoMode=3, msInSndCardBuf=40, "
                        f"steady state (startupState=2 after {PREP_BLOCKS}+ warm-up blocks)",
            "step": f"{cps} lockstep 10 ms call(s) (BufferFarend+Process) for every instance; one kernel launch per GPU on our arm",
            "total_batch": total, "batch_per_gpu": B, "scaling": scaling, "calls_per_step": cps, "fs": fs,
            "l2": "inputs fresh every step and larger than L2 at this batch; L2 not flushed" if args.no_flush
                  else "inputs fresh every step; L2 flushed (256 MiB memset) between steps"}


def prep_calls(spc):
    """10 ms calls needed to put every instance into steady state (a multiple of 40 for the chunked warm-up)."""
    calls = int(np.ceil(PREP_BLOCKS * 64 / spc))
    return (calls + 39) // 40 * 40


def load_profile(batch, fs, cps):
    """ncu-measured figures for this workload (profiles/traffic.json): DRAM bytes and warp instructions per launch."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)[f"batch{batch}_fs{fs}_cps{cps}"]
    except (OSError, KeyError, ValueError):
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms in the background; `window()` returns the samples
    whose timestamps fall inside the timed region (nvidia-smi needs ~1 s to start, so it is launched early)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def window(self, t0: float, t1: float):
        import datetime

        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.05)
        self.proc.terminate()  # the exact PID we started
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        rows = []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), [n for n, v in zip(self.NAMES, f[4:8]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        inside = [r for r in rows if t0 - 0.02 <= r[0] <= t1 + 0.02]
        note = "samples inside the timed region"
        if not inside and rows:  # region shorter than the sampling period: take the nearest sample
            inside = [min(rows, key=lambda r: abs(r[0] - 0.5 * (t0 + t1)))]
            note = "timed region shorter than the 20 ms sampling period: nearest sample"
        reasons = sorted({n for r in inside for n in r[3]})
        return {"sm_mhz": float(np.median([r[1] for r in inside])) if inside else None,
                "sm_max_mhz": max(r[2] for r in inside) if inside else None, "reasons": reasons,
                "samples": len(inside), "note": note}


class NvmlSampler:
    """SM clock and throttle reasons read through NVML from a thread of this process every millisecond: at 8 GPUs the timed
    region of a 20-step run is a few milliseconds, shorter than anything nvidia-smi can sample.  Same interface as
    ClockSampler, which remains the fall-back when NVML is not usable."""
    REASONS = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.thread = None
        self.stop = False
        self.fallback = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            if vis and all(p.strip().isdigit() for p in vis.split(",")):
                idx = int(vis.split(",")[self.gpu])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            get_reasons(h)
            float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:  # no NVML in this environment: the nvidia-smi sampler
            self.fallback = ClockSampler(self.gpu)
            self.fallback.start()
            return

        def loop():
            while not self.stop:
                try:
                    self.rows.append((time.time(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), mx, int(get_reasons(h))))
                except Exception:
                    pass
                time.sleep(0.001)

        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    def window(self, t0: float, t1: float):
        if self.fallback is not None:
            return self.fallback.window(t0, t1)
        self.stop = True
        self.thread.join(timeout=2)
        rows = list(self.rows)
        inside = [r for r in rows if t0 <= r[0] <= t1]
        note = "NVML samples inside the timed region (1 ms period)"
        if not inside and rows:
            inside = [min(rows, key=lambda r: abs(r[0] - 0.5 * (t0 + t1)))]
            note = "no NVML sample fell inside the timed region: nearest sample"
        bits = 0
        for r in inside:
            bits |= r[3]
        return {"sm_mhz": float(np.median([r[1] for r in inside])) if inside else None,
                "sm_max_mhz": max(r[2] for r in inside) if inside else None,
                "reasons": [n for n, m in self.REASONS if bits & m], "samples": len(inside), "note": note}


class HostPlanes:
    """Pinned host memory from the library's own allocator (cudaMallocHost), viewed as numpy int16 arrays.  First touched
    by the calling thread, which bench.py has bound to the GPU's NUMA node."""

    def __init__(self, lib, shape):
        self.lib = lib
        self.nbytes = int(np.prod(shape)) * 2
        self.ptr = lib.AecmB200_HostAlloc(self.nbytes)
        if not self.ptr:
            raise MemoryError(f"AecmB200_HostAlloc({self.nbytes}) failed")
        self.a = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_int16)), shape=(self.nbytes // 2,)).reshape(shape)
        self.a.reshape(-1)[::2048] = 0  # touch every page

    def free(self):
        if self.ptr:
            self.a = None
            self.lib.AecmB200_HostFree(self.ptr)
            self.ptr = None


def pace_until(t):
    """Sleep, then spin for the last half millisecond."""
    while True:
        d = t - time.perf_counter()
        if d <= 0:
            return
        if d > 0.0007:
            time.sleep(d - 0.0005)


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    # torch's intra-op OpenMP pool spin-waits on every core after each CPU op and starves the thread that
    # drives the GPU (measured: 3x slower host-plane steps); torchrun sets this for its workers, do the same here
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import torch

    torch.set_num_threads(1)
    import torch.distributed as dist

    from webrtc_aecm_b200 import AecmBatch, api
    from webrtc_aecm_b200.synth import synth_streams, synth_streams_torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the AECM engine has no CPU path)")
    from webrtc_aecm_b200.numa import bind_to_gpu

    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu(local, world)  # host planes are first-touched near the GPU; see numa.py for the degenerate topologies
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.cuda.set_stream(torch.cuda.Stream(dev))  # a real (non-blocking) stream, not the legacy default stream

    B, scaling = workload(args, world)
    fs, cps = args.fs, args.calls_per_step
    spc = 160 if fs == 16000 else 80
    PREP_CALLS = prep_calls(spc)
    K, W = args.steps, args.warmup
    first_inst = rank * B  # every rank owns B instances of one global instance numbering
    L = api.lib()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    host_ranks = [numa]
    if world > 1:  # every rank's placement goes into the line: the 8-GPU end-to-end number depends on it
        host_ranks = [None] * world
        dist.all_gather_object(host_ranks, numa)

    sampler = NvmlSampler(local) if args.clock_sampler == "nvml" else ClockSampler(local)
    sampler.start()  # (the nvidia-smi fall-back takes about a second before its first sample: start early)
    batch = AecmBatch(B, local)
    stream = torch.cuda.current_stream(dev)
    batch.adopt_cuda_stream(stream.cuda_stream)  # so torch.cuda.Event brackets exactly our launches
    assert batch.init(fs) == 0
    rc, st = batch.set_config(1, 3)
    assert rc == 0 and not st.any()
    d_status = torch.zeros(B, dtype=torch.int32, device=dev)

    # sampled rows for the parity replay: spread over the shard, always including its first and last instance
    n_par = 0 if args.no_parity else max(2, -(-args.parity_sample // world))
    par_rows = np.unique(np.linspace(0, B - 1, n_par).astype(np.int64)) if n_par else np.zeros(0, dtype=np.int64)
    par_rows_t = torch.as_tensor(par_rows, device=dev)

    # ---- untimed preparation: run the streams into steady state, chunked so HBM use stays small ----
    sample0 = 0
    chunk_calls = 40 if B <= 32768 else 8
    gen_rows = max(1, (1 << 25) // (chunk_calls * spc))  # generate in pieces: the generator's int64 temporaries are 8x the plane size
    far = torch.empty((B, chunk_calls * spc), dtype=torch.int16, device=dev)
    near = torch.empty_like(far)
    out = torch.empty_like(far)
    for c0 in range(0, PREP_CALLS, chunk_calls):
        for r0 in range(0, B, gen_rows):
            r1 = min(B, r0 + gen_rows)
            far[r0:r1], near[r0:r1] = synth_streams_torch(r1 - r0, chunk_calls * spc, fs, SEED, 40, "echo", first_inst + r0, sample0, dev)
        rc = batch.process_stream_device(far, near, None, out, row_stride=chunk_calls * spc, samples_per_call=spc,
                                         n_calls=chunk_calls, d_status=d_status)
        assert rc == 0
        torch.cuda.synchronize(dev)  # the planes are refilled next
        sample0 += chunk_calls * spc
    assert int(d_status.abs().max().item()) == 0
    st0 = batch.read_state(0)
    assert st0.s.startup_state == 2 and st0.s.ec_startup == 0
    del far, near, out

    # ---- timed streams: (W + K) steps of fresh audio, resident in HBM ----
    n_calls_total = (W + K) * cps
    row_stride = n_calls_total * spc
    far = torch.empty((B, row_stride), dtype=torch.int16, device=dev)
    near = torch.empty_like(far)
    out = torch.empty_like(far)
    gen_rows = max(1, (1 << 25) // row_stride)
    for r0 in range(0, B, gen_rows):
        r1 = min(B, r0 + gen_rows)
        far[r0:r1], near[r0:r1] = synth_streams_torch(r1 - r0, row_stride, fs, SEED, 40, "echo", first_inst + r0, sample0, dev)
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(k):
        return batch.process_stream_device(far, near, None, out, row_stride=row_stride, samples_per_call=spc,
                                           n_calls=cps, d_status=d_status, sample_offset=k * cps * spc)

    for k in range(W):
        if flush is not None:
            flush.zero_()
        assert step(k) == 0
    torch.cuda.synchronize(dev)
    blocks_before = batch.read_state(0).s.tot_count
    launches_before = L.AecmB200_KernelLaunchCount()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    wall0 = time.perf_counter()
    epoch0 = time.time()
    for k in range(K):
        if flush is not None:
            flush.zero_()  # > L2 (126 MB): every step starts with a cold L2
        ev0[k].record(stream)
        rc = step(W + k)
        ev1[k].record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    wall1 = time.perf_counter()
    clocks = sampler.window(epoch0, time.time())
    assert rc == 0 and int(d_status.abs().max().item()) == 0
    launches = L.AecmB200_KernelLaunchCount() - launches_before
    step_ms = np.array([a.elapsed_time(b) for a, b in zip(ev0, ev1)])
    dev_ms = float(step_ms.sum())
    blocks_per_inst = batch.read_state(0).s.tot_count - blocks_before
    max_ms = allmax(dev_ms)
    total_blocks = blocks_per_inst * B * world
    value = total_blocks / (max_ms * 1e-3)
    # what the parity replay needs from the device-resident section: the sampled rows of its output (and input, to pin the
    # device-side generator against the numpy one)
    dev_out_rows = out[par_rows_t].cpu().numpy() if n_par else None
    dev_far_rows = far[par_rows_t].cpu().numpy() if n_par else None
    dev_near_rows = near[par_rows_t].cpu().numpy() if n_par else None
    sample_e2e = sample0 + row_stride
    del far, near, out, flush
    torch.cuda.empty_cache()

    # ---- e2e: host planes through the public C ABI, pipelined over the library's stream slots ----
    e2e = realtime = None
    e2e_out_rows = None
    Ke, Kw = min(K, 32), 4
    if not args.no_e2e:
        width = cps * spc
        # continue the same streams: fresh audio for every e2e step, laid out [step][far|near][instance][samples] in PINNED
        # host memory; per step the far plane is immediately followed by the near plane: the library moves both in one DMA
        hin = HostPlanes(L, (Ke + Kw, 2, B, width))
        hout = HostPlanes(L, (Ke + Kw, B, width))
        tmp_rows = max(1, (1 << 25) // ((Ke + Kw) * width))
        for r0 in range(0, B, tmp_rows):
            r1 = min(B, r0 + tmp_rows)
            f, n_ = synth_streams_torch(r1 - r0, (Ke + Kw) * width, fs, SEED, 40, "echo", first_inst + r0, sample_e2e, dev)
            hin.a[:, 0, r0:r1] = f.view(r1 - r0, Ke + Kw, width).permute(1, 0, 2).cpu().numpy()
            hin.a[:, 1, r0:r1] = n_.view(r1 - r0, Ke + Kw, width).permute(1, 0, 2).cpu().numpy()
        del f, n_
        plane = B * width * 2  # bytes per step and plane
        S = int(L.AecmB200_NumStreamSlots())  # calls in flight: H2D of one, kernel of another, D2H of a third overlap
        hstatus = [np.zeros(B, dtype=np.int32) for _ in range(S)]

        def submit(k, slot=None):
            slot = k % S if slot is None else slot
            a = api.StreamArgs(hin.ptr + (2 * k) * plane, hin.ptr + (2 * k + 1) * plane, None, hout.ptr + k * plane,
                               width, 0, B, spc, cps, None, 40, None, hstatus[slot].ctypes.data)
            assert L.AecmB200_SubmitStream(batch._h, C.byref(a), slot) == 0

        def wait(k, slot=None):
            slot = k % S if slot is None else slot
            assert L.AecmB200_WaitStream(batch._h, slot) == 0 and not hstatus[slot].any()

        for k in range(Kw):  # warm-up, unpipelined
            submit(k)
            wait(k)
        b0 = batch.read_state(0).s.tot_count
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for k in range(Kw, Kw + Ke):
            submit(k)                    # slot k % S was released by the wait of step k - S
            if k - (S - 1) >= Kw:
                wait(k - (S - 1))        # results of the oldest step in flight are in host memory
        for k in range(max(Kw, Kw + Ke - (S - 1)), Kw + Ke):
            wait(k)
        secs = allmax(time.perf_counter() - t0)
        eb = (batch.read_state(0).s.tot_count - b0) * B * world
        h2d, d2h = 2 * B * width * 2, B * width * 2 + B * 4
        e2e = {"value": eb / secs, "unit": "blocks/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
               "steps": Ke, "ms_per_step": secs * 1e3 / Ke,
               "host_link_gbs": {"h2d": h2d * world * Ke / secs / 1e9, "d2h": d2h * world * Ke / secs / 1e9,
                                 "note": "aggregate over all ranks, achieved during the e2e steps (both directions run concurrently)"},
               "note": f"AecmB200_SubmitStream/WaitStream on pinned host planes, {S} slots round-robin: per step and GPU one H2D of "
                       "far+near, the fused kernel, D2H of out and status, all inside the timed region; bytes are whole-job"}
        if n_par:
            e2e_out_rows = hout.a[:, par_rows].transpose(1, 0, 2).reshape(len(par_rows), -1).copy()

        # ---- realtime: the same batch, one tick per 10 ms (x calls per step), host planes in, host planes out ----
        if not args.no_realtime:
            period = 0.010 * cps
            ticks = max(1, int(round(args.realtime_seconds / period)))
            lat = np.zeros(ticks)
            ring = Ke + Kw  # the audio of the e2e steps, replayed cyclically (content does not change the work per tick)
            if world > 1:
                dist.barrier()
            start = time.perf_counter() + 0.02
            late_release = 0
            for k in range(ticks):
                due = start + k * period
                if time.perf_counter() > due + 0.0005:
                    late_release += 1  # the previous tick overran its period
                pace_until(due)
                kk = k % ring
                for j in range(S):  # the tick's channels in S sub-batches, one per slot: the H2D of sub-batch j+1 runs behind
                    r0, r1 = B * j // S, B * (j + 1) // S  # the kernel of sub-batch j, its D2H behind the next kernel
                    a = api.StreamArgs(hin.ptr + (2 * kk) * plane + r0 * width * 2, hin.ptr + (2 * kk + 1) * plane + r0 * width * 2,
                                       None, hout.ptr + kk * plane + r0 * width * 2, width, r0, r1 - r0, spc, cps, None, 40, None,
                                       hstatus[j].ctypes.data)
                    assert L.AecmB200_SubmitStream(batch._h, C.byref(a), j) == 0
                for j in range(S):
                    assert L.AecmB200_WaitStream(batch._h, j) == 0 and not hstatus[j][:B * (j + 1) // S - B * j // S].any()
                lat[k] = time.perf_counter() - due
            misses = int((lat > period).sum())
            realtime = {"channels": B * world, "tick_ms": period * 1e3, "ticks": ticks, "seconds": ticks * period,
                        "deadline_misses": int(allsum(misses)), "late_releases": int(allsum(late_release)),
                        "latency_ms": {"p50": float(np.percentile(lat, 50) * 1e3), "p99": float(np.percentile(lat, 99) * 1e3),
                                       "max_over_ranks": allmax(float(lat.max()) * 1e3)},
                        "realtime_factor": period / float(np.percentile(lat, 50)),
                        "sub_batches_per_tick": S,
                        "note": "per tick: audio of all channels released at t0 + k*tick in pinned host memory -> AecmB200_SubmitStream "
                                f"(H2D far+near, kernel, D2H out+status) for {S} sub-batches of the GPU's channels, one per stream slot, so "
                                "copies and kernels of neighbouring sub-batches overlap -> AecmB200_WaitStream on every slot; a tick misses "
                                "its deadline when its output is not back in host memory before the next tick is due; latency = release "
                                "to last output in host memory; percentiles are rank 0's"}
        hin.free()
        hout.free()

    # ---- parity: replay the sampled instances on the CPU reference from the first sample of the stream ----
    parity = None
    if n_par:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as ol

        impl = "ref" if ol.ref_path() else "oracle"
        n_e2e = (Ke + Kw) * cps * spc if e2e is not None else 0
        total_samples = sample_e2e + n_e2e
        gids = first_inst + par_rows
        cfar, cnear = synth_streams(len(gids), total_samples, fs, SEED, 40, "echo", instances=gids)
        gen_ok = bool(np.array_equal(cfar[:, sample0:sample_e2e], dev_far_rows) and np.array_equal(cnear[:, sample0:sample_e2e], dev_near_rows))
        want, ret, _ = ol.run_streams(cfar, cnear, fs=fs, samples_per_call=spc, impl=impl, threads=min(len(gids), len(all_cpus)))
        bad_dev = int((want[:, sample0:sample_e2e] != dev_out_rows).any(axis=1).sum())
        bad_e2e = int((want[:, sample_e2e:] != e2e_out_rows).any(axis=1).sum()) if e2e_out_rows is not None else 0
        ok = gen_ok and not ret.any() and bad_dev == 0 and bad_e2e == 0
        n_ok = int(allsum(len(gids) if ok else 0))
        n_all = int(allsum(len(gids)))
        parity = {"checked": n_all, "bit_exact": n_ok, "against": "oracle/_ref (the unmodified reference, compiled)" if impl == "ref"
                  else "oracle/ (C restatement; oracle/_ref not present)",
                  "covers": f"every sample of the {W}+{K} device-resident steps" + (f" and of the {Kw}+{Ke} e2e steps" if e2e_out_rows is not None else "")
                            + f", after {PREP_CALLS} warm-up calls replayed from sample 0",
                  "generator_twin_equal": gen_ok}
        if n_ok != n_all:
            raise SystemExit(f"bench.py: PARITY FAILURE on rank {rank}: generator_equal={gen_ok} ret={ret.any()} "
                             f"rows differing: device-resident {bad_dev}, e2e {bad_e2e} of {len(gids)}")

    # ---- CPU baseline beside it (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)  # the CPU reference gets every host core, not just the GPU's socket
        cpu = cpu_reference_throughput(fs, spc, n_inst=1024, prep_calls=PREP_CALLS, timed_calls=160)

    if rank == 0:
        peak, peak_src = load_peaks()
        blocks_per_launch = blocks_per_inst * B / K
        kblk = blocks_per_inst / K  # blocks per instance per launch: the record stays on-chip for this many blocks
        kernel_s = dev_ms * 1e-3 / K  # one launch per step
        algo_per_block = ALGO_STREAM_BYTES + ALGO_STATE_BYTES / kblk
        achieved = blocks_per_launch * algo_per_block / kernel_s / 1e9
        prof = load_profile(B, fs, cps)
        traffic = prof["bytes_per_launch"] if prof else None
        sm_clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
        issue = None
        if prof and prof.get("warp_inst_per_block"):
            issue = {"warp_inst_per_block": prof["warp_inst_per_block"],
                     "frac": prof["warp_inst_per_block"] * (blocks_per_launch / kernel_s) / (148 * 4 * sm_clock),
                     "peak": "148 SMs x 4 schedulers x SM clock warp-instructions/s",
                     "note": "the binding resource: the kernel is integer-issue bound (DESIGN.md 4.1)"}
        line = {
            "metric": METRIC,
            "value": value, "unit": "blocks/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": max_ms / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "int16/int32 Q-format", "data": "synthetic",
            "config": config_dict(args, world),
            "blocks_per_instance_per_step": blocks_per_inst / K, "realtime_channels": value / (250.0 if fs == 16000 else 125.0),
            "parity_checked": parity["bit_exact"] if parity else 0, "parity": parity,
            "clocks": clocks, "gpu_launches": int(launches), "host": numa, "host_ranks": host_ranks if world > 1 else None,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                         "traffic_gbs": traffic / kernel_s / 1e9 if traffic else None,
                         "traffic_frac": traffic / kernel_s / 1e9 / peak if traffic else None,
                         "traffic_source": prof.get("source") if prof else None,
                         "algorithmic_bytes_per_launch": blocks_per_launch * algo_per_block,
                         "algorithmic_bytes_per_block": algo_per_block, "blocks_per_instance_per_launch": kblk,
                         "algorithmic_rule": f"SURVEY 8(d): {ALGO_STREAM_BYTES} B stream+history per block + {ALGO_STATE_BYTES} B state / K, "
                                             f"K = blocks per instance per launch (record held on-chip)",
                         "nominal_frac_one_block_per_launch": blocks_per_launch * ALGO_BYTES_PER_BLOCK_NOMINAL / kernel_s / 1e9 / peak,
                         "peak_source": peak_src, "blocks_per_launch": blocks_per_launch, "kernel_ms": kernel_s * 1e3,
                         "issue": issue},
            "wall_ms_per_step": (wall1 - wall0) * 1e3 / K,
        }
        if e2e:
            line["e2e"] = e2e
        if realtime:
            line["realtime"] = realtime
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
def run_group(args):
    """One process, all GPUs of the run through the library's own multi-GPU entry points (AecmB200_Group*): end to end
    with host planes only (the group API has no device-plane call).  --group-threads 1 submits every shard from its
    own host thread through the shard's batch handle instead of from one thread."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import torch

    torch.set_num_threads(1)
    from webrtc_aecm_b200 import api
    from webrtc_aecm_b200.synth import synth_streams_torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the AECM engine has no CPU path)")
    n_dev = args.gpus
    B, scaling = workload(args, n_dev)
    total = B * n_dev
    fs, cps = args.fs, args.calls_per_step
    spc = 160 if fs == 16000 else 80
    width = cps * spc
    PREP_CALLS = prep_calls(spc)
    L = api.lib()
    g = L.AecmB200_GroupCreate(total, n_dev, None)
    if not g:
        raise SystemExit("AecmB200_GroupCreate failed")
    assert L.AecmB200_GroupInit(g, fs) == 0
    assert L.AecmB200_GroupSetConfig(g, api.AecmConfig(1, 3), None) == 0
    Ke, Kw = min(args.steps, 32), 4
    ring = 8  # distinct host planes for the warm-up into steady state (refilled chunk by chunk), then Ke + Kw timed ones
    hfar = HostPlanes(L, (max(ring, Ke + Kw), total, width))
    hnear = HostPlanes(L, (max(ring, Ke + Kw), total, width))
    hout = HostPlanes(L, (max(ring, Ke + Kw), total, width))
    dev = torch.device("cuda", 0)
    S = int(L.AecmB200_NumStreamSlots())
    hstatus = [np.zeros(total, dtype=np.int32) for _ in range(S)]
    plane = total * width * 2

    def fill(n_steps, sample0):
        rows = max(1, (1 << 25) // (n_steps * width))
        for r0 in range(0, total, rows):
            r1 = min(total, r0 + rows)
            f, n_ = synth_streams_torch(r1 - r0, n_steps * width, fs, SEED, 40, "echo", r0, sample0, dev)
            hfar.a[:n_steps, r0:r1] = f.view(r1 - r0, n_steps, width).permute(1, 0, 2).cpu().numpy()
            hnear.a[:n_steps, r0:r1] = n_.view(r1 - r0, n_steps, width).permute(1, 0, 2).cpu().numpy()

    shards = []
    for i in range(n_dev):
        f0, cnt = C.c_int32(0), C.c_int32(0)
        h = L.AecmB200_GroupShard(g, i, C.byref(f0), C.byref(cnt))
        shards.append((h, f0.value, cnt.value))

    def args_for(k, slot, first=0, count=total, local_first=0):
        off = k * plane + first * width * 2
        return api.StreamArgs(hfar.ptr + off, hnear.ptr + off, None, hout.ptr + off, width, local_first, count, spc, cps, None, 40,
                              None, hstatus[slot].ctypes.data + first * 4)

    def submit(k):
        a = args_for(k, k % S)
        assert L.AecmB200_GroupSubmitStream(g, C.byref(a), k % S) == 0

    def wait(k):
        assert L.AecmB200_GroupWaitStream(g, k % S) == 0 and not hstatus[k % S].any()

    sample0 = 0
    for c0 in range(0, PREP_CALLS // cps, ring):  # warm-up into steady state
        n = min(ring, PREP_CALLS // cps - c0)
        fill(n, sample0)
        for k in range(n):
            submit(k)
            wait(k)
        sample0 += n * width
    fill(Ke + Kw, sample0)
    for k in range(Kw):
        submit(k)
        wait(k)

    if args.group_threads:
        # one host thread per shard: each pipelines its own shard's submissions through the shard's batch handle
        barrier = threading.Barrier(n_dev + 1)
        errs = []

        def worker(i):
            h, f0, cnt = shards[i]
            torch.cuda.set_device(i)
            barrier.wait()
            try:
                for k in range(Kw, Kw + Ke):
                    a = args_for(k, k % S, f0, cnt, 0)
                    assert L.AecmB200_SubmitStream(h, C.byref(a), k % S) == 0
                    if k - (S - 1) >= Kw:
                        assert L.AecmB200_WaitStream(h, (k - (S - 1)) % S) == 0
                for k in range(max(Kw, Kw + Ke - (S - 1)), Kw + Ke):
                    assert L.AecmB200_WaitStream(h, k % S) == 0
            except AssertionError as e:  # pragma: no cover
                errs.append((i, e))
            barrier.wait()

        th = [threading.Thread(target=worker, args=(i,)) for i in range(n_dev)]
        for t in th:
            t.start()
        barrier.wait()
        t0 = time.perf_counter()
        barrier.wait()
        secs = time.perf_counter() - t0
        for t in th:
            t.join()
        assert not errs, errs
        assert not any(s.any() for s in hstatus)
    else:
        t0 = time.perf_counter()
        for k in range(Kw, Kw + Ke):
            submit(k)
            if k - (S - 1) >= Kw:
                wait(k - (S - 1))
        for k in range(max(Kw, Kw + Ke - (S - 1)), Kw + Ke):
            wait(k)
        secs = time.perf_counter() - t0
    blocks = total * Ke * width / 64.0
    assert int(np.abs(hout.a[Kw + Ke - 1]).max()) > 0
    print(json.dumps({"mode": "group", "metric": METRIC, "value": blocks / secs, "unit": "blocks/s", "n_gpus": n_dev,
                      "steps": Ke, "ms_per_step": secs * 1e3 / Ke, "scaling": scaling, "config": config_dict(args, n_dev),
                      "submit_threads": n_dev if args.group_threads else 1,
                      "e2e": {"value": blocks / secs, "unit": "blocks/s", "h2d_bytes_per_step": 2 * plane, "d2h_bytes_per_step": plane + total * 4,
                              "host_link_gbs": {"h2d": 2 * plane * Ke / secs / 1e9, "d2h": plane * Ke / secs / 1e9}},
                      "note": "single process: AecmB200_GroupSubmitStream/GroupWaitStream over all devices" if not args.group_threads
                      else "single process, one host thread per shard: AecmB200_SubmitStream/WaitStream on AecmB200_GroupShard handles"}))
    L.AecmB200_GroupFree(g)


def usable_cpus():
    """Host threads this process can really run at once: the affinity mask, capped by the cgroup CPU quota (a container
    that sees 128 CPUs but is given 16 CPUs' worth of time gets slower, not faster, with 128 busy threads)."""
    n = len(os.sched_getaffinity(0))
    quota = None
    try:
        q, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]  # cgroup v2
        if q != "max":
            quota = float(q) / float(period)
    except (OSError, ValueError):
        try:
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())  # cgroup v1
            if q > 0:
                quota = q / float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        except (OSError, ValueError):
            pass
    if quota:
        n = max(1, min(n, int(quota + 0.999)))
    return n


def cpu_pool_throughput(fs, spc, n_inst, prep, warm, timed, threads):
    """Blocks/s of the reference's C path (oracle/_ref; the oracle port if it is absent) on `threads` host threads:
    n_inst persistent instances, instance i fed stream i % 64 of the bench's synthetic family, `prep` + `warm` untimed
    calls into steady state, then `timed` timed calls."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    from webrtc_aecm_b200.synth import synth_streams

    kind = "reference" if ol.ref_path() else "port"
    pool = ol.CpuPool(n_inst, fs=fs, impl="ref" if kind == "reference" else "oracle", cng_mode=1, echo_mode=3, threads=threads)
    base = min(n_inst, 64)  # numpy generation is slow: 64 distinct streams, tiled over the pool inside the C runner
    far, near = synth_streams(base, (prep + warm + timed) * spc, fs, SEED, 40, "echo")
    c0 = (prep + warm) * spc
    pool.run_tiled(np.ascontiguousarray(far[:, :c0]), np.ascontiguousarray(near[:, :c0]), samples_per_call=spc)
    acc, secs = pool.run_tiled(np.ascontiguousarray(far[:, c0:]), np.ascontiguousarray(near[:, c0:]), samples_per_call=spc)
    pool.close()
    blocks = n_inst * timed * spc / 64.0
    return kind, blocks, secs, acc


def cpu_reference_throughput(fs, spc, n_inst, prep_calls, timed_calls, threads=None):
    """The `cpu_baseline` block of our arm's line (bounded sample)."""
    threads = threads or usable_cpus()
    kind, blocks, secs, acc = cpu_pool_throughput(fs, spc, n_inst, prep_calls, 0, timed_calls, threads)
    return {"value": blocks / secs, "unit": "blocks/s", "cores": threads, "kind": kind,
            "sample": f"{n_inst} instances x {timed_calls} steady-state 10 ms calls ({blocks:.0f} blocks) after "
                      f"{prep_calls} warm-up calls, {threads} threads (= usable CPUs: affinity {len(os.sched_getaffinity(0))}, "
                      f"capped by the cgroup quota), instances statically sharded",
            "seconds": secs, "return_codes_or": int(acc)}


def run_reference(args):
    """The reference arm: the reference's own C implementation on the host cores, same workload/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    B, scaling = workload(args, world)
    fs, cps = args.fs, args.calls_per_step
    spc = 160 if fs == 16000 else 80
    K, W = args.steps, args.warmup
    PREP_CALLS = prep_calls(spc)
    threads = usable_cpus()
    # bounded sample of the workload: a slice of the batch sized for roughly 20 s of CPU work in total (warm-up into steady
    # state included) at ~1e5 blocks/s per thread, at least 8 instances per thread, at most the batch itself
    calls = PREP_CALLS + (W + K) * cps
    n_inst = int(min(B * world, max(threads * 8, 20.0 * 1.0e5 * threads / (calls * spc / 64.0))))
    kind, blocks, secs, acc = cpu_pool_throughput(fs, spc, n_inst, PREP_CALLS, W * cps, K * cps, threads)
    value = blocks / secs
    cpu = {"value": value, "unit": "blocks/s", "cores": threads, "kind": kind,
           "sample": f"{n_inst} of {B * world} instances x {K * cps} steady-state calls ({blocks:.0f} blocks) after {PREP_CALLS} + {W * cps} "
                     f"untimed calls, {threads} threads (= usable CPUs: affinity {len(os.sched_getaffinity(0))}, capped by the cgroup quota)"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC,
        "value": value, "unit": "blocks/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": secs * 1e3 / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "int16/int32 Q-format", "data": "synthetic",
        "config": config_dict(args, world),
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "blocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "return_codes_or": int(acc)}))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif a.group:
        run_group(a)
    else:
        run_ours(a)
