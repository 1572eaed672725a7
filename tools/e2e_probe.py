#!/usr/bin/env python3
"""Break the host-plane (e2e) step time down: CPU time inside Submit / Wait, synchronous vs double buffered,
own stream vs adopted torch stream.  Diagnostic, run on the GPU box."""
import ctypes as C
import os
import sys
import time

import numpy as np
os.environ.setdefault('OMP_NUM_THREADS', '1')
import torch
torch.set_num_threads(1)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from webrtc_aecm_b200 import AecmBatch, api  # noqa: E402

B, spc, K = 4096, 160, 80
L = api.lib()
if os.environ.get("BIND") == "1":
    from webrtc_aecm_b200.numa import bind_to_gpu
    print("bound to", bind_to_gpu(0), "gpu-local cpus")
print("affinity size", len(os.sched_getaffinity(0)))
# copy bandwidth for 2.6 MB slices of a LARGE pinned buffer (not cache resident)
big = torch.empty(200 << 20, dtype=torch.uint8).pin_memory(); big.zero_()
dbuf = torch.empty(2621440, dtype=torch.uint8, device="cuda")
for direction in ("h2d", "d2h"):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(40):
        sl = big[i * 4000000:i * 4000000 + 2621440]
        (dbuf.copy_(sl, non_blocking=True) if direction == "h2d" else sl.copy_(dbuf, non_blocking=True))
    e1.record(); torch.cuda.synchronize()
    print(f"{direction} 2.6 MB slices of a 200 MB pinned buffer: {e0.elapsed_time(e1) / 40 * 1e3:.1f} us each")
for adopt in (False, True):
    b = AecmBatch(B, 0)
    if adopt:
        b.adopt_cuda_stream(torch.cuda.current_stream().cuda_stream)
    b.init(16000)
    rng = np.random.default_rng(0)
    hfar = torch.from_numpy(rng.integers(-3000, 3000, size=(K, B, spc)).astype(np.int16)).pin_memory()
    hnear = torch.from_numpy(rng.integers(-3000, 3000, size=(K, B, spc)).astype(np.int16)).pin_memory()
    hout = torch.empty_like(hnear).pin_memory()
    st = [np.zeros(B, dtype=np.int32) for _ in range(2)]
    plane = B * spc * 2

    def args(k):
        return api.StreamArgs(hfar.data_ptr() + k * plane, hnear.data_ptr() + k * plane, None, hout.data_ptr() + k * plane,
                              spc, 0, B, spc, 1, None, 40, None, st[k % 2].ctypes.data)

    for k in range(10):  # through start-up
        a = args(k)
        L.AecmB200_ProcessStream(b._h, C.byref(a))
    t0 = time.perf_counter()
    for k in range(10, 40):
        a = args(k)
        L.AecmB200_ProcessStream(b._h, C.byref(a))
    t_sync = (time.perf_counter() - t0) / 30
    ts = tw = 0.0
    t0 = time.perf_counter()
    a = args(40)
    L.AecmB200_SubmitStream(b._h, C.byref(a), 0)
    for k in range(41, K):
        a = args(k)
        t1 = time.perf_counter()
        L.AecmB200_SubmitStream(b._h, C.byref(a), k % 2)
        t2 = time.perf_counter()
        L.AecmB200_WaitStream(b._h, (k - 1) % 2)
        t3 = time.perf_counter()
        ts += t2 - t1
        tw += t3 - t2
    L.AecmB200_WaitStream(b._h, (K - 1) % 2)
    t_db = (time.perf_counter() - t0) / (K - 40)
    print(f"adopt_torch_stream={adopt}: synchronous {t_sync * 1e6:.0f} us/step; double-buffered {t_db * 1e6:.0f} us/step "
          f"(cpu in Submit {ts / (K - 41) * 1e6:.0f} us, in Wait {tw / (K - 41) * 1e6:.0f} us)")
    b.close()
