#!/usr/bin/env python3
"""Does this box overlap PCIe copies with kernels and with each other?  (torch only; diagnostic)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from webrtc_aecm_b200.numa import bind_to_gpu
print("bound cpus:", bind_to_gpu(0))
p = torch.cuda.get_device_properties(0)
print("asyncEngineCount-ish:", getattr(p, "multi_processor_count", None), p.name)
import ctypes
rt = ctypes.CDLL("libcudart.so.12")
val = ctypes.c_int()
for name, attr in (("asyncEngineCount", 40), ("concurrentKernels", 31), ("gpuOverlap", 15)):
    rt.cudaDeviceGetAttribute(ctypes.byref(val), attr, 0); print(name, val.value)
h1 = torch.empty(40 << 20, dtype=torch.uint8).pin_memory(); h2 = torch.empty(40 << 20, dtype=torch.uint8).pin_memory()
d1 = torch.empty(40 << 20, dtype=torch.uint8, device="cuda"); d2 = torch.empty_like(d1)
x = torch.randn(4096, 4096, device="cuda")
sA, sB, sC = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def h2d():
    with torch.cuda.stream(sA): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(sB): h2.copy_(d2, non_blocking=True)
def kern():
    with torch.cuda.stream(sC):
        y = x
        for _ in range(6): y = y @ x
print(f"h2d 40MB {timed(h2d):.2f} ms, d2h 40MB {timed(d2h):.2f} ms, kernel {timed(kern):.2f} ms")
print(f"h2d + d2h together {timed(lambda: (h2d(), d2h())):.2f} ms")
print(f"h2d + kernel together {timed(lambda: (h2d(), kern())):.2f} ms")
print(f"all three together {timed(lambda: (h2d(), d2h(), kern())):.2f} ms")
