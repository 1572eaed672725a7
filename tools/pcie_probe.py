import torch, time
dev=torch.device('cuda',0)
for mb in (0.32, 1.3, 2.6, 16, 64):
    n=int(mb*1e6)
    h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device=dev)
    for direction in ('h2d','d2h'):
        for _ in range(3): (d.copy_(h,non_blocking=True) if direction=='h2d' else h.copy_(d,non_blocking=True)); torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): (d.copy_(h,non_blocking=True) if direction=='h2d' else h.copy_(d,non_blocking=True))
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/20
        t0=time.perf_counter()
        for _ in range(20):
            (d.copy_(h,non_blocking=True) if direction=='h2d' else h.copy_(d,non_blocking=True)); torch.cuda.synchronize()
        wall=(time.perf_counter()-t0)/20*1e3
        print(f'{direction} {mb:6.2f} MB: {ms*1e3:8.1f} us device ({n/ms/1e6:6.1f} GB/s), sync round trip {wall*1e3:8.1f} us')
