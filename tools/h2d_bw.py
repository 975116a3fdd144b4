import torch, time
x=[torch.empty(4,160,160,160, pin_memory=True) for _ in range(4)]
d=[torch.empty(4,160,160,160, device='cuda') for _ in range(4)]
s=torch.cuda.Stream()
for rep in range(3):
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record(s)
        for a,b in zip(d,x): a.copy_(b, non_blocking=True)
        e1.record(s)
    torch.cuda.synchronize()
    ms=e0.elapsed_time(e1); print("H2D 262 MB pinned: %.2f ms = %.1f GB/s"%(ms, 0.262144/ms*1e3))
