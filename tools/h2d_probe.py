import torch, time
n=299_030_000
h=torch.empty(n,dtype=torch.uint8).pin_memory()
d=torch.empty(n,dtype=torch.uint8,device='cuda')
for _ in range(3): d.copy_(h,non_blocking=True)
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): d.copy_(h,non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print("1D pinned H2D 299MB: %.3f ms  %.1f GB/s"%(ms, n/ms/1e6))
# 2D: rows 10000 x 29903 -> pitch 29952
h2=h[:10000*29903].view(10000,29903)
d2=torch.zeros(10000,29952,dtype=torch.uint8,device='cuda')
for _ in range(3): d2[:, :29903].copy_(h2,non_blocking=True)
torch.cuda.synchronize()
e0.record()
for _ in range(10): d2[:, :29903].copy_(h2,non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print("2D pinned H2D (torch strided): %.3f ms  %.1f GB/s"%(ms, n/ms/1e6))
# chunks of 48MB
import math
chunk=48<<20
e0.record()
for _ in range(10):
    for o in range(0,n,chunk):
        d[o:o+chunk].copy_(h[o:o+chunk],non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print("1D in 48MB chunks: %.3f ms  %.1f GB/s"%(ms, n/ms/1e6))
# two streams, each copying half of the buffer at the same time (do two copy engines add up on the H2D direction?)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
half = n // 2
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    for _ in range(10):
        with torch.cuda.stream(s1): d[:half].copy_(h[:half], non_blocking=True)
        with torch.cuda.stream(s2): d[half:].copy_(h[half:], non_blocking=True)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / 10
print("1D, two streams x half: %.3f ms  %.1f GB/s (wall clock)" % (ms, n / ms / 1e6))
t0 = time.perf_counter()
for _ in range(10): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3 / 10
print("1D, one stream: %.3f ms  %.1f GB/s (wall clock)" % (ms, n / ms / 1e6))
