import torch, time
n=19660800
g=torch.empty(n,dtype=torch.uint8,device='cuda'); h=torch.empty(n,dtype=torch.uint8).pin_memory()
for _ in range(3): h.copy_(g,non_blocking=True); torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(20): h.copy_(g,non_blocking=True); torch.cuda.synchronize()
dt=(time.perf_counter()-t0)/20
print(f"D2H 19.66 MB: {dt*1e6:.0f} us = {n/dt/1e9:.1f} GB/s")
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
half=n//2
t0=time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(s1): h[:half].copy_(g[:half],non_blocking=True)
    with torch.cuda.stream(s2): h[half:].copy_(g[half:],non_blocking=True)
    torch.cuda.synchronize()
dt=(time.perf_counter()-t0)/20
print(f"D2H split over 2 streams: {dt*1e6:.0f} us = {n/dt/1e9:.1f} GB/s")
a=torch.empty(524288//4,dtype=torch.int32).pin_memory(); ga=torch.empty_like(a,device='cuda')
t0=time.perf_counter()
for _ in range(50): ga.copy_(a,non_blocking=True); torch.cuda.synchronize()
print(f"H2D 0.5 MB: {(time.perf_counter()-t0)/50*1e6:.0f} us")
