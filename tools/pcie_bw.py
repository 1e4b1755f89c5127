import torch, time
n=256<<20
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
h2=torch.empty(n,dtype=torch.uint8).pin_memory(); d2=torch.empty(n,dtype=torch.uint8,device='cuda')
for name,fn in (("h2d",lambda: d.copy_(h,non_blocking=True)),("d2h",lambda: h.copy_(d,non_blocking=True))):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print(name, 10*n/dt/1e9,"GB/s")
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print("both directions at once, each", 10*n/dt/1e9,"GB/s")
