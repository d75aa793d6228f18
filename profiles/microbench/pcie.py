import torch, time
dev=torch.device('cuda:0')
n=1<<30
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device=dev)
h2=torch.empty(n,dtype=torch.uint8).pin_memory(); d2=torch.empty(n,dtype=torch.uint8,device=dev)
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
for name,fn in (("D2H",lambda: h.copy_(d,non_blocking=True)),("H2D",lambda: d.copy_(h,non_blocking=True))):
    fn(); torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); print(name, 5*n/(time.perf_counter()-t)/1e9,"GB/s")
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): h.copy_(d,non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2,non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t; print("both directions:", 5*n/dt/1e9,"GB/s each")
