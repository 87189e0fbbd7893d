import torch, time
n = 4096*4096
d = torch.zeros(n, dtype=torch.float32, device="cuda")
h = torch.empty(n, dtype=torch.float32).pin_memory()
def t_one():
    torch.cuda.synchronize(); t0=time.perf_counter()
    h.copy_(d, non_blocking=True); torch.cuda.synchronize(); return time.perf_counter()-t0
def t_k(k):
    ss=[torch.cuda.Stream() for _ in range(k)]
    torch.cuda.synchronize(); t0=time.perf_counter()
    m=n//k
    for i,s in enumerate(ss):
        with torch.cuda.stream(s):
            h[i*m:(i+1)*m].copy_(d[i*m:(i+1)*m], non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter()-t0
for _ in range(3): t_one(); t_k(2)
print("one stream ms", min(t_one() for _ in range(10))*1e3)
for k in (2,4,8): print(k, "streams ms", min(t_k(k) for _ in range(10))*1e3)
h2 = torch.empty(n, dtype=torch.float32).pin_memory()
torch.cuda.synchronize(); t0=time.perf_counter(); d.copy_(h2, non_blocking=True); torch.cuda.synchronize(); print("h2d ms", (time.perf_counter()-t0)*1e3)
