import torch, time
n = 370*1024*1024
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device='cuda')
h2 = torch.empty(n//4, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(n//4, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, fn in [("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))]:
    for _ in range(2): fn()
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print(name, "%.1f GB/s" % (10*n/dt/1e9))
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print("h2d with concurrent d2h(25%%): %.1f GB/s h2d" % (10*n/dt/1e9))
# two h2d streams concurrently (left/right eye)
ha = torch.empty(n//2, dtype=torch.uint8, pin_memory=True); da = torch.empty(n//2, dtype=torch.uint8, device='cuda')
hb = torch.empty(n//2, dtype=torch.uint8, pin_memory=True); db = torch.empty(n//2, dtype=torch.uint8, device='cuda')
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): da.copy_(ha, non_blocking=True)
    with torch.cuda.stream(s2): db.copy_(hb, non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print("two concurrent h2d streams: %.1f GB/s total" % (10*n/dt/1e9))
