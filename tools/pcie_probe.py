import torch, time
n = 49152000 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(name, f"{ms:.3f} ms  {n*4/ms/1e6:.1f} GB/s")
# fresh pinned allocation cost
t=time.time(); x = torch.empty(n, dtype=torch.float32).pin_memory(); print("pin alloc", time.time()-t)
