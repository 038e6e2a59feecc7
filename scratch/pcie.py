import torch, time
n = 2 * 1024**3 // 4
h1 = torch.empty(n, dtype=torch.float32, pin_memory=True); h2 = torch.empty(n, dtype=torch.float32, pin_memory=True)
d1 = torch.empty(n, dtype=torch.float32, device='cuda'); d2 = torch.empty(n, dtype=torch.float32, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
gb = n * 4 / 1e9
a = t(lambda: d1.copy_(h1, non_blocking=True)); b = t(lambda: h2.copy_(d2, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both)
print(f"PCIe pinned 2 GiB: H2D {gb/a:.1f} GB/s, D2H {gb/b:.1f} GB/s, both at once {gb/c:.1f} + {gb/c:.1f} GB/s")
