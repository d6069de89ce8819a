"""pinned host <-> device copy bandwidth on this box (the bound of bench.py's e2e leg)"""
import torch, time
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): h2d(); d2h()
print("h2d GB/s", n / t(h2d) / 1e9, "d2h GB/s", n / t(d2h) / 1e9, "duplex each GB/s", n / t(both) / 1e9)
for ch in (4 << 20, 32 << 20):
    def chunks():
        with torch.cuda.stream(s1):
            for o in range(0, n, ch): d[o:o + ch].copy_(h[o:o + ch], non_blocking=True)
    print("h2d chunk", ch >> 20, "MiB:", n / t(chunks) / 1e9)
