"""pinned host <-> device bandwidth with all GPUs of the box copying at the same time
(one process per GPU: for i in 0..7: CUDA_VISIBLE_DEVICES=i python tools/pcie_bw_multi.py i 8 &)"""
import os, sys, time, glob, torch
rank, world = int(sys.argv[1]), int(sys.argv[2])
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def h2d():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
h2d(); d2h(); torch.cuda.synchronize()
open("/tmp/pcie_ready_%d" % rank, "w").close()
while len(glob.glob("/tmp/pcie_ready_*")) < world: time.sleep(0.01)
def t(fn, secs=2.0):
    t0 = time.perf_counter(); k = 0
    while time.perf_counter() - t0 < secs:
        fn(); torch.cuda.synchronize(); k += 1
    return n * k / (time.perf_counter() - t0) / 1e9
a = t(h2d); b = t(d2h); c = t(lambda: (h2d(), d2h()))
print("rank %d of %d: h2d %.1f GB/s  d2h %.1f GB/s  both %.1f GB/s each" % (rank, world, a, b, c), flush=True)
