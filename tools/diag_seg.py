"""development: time only the compress kernels of the library named by FSE_B200_LIB on a c4-shaped input (output not checked)"""
import sys, torch
sys.path.insert(0, ".")
import entropy_coders_b200 as E
kind, bs, seg = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
n = int(sys.argv[4]) << 20 if len(sys.argv) > 4 else 1024 << 20
ctx = E.Context(0)
src = ctx.generate(kind, 0xC0FFEE04, n)
p = ctx.params(bs, 0, 128, 0, seg)
nb = ctx.num_streams(n, p)
dst = torch.empty(ctx.bound(n, p), dtype=torch.uint8, device="cuda")
off = torch.empty(nb + 1, dtype=torch.int64, device="cuda"); st = torch.empty(nb, dtype=torch.int32, device="cuda")
for _ in range(3): ctx.compress_blocks_async(src, p, dst, off, st)
ctx.sync(); ctx.set_timing(True)
for _ in range(10): ctx.compress_blocks_async(src, p, dst, off, st)
ctx.sync()
print(sys.argv[1:], {k: round(v[0] / max(v[1], 1), 4) for k, v in ctx.get_timing().items()})
