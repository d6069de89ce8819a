"""development: time the decode kernel for a workload shape (kind, block size, MiB); with a -DFSE_DEV build (FSE_B200_LIB),
FSE_B200_DECODE128_WIDE=0|1 forces the table form and FSE_B200_WPC caps the warps per CTA"""
import sys, torch
sys.path.insert(0, ".")
import entropy_coders_b200 as E
kind, bs, n = sys.argv[1], int(sys.argv[2]), int(float(sys.argv[3]) * (1 << 20))
ctx = E.Context(0)
src = ctx.generate(kind, 0xC0FFEE02, n)
d, off, st, total = ctx.compress_blocks(src, bs, 0, 128)
p = ctx.params(bs, 0, 128, 0)
nb = ctx.num_blocks(n, bs)
out = torch.empty(n, dtype=torch.uint8, device="cuda"); st2 = torch.empty(nb, dtype=torch.int32, device="cuda")
for _ in range(3): ctx.decompress_blocks_async(d, total, off, nb, p, out, n, st2)
ctx.sync(); ctx.set_timing(True)
for _ in range(10): ctx.decompress_blocks_async(d, total, off, nb, p, out, n, st2)
ctx.sync()
t = ctx.get_timing()["decode"]
assert torch.equal(out, src)
print(kind, bs, n >> 20, "MiB: decode %.4f ms  %.1f GB/s" % (t[0] / t[1], n / (t[0] / t[1]) / 1e6))
