import sys, torch
sys.path.insert(0, ".")
import entropy_coders_b200 as E
ctx = E.Context(0)
def timed(fn, reps=3):
    fn(); ctx.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
SIZES = ((131072, 512), (131072, 1024), (131072, 2048), (65536, 1024), (16384, 512), (4096, 256))
if len(sys.argv) > 1:                                    # tps_sweep.py 131072:8192 65536:1024 ...
    SIZES = tuple(tuple(int(v) for v in a.split(":")) for a in sys.argv[1:])
for bs, mib in SIZES:
    n = mib << 20
    src = ctx.generate("geo", 0xC0FFEE04, n)
    for ns in (2, 1):
        p = ctx.params(bs, 0, ns, 0)
        nb = ctx.num_streams(n, p)
        dst = torch.empty(ctx.bound(n, p), dtype=torch.uint8, device="cuda")
        off = torch.empty(nb + 1, dtype=torch.int64, device="cuda"); st = torch.empty(nb, dtype=torch.int32, device="cuda")
        out = torch.empty(n, dtype=torch.uint8, device="cuda"); st2 = torch.empty(nb, dtype=torch.int32, device="cuda")
        e = timed(lambda: ctx.compress_blocks_async(src, p, dst, off, st))
        total = int(off[nb].item())
        d = timed(lambda: ctx.decompress_blocks_async(dst, total, off, nb, p, out, n, st2))
        ok = bool(torch.equal(out, src))
        print("bs %6d blocks %6d N=%d  compress %7.2f ms %6.1f GB/s   decompress %7.2f ms %6.1f GB/s  ok=%s" % (bs, nb, ns, e, n / e / 1e6, d, n / d / 1e6, ok), flush=True)
