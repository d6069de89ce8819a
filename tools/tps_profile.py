"""One compress + one decompress of 1 GiB geometric bytes in 128 KiB blocks at two states (and one), for an ncu capture of
the thread-per-stream kernels: ncu --set full --import-source on --kernel-name regex:k_tps -c 8 python tools/tps_profile.py"""
import sys, torch
sys.path.insert(0, ".")
import entropy_coders_b200 as E
ctx = E.Context(0)
n, bs = 1 << 30, 131072
src = ctx.generate("geo", 0xC0FFEE04, n)
for ns in (2, 1):
    p = ctx.params(bs, 0, ns, 0)
    nb = ctx.num_streams(n, p)
    dst = torch.empty(ctx.bound(n, p), dtype=torch.uint8, device="cuda")
    off = torch.empty(nb + 1, dtype=torch.int64, device="cuda"); st = torch.empty(nb, dtype=torch.int32, device="cuda")
    out = torch.empty(n, dtype=torch.uint8, device="cuda"); st2 = torch.empty(nb, dtype=torch.int32, device="cuda")
    ctx.compress_blocks_async(src, p, dst, off, st); ctx.sync()
    total = int(off[nb].item())
    ctx.decompress_blocks_async(dst, total, off, nb, p, out, n, st2); ctx.sync()
    print(ns, total, bool(torch.equal(out, src)))
