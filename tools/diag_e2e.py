"""development: host-buffer round trip (fse_b200_compress_host + decompress_host) timing; FSE_B200_PIPE_CHUNK_MB sets the chunk"""
import sys, time, torch, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import entropy_coders_b200 as E
kind, bs, n = sys.argv[1], int(sys.argv[2]), int(float(sys.argv[3]) * (1 << 20))
ctx = E.Context(0)
hsrc = ctx.generate(kind, 0xC0FFEE02, n).cpu().pin_memory()
p = ctx.params(bs, 0, 128, 0)
hdst = torch.empty(ctx.bound(n, p), dtype=torch.uint8).pin_memory(); hout = torch.empty(n, dtype=torch.uint8).pin_memory()
def step():
    _, offs, st, tot = ctx.compress_host(hsrc, bs, 0, 128, 0, dst=hdst)
    t1 = time.perf_counter()
    ctx.decompress_host(hdst, tot, offs, n, bs, 0, 128, 0, dst=hout)
    return t1
for _ in range(2): step()
tc = td = 0.0
for _ in range(5):
    t0 = time.perf_counter(); t1 = step(); t2 = time.perf_counter(); tc += t1 - t0; td += t2 - t1
assert torch.equal(hout, hsrc)
print("compress %.2f ms  decompress %.2f ms  round trip %.1f GB/s" % (tc / 5 * 1e3, td / 5 * 1e3, n / ((tc + td) / 5) / 1e9))
