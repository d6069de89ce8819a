"""development: wall-clock breakdown of the global-table set-up of bench.py's c5 step (one GPU, no NCCL)"""
import sys, time, torch
sys.path.insert(0, ".")
import entropy_coders_b200 as E
n = 1 << 30
ctx = E.Context(0); ctx2 = E.Context(0)
src = ctx.generate("geo", 0xC0FFEE05, n)
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, r
ms, counts = t(lambda: ctx.histogram_global(src)); print("histogram_global %.3f ms" % ms)
ms, (hdr, _) = t(lambda: ctx.set_global_table(counts, 11)); print("set_global_table %.3f ms" % ms)
ms, _ = t(lambda: ctx2.set_global_table_from_header(hdr)); print("set_global_table_from_header %.3f ms" % ms)
