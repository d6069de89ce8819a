"""Small end-to-end run for compute-sanitizer (one tool per gpurun call): every kernel, small sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
import entropy_coders_b200 as E
import oracle_lib as O

ctx = E.Context(0)
for kind, n, bs, ns, tl in [("text", 5 * 8192 + 77, 8192, 64, 0), ("geo", 3 * 4096 + 5, 4096, 32, 0), ("few", 9000, 3000, 2, 9),
                            ("uniform", 70000, 65536, 64, 12), ("text", 3001, 1000, 1, 0)]:
    src = O.generate(kind, 5, n)
    d = torch.from_numpy(src).to(ctx.device)
    comp, off, st, total = ctx.compress_blocks(d, bs, tl, ns)
    out, st2 = ctx.decompress_blocks(comp, total, off, n, bs, tl, ns)
    assert (st.cpu().numpy() >= 0).all() and (st2.cpu().numpy() >= 0).all() and np.array_equal(out.cpu().numpy(), src)
c = ctx.histogram_global(d)
hdr, l2 = ctx.set_global_table(c, 11)
comp, off, st, total = ctx.compress_blocks(d, 1000, 11, 64, table_mode=1)
out, st2 = ctx.decompress_blocks(comp, total, off, d.numel(), 1000, 11, 64, table_mode=1)
assert np.array_equal(out.cpu().numpy(), src)
print("sanitize run ok")
