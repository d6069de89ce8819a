"""compute-sanitizer target: small invocations of the round-2 kernels (shared-table global mode, segmented per-block mode,
bit I/O entry points, zstd normaliser).  compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_r2.py"""
import sys
import numpy as np
import torch
sys.path[:0] = [".", "tests"]
import entropy_coders_b200 as E
import oracle_lib as O

ctx = E.Context(0)
src = O.generate("geo", 3, 20 * 16384 + 77)
d = torch.from_numpy(src).cuda()
# global table: CTA-owned replicated tables
hdr, log2 = ctx.set_global_table(ctx.histogram_global(d), 11)
c, off, st, tot = ctx.compress_blocks(d, 16384, 11, 128, table_mode=1)
out, st2 = ctx.decompress_blocks(c, tot, off, src.size, 16384, 11, 128, table_mode=1)
assert np.array_equal(out.cpu().numpy(), src) and (st2.cpu().numpy() >= 0).all()
hdr, log2 = ctx.set_global_table(ctx.histogram_global(d), 9)
c, off, st, tot = ctx.compress_blocks(d, 16384, 9, 128, table_mode=1)
out, st2 = ctx.decompress_blocks(c, tot, off, src.size, 16384, 9, 128, table_mode=1)
assert np.array_equal(out.cpu().numpy(), src)
# segmented per-block mode: builder warp + coder warps
txt = O.generate("text", 4, 5 * 32768 + 5000)
d = torch.from_numpy(txt).cuda()
c, off, st, tot = ctx.compress_blocks(d, 32768, 0, 128, segment_size=4096)
out, st2 = ctx.decompress_blocks(c, tot, off, txt.size, 32768, 0, 128, segment_size=4096)
assert np.array_equal(out.cpu().numpy(), txt) and (st2.cpu().numpy() >= 0).all()
# private tables (unchanged kernels, shared prologue)
c, off, st, tot = ctx.compress_blocks(d, 32768, 0, 128)
out, st2 = ctx.decompress_blocks(c, tot, off, txt.size, 32768, 0, 128)
assert np.array_equal(out.cpu().numpy(), txt)
c, off, st, tot = ctx.compress_blocks(d, 32768, 0, 128, flags=1)
# bit I/O
rng = np.random.default_rng(1)
bits = rng.integers(1, 17, size=700).astype(np.uint8)
vals = (rng.integers(0, 65536, size=700) & ((1 << bits.astype(np.int64)) - 1)).astype(np.uint32)
o, nb = ctx.bitstack_write(torch.from_numpy(vals).cuda(), torch.from_numpy(bits).cuda(), True)
g, s = ctx.bitstack_read(o, torch.from_numpy(bits).cuda())
assert s == 0 and np.array_equal(g.cpu().numpy().view(np.uint32), vals)
# zstd normaliser
cnt = torch.from_numpy(np.stack([np.bincount(src, minlength=256), np.bincount(txt, minlength=256)]).astype(np.int64)).cuda()
ctx.normalize_zstd(cnt, 11, True)
ctx.close()
print("sanitize target ok")
