"""Block-size sweep: compress / decompress GB/s of one B200 on 256 MiB and 1 GiB of text-like bytes for
block sizes 4 KiB .. 4 MiB, one stream per block (one warp per block) next to the segmented per-block mode (one CTA per
block, 16 coder warps).  Every round trip is checked.  Prints one JSON object per line.
    python tools/sweep_block_size.py [kind] > gpurun_out/sweep.jsonl
"""
import json
import sys

import torch

sys.path.insert(0, ".")
import entropy_coders_b200 as E

kind = sys.argv[1] if len(sys.argv) > 1 else "text"
ctx = E.Context(0)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ctx.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for n_mib in (256, 1024):
    n = n_mib << 20
    src = ctx.generate(kind, 0xC0FFEE02, n)
    for bs in (4 << 10, 16 << 10, 32 << 10, 64 << 10, 128 << 10, 256 << 10, 1 << 20, 4 << 20):
        for seg in (0, max(8 << 10, bs // 64) if bs >= 64 << 10 else None):
            if seg is None:
                continue
            p = ctx.params(bs, 0, 128, 0, seg)
            ns = ctx.num_streams(n, p)
            dst = torch.empty(ctx.bound(n, p), dtype=torch.uint8, device="cuda")
            off = torch.empty(ns + 1, dtype=torch.int64, device="cuda")
            st = torch.empty(ns, dtype=torch.int32, device="cuda")
            out = torch.empty(n, dtype=torch.uint8, device="cuda")
            st2 = torch.empty(ns, dtype=torch.int32, device="cuda")
            enc_ms = timed(lambda: ctx.compress_blocks_async(src, p, dst, off, st))
            total = int(off[ns].item())
            dec_ms = timed(lambda: ctx.decompress_blocks_async(dst, total, off, ns, p, out, n, st2))
            ok = bool(torch.equal(out, src)) and not bool(st.any().item()) and not bool(st2.any().item())
            print(json.dumps({"kind": kind, "MiB": n_mib, "block_size": bs, "segment_size": seg, "blocks": ctx.num_blocks(n, bs),
                              "compress_ms": round(enc_ms, 4), "decompress_ms": round(dec_ms, 4),
                              "compress_GBps": round(n / enc_ms / 1e6, 1), "decompress_GBps": round(n / dec_ms / 1e6, 1),
                              "ratio": round(total / n, 4), "roundtrip_ok": ok}), flush=True)
            del dst, off, st, out, st2
    del src
