"""Multi-GPU parity on REAL GPUs (run under torchrun, one rank per GPU; tests/test_multi_gpu.py launches it):
block-range sharding of one logical stream, NCCL for the two tiny exchanges only.

  per-block tables (config 4's shape): every rank codes its contiguous block range; the concatenation of the ranks' outputs,
      placed with the all-gathered totals, equals the oracle's per-block streams of the whole input, byte for byte;
  global table (config 5's shape): the all-reduced 64-bit histogram gives every rank the same table; the one header equals
      the oracle's header of the whole input's histogram and every block equals the oracle's header-less 128-state stream.
Each rank then decodes its own range and the gathered result equals the input."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import entropy_coders_b200 as E  # noqa: E402
from entropy_coders_b200 import sharding as S  # noqa: E402
import oracle_lib as O  # noqa: E402


def gather_bytes(t, world, dev):
    """all-gather of variable-length uint8 tensors -> list of numpy arrays (rank order)"""
    n = torch.tensor([t.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n)
    m = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros(m, dtype=torch.uint8, device=dev)
    pad[: t.numel()] = t
    outs = [torch.zeros(m, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(outs, pad)
    return [o[: int(s.item())].cpu().numpy() for o, s in zip(outs, sizes)]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = E.Context(local)
    total_bytes, bs = 96 * 16384 + 5000, 16384           # 97 blocks: an uneven split and a ragged last block
    kind, seed = "geo", 0xC0FFEE05
    first_block, nblocks, first, nbytes = S.shard_blocks(total_bytes, bs, rank, world)
    src = ctx.generate(kind, seed, nbytes, first_index=first)
    whole = O.generate(kind, seed, total_bytes) if rank == 0 else None

    for mode in (0, 1):
        tl = 11 if mode else 0
        header = None
        if mode == 1:
            counts = ctx.histogram_global(src)
            S.allreduce_histogram(counts)                # NCCL all-reduce of uint64[256]
            header, log2 = ctx.set_global_table(counts, tl)
        d, off, st, total = ctx.compress_blocks(src, bs, tl, 128, table_mode=mode)
        assert (st.cpu().numpy() >= 0).all()
        totals = S.gather_totals(off[nblocks:nblocks + 1], dev)          # NCCL all-gather of one int64 per rank
        base = S.base_offsets(totals)
        assert int(totals[rank].item()) == total
        parts = gather_bytes(d[:total], world, dev)
        offs = gather_bytes(S.global_block_offsets(off[:nblocks], base[rank]).view(torch.uint8), world, dev)
        out, st2 = ctx.decompress_blocks(d, total, off, nbytes, bs, tl, 128, table_mode=mode)
        assert (st2.cpu().numpy() >= 0).all()
        decoded = gather_bytes(out, world, dev)
        if rank == 0:
            logical = np.concatenate(parts)
            goff = np.concatenate([o.view(np.int64) for o in offs] + [np.array([logical.size], dtype=np.int64)])
            nb_all = (total_bytes + bs - 1) // bs
            assert len(goff) == nb_all + 1 and goff[0] == 0 and (np.diff(goff) > 0).all()
            if mode == 1:
                h = O.histogram(whole)
                rc, nh = O.normalize(h, tl)
                assert rc >= 0 and header == O.ncount_write(nh)[0], "global header differs from the oracle's"
                et = O.enc_table(nh)
            for b in range(nb_all):
                blk = whole[b * bs:(b + 1) * bs]
                exp = O.encode_payload(et, blk, 128)[0] if mode == 1 else O.compress_n(blk, tl, 128)[0]
                assert logical[goff[b]:goff[b + 1]].tobytes() == exp, (mode, b)
            assert np.array_equal(np.concatenate(decoded), whole)
            print("multi-GPU parity ok: world %d, mode %s, %d blocks, %d -> %d bytes" % (world, "global" if mode else "per-block", nb_all, total_bytes, logical.size))
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
