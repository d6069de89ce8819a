"""world_size-2 CPU test (gloo) of the multi-GPU host logic: block-range sharding, the all-gather +
exclusive scan that places each rank's output, and the histogram all-reduce of the global-table mode.
The per-rank coding is done by the oracle here (no GPU); on the B200 box the same functions run over NCCL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, bs, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    from entropy_coders_b200 import sharding as S
    first, count, b0, nbytes = S.shard_blocks(n, bs, rank, world)
    src = O.generate("geo", 0xC0FFEE04, nbytes, first_index=b0)          # each rank generates its own slice
    # per-block tables: code the local range, then place it
    scratch, sizes, status = O.compress_blocks(src, bs, 0, 32, threads=2)
    assert not status.any()
    local = np.concatenate([scratch[i, :int(s)] for i, s in enumerate(sizes)]) if count else np.zeros(0, np.uint8)
    local_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    totals = S.gather_totals(int(local_off[-1]), torch.device("cpu"))
    base = S.base_offsets(totals)
    goff = S.global_block_offsets(torch.from_numpy(local_off), base[rank])
    # global table: all-reduce of the 64-bit histogram
    counts = torch.from_numpy(np.bincount(src, minlength=256).astype(np.int64))
    S.allreduce_histogram(counts)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), local)
    np.save(os.path.join(out_dir, "o%d.npy" % rank), goff.numpy())
    np.save(os.path.join(out_dir, "h%d.npy" % rank), counts.numpy())
    np.save(os.path.join(out_dir, "t%d.npy" % rank), totals.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    import oracle_lib as O
    from entropy_coders_b200 import sharding as S
    world, bs = 2, 4096
    n = 37 * bs + 1234                                                    # odd block count, ragged tail
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n, bs, str(tmp_path)), nprocs=world, join=True)
    # the ranges tile the input
    cover = [S.shard_blocks(n, bs, r, world) for r in range(world)]
    assert cover[0][0] == 0 and cover[0][0] + cover[0][1] == cover[1][0]
    assert cover[0][3] + cover[1][3] == n and cover[1][2] == cover[0][3]
    full = O.generate("geo", 0xC0FFEE04, n)
    # the logical output = rank streams placed at their base offsets = the single-process stream
    scratch, sizes, status = O.compress_blocks(full, bs, 0, 32)
    ref = np.concatenate([scratch[i, :int(s)] for i, s in enumerate(sizes)])
    ref_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    parts = [np.load(tmp_path / ("r%d.npy" % r)) for r in range(world)]
    offs = [np.load(tmp_path / ("o%d.npy" % r)) for r in range(world)]
    totals = np.load(tmp_path / "t0.npy")
    assert np.array_equal(totals, [len(p) for p in parts])
    assert np.array_equal(np.concatenate(parts), ref)
    assert np.array_equal(np.concatenate([offs[0][:-1], offs[1]]), ref_off)
    # the all-reduced histogram is the whole-input histogram on every rank
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("h%d.npy" % r)), np.bincount(full, minlength=256))


def test_shard_blocks_edge_cases():
    from entropy_coders_b200 import sharding as S
    for n, bs, world in [(0, 64, 4), (1, 64, 4), (64 * 3, 64, 4), (64 * 9 + 1, 64, 4), (1 << 20, 65536, 8)]:
        nb = (n + bs - 1) // bs
        got = [S.shard_blocks(n, bs, r, world) for r in range(world)]
        assert sum(g[1] for g in got) == nb and sum(g[3] for g in got) == n
        pos = 0
        for first, count, b0, nbytes in got:
            assert first * bs == b0 or nbytes == 0
            assert b0 == pos or nbytes == 0
            pos += nbytes
