"""Multi-GPU partitioning of the block path: contiguous block ranges per rank, no
collective on the data path.  The only exchanges are (i) an all-gather of each rank's compressed
total, turned into the rank's base offset in the logical output by an exclusive scan, and (ii) in
global-table mode an all-reduce (sum) of the 256-bin 64-bit histogram.  Backend agnostic: NCCL on
GPUs, gloo in the CPU tests."""
import torch
import torch.distributed as dist


def shard_blocks(n_bytes, block_size, rank, world):
    """Contiguous block range of `rank`: (first_block, n_blocks, first_byte, n_bytes).
    Blocks are dealt out as evenly as possible; earlier ranks take the remainder."""
    nb = (n_bytes + block_size - 1) // block_size
    base, rem = divmod(nb, world)
    first = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    b0 = min(first * block_size, n_bytes)
    b1 = min((first + count) * block_size, n_bytes)
    return first, count, b0, b1 - b0


def gather_totals(local_total, device, group=None):
    """all-gather of one int64 per rank -> tensor[world] (8 bytes per rank on the wire)"""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    mine = torch.as_tensor([int(local_total)], dtype=torch.int64, device=device) if not torch.is_tensor(local_total) \
        else local_total.reshape(1).to(torch.int64)
    if world == 1:
        return mine.clone()
    out = torch.empty(world, dtype=torch.int64, device=mine.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out


def base_offsets(totals):
    """exclusive scan of per-rank totals -> each rank's base offset in the logical output"""
    return torch.cumsum(totals, 0) - totals


def allreduce_histogram(counts64, group=None):
    """sum the uint64[256] histograms of all ranks in place (global-table mode)"""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts64, op=dist.ReduceOp.SUM, group=group)
    return counts64


def global_block_offsets(local_offsets, base):
    """this rank's uint64 block offsets shifted into the logical output"""
    return local_offsets + base
