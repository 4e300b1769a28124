"""Multi-GPU sharding of the encode path: one process per GPU, blocks split into contiguous ranges of whole
8-block groups (a group is one reference call -- reference ConvectionKernels.h:71 -- and must never be split, because
the reference takes some decisions jointly for its 8 blocks).  The path has no data dependency between groups, so the
only collectives are the distribution of input ranges and the collection of the encoded ranges:
one scatter and one gather over torch.distributed (NCCL on GPUs; gloo in the CPU tests).
"""
import torch
import torch.distributed as dist

GROUP = 8


def shard_ranges(n_blocks, world_size):
    """[(first_block, n_blocks_of_rank)] * world_size; boundaries are multiples of 8; sizes differ by at most one group."""
    if n_blocks % GROUP:
        raise ValueError("n_blocks must be a multiple of %d" % GROUP)
    groups = n_blocks // GROUP
    out = []
    for r in range(world_size):
        g0 = groups * r // world_size
        g1 = groups * (r + 1) // world_size
        out.append((g0 * GROUP, (g1 - g0) * GROUP))
    return out


def scatter_blocks(blocks, n_blocks, block_bytes, src=0, device=None, group=None):
    """Rank `src` holds `blocks` (uint8 tensor, n_blocks * block_bytes); every rank receives its contiguous range.
    Ranges may differ by one group, so the transfer is padded to the largest range and trimmed."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ranges = shard_ranges(n_blocks, world)
    pad = max(n for _, n in ranges) * block_bytes
    recv = torch.empty(pad, dtype=torch.uint8, device=device)
    if rank == src:
        flat = blocks.reshape(-1)
        chunks = []
        for first, n in ranges:
            if n * block_bytes == pad:          # equal ranges (the usual case): send views, no staging copy
                chunks.append(flat[first * block_bytes:(first + n) * block_bytes])
                continue
            c = torch.zeros(pad, dtype=torch.uint8, device=device)
            c[: n * block_bytes] = flat[first * block_bytes:(first + n) * block_bytes]
            chunks.append(c)
        dist.scatter(recv, chunks, src=src, group=group)
    else:
        dist.scatter(recv, None, src=src, group=group)
    return recv[: ranges[rank][1] * block_bytes]


def gather_encoded(encoded, n_blocks, out_block_bytes, dst=0, group=None):
    """Inverse of scatter_blocks for the encoded ranges; returns the whole (n_blocks, out_block_bytes) tensor on `dst`,
    None elsewhere."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ranges = shard_ranges(n_blocks, world)
    pad = max(n for _, n in ranges) * out_block_bytes
    flat = encoded.reshape(-1)
    if flat.numel() == pad:
        send = flat
    else:
        send = torch.zeros(pad, dtype=torch.uint8, device=encoded.device)
        send[: flat.numel()] = flat
    if rank == dst:
        parts = [torch.empty(pad, dtype=torch.uint8, device=encoded.device) for _ in range(world)]
        dist.gather(send, parts, dst=dst, group=group)
        return torch.cat([p[: n * out_block_bytes] for p, (_, n) in zip(parts, ranges)]).reshape(n_blocks, out_block_bytes)
    dist.gather(send, None, dst=dst, group=group)
    return None


def encode_sharded(encode_fn, blocks, n_blocks, in_block_bytes, out_block_bytes, src=0, device=None, group=None):
    """scatter -> encode_fn(local uint8 tensor of whole groups) -> gather.  `encode_fn` returns a uint8 tensor of
    n_local * out_block_bytes bytes.  Returns the full result on `src`, None elsewhere."""
    local = scatter_blocks(blocks, n_blocks, in_block_bytes, src=src, device=device, group=group)
    enc = encode_fn(local)
    return gather_encoded(enc, n_blocks, out_block_bytes, dst=src, group=group)
