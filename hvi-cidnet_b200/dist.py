"""Multi-GPU plumbing for the CIDNet forward: images are independent units (the attention is per
image), so a batch is partitioned contiguously over the ranks of one node, every rank
runs the single-GPU path on its share and NO data-path collective is needed.  The only collective is
the optional gather of the outputs onto every rank."""
import torch
import torch.distributed as dist


def shard_range(n_items, world_size, rank):
    """Contiguous, balanced partition of `n_items` over `world_size` ranks: (start, count).
    The first `n_items % world_size` ranks get one extra item; empty shards are allowed."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world {world_size}")
    base, extra = divmod(int(n_items), int(world_size))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def forward_sharded(model, x_full, group=None, gather=True):
    """Run `model` on this rank's share of the batch `x_full` ([B,3,H,W], present on every rank, on the
    rank's own device).  With `gather` the per-rank outputs are all-gathered (ragged shards supported)
    and the full [B,3,H,W] result is returned on every rank; otherwise only the local shard's output."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    start, count = shard_range(x_full.shape[0], world, rank)
    local = x_full[start:start + count]
    out = model(local) if count > 0 else x_full.new_empty((0,) + tuple(x_full.shape[1:]))
    if not gather or world == 1:
        return out
    counts = [shard_range(x_full.shape[0], world, r)[1] for r in range(world)]
    cmax = max(counts)
    pad = out.new_zeros((cmax,) + tuple(out.shape[1:]))
    pad[:count] = out
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
