"""Multi-GPU plumbing of the forward path: images are independent units, so inference shards the batch
across ranks with NO data-path collective.  torch.distributed is used only to place
ranks, to barrier around the timed region and to reduce timings (max over ranks)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced shard [lo, hi) of `n_items` for `rank` (first n % world ranks get one extra)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(values, device="cpu"):
    """Element-wise maximum of a list of floats over all ranks (identity when not distributed)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def gather_to_rank0(tensor):
    """Concatenate per-rank result shards on rank 0 (used by tests / offline evaluation, never inside a
    timed region).  Returns None on other ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensor
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([tensor.shape[0]], dtype=torch.int64))
    counts = [int(s.item()) for s in sizes]
    n_max = max(counts)
    mine = tensor.cpu().contiguous()
    if mine.shape[0] < n_max:  # gather needs equal shapes: pad ragged shards, trim on rank 0
        pad = torch.zeros((n_max - mine.shape[0],) + tuple(mine.shape[1:]), dtype=mine.dtype)
        mine = torch.cat([mine, pad], 0)
    out = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, out, dst=0)
    return torch.cat([o[:c] for o, c in zip(out, counts)], 0) if rank == 0 else None
