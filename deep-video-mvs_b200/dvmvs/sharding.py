"""Clip sharding for multi-GPU inference: sequences are independent (state is per clip), so
rank r of N owns clips r, r+N, ...; ranks never exchange activations.  torch.distributed (NCCL on GPUs, gloo in the CPU
tests) is used only to agree on timings (max over ranks) and, optionally, to gather per-clip results on rank 0."""
import torch
import torch.distributed as dist


def clips_of_rank(n_clips, rank, world_size):
    """Round-robin assignment; every clip belongs to exactly one rank."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return list(range(rank, n_clips, world_size))


def max_over_ranks(value, device=None):
    """Whole-job time of a step = the slowest rank's device time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def sum_over_ranks(value, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])


def gather_clip_results(local_results, n_clips, device=None):
    """local_results: {clip_id: tensor} of this rank (identical shapes everywhere).  Returns on every rank the list of
    n_clips tensors in clip order (all_gather of the small per-clip outputs, e.g. 256 KiB depth maps)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local_results[c] for c in range(n_clips)]
    world, rank = dist.get_world_size(), dist.get_rank()
    per_rank = (n_clips + world - 1) // world
    sample = next(iter(local_results.values()))
    buf = torch.zeros((per_rank,) + tuple(sample.shape), dtype=sample.dtype, device=sample.device)
    for i, c in enumerate(clips_of_rank(n_clips, rank, world)):
        buf[i] = local_results[c]
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return [out[c % world][c // world] for c in range(n_clips)]
