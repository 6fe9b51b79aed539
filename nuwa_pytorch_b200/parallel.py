"""Multi-GPU plumbing of the hot path.  The path shards by batch with NO data-path collective (inference
replicas): each rank owns a contiguous slice of the samples.  The only reductions are the
benchmark's max-over-ranks timing and (for callers that want the outputs in one place) an optional gather
outside any timed region."""
import torch


def rank_slice(n_items, rank, world):
    """Contiguous [start, stop) of `n_items` independent samples owned by `rank` (remainder to the first ranks)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def max_over_ranks(seconds, dist=None, device="cpu"):
    """Whole-job time of a step = the slowest rank's device time."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(seconds)
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_indices(local_idx, dist=None):
    """Collect per-rank token-id tensors (same shape on every rank) on every rank, in rank order."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_idx
    out = [torch.empty_like(local_idx) for _ in range(dist.get_world_size())]
    dist.all_gather(out, local_idx.contiguous())
    return torch.cat(out, dim=0)
