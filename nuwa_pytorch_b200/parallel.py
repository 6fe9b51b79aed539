"""Multi-GPU plumbing of the hot path.  The path shards by batch: each rank owns a contiguous
slice of the samples.  Inference (VAE, generate) has NO data-path collective -- only the benchmark's max-over-ranks
timing and an optional gather outside any timed region.  Training has exactly ONE: the gradient all-reduce
(sum -> mean) of the flat fp32 gradient buffer, issued per finished sub-block so that NCCL over NVLink overlaps the
rest of the backward (GradAllReduce)."""
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


class GradAllReduce:
    """Data-parallel gradient averaging over the flat fp32 gradient buffer of train.GradStore.

    `ready(lo, hi)` launches an asynchronous all-reduce of flat[lo:hi] as soon as the backward has finished writing
    that range (the collective runs on the backend's own stream and overlaps the remaining backward kernels);
    `finish()` reduces whatever was not announced and waits for everything.  NCCL averages in the collective
    (ReduceOp.AVG); backends without AVG (gloo, used by the CPU tests) sum and scale."""

    def __init__(self, dist, group=None, max_bucket_elems=32 * 1024 * 1024, min_bucket_elems=4 * 1024 * 1024):
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group) if dist is not None and dist.is_initialized() else 1
        self.max_bucket = max_bucket_elems
        # announced ranges are merged with their neighbours until a bucket holds min_bucket elements (16 MB): a decoder
        # sub-block is 0.8-2 M parameters, and ~50 collectives of 3-8 MB per backward cost more launch / ring latency
        # (and SM time taken from the backward kernels) than a dozen of 16-32 MB
        self.min_bucket = min_bucket_elems
        self.avg = self.world > 1 and dist.get_backend(group) == "nccl"
        self.flat, self.done, self.handles, self.pending = None, [], [], None

    def begin(self, flat):
        self.flat, self.done, self.handles, self.pending = flat, [], [], None

    def _launch(self, lo, hi):
        d = self.dist
        for s in range(lo, hi, self.max_bucket):
            chunk = self.flat[s:min(hi, s + self.max_bucket)]
            op = d.ReduceOp.AVG if self.avg else d.ReduceOp.SUM
            self.handles.append((d.all_reduce(chunk, op=op, group=self.group, async_op=True), chunk))

    def ready(self, lo, hi):
        if self.world == 1 or hi <= lo:
            return
        self.done.append((lo, hi))
        if self.pending is not None:
            plo, phi = self.pending
            if hi == plo or lo == phi:                     # adjacent (the backward finishes sub-blocks in layer order)
                lo, hi = min(lo, plo), max(hi, phi)
            else:
                self._launch(plo, phi)
        self.pending = (lo, hi)
        if hi - lo >= self.min_bucket:
            self._launch(lo, hi)
            self.pending = None

    def finish(self):
        if self.world == 1:
            return
        if self.pending is not None:
            self._launch(*self.pending)
            self.pending = None
        pos = 0
        for lo, hi in sorted(self.done):  # complement of the announced ranges
            if lo > pos:
                self._launch(pos, lo)
            pos = max(pos, hi)
        if pos < self.flat.numel():
            self._launch(pos, self.flat.numel())
        for h, chunk in self.handles:
            h.wait()
            if not self.avg:
                chunk.div_(self.world)
        self.handles = []
