"""The training step of the decoder path end to end: the body of `NUWATrainer.train_step`
(reference train_nuwa.py:237-258) --

    for _ in range(grad_accum_every):                       loss = nuwa(text=, video=, return_loss=True)
                                                            (loss / grad_accum_every).backward()
    clip_grad_norm_(nuwa.parameters(), max_grad_norm);      optim.step();  optim.zero_grad()

with `optim = get_optimizer(nuwa.parameters(), lr, wd)` (optimizer.py:11-31: AdamW, tensors with ndim < 2 not decayed) --
as ONE object that owns the B200-side data layout:

  * every trainable parameter lives in one flat fp32 master buffer (optim.FusedAdamW), every gradient in one flat fp32
    accumulator with the same layout (train.GradStore) that is each parameter's `.grad` for the whole run: the
    weight-gradient kernels ADD into it, so micro-batches accumulate in place (no per-parameter add kernels) and the
    optimizer consumes and zeroes it in the same pass;
  * data parallel (one process per GPU, batch sharded): the step's ONE collective, the mean all-reduce of that buffer over
    NCCL, is issued during the LAST micro-batch's backward, per finished sub-block, so that NVLink traffic overlaps the
    remaining backward kernels (parallel.GradAllReduce); clip + AdamW then run identically on every rank;
  * `capture()` records the whole step -- packed-weight refresh, `grad_accum_every` forward + backward passes, the
    all-reduce, clip + AdamW + zero-grad -- into one CUDA graph (streams and graphs instead of a tracing compiler).

Out of scope (NUWATrainer's non-arithmetic shell): the DataLoader, periodic sampling / checkpoint files, console prompts.
"""
import torch

from . import _lib
from .optim import FusedAdamW, trainable_parameters
from .parallel import GradAllReduce
from .train import GradStore


class TrainStep:
    """train_step of NUWATrainer for NUWA or NUWASketch.  `step(batches)` takes `grad_accum_every` keyword dicts for
    `model(**batch, return_loss=True)` and returns (mean loss, pre-clip gradient norm) as 0-d device tensors."""

    def __init__(self, model, *, lr=3e-4, wd=0.01, grad_accum_every=8, max_grad_norm=0.5, dist=None, forward_kwargs=None):
        self.model = model
        self.accum = int(grad_accum_every)
        self.forward_kwargs = dict(forward_kwargs or {})
        self.params = trainable_parameters(model)
        if not self.params or not self.params[0].is_cuda:
            raise _lib.NuwaB200Error('TrainStep needs a model on a CUDA device (nuwa_pytorch_b200 has no CPU path)')
        self.opt = FusedAdamW(self.params, lr=lr, wd=wd, max_grad_norm=max_grad_norm)
        self.store = GradStore(self.params)                 # persistent accumulator, same layout as the master buffer
        assert self.store.offsets == self.opt.layout.offsets
        for p in self.params:
            p.grad = self.store(p)
        model._grad_store = self.store
        world = dist.get_world_size() if (dist is not None and dist.is_initialized()) else 1
        self.reducer = GradAllReduce(dist) if world > 1 else None
        self.graph = None
        self.steps = 0

    # ---- eager -----------------------------------------------------------------------------------
    def _micro(self, batch, last):
        self.model._grad_reducer = self.reducer if last else None   # the collective rides on the last backward only
        loss = self.model(**batch, return_loss=True, **self.forward_kwargs)
        (loss / self.accum).backward()
        self.model._grad_reducer = None
        return loss.detach()

    def _body(self, batches):
        self.model.train()
        total = None
        for i, b in enumerate(batches):
            lval = self._micro(b, last=(i == len(batches) - 1)) / self.accum
            total = lval if total is None else total + lval
        norm = self.opt.step(grads_flat=self.store.flat)            # clip + AdamW + zero_grad, one fused pass
        return total, norm

    def step(self, batches):
        assert len(batches) == self.accum, f'expected {self.accum} micro-batches, got {len(batches)}'
        if self.graph is not None:
            return self._replay(batches)
        out = self._body(batches)
        self.steps += 1
        return out

    # ---- whole step as one CUDA graph --------------------------------------------------------------
    def capture(self, example_batches, warmup=2):
        """Record the step for batches shaped like `example_batches`; later `step()` calls copy their tensors into the
        static inputs and replay.  Warm-up steps are rolled back (parameters, moments, step counter)."""
        assert len(example_batches) == self.accum
        self.static = [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in b.items()} for b in example_batches]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        snap = self.opt.state_snapshot()

        def body():
            self.model.refresh_packed_weights()                     # the weights changed since the last replay
            return self._body(self.static)

        with torch.cuda.stream(side):
            for _ in range(warmup):
                body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.opt.state_restore(snap)
        self.store.flat.zero_()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            self.static_out = body()
        return self

    def _replay(self, batches):
        for dst, src in zip(self.static, batches):
            for k, v in src.items():
                if torch.is_tensor(v) and dst[k].data_ptr() != v.data_ptr():
                    dst[k].copy_(v, non_blocking=True)
        self.graph.replay()
        _lib.WEIGHTS_EPOCH[0] += 1      # host-side mirror of the captured update (eager callers refresh their packs)
        self.steps += 1
        return self.static_out
