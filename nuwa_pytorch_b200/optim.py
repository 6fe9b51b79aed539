"""Trainer step of the training path: `clip_grad_norm_` + AdamW + `zero_grad` of
NUWATrainer.train_step (train_nuwa.py:237-260) with the optimizer of optimizer.py:11-31 (AdamW; parameters with
ndim < 2 are not decayed), as two HBM-bound CUDA passes over flat fp32 buffers (csrc/optim.cu).

Layout: parameters, gradients and both Adam moments share the flat layout of `train.GradStore` (every tensor starts
on a 16-byte boundary, model.named_parameters() order).  The nn.Parameters are re-pointed to views of the flat master
buffer (names and state_dict unchanged), so the update needs no gather / scatter, and the gradient buffer the CUDA
backward fills is consumed in place.  The data-parallel mean (parallel.GradAllReduce) has already been applied to
that buffer, so the norm and the update are identical on every rank without a further collective.
"""
import ctypes

import torch

from . import _lib
from ._lib import check, lib, ptr, stream
from .train import GradStore

CHUNK = 8192  # elements per CTA of the AdamW kernel


def trainable_parameters(model):
    """The parameter list the CUDA backward differentiates (the frozen VAE copies are excluded)."""
    return [p for n, p in model.named_parameters() if p.requires_grad and not n.startswith(('vae.', 'sketch_vae.'))]


class FusedAdamW:
    """get_optimizer(params, lr, wd) + clip_grad_norm_(params, max_grad_norm) + zero_grad as one fused step.

    step() returns the (pre-clip) global gradient norm as a 0-d device tensor, like clip_grad_norm_."""

    def __init__(self, params, lr=3e-4, wd=0.01, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=0.5):
        self.params = [p for p in params if p.requires_grad]
        if not self.params or not self.params[0].is_cuda:
            raise _lib.NuwaB200Error('FusedAdamW needs CUDA parameters (nuwa_pytorch_b200 has no CPU path)')
        assert all(p.dtype == torch.float32 for p in self.params)
        self.lr, self.wd, self.betas, self.eps, self.max_grad_norm = lr, wd, betas, eps, max_grad_norm
        self.layout = GradStore(self.params)      # offsets only; its zero buffer doubles as the fallback gradient buffer
        dev = self.params[0].device
        n = self.layout.flat.numel()
        self.master = torch.zeros(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, off in zip(self.params, self.layout.offsets):
                view = self.master[off:off + p.numel()].view(p.shape)
                view.copy_(p)
                p.data = view                        # the parameter now lives in the flat master buffer
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step_dev = torch.ones(1, dtype=torch.int32, device=dev)   # device-side step counter (graph replayable)
        self.sqnorm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.nparts = 4 * torch.cuda.get_device_properties(dev).multi_processor_count
        self.partials = torch.empty(self.nparts, dtype=torch.float32, device=dev)
        chunks = []
        for p, off in zip(self.params, self.layout.offsets):
            decay = int(wd != 0 and p.ndim >= 2)     # optimizer.py:6-9
            for s in range(0, p.numel(), CHUNK):
                chunks.append((off + s, min(CHUNK, p.numel() - s), decay))
        arr = (_lib.OptChunk * len(chunks))()
        for i, (o, ln, d) in enumerate(chunks):
            arr[i].offset, arr[i].len, arr[i].weight_decay = o, ln, d
        self.chunks_dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        self.nchunks = len(chunks)

    # ---- gradient buffer -----------------------------------------------------------------------
    def _flat_grads(self):
        """(flat, aliased): the flat buffer the parameters' .grad tensors alias (train.GradStore layout, aliased=True) --
        or, when the gradients were produced some other way, a gathered copy (aliased=False).

        Aliasing is detected by ADDRESS: autograd's AccumulateGrad stores `grad.detach()`, which drops `._base`, so the
        views handed out by GradStore arrive here as base-less tensors that still point into the flat buffer."""
        p0 = self.params[0]
        if p0.grad is None:
            raise _lib.NuwaB200Error('FusedAdamW.step(): no gradients (call loss.backward() first)')
        g0 = p0.grad
        n = self.layout.flat.numel()
        ok = g0.dtype == torch.float32 and g0.is_cuda
        if ok:
            st = g0.untyped_storage()
            b0 = g0.data_ptr()
            ok = (b0 - st.data_ptr()) % 4 == 0 and st.data_ptr() + st.nbytes() >= b0 + 4 * n
            if ok:
                for p, off in zip(self.params, self.layout.offsets):
                    g = p.grad
                    if (g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.data_ptr() != b0 + 4 * off
                            or g.untyped_storage().data_ptr() != st.data_ptr()):
                        ok = False
                        break
        if ok:
            flat = torch.empty(0, dtype=torch.float32, device=g0.device).set_(st, (b0 - st.data_ptr()) // 4, (n,), (1,))
            return flat, True
        flat = self.layout.flat
        flat.zero_()
        for p, off in zip(self.params, self.layout.offsets):
            if p.grad is not None:
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
        return flat, False

    @torch.no_grad()
    def step(self, grads_flat=None, grad_scale=1.0, zero_grad=True):
        aliased = True
        if grads_flat is not None:
            g = grads_flat
        else:
            g, aliased = self._flat_grads()
        assert g.dtype == torch.float32 and g.numel() == self.master.numel() and g.is_contiguous()
        check(lib().nuwa_sqnorm_f32(ptr(g), g.numel(), ptr(self.partials), self.nparts, ptr(self.sqnorm), 0, stream()),
              "nuwa_sqnorm_f32")
        a = _lib.AdamWParams()
        a.p, a.g, a.m, a.v = ptr(self.master), ptr(g), ptr(self.exp_avg), ptr(self.exp_avg_sq)
        a.chunks, a.nchunks = ptr(self.chunks_dev), self.nchunks
        a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = self.lr, self.betas[0], self.betas[1], self.eps, self.wd
        a.max_norm = float(self.max_grad_norm) if self.max_grad_norm else 0.0
        a.grad_scale, a.sqnorm = float(grad_scale), ptr(self.sqnorm)
        a.step, a.step_ptr, a.zero_grad = 0, ptr(self.step_dev), int(bool(zero_grad))
        check(lib().nuwa_adamw_step(ctypes.byref(a), stream()), "nuwa_adamw_step")
        check(lib().nuwa_step_increment(ptr(self.step_dev), stream()), "nuwa_step_increment")
        _lib.WEIGHTS_EPOCH[0] += 1                    # the weights changed: packed bf16 copies are refreshed lazily
        if zero_grad and not aliased:
            # the kernel zeroed the gathered scratch copy, not the tensors autograd accumulates into
            for p in self.params:
                if p.grad is not None:
                    p.grad.zero_()
        return self.sqnorm.sqrt() * grad_scale

    # ---- state (checkpointing, graph warm-up) --------------------------------------------------
    def state_snapshot(self):
        return tuple(t.clone() for t in (self.master, self.exp_avg, self.exp_avg_sq, self.step_dev))

    def state_restore(self, snap):
        with torch.no_grad():
            for dst, src in zip((self.master, self.exp_avg, self.exp_avg_sq, self.step_dev), snap):
                dst.copy_(src)
        _lib.WEIGHTS_EPOCH[0] += 1

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()
