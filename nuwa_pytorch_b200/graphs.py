"""CUDA-graph capture of a hot-path call (streams and graphs instead of a tracing compiler).

The transformer paths launch ~150 small kernels per layer-stack pass through ctypes; replaying them from a
captured CUDA graph removes the per-launch host cost.  Everything our modules do inside forward() is capture
safe: raw kernel launches on the current stream, caching-allocator allocations, no host synchronisation.
"""
import torch


class GraphedCall:
    """Capture `fn(*static_inputs)` once; `__call__(*inputs)` copies new inputs into the static buffers and replays.

    `fn` must be shape-static and free of host syncs (e.g. NUWA.forward with return_loss=True, VQGanVAE.forward).
    The returned tensors are static output buffers that are overwritten by the next call."""

    def __init__(self, fn, *example_inputs, warmup=2):
        self.static_inputs = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):  # builds packed weights / sets kernel attributes outside the capture
                fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_outputs = fn(*self.static_inputs)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_inputs, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_outputs


class GraphedTrainStep:
    """Whole-step capture of `loss = fn(*inputs); loss.backward()` [+ `optimizer.step()`]: one graph replay runs the
    CUDA forward, the CUDA backward and the writes of every `p.grad` (static buffers owned by the graph's memory pool).

    Weights may change between replays (that is the point of a training step), so the captured region starts by
    re-deriving the packed bf16 GEMM operands from the fp32 parameters in place (`model.refresh_packed_weights()`:
    fixed addresses, device copies only).  Pass `model=` whenever the parameters are updated between replays -- by a
    captured `optimizer=` (optim.FusedAdamW: clip + AdamW + zero-grad become part of the same graph) or by any
    optimizer stepping outside it.  Without `model=` the graph assumes frozen weights and raises if they moved.

    `grad_hook`, if given, is called after backward inside the capture (e.g. a NCCL all-reduce of the flat gradient
    buffer enqueued on the capture stream)."""

    def __init__(self, fn, params, *example_inputs, warmup=3, model=None, optimizer=None, grad_hook=None):
        from . import _lib
        self._lib = _lib
        self.params = [p for p in params if p.requires_grad]
        self.model, self.optimizer = model, optimizer
        self.static_inputs = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())

        def body():
            if model is not None:
                model.refresh_packed_weights()
            loss = fn(*self.static_inputs)
            loss.backward()
            if grad_hook is not None:
                grad_hook()
            if optimizer is not None:
                optimizer.step()
            return loss

        snapshot = None
        if optimizer is not None:  # warm-up steps must not train: restore parameters / moments / step counter afterwards
            snapshot = optimizer.state_snapshot()
        with torch.cuda.stream(side):
            for _ in range(warmup):  # packs weights, sets kernel attributes, warms the allocator outside the capture
                for p in self.params:
                    p.grad = None
                body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if snapshot is not None:
            optimizer.state_restore(snapshot)
            if model is not None:
                model.refresh_packed_weights()
        for p in self.params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        # capture on the warm-up stream: autograd binds each parameter's AccumulateGrad node to the stream it was created
        # on, and a node that outlived an earlier eager step on another stream would invalidate the capture
        with torch.cuda.graph(self.graph, stream=side):
            self.loss = body()
        self._versions = self._weights_signature()

    def _weights_signature(self):
        return (self._lib.WEIGHTS_EPOCH[0],) + tuple((p.data_ptr(), p._version) for p in self.params)

    def __call__(self, *inputs):
        if self.model is None and self._weights_signature() != self._versions:
            raise self._lib.NuwaB200Error(
                'GraphedTrainStep was captured without model=: its graph holds the packed copies of the weights as they '
                'were at capture time, and the parameters have changed since.  Re-create it with model=<the module> so '
                'the packed weights are refreshed inside the graph')
        for dst, src in zip(self.static_inputs, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        if self.optimizer is not None:
            self._lib.WEIGHTS_EPOCH[0] += 1  # host-side mirror of the captured update: eager callers refresh their packs
        return self.loss
