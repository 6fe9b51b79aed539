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
    """Whole-step capture of `loss = fn(*inputs); loss.backward()`: one graph replay runs the CUDA forward, the CUDA
    backward and the writes of every `p.grad` (static buffers owned by the graph's memory pool; each replay overwrites
    them, so accumulate / apply the gradients before the next replay)."""

    def __init__(self, fn, params, *example_inputs, warmup=3):
        self.params = [p for p in params if p.requires_grad]
        self.static_inputs = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):  # packs weights, sets kernel attributes, warms the allocator outside the capture
                for p in self.params:
                    p.grad = None
                fn(*self.static_inputs).backward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for p in self.params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        # capture on the warm-up stream: autograd binds each parameter's AccumulateGrad node to the stream it was created
        # on, and a node that outlived an earlier eager step on another stream would invalidate the capture
        with torch.cuda.graph(self.graph, stream=side):
            self.loss = fn(*self.static_inputs)
            self.loss.backward()

    def __call__(self, *inputs):
        for dst, src in zip(self.static_inputs, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss
