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
