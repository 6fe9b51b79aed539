"""Deterministic stand-ins for the data-format tests (no GPU, no reference needed at test time)."""
import torch


class StubVAE:
    """Quacks like VQGanVAE for the dataset code: image_size, num_layers, parameters(), get_video_indices()."""

    def __init__(self, image_size, num_layers, codebook):
        self.image_size, self.num_layers, self.codebook = image_size, num_layers, codebook
        self._p = torch.nn.Parameter(torch.zeros(1))
        self.calls = []

    def parameters(self):
        return iter([self._p])

    def get_video_indices(self, video):
        # (b, f, c, h, w) -> (b, f, fmap, fmap): a deterministic function of each frame's pixels only, so that batching
        # videos differently cannot change the result
        b, f = video.shape[:2]
        fmap = self.image_size // (self.num_layers ** 2)
        self.calls.append(b)
        pooled = torch.nn.functional.adaptive_avg_pool2d(video.reshape(b * f, *video.shape[2:]).float(), fmap).sum(1)
        return (pooled * 1000).round().long().remainder(self.codebook).reshape(b, f, fmap, fmap)


class StubVideos:
    def __init__(self, n, frames, channels, size, seed):
        g = torch.Generator().manual_seed(seed)
        self.v = torch.rand(n, frames, channels, size, size, generator=g)

    def __len__(self):
        return self.v.shape[0]

    def __getitem__(self, i):
        return torch.tensor([i]), self.v[i]
