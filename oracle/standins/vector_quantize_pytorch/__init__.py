"""TEST INFRASTRUCTURE ONLY -- stand-in for the un-vendored third-party package
`vector-quantize-pytorch` (reference dependency, setup.py:29, ">=0.4.10"; absent from /root/reference and
from this image).  Used solely to import the UNMODIFIED reference for golden generation.

Contract restated from the published v0.4.10 behaviour (call sites vqgan_vae.py:6,368-378,429,435;
nuwa_pytorch.py:1910,2507):  VectorQuantize(dim, codebook_size, codebook_dim, decay, commitment_weight,
accept_image_fmap, kmeans_init, use_cosine_sim) ; forward (B,C,h,w) -> 'b (h w) c' -> project_in (Linear iff
codebook_dim != dim) -> codebook -> [train: straight-through, commit loss] -> project_out -> 'b c h w';
returns (quantize, indices (B,h,w) int64, loss (1,)).
 * cosine codebook: x^=l2norm(x), e^=l2norm(embed), ind = argmax(x^ @ e^.T) (first max wins), quantize =
   embed[ind] (raw, un-normalised buffer).  Buffers initted (1,), cluster_size (K,), embed (K,D).
 * euclidean codebook: ind = argmax(-(|x|^2 - 2 x.e + |e|^2)); extra buffer embed_avg.
 * training mode additionally EMA-updates cluster_size / embed (decay) -- needed only so that the reference
   module runs; the hot path (all BASELINE configs) is eval-mode.
Parity of this restatement against the real package is UNPINNED (package source unavailable).
"""
import torch
import torch.nn.functional as F
from torch import nn


def _l2norm(t):
    return F.normalize(t, p=2, dim=-1)


def _ema_inplace(avg, new, decay):
    avg.mul_(decay).add_(new, alpha=1 - decay)


def _kmeans(samples, k, iters=10, cosine=False):
    n = samples.shape[0]
    idx = torch.randperm(n)[:k] if n >= k else torch.randint(0, n, (k,))
    means = samples[idx].clone()
    for _ in range(iters):
        d = samples @ means.t() if cosine else -torch.cdist(samples, means)
        b = d.argmax(-1)
        cnt = torch.bincount(b, minlength=k)
        new = torch.zeros_like(means).index_add_(0, b, samples) / cnt.clamp(min=1)[:, None]
        if cosine:
            new = _l2norm(new)
        means = torch.where((cnt == 0)[:, None], means, new)
    return means, cnt


class _Codebook(nn.Module):
    def __init__(self, dim, codebook_size, kmeans_init, kmeans_iters, decay, eps, cosine):
        super().__init__()
        self.decay, self.eps, self.cosine = decay, eps, cosine
        self.codebook_size, self.kmeans_iters = codebook_size, kmeans_iters
        if kmeans_init:
            embed = torch.zeros(codebook_size, dim)
        else:
            embed = _l2norm(torch.randn(codebook_size, dim)) if cosine else torch.randn(codebook_size, dim)
        self.register_buffer('initted', torch.Tensor([not kmeans_init]))
        self.register_buffer('cluster_size', torch.zeros(codebook_size))
        self.register_buffer('embed', embed)
        if not cosine:
            self.register_buffer('embed_avg', embed.clone())

    def _maybe_init(self, flat):
        if self.initted.item():
            return
        means, cnt = _kmeans(flat, self.codebook_size, self.kmeans_iters, self.cosine)
        self.embed.data.copy_(means)
        self.cluster_size.data.copy_(cnt.float())
        self.initted.data.copy_(torch.Tensor([True]))

    def forward(self, x):
        shape, dtype = x.shape, x.dtype
        flat = x.reshape(-1, shape[-1])
        if self.cosine:
            flat = _l2norm(flat)
        self._maybe_init(flat)
        if self.cosine:
            dist = flat @ _l2norm(self.embed).t()
        else:
            e = self.embed.t()
            dist = -(flat.pow(2).sum(1, keepdim=True) - 2 * flat @ e + e.pow(2).sum(0, keepdim=True))
        ind = dist.max(dim=-1).indices
        onehot = F.one_hot(ind, self.codebook_size).type(dtype)
        ind = ind.view(*shape[:-1])
        quantize = F.embedding(ind, self.embed)
        if self.training:
            bins = onehot.sum(0)
            if self.cosine:
                _ema_inplace(self.cluster_size, bins, self.decay)
                zero = bins == 0
                bins = bins.masked_fill(zero, 1.)
                embed_norm = _l2norm((flat.t() @ onehot / bins[None]).t())
                embed_norm = torch.where(zero[:, None], self.embed, embed_norm)
                _ema_inplace(self.embed, embed_norm, self.decay)
            else:
                _ema_inplace(self.cluster_size, bins, self.decay)
                _ema_inplace(self.embed_avg, (flat.t() @ onehot).t(), self.decay)
                n = self.cluster_size.sum()
                cs = (self.cluster_size + self.eps) / (n + self.codebook_size * self.eps) * n
                self.embed.data.copy_(self.embed_avg / cs[:, None])
        return quantize, ind


class VectorQuantize(nn.Module):
    def __init__(self, dim, codebook_size, codebook_dim=None, decay=0.8, eps=1e-5, kmeans_init=False,
                 kmeans_iters=10, use_cosine_sim=False, channel_last=True, accept_image_fmap=False,
                 commitment_weight=1., **kwargs):
        super().__init__()
        codebook_dim = codebook_dim if codebook_dim is not None else dim
        proj = codebook_dim != dim
        self.project_in = nn.Linear(dim, codebook_dim) if proj else nn.Identity()
        self.project_out = nn.Linear(codebook_dim, dim) if proj else nn.Identity()
        self.commitment_weight = commitment_weight
        self.accept_image_fmap = accept_image_fmap
        self.channel_last = channel_last
        self.codebook_size = codebook_size
        self._codebook = _Codebook(codebook_dim, codebook_size, kmeans_init, kmeans_iters, decay, eps, use_cosine_sim)

    @property
    def codebook(self):
        return self._codebook.embed

    def forward(self, x):
        device = x.device
        if self.accept_image_fmap:
            b, c, h, w = x.shape
            x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
        elif not self.channel_last:
            x = x.transpose(1, 2)
        x = self.project_in(x)
        quantize, ind = self._codebook(x)
        if self.training:
            quantize = x + (quantize - x).detach()
        loss = torch.tensor([0.], device=device, requires_grad=self.training)
        if self.training and self.commitment_weight > 0:
            loss = loss + F.mse_loss(quantize.detach(), x) * self.commitment_weight
        quantize = self.project_out(quantize)
        if self.accept_image_fmap:
            quantize = quantize.reshape(b, h, w, -1).permute(0, 3, 1, 2)
            ind = ind.reshape(b, h, w)
        elif not self.channel_last:
            quantize = quantize.transpose(1, 2)
        return quantize, ind, loss
