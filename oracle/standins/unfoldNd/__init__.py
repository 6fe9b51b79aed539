"""TEST INFRASTRUCTURE ONLY -- stand-in for the un-vendored third-party package `unfoldNd`
(reference dependency, setup.py:28, unpinned; absent from /root/reference and from this image).

It exists so that the UNMODIFIED reference (/root/reference/nuwa_pytorch) can be imported in the build
container to generate golden vectors (oracle/make_golden.py).  It is never imported by the product.

Contract restated (published behaviour of unfoldNd.unfoldNd; call sites nuwa_pytorch.py:447,526,662):
N-d im2col with torch.nn.functional.unfold's conventions -- input (B, C, *spatial), output
(B, C*prod(kernel), L); output channel = c*J + j with j row-major over kernel offsets; L row-major over
output positions; zero padding.  Validated bit-equal against F.unfold in 2-D (tests/test_oracle_cpu.py).
Parity of this restatement against the real package is UNPINNED (package source unavailable).
"""
import itertools

import torch
import torch.nn.functional as F


def _tuple(v, n):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


def unfoldNd(input, kernel_size, dilation=1, padding=0, stride=1):
    nd = input.dim() - 2
    ks, dil, pad, st = (_tuple(v, nd) for v in (kernel_size, dilation, padding, stride))
    if any(pad):
        flat = []
        for p in reversed(pad):
            flat += [p, p]
        input = F.pad(input, flat)
    spatial = input.shape[2:]
    out_sizes = [(spatial[i] - dil[i] * (ks[i] - 1) - 1) // st[i] + 1 for i in range(nd)]
    cols = []
    for offs in itertools.product(*[range(k) for k in ks]):
        sl = [slice(None), slice(None)]
        for i in range(nd):
            start = offs[i] * dil[i]
            sl.append(slice(start, start + (out_sizes[i] - 1) * st[i] + 1, st[i]))
        cols.append(input[tuple(sl)].reshape(input.shape[0], input.shape[1], -1))
    out = torch.stack(cols, dim=2)  # (B, C, J, L)
    return out.reshape(input.shape[0], -1, out.shape[-1])
