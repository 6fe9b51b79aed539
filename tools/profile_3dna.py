"""One launch of each Sparse3DNA forward kernel variant at the cfg-3 shape, for ncu (tools: see profiles/)."""
import sys

import torch

sys.path.insert(0, '.')
from nuwa_pytorch_b200 import ops  # noqa: E402

dev = torch.device('cuda')
B, H, dh, nv = 8, 8, 64, 2559
inner, n = H * dh, nv + 1
qkv = torch.randn(B, n, 3 * inner, device=dev).bfloat16()
talk = torch.randn(H, H, device=dev) / 2
o = torch.empty(B, n, inner, dtype=torch.bfloat16, device=dev)
variants = sys.argv[1:] or ['halo']
for dil in (1, 2, 4):
    for name in variants:
        for _ in range(2):
            ops.attn_sparse3dna(qkv, o, B=B, nq=n, t0=0, npos=n, H=H, dh=dh, talk=talk, fmap=16, max_frames=10, nv=nv,
                                kernel=(5, 3, 3), dilation=(dil,) * 3, causal=True, variant=name)
torch.cuda.synchronize()
