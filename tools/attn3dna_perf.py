"""Sparse3DNA attention core at the cfg-3 shape (batch 8, 2560 video tokens, 8 heads x 64, kernel (5,3,3)):
gather kernel vs the halo-tiled mma.sync kernel vs the tcgen05 / TMEM kernel, per dilation.  CUDA events, L2 flushed."""
import json
import sys

import torch

sys.path.insert(0, '.')
from nuwa_pytorch_b200 import ops  # noqa: E402

dev = torch.device('cuda')
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
B, H, dh, nv = 8, 8, 64, 2559
inner = H * dh
n = nv + 1
qkv = torch.randn(B, n, 3 * inner, device=dev).bfloat16()
qkv_src = qkv.clone()
talk = torch.randn(H, H, device=dev) / 2
o = torch.empty(B, n, inner, dtype=torch.bfloat16, device=dev)
rows = []
for dil in (1, 2, 4):
    for name in (sys.argv[1:] or ['gather', 'halo', 'umma']):
        def run():
            ops.attn_sparse3dna(qkv, o, B=B, nq=n, t0=0, npos=n, H=H, dh=dh, talk=talk, fmap=16, max_frames=10, nv=nv,
                                kernel=(5, 3, 3), dilation=(dil,) * 3, causal=True, variant=name)
        for _ in range(2):
            run()
        ts, tw = [], []
        for _ in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            run()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        for _ in range(5):   # L2-warm: q|k|v (63 MB) rewritten just before, as the projection GEMM leaves it in the real step
            qkv.copy_(qkv_src)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            run()
            b.record()
            torch.cuda.synchronize()
            tw.append(a.elapsed_time(b))
        ts.sort()
        tw.sort()
        us = ts[2] * 1e3
        gbs = B * nv * 4096 / (us * 1e-6) / 1e9  # algorithmic bytes: 4096 B per token
        rows.append(dict(dilation=dil, kernel=name, us=round(us, 1), algorithmic_GBps=round(gbs, 1),
                         us_l2_warm=round(tw[2] * 1e3, 1)))
        print(rows[-1], flush=True)
json.dump(rows, open('gpurun_out/attn3dna_perf.json', 'w'), indent=1)
