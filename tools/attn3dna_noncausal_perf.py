"""Centred (non-causal) Sparse3DNA at the NUWASketch sketch-encoder shape (batch 4 and 32, 768 sketch tokens = 3 frames,
8 heads x 64, kernel (5,3,3)): gather kernel vs the tcgen05 / TMEM kernel, per dilation.  CUDA events, L2 flushed."""
import json
import sys

import torch

sys.path.insert(0, '.')
from nuwa_pytorch_b200 import ops  # noqa: E402

dev = torch.device('cuda')
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
H, dh, nv = 8, 64, 767
inner, n = H * dh, nv + 1
rows = []
for B in (4, 32):
    qkv = torch.randn(B, n, 3 * inner, device=dev).bfloat16()
    talk = torch.randn(H, H, device=dev) / 2
    o = torch.empty(B, n, inner, dtype=torch.bfloat16, device=dev)
    for dil in (1, 2, 4):
        for name in ('gather', 'umma'):
            def run():
                ops.attn_sparse3dna(qkv, o, B=B, nq=n, t0=0, npos=n, H=H, dh=dh, talk=talk, fmap=16, max_frames=3, nv=nv,
                                    kernel=(5, 3, 3), dilation=(dil,) * 3, causal=False, variant=name)
            for _ in range(2):
                run()
            ts = []
            for _ in range(5):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                run()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ts.sort()
            rows.append(dict(batch=B, dilation=dil, kernel=name, us=round(ts[2] * 1e3, 1)))
            print(rows[-1], flush=True)
json.dump(rows, open('gpurun_out/attn3dna_noncausal_perf.json', 'w'), indent=1)
