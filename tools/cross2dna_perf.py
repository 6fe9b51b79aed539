"""SparseCross2DNA attention core at the cfg-5 shape (batch 4, 2560 video queries, 3 sketch frames = 768 context tokens,
8 heads x 64, 3 x 3 window): gather kernel vs the tcgen05 / TMEM kernel, per dilation.  CUDA events, L2 flushed."""
import json
import sys

import torch

sys.path.insert(0, '.')
from nuwa_pytorch_b200 import ops  # noqa: E402

dev = torch.device('cuda')
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
H, dh, frames, fmap = 8, 64, 3, 16
inner, nk, n = H * dh, frames * fmap * fmap, 2561
rows = []
for B in (4, 8):
    q = torch.randn(B, n, inner, device=dev).bfloat16()
    kv = torch.randn(B, nk, 2 * inner, device=dev).bfloat16()
    o = torch.empty(B, n, inner, dtype=torch.bfloat16, device=dev)
    talk = torch.randn(H, H, device=dev) / 2
    null_k, null_v = torch.randn(inner, device=dev), torch.randn(inner, device=dev)
    mask = (torch.rand(B, nk, device=dev) > 0.1).to(torch.uint8)
    for dil in (1, 2, 4):
        for name in (sys.argv[1:] or ['gather', 'umma']):
            def run():
                ops.attn_cross2dna(q.data_ptr() + inner * 2, kv.data_ptr(), kv.data_ptr() + inner * 2, o.data_ptr() + inner * 2,
                                   B=B, nq=n - 1, t0=1, H=H, dh=dh, q_bs=n * inner, q_rs=inner, k_bs=nk * 2 * inner,
                                   k_rs=2 * inner, v_bs=nk * 2 * inner, v_rs=2 * inner, o_bs=n * inner, o_rs=inner, talk=talk,
                                   null_k=null_k, null_v=null_v, key_mask=mask, fmap=fmap, frames=frames, ck=3, cdil=dil,
                                   variant=name)
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
            rows.append(dict(B=B, dilation=dil, kernel=name, us=round(ts[2] * 1e3, 1)))
            print(rows[-1], flush=True)
json.dump(rows, open('gpurun_out/cross2dna_perf.json', 'w'), indent=1)
