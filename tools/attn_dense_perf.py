"""Dense attention core at the cfg-3 cross-attention shape (batch 8, 2560 queries, 256 text keys + null, 8 heads x 64,
key mask, talking heads) and at the text-encoder shape (256 x 256): CUDA events, L2 flushed between iterations."""
import json
import sys

import torch

sys.path.insert(0, '.')
from nuwa_pytorch_b200 import ops  # noqa: E402

dev = torch.device('cuda')
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
H, dh = 8, 64
inner = H * dh
rows = []
for (B, nq, nk, variant) in ((8, 2560, 256, 'lib'), (8, 2560, 256, 'pres'), (8, 256, 256, 'lib'), (8, 256, 256, 'pres'), (4, 2560, 768, 'lib')):
    q = torch.randn(B, nq, inner, device=dev).bfloat16()
    kv = torch.randn(B, nk, 2 * inner, device=dev).bfloat16()
    talk = torch.randn(H, H, device=dev) / 2
    nk_, nv_ = torch.randn(inner, device=dev), torch.randn(inner, device=dev)
    mask = (torch.rand(B, nk, device=dev) > 0.2).to(torch.uint8)
    o = torch.empty(B, nq, inner, dtype=torch.bfloat16, device=dev)

    def run():
        ops.attn_dense(q.data_ptr(), kv.data_ptr(), kv.data_ptr() + inner * 2, o, B=B, nq=nq, nk=nk, H=H, dh=dh,
                       q_bs=nq * inner, q_rs=inner, k_bs=nk * 2 * inner, k_rs=2 * inner, v_bs=nk * 2 * inner, v_rs=2 * inner,
                       o_bs=nq * inner, o_rs=inner, talk=talk, null_k=nk_, null_v=nv_, key_mask=mask, variant=variant)
    for _ in range(3):
        run()
    ts = []
    for _ in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    us = ts[len(ts) // 2] * 1e3
    flops = 2 * 2 * B * H * nq * (nk + 1) * dh
    rows.append(dict(B=B, nq=nq, nk=nk, variant=variant, us=round(us, 1), tflops=round(flops / us / 1e6, 1)))
    print(rows[-1], flush=True)
json.dump(rows, open('gpurun_out/attn_dense_perf.json', 'w'), indent=1)
