"""One dense-attention launch at the cfg-3 cross-attention shape for an ncu --set full capture."""
import sys

import torch

sys.path.insert(0, '.')
from nuwa_pytorch_b200 import ops  # noqa: E402

dev = torch.device('cuda')
H, dh, B, nq, nk = 8, 64, 8, 2560, 256
inner = H * dh
q = torch.randn(B, nq, inner, device=dev).bfloat16()
kv = torch.randn(B, nk, 2 * inner, device=dev).bfloat16()
talk = torch.randn(H, H, device=dev) / 2
nk_, nv_ = torch.randn(inner, device=dev), torch.randn(inner, device=dev)
mask = (torch.rand(B, nk, device=dev) > 0.2).to(torch.uint8)
o = torch.empty(B, nq, inner, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    ops.attn_dense(q.data_ptr(), kv.data_ptr(), kv.data_ptr() + inner * 2, o, B=B, nq=nq, nk=nk, H=H, dh=dh, q_bs=nq * inner,
                   q_rs=inner, k_bs=nk * 2 * inner, k_rs=2 * inner, v_bs=nk * 2 * inner, v_rs=2 * inner, o_bs=nq * inner,
                   o_rs=inner, talk=talk, null_k=nk_, null_v=nv_, key_mask=mask)
torch.cuda.synchronize()
print(float(o.float().abs().mean()))
