"""Dense attention backward at the cfg-3 cross-attention shape (B 8, 2560 queries, 256 text keys + null, 8 x 64):
fused probability stage (attention_dense_bwd.cu) vs the materialised-logits path, CUDA events, and the error of either
against fp32 autograd on the same bf16 operands (smaller B)."""
import sys

import torch

sys.path.insert(0, '.')
sys.path.insert(0, 'oracle')
from nuwa_pytorch_b200 import ops_bwd  # noqa: E402

dev = torch.device('cuda')
H, dh = 8, 64
inner = H * dh


def run(B, nq, nk, fused, reps=5, check=False):
    g = torch.Generator().manual_seed(5)
    q = torch.randn(B, nq, inner, generator=g).bfloat16().to(dev)
    kv = torch.randn(B, nk, 2 * inner, generator=g).bfloat16().to(dev)
    do = (torch.randn(B, nq, inner, generator=g) / 8).bfloat16().to(dev)
    talk = (torch.randn(H, H, generator=g) / 2).to(dev)
    nkp, nvp = torch.randn(inner, generator=g).to(dev), torch.randn(inner, generator=g).to(dev)
    ops_bwd.DENSE_BWD_FUSED = fused
    ts = []
    for it in range(reps + 2):
        dtalk = torch.zeros(H, H, device=dev)
        dnk, dnv = torch.zeros(inner, device=dev), torch.zeros(inner, device=dev)
        dq = torch.empty(B, nq, inner, dtype=torch.bfloat16, device=dev)
        dkv = torch.empty(B, nk, 2 * inner, dtype=torch.bfloat16, device=dev)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops_bwd.attn_dense_bwd(q.data_ptr(), kv.data_ptr(), kv.data_ptr() + inner * 2, do, B=B, nq=nq, nk=nk, H=H, dh=dh,
                               q_bs=nq * inner, q_rs=inner, kv_bs=nk * 2 * inner, kv_rs=2 * inner, talk=talk, dtalk=dtalk,
                               null_k=nkp, null_v=nvp, dnull_k=dnk, dnull_v=dnv, key_mask=None, dq_out=dq, dq_bs=nq * inner,
                               dq_rs=inner, dk_ptr=dkv.data_ptr(), dv_ptr=dkv.data_ptr() + inner * 2, dkv_bs=nk * 2 * inner,
                               dkv_rs=2 * inner, out_f32=False)
        b.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(a.elapsed_time(b))
    ops_bwd.DENSE_BWD_FUSED = True
    out = dict(fused=fused, B=B, nq=nq, nk=nk, ms=round(sorted(ts)[len(ts) // 2], 3))
    if check:
        qf, kvf, tf = q.float().requires_grad_(), kv.float().requires_grad_(), talk.clone().requires_grad_()
        nkf, nvf = nkp.clone().requires_grad_(), nvp.clone().requires_grad_()
        qh = qf.view(B, nq, H, dh).transpose(1, 2) * dh ** -0.5
        k, v = kvf.chunk(2, -1)
        kh = torch.cat([nkf.view(1, H, 1, dh).expand(B, -1, -1, -1), k.reshape(B, nk, H, dh).transpose(1, 2)], 2)
        vh = torch.cat([nvf.view(1, H, 1, dh).expand(B, -1, -1, -1), v.reshape(B, nk, H, dh).transpose(1, 2)], 2)
        attn = (qh @ kh.transpose(-1, -2)).softmax(-1)
        attn = torch.einsum('gh,bhqj->bgqj', tf, attn)
        o = (attn @ vh).transpose(1, 2).reshape(B, nq, inner)
        o.backward(do.float())
        rel = lambda a_, b_: ((a_ - b_).norm() / b_.norm()).item()
        out.update(dq=rel(dq.float(), qf.grad), dkv=rel(dkv.float(), kvf.grad), dtalk=rel(dtalk, tf.grad),
                   dnull_k=rel(dnk, nkf.grad), dnull_v=rel(dnv, nvf.grad))
    return out


for fused in (False, True):
    print(run(2, 512, 256, fused, check=True), flush=True)
for fused in (False, True, False, True):
    print(run(8, 2560, 256, fused), flush=True)
