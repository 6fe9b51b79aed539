import sys, torch
sys.path.insert(0, '.')
from nuwa_pytorch_b200 import ops_bwd
dev = torch.device('cuda'); H, dh = 8, 64; inner = 512; B, nq, nk = 8, 2560, 256
g = torch.Generator().manual_seed(5)
q = torch.randn(B, nq, inner, generator=g).bfloat16().to(dev); kv = torch.randn(B, nk, 2 * inner, generator=g).bfloat16().to(dev)
do = (torch.randn(B, nq, inner, generator=g) / 8).bfloat16().to(dev); talk = (torch.randn(H, H, generator=g) / 2).to(dev)
nkp, nvp = torch.randn(inner, generator=g).to(dev), torch.randn(inner, generator=g).to(dev)
ops_bwd.DENSE_BWD_FUSED = True
for it in range(3):
    dtalk = torch.zeros(H, H, device=dev); dnk, dnv = torch.zeros(inner, device=dev), torch.zeros(inner, device=dev)
    dq = torch.empty(B, nq, inner, dtype=torch.bfloat16, device=dev); dkv = torch.empty(B, nk, 2 * inner, dtype=torch.bfloat16, device=dev)
    ops_bwd.attn_dense_bwd(q.data_ptr(), kv.data_ptr(), kv.data_ptr() + inner * 2, do, B=B, nq=nq, nk=nk, H=H, dh=dh, q_bs=nq * inner, q_rs=inner, kv_bs=nk * 2 * inner, kv_rs=2 * inner, talk=talk, dtalk=dtalk, null_k=nkp, null_v=nvp, dnull_k=dnk, dnull_v=dnv, key_mask=None, dq_out=dq, dq_bs=nq * inner, dq_rs=inner, dk_ptr=dkv.data_ptr(), dv_ptr=dkv.data_ptr() + inner * 2, dkv_bs=nk * 2 * inner, dkv_rs=2 * inner, out_f32=False)
torch.cuda.synchronize()
