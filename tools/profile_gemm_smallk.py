"""Small-K GEMM (decoder q|k|v projection shape: M=20480, N=1536, K=512, bf16 out) for an ncu source-level
capture of gemm_tcgen05_kernel's epilogue."""
import sys

import torch

sys.path.insert(0, '.')
from nuwa_pytorch_b200 import ops  # noqa: E402

dev = torch.device('cuda')
a = torch.randn(20480, 512, device=dev).bfloat16()
w = torch.randn(1536, 512, device=dev).bfloat16()
for _ in range(3):
    y = ops.gemm(a, w, out_dtype=torch.bfloat16)
torch.cuda.synchronize()
print(float(y.float().abs().mean()))
