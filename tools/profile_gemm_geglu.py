"""FF1 GEMM of cfg 3 (20480 x 2752 x 512, GEGLU epilogue, bf16 out) for an ncu --set full capture."""
import sys

import torch

sys.path.insert(0, '.')
from nuwa_pytorch_b200 import ops  # noqa: E402

dev = torch.device('cuda')
a = torch.randn(20480, 512, device=dev).bfloat16()
w = torch.randn(2752, 512, device=dev).bfloat16() / 512 ** 0.5
for _ in range(3):
    y = ops.gemm(a, w, act=sys.argv[1] if len(sys.argv) > 1 else 'geglu', out_dtype=torch.bfloat16)
torch.cuda.synchronize()
print(float(y.float().abs().mean()))
