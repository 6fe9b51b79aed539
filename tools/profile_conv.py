"""One representative tcgen05 implicit-GEMM conv (cfg-2 ResBlock conv: 4096->4096 3x3 @16x16, batch 64) for an
`ncu --set full` capture of gemm_tcgen05_kernel."""
import sys

import torch

sys.path.insert(0, '.')
from nuwa_pytorch_b200 import ops  # noqa: E402

dev = torch.device('cuda')
x = torch.randn(64, 16, 16, 4096, device=dev).bfloat16()
w = torch.randn(4096, 4096, 3, 3, device=dev) / (4096 * 9) ** 0.5
wp = ops.pack_conv_weight(w)
del w
b = torch.randn(4096, device=dev)
for _ in range(4):
    y = ops.conv2d_nhwc(x, wp, Cin=4096, ksize=3, bias=b, out_dtype=torch.float32)
torch.cuda.synchronize()
print(float(y.float().abs().mean()))
