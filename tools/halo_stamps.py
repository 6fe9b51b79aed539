"""Where one CTA of the halo Sparse3DNA kernel spends its time: clock64 stamps of CTA 0 (heaviest frame), warp 0."""
import ctypes
import sys

import torch

sys.path.insert(0, '.')
from nuwa_pytorch_b200 import _lib, ops  # noqa: E402

dev = torch.device('cuda')
B, H, dh, nv = 8, 8, 64, 2559
inner, n = H * dh, nv + 1
qkv = torch.randn(B, n, 3 * inner, device=dev).bfloat16()
talk = torch.randn(H, H, device=dev) / 2
o = torch.empty(B, n, inner, dtype=torch.bfloat16, device=dev)
stamps = torch.zeros(24, dtype=torch.int64, device=dev)
L = _lib.lib()
L.nuwa_debug_halo_stamps.argtypes = [ctypes.c_void_p]
L.nuwa_debug_halo_stamps.restype = None
for dil in (1, 4):
    for it in range(3):
        L.nuwa_debug_halo_stamps(stamps.data_ptr() if it == 2 else None)
        ops.attn_sparse3dna(qkv, o, B=B, nq=n, t0=0, npos=n, H=H, dh=dh, talk=talk, fmap=16, max_frames=10, nv=nv,
                            kernel=(5, 3, 3), dilation=(dil,) * 3, causal=True, variant='halo')
    torch.cuda.synchronize()
    L.nuwa_debug_halo_stamps(None)
    s = stamps.cpu().tolist()
    t0 = s[0]
    print(f'dilation {dil}: steps {s[22]}  init {s[1] - t0}  phase-1 head ends {[x - t0 for x in s[2:10]]}  mix end {s[10] - t0}'
          f'  phase-3 head ends {[x - t0 for x in s[11:19]]}  end {s[19] - t0}  (cycles)')
    print(f'   warp 0 waited for data: phase 1 {s[20]} cycles, phase 3 {s[21]} cycles; producer waited for free stages {s[23]} cycles')
