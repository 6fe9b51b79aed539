"""Where one CTA of the tcgen05 Sparse3DNA kernel spends its time: clock64 stamps of CTA 0's first tile (the heaviest
frame), lane 0 of the first warp of each warpgroup."""
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
stamps = torch.zeros(512, dtype=torch.int64, device=dev)
L = _lib.lib()
L.nuwa_debug_umma_stamps.argtypes = [ctypes.c_void_p]
L.nuwa_debug_umma_stamps.restype = None
for dil in (1, 4):
    for it in range(3):
        L.nuwa_debug_umma_stamps(stamps.data_ptr() if it == 2 else None)
        ops.attn_sparse3dna(qkv, o, B=B, nq=n, t0=0, npos=n, H=H, dh=dh, talk=talk, fmap=16, max_frames=10, nv=nv,
                            kernel=(5, 3, 3), dilation=(dil,) * 3, causal=True, variant='umma')
    torch.cuda.synchronize()
    L.nuwa_debug_umma_stamps(None)
    s = stamps.cpu().tolist()
    for w in (0, 1):
        z = s[32 * w:32 * w + 11]
        t0 = z[0]
        print(f'dilation {dil} warpgroup {w}: phase-1 head ends {[x - t0 for x in z[1:5]]}  barrier {z[5] - t0}  mix end {z[6] - t0}'
              f'  phase-2 head ends {[x - t0 for x in z[7:11]]}  (cycles after the tile-ready arrive)')
    t0 = s[0]
    print('  phase 1, issuer 0 (unit: at, +sempty, +full, +issue):',
          [(s[320 + 4 * i] - t0, s[321 + 4 * i] - s[320 + 4 * i], s[322 + 4 * i] - s[321 + 4 * i], s[323 + 4 * i] - s[322 + 4 * i])
           for i in range(1, 13)])
    print('  phase 1, warpgroup 0 (unit: at, +sfull wait, +ld/arrive, +store):',
          [(s[416 + 4 * i] - t0, s[417 + 4 * i] - s[416 + 4 * i], s[418 + 4 * i] - s[417 + 4 * i], s[419 + 4 * i] - s[418 + 4 * i])
           for i in range(1, 13)])
    print('  phase 2, warpgroup 0 (unit: window ready at, +aempty wait, +st/arrive):',
          [(s[64 + 4 * i] - t0, s[65 + 4 * i] - s[64 + 4 * i], s[66 + 4 * i] - s[65 + 4 * i]) for i in range(12)])
    print('  phase 2, issuer 0 (unit: at, +full, +afull, +issue):',
          [(s[192 + 4 * i] - t0, s[193 + 4 * i] - s[192 + 4 * i], s[194 + 4 * i] - s[193 + 4 * i], s[195 + 4 * i] - s[194 + 4 * i]) for i in range(12)])
