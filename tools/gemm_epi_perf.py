"""Epilogue cost of the decoder-shaped GEMMs: GEGLU vs GLU vs plain at the FF1 shape, N tile choice at N = 512."""
import sys

import torch

sys.path.insert(0, '.')
from nuwa_pytorch_b200 import ops  # noqa: E402

dev = torch.device('cuda')
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=7, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


M = 20480
for (N, K, act, bn, od) in [(2752, 512, 'geglu', 0, torch.bfloat16), (2752, 512, 'glu', 0, torch.bfloat16), (2752, 512, None, 0, torch.bfloat16),
                            (2752, 512, 'geglu', 128, torch.bfloat16), (2752, 512, 'geglu', 64, torch.bfloat16),
                            (512, 512, None, 0, torch.float32), (512, 512, None, 128, torch.float32), (512, 512, None, 64, torch.float32),
                            (512, 1376, None, 0, torch.float32), (512, 1376, None, 128, torch.float32),
                            (1536, 512, None, 0, torch.bfloat16), (1536, 512, None, 128, torch.bfloat16)]:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = torch.randn(N, K, device=dev).bfloat16() / K ** 0.5
    us = timeit(lambda: ops.gemm(a, w, act=act, out_dtype=od, force_bn=bn))
    print(dict(N=N, K=K, act=act, bn=bn, out=str(od)[6:], us=round(us, 1), tflops=round(2 * M * N * K / us / 1e6, 1)), flush=True)
