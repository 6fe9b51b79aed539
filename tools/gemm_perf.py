"""Micro-benchmark of the tcgen05 GEMM / implicit-GEMM conv kernel on the shapes of BASELINE configs 2 and 3.
Prints achieved TFLOP/s per shape (CUDA events, L2 flushed between iterations).  Run on the GPU box."""
import json
import sys

import torch

sys.path.insert(0, '.')
from nuwa_pytorch_b200 import ops  # noqa: E402

dev = torch.device('cuda')
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=5, warm=2):
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
    return ts[len(ts) // 2]


rows = []
for (M, N, K, act, bn) in [(20480, 512, 512, None, 0), (20480, 1536, 512, None, 0), (20480, 2752, 512, 'geglu', 0),
                           (20480, 512, 1376, None, 0), (20480, 8192, 512, None, 0), (16384, 4096, 4096, None, 256),
                           (16384, 4096, 4096, None, 128), (8192, 8192, 8192, None, 256)]:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = torch.randn(N, K, device=dev).bfloat16()
    ms = timeit(lambda: ops.gemm(a, w, act=act, out_dtype=torch.bfloat16, force_bn=bn))
    rows.append(dict(op='gemm', M=M, N=N, K=K, act=act, bn=bn, ms=round(ms, 4), tflops=round(2 * M * N * K / ms / 1e9, 1)))
    print(rows[-1], flush=True)
for (B, H, Cin, Cout, k, s, act) in [(64, 16, 4096, 4096, 3, 1, None), (64, 16, 4096, 8192, 3, 1, 'glu'),
                                     (8, 256, 512, 512, 3, 1, 'leaky'), (8, 256, 512, 512, 4, 2, 'leaky'),
                                     (16, 64, 2048, 1024, 3, 1, 'leaky'), (16, 32, 2048, 4096, 4, 2, 'leaky')]:
    x = torch.randn(B, H, H, Cin, device=dev).bfloat16()
    w = torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5
    wp = ops.pack_conv_weight(w, pairs=(act == 'glu'))
    del w
    ms = timeit(lambda: ops.conv2d_nhwc(x, wp, Cin=Cin, ksize=k, stride=s, act=act), iters=3, warm=1)
    Ho = H // s
    fl = 2 * B * Ho * Ho * Cout * Cin * k * k
    rows.append(dict(op='conv', B=B, H=H, Cin=Cin, Cout=Cout, k=k, s=s, act=act, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1)))
    print(rows[-1], flush=True)
    del x, wp
json.dump(rows, open('gpurun_out/gemm_perf.json', 'w'), indent=1)
