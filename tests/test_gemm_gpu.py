"""tcgen05 GEMM / implicit-GEMM conv vs a CPU fp32 evaluation on the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _gemm(a, w, **kw):
    from nuwa_pytorch_b200 import ops
    return ops.gemm(a, w, **kw)


@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (256, 128, 128, 128), (384, 512, 512, 256),
                                      (1000, 1536, 512, 0), (77, 200, 136, 0), (4096, 2752, 512, 0),
                                      (2048, 512, 1376, 0), (20480, 512, 512, 0)])
def test_gemm_plain(cuda_device, M, N, K, bn):
    g = torch.Generator().manual_seed(M * 7 + N)
    a = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, generator=g)
    ref = a.float() @ w.float().t() + bias
    out = _gemm(a.to(cuda_device), w.to(cuda_device), bias=bias.to(cuda_device), out_dtype=torch.float32, force_bn=bn)
    torch.cuda.synchronize()
    assert _rel(out.cpu(), ref) < 1e-5  # fp32 accumulate of identical bf16 operands: order-of-sum error only


def test_gemm_epilogues(cuda_device):
    from nuwa_pytorch_b200 import ops
    g = torch.Generator().manual_seed(1)
    M, K, inner = 300, 192, 80
    a = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(2 * inner, K, generator=g) / K ** 0.5).bfloat16()
    res = torch.randn(M, inner, generator=g)
    h = a.float() @ w.float().t()
    val, gate = h[:, :inner], h[:, inner:]
    wp = ops.pack_pairs(w.to(cuda_device))  # (ceil(inner/16)*32, K)
    for act, fn in (("geglu", lambda v, g_: v * F.gelu(g_)), ("glu", lambda v, g_: v * torch.sigmoid(g_))):
        out = ops.gemm(a.to(cuda_device), wp, act=act, residual=res.to(cuda_device), out_dtype=torch.float32)
        ref = fn(val, gate) + res
        assert _rel(out.cpu()[:, :inner], ref) < 2e-5, act
    # leaky + bf16 out
    w2 = (torch.randn(96, K, generator=g) / K ** 0.5).bfloat16()
    out = ops.gemm(a.to(cuda_device), w2.to(cuda_device), act="leaky", out_dtype=torch.bfloat16)
    ref = F.leaky_relu(a.float() @ w2.float().t(), 0.1)
    assert _rel(out.float().cpu(), ref) < 4e-3  # one bf16 rounding of the output


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,s", [(4, 8, 8, 64, 64, 3, 1), (2, 16, 16, 128, 256, 3, 1),
                                                (1, 32, 32, 64, 128, 3, 1), (3, 16, 16, 64, 128, 4, 2),
                                                (2, 64, 64, 64, 64, 4, 2), (2, 16, 16, 256, 64, 1, 1),
                                                (1, 256, 256, 64, 64, 3, 1), (5, 16, 16, 48, 80, 3, 1),
                                                (2, 24, 24, 64, 64, 3, 1)])
def test_conv_nhwc(cuda_device, B, H, W, Cin, Cout, k, s):
    from nuwa_pytorch_b200 import ops
    g = torch.Generator().manual_seed(B * 100 + H + Cin + k)
    x = torch.randn(B, Cin, H, W, generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).bfloat16()
    bias = torch.randn(Cout, generator=g)
    pad = {3: 1, 1: 0, 4: 1}[k]
    ref = F.conv2d(x.float(), w.float(), bias, stride=s, padding=pad)  # NCHW fp32 on CPU
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(cuda_device)
    wp = ops.pack_conv_weight(w.to(cuda_device))
    out = ops.conv2d_nhwc(x_nhwc, wp, Cin=Cin, ksize=k, stride=s, bias=bias.to(cuda_device),
                          out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert out.shape == (B, ref.shape[2], ref.shape[3], Cout)
    assert _rel(out.permute(0, 3, 1, 2).cpu(), ref) < 1e-5
