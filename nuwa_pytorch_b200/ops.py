"""Thin Python wrappers over the C-ABI: allocate outputs with torch, pass raw pointers, check status.

Weight "packing" helpers here are one-off layout transforms done when weights are loaded (plumbing);
all arithmetic of the hot path happens inside libnuwa_b200.so.
"""
import torch

from . import _lib
from ._lib import check, lib, ptr, stream

ACT = {None: 0, "none": 0, "leaky": 1, "glu": 2, "geglu": 3}


def _round_up(x, m):
    return (x + m - 1) // m * m


# ------------------------------------------------------------------------------------------------
# weight packing
# ------------------------------------------------------------------------------------------------
def pack_pairs(w):
    """(2*inner, ...) [value rows | gate rows] -> pair-packed (2*roundup(inner,16), ...): every block of 32 rows
    holds 16 value rows followed by their 16 gate rows (zero rows pad the tail)."""
    two_inner = w.shape[0]
    assert two_inner % 2 == 0
    inner = two_inner // 2
    ip = _round_up(inner, 16)
    val, gate = w[:inner], w[inner:]
    pad_shape = (ip - inner,) + tuple(w.shape[1:])
    if ip != inner:
        z = torch.zeros(pad_shape, dtype=w.dtype, device=w.device)
        val, gate = torch.cat([val, z]), torch.cat([gate, z])
    val = val.reshape(ip // 16, 16, *w.shape[1:])
    gate = gate.reshape(ip // 16, 16, *w.shape[1:])
    return torch.stack([val, gate], dim=1).reshape(2 * ip, *w.shape[1:]).contiguous()


def pack_conv_weight(w, pairs=False):
    """(Cout, Cin, KH, KW) -> bf16 (Cout', KH*KW*Cin_pad), K ordered (tap, channel), Cin zero padded to 64."""
    cout, cin, kh, kw = w.shape
    cp = _round_up(cin, 64)
    wt = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
    if cp != cin:
        wt = torch.cat([wt, torch.zeros(cout, kh * kw, cp - cin, dtype=w.dtype, device=w.device)], dim=2)
    wt = wt.reshape(cout, kh * kw * cp)
    if pairs:
        wt = pack_pairs(wt)
    return wt.to(torch.bfloat16).contiguous()


# ------------------------------------------------------------------------------------------------
# GEMM / conv
# ------------------------------------------------------------------------------------------------
def gemm(a, w, bias=None, residual=None, act=None, out_dtype=torch.float32, out=None, force_bn=0, also_bf16=False):
    """out = act(a @ w.T + bias) + residual.  a: (M,K) bf16 (row stride may exceed K), w: (N,K) bf16."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 2 and w.dim() == 2
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K or (w.shape[1] >= K)
    n_out = N // 2 if ACT[act] >= 2 else N
    out_f32 = out_bf16 = None
    if out is None:
        out = torch.empty(M, n_out, dtype=out_dtype, device=a.device)
    assert out.stride(1) == 1
    if out.dtype == torch.float32:
        out_f32 = out
    else:
        out_bf16 = out
    extra = None
    if also_bf16:
        assert out_f32 is not None
        extra = torch.empty(M, n_out, dtype=torch.bfloat16, device=a.device)
        assert extra.stride(0) == out.stride(0)
        out_bf16 = extra
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= N
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.stride(1) == 1
    code = lib().nuwa_gemm_bf16(ptr(a), a.stride(0), ptr(w), w.stride(0), M, N, K, ptr(bias), ptr(residual),
                                residual.stride(0) if residual is not None else 0, ptr(out_f32), ptr(out_bf16),
                                out.stride(0), ACT[act], force_bn, stream())
    check(code, "nuwa_gemm_bf16")
    return (out, extra) if also_bf16 else out


def conv2d_nhwc(x, wp, Cin, ksize, stride=1, bias=None, residual=None, act=None, out_dtype=torch.bfloat16,
                force_bn=0, also_bf16=False):
    """x: (B,H,W,Cin) bf16 contiguous; wp: packed weights from pack_conv_weight.  Returns NHWC."""
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 4 and x.shape[3] == Cin
    B, Hin, Win, _ = x.shape
    cout = wp.shape[0]
    H, W = (Hin // 2, Win // 2) if stride == 2 else (Hin, Win)
    n_out = cout // 2 if ACT[act] >= 2 else cout
    out = torch.empty(B, H, W, n_out, dtype=out_dtype, device=x.device)
    out_f32 = out if out_dtype == torch.float32 else None
    out_bf16 = out if out_dtype == torch.bfloat16 else None
    extra = None
    if also_bf16:
        extra = torch.empty(B, H, W, n_out, dtype=torch.bfloat16, device=x.device)
        out_bf16 = extra
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.is_contiguous()
    code = lib().nuwa_conv2d_nhwc_bf16(ptr(x), ptr(wp), B, Hin, Win, Cin, cout, ksize, stride, ptr(bias),
                                       ptr(residual), ptr(out_f32), ptr(out_bf16), ACT[act], force_bn, stream())
    check(code, "nuwa_conv2d_nhwc_bf16")
    return (out, extra) if also_bf16 else out
