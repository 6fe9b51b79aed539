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


# ------------------------------------------------------------------------------------------------
# norms
# ------------------------------------------------------------------------------------------------
def sandwich_ln(B, nt, D, *, y=None, post=None, res_in=None, x_out=None, x_out_bf16=None, pre=None, a_out=None,
                a_bs=0, a_rs=0, a_t0=0, a_npos=0, shift=False, fmap=0, t0=0, t_dev=None, gather=False, shift_cache=None,
                sc_bs=0):
    """See nuwa_sandwich_ln in include/nuwa_b200.h.  post / pre: (weight, bias) fp32 tensors or None.
    t_dev: int32 device scalar holding the position (graph-replayed decode); gather: decode form of the shift."""
    p = _lib.LnParams()
    p.t0_ptr, p.gather, p.shift_cache, p.sc_bs = ptr(t_dev), int(bool(gather)), ptr(shift_cache), sc_bs
    p.y, p.res_in, p.x_out, p.x_out_bf16 = ptr(y), ptr(res_in), ptr(x_out), ptr(x_out_bf16)
    p.post_w, p.post_b = (ptr(post[0]), ptr(post[1])) if post is not None else (None, None)
    p.pre_w, p.pre_b = (ptr(pre[0]), ptr(pre[1])) if pre is not None else (None, None)
    p.a_out, p.a_bs, p.a_rs, p.a_t0, p.a_npos = ptr(a_out), a_bs, a_rs, a_t0, a_npos
    p.shift, p.fmap, p.t0, p.B, p.nt, p.D, p.eps = int(bool(shift)), fmap, t0, B, nt, D, 1e-5
    check(lib().nuwa_sandwich_ln(p, stream()), "nuwa_sandwich_ln")


def stable_ln(a, w, b, b2=None, want_f32=True, want_bf16=False):
    rows, D = a.numel() // a.shape[-1], a.shape[-1]
    o32 = torch.empty_like(a) if want_f32 else None
    o16 = torch.empty(a.shape, dtype=torch.bfloat16, device=a.device) if want_bf16 else None
    check(lib().nuwa_stable_ln(ptr(a), ptr(b2), ptr(w), ptr(b), ptr(o32), ptr(o16), rows, D, stream()),
          "nuwa_stable_ln")
    return o32, o16


# ------------------------------------------------------------------------------------------------
# attention cores
# ------------------------------------------------------------------------------------------------
def _attn_base(q, k, v, o, B, nq, t0, H, dh, q_bs, k_bs, v_bs, o_bs, q_rs, k_rs, v_rs, o_rs, talk):
    p = _lib.AttnParams()
    p.q, p.k, p.v, p.o = q, k, v, o
    p.q_bs, p.k_bs, p.v_bs, p.o_bs = q_bs, k_bs, v_bs, o_bs
    p.q_rs, p.k_rs, p.v_rs, p.o_rs = q_rs, k_rs, v_rs, o_rs
    p.B, p.nq, p.t0, p.H, p.dh = B, nq, t0, H, dh
    p.qscale = dh ** -0.5
    p.talk = ptr(talk)
    return p


# Which Sparse3DNA kernel 'auto' tries first (each tensor-core kernel declines calls outside its envelope, the gather
# kernel takes everything).  Measured at the cfg-3 / cfg-5 shapes (profiles/r02_attn3dna_perf*.json):
#   causal full pass, B=8 x 2560 tokens:  halo (mma.sync) 61 / 57 / 49 us, umma (tcgen05) 66 / 64 / 59 us, gather 233 / 199 / 158
#   centred (sketch encoder), 768 tokens:  B=32: umma 86 / 72 / 59 us vs gather 229 / 184 / 160;  B=4: 49 / 41 / 35 us both
#   (24 tiles of 128 queries cannot fill 148 SMs; the gather kernel spreads over every SM)
def _prefer_3dna(causal, B, nv):
    if causal:
        return ('halo', 'umma')
    tiles = B * ((nv + 255) // 256) * 2
    return ('umma',) if tiles >= 64 else ()


def attn_sparse3dna(qkv, o, *, B, nq, t0, npos, H, dh, talk, fmap, max_frames, nv, kernel, dilation, causal,
                    o_bs=None, variant='auto'):
    """qkv: bf16 buffer (B, npos, 3*H*dh) holding q|k|v rows for positions [0, npos); queries are positions
    [t0, t0+nq).  o: bf16 (B, nq, H*dh).  nv = number of video tokens present (positions 1..nv).
    variant: 'auto' = the fastest measured kernel whose envelope holds the call (_prefer_3dna): the halo-tiled mma.sync
    kernel (attention_3dna_halo.cu) for causal full passes, the tcgen05 / TMEM kernel (attention_3dna_umma.cu) for
    centred windows with enough tiles to fill the chip and for causal shapes the halo kernel declines, else the gather
    kernel; 'umma' / 'halo' / 'gather' pin one (a pinned kernel outside its envelope raises)."""
    inner = H * dh
    esz = 2
    base = qkv.data_ptr()
    p = _attn_base(base + t0 * 3 * inner * esz, base + inner * esz, base + 2 * inner * esz, ptr(o), B, nq, t0, H, dh,
                   npos * 3 * inner, npos * 3 * inner, npos * 3 * inner, o_bs if o_bs is not None else nq * inner,
                   3 * inner, 3 * inner, 3 * inner, inner, talk)
    p.fmap, p.max_frames, p.nv = fmap, max_frames, nv
    p.kt, p.kh, p.kw = kernel
    p.dt, p.dh_, p.dw = dilation
    p.causal = int(bool(causal))
    p.jmax = 1 + kernel[0] * kernel[1] * kernel[2]
    order = {'auto': _prefer_3dna(causal, B, nv), 'umma': ('umma',), 'halo': ('halo',), 'gather': ()}[variant]
    for name in order:
        fn = lib().nuwa_attn_sparse3dna_umma if name == 'umma' else lib().nuwa_attn_sparse3dna_halo
        rc = fn(p, stream())
        if rc == 0:
            return
        if variant == name or rc != _lib.NUWA_ERR_INVALID:
            check(rc, "nuwa_attn_sparse3dna_" + name)
    check(lib().nuwa_attn_sparse3dna(p, None, stream()), "nuwa_attn_sparse3dna")


def attn_sparse3dna_decode(q_row, cache, o, t_dev, *, B, npos, H, dh, talk, fmap, max_frames, kernel, dilation, causal):
    """Graph-replayable decode step: q_row (B, 3*H*dh) bf16 holds the new token's q|k|v (its k|v must already be in
    `cache` (B, npos, 3*H*dh)); the position is read from the int32 device scalar t_dev."""
    inner = H * dh
    base = cache.data_ptr()
    p = _attn_base(q_row.data_ptr(), base + inner * 2, base + 2 * inner * 2, ptr(o), B, 1, 0, H, dh, 3 * inner,
                   npos * 3 * inner, npos * 3 * inner, inner, 3 * inner, 3 * inner, 3 * inner, inner, talk)
    p.t0_ptr = ptr(t_dev)
    p.fmap, p.max_frames, p.nv = fmap, max_frames, 0
    p.kt, p.kh, p.kw = kernel
    p.dt, p.dh_, p.dw = dilation
    p.causal = int(bool(causal))
    p.jmax = 1 + kernel[0] * kernel[1] * kernel[2]
    check(lib().nuwa_attn_sparse3dna(p, None, stream()), "nuwa_attn_sparse3dna")


def cache_append(row, cache, t_dev):
    """cache[b, *t_dev, :] = row[b, :]  (bf16)."""
    B, npos, width = cache.shape
    check(lib().nuwa_cache_append(ptr(row), ptr(cache), npos * width, width, B, ptr(t_dev), stream()), "nuwa_cache_append")


def step_increment(t_dev):
    check(lib().nuwa_step_increment(ptr(t_dev), stream()), "nuwa_step_increment")


def sample_topk_gumbel_at(cond, uncond, noise_all, out_seq, t_dev, k, cond_scale, temperature):
    """noise_all: (steps, B, V) fp32; out_seq: (B, steps) int64; the step index is read from t_dev on the device."""
    B, V = cond.shape
    check(lib().nuwa_sample_topk_gumbel_at(ptr(cond), ptr(uncond), ptr(noise_all), ptr(out_seq), out_seq.stride(0),
                                           ptr(t_dev), B, V, k, float(cond_scale), float(temperature), stream()),
          "nuwa_sample_topk_gumbel_at")


def attn_dense(q_ptr, k_ptr, v_ptr, o, *, B, nq, nk, H, dh, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs, talk=None,
               null_k=None, null_v=None, key_mask=None, head_scale=None, bias=None, qscale=None, t0=0, use_mma=True,
               variant='auto'):
    p = _attn_base(q_ptr, k_ptr, v_ptr, ptr(o) if torch.is_tensor(o) else o, B, nq, t0, H, dh, q_bs, k_bs, v_bs, o_bs,
                   q_rs, k_rs, v_rs, o_rs, talk)
    if qscale is not None:
        p.qscale = qscale
    p.null_k, p.null_v = ptr(null_k), ptr(null_v)
    if key_mask is not None:
        assert key_mask.dtype == torch.uint8 and key_mask.is_contiguous()
        p.key_mask, p.mask_bs = ptr(key_mask), key_mask.shape[1]
    p.head_scale = ptr(head_scale)
    if bias is not None:
        p.bias, p.bias_nq, p.bias_nk = ptr(bias), bias.shape[1], bias.shape[2]
    p.jmax = nk + (1 if null_k is not None else 0)
    if variant == 'auto' and nq == 1 and talk is None and dh == 64 and head_scale is None and bias is None and t0 == 0:
        # one query per sample without talking heads (the bos query of SparseCross2DNA): one CTA per (sample, head)
        rc = lib().nuwa_attn_dense_q1(p, nk, stream())
        if rc == 0:
            return
        if rc != _lib.NUWA_ERR_INVALID:
            check(rc, "nuwa_attn_dense_q1")
    # variant: 'auto' = probability-resident kernel (attention_dense_pres.cu) inside its envelope (8 x 64 heads, <= 256
    # keys, >= 16 queries), else the library's own choice (attention_x64.cu / attention_mma.cu / generic); 'pres' pins it.
    # 'generic' = none of the specialised kernels (A/B tests)
    if variant == 'pres' or (variant == 'auto' and use_mma and nq >= 16):
        rc = lib().nuwa_attn_dense_pres(p, stream())
        if rc == 0:
            return
        if variant == 'pres' or rc != _lib.NUWA_ERR_INVALID:
            check(rc, "nuwa_attn_dense_pres")
    ws = None
    if use_mma and nq >= 8 and nk <= 256 and dh in (32, 64) and H <= 8:
        dev = o.device if torch.is_tensor(o) else torch.device('cuda')
        ws = torch.empty(B * H * dh * _round_up(nk, 64), dtype=torch.bfloat16, device=dev)
    check(lib().nuwa_attn_dense(p, ptr(ws), stream()), "nuwa_attn_dense")


def attn_cross2dna(q_ptr, k_ptr, v_ptr, o_ptr, *, B, nq, t0, H, dh, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs,
                   talk, null_k, null_v, key_mask, fmap, frames, ck, cdil, variant='auto'):
    """SparseCross2DNA non-bos queries.  variant 'auto': the tcgen05 / TMEM kernel when the call is inside its envelope and
    has enough 128-query tiles to occupy the machine, else the gather kernel; 'umma' / 'gather' force one (tests, tools)."""
    p = _attn_base(q_ptr, k_ptr, v_ptr, o_ptr, B, nq, t0, H, dh, q_bs, k_bs, v_bs, o_bs, q_rs, k_rs, v_rs, o_rs, talk)
    p.null_k, p.null_v = ptr(null_k), ptr(null_v)
    if key_mask is not None:
        assert key_mask.dtype == torch.uint8 and key_mask.is_contiguous()
        p.key_mask, p.mask_bs = ptr(key_mask), key_mask.shape[1]
    p.fmap, p.ck, p.cdil = fmap, ck, cdil
    p.jmax = 1 + frames * ck * ck
    if variant == 'umma' or (variant == 'auto' and nq > 1 and B * ((nq + 255) // 256) * 8 >= 64):
        rc = lib().nuwa_attn_cross2dna_umma(p, stream())
        if rc != _lib.NUWA_ERR_INVALID or variant == 'umma':
            check(rc, "nuwa_attn_cross2dna_umma")
            return
    check(lib().nuwa_attn_cross2dna(p, stream()), "nuwa_attn_cross2dna")


# ------------------------------------------------------------------------------------------------
# token level
# ------------------------------------------------------------------------------------------------
def embed_tokens(idx, table, *, nt, t0=0, bos=None, axials=(None, None, None), dims=(1, 1, 1), t_dev=None):
    """out[b, tl] for absolute position t0+tl (see nuwa_embed_tokens).  idx: int64 (B, n_idx) contiguous."""
    B = idx.shape[0]
    D = table.shape[1]
    out = torch.empty(B, nt, D, dtype=torch.float32, device=table.device)
    p = _lib.EmbedParams()
    p.out, p.idx, p.idx_bs, p.table, p.bos = ptr(out), ptr(idx) if idx.numel() else None, idx.stride(0), ptr(table), ptr(bos)
    p.ax1, p.ax2, p.ax3 = (ptr(a) for a in axials)
    p.d2, p.d3 = dims[1], dims[2]
    p.has_bos, p.t0, p.B, p.nt, p.D = int(bos is not None), t0, B, nt, D
    p.t0_ptr = ptr(t_dev)
    check(lib().nuwa_embed_tokens(p, stream()), "nuwa_embed_tokens")
    return out


def rotary_to_bf16(qkv_f32, inv_freq, n, H, dh, rot):
    out = torch.empty(qkv_f32.shape, dtype=torch.bfloat16, device=qkv_f32.device)
    rows = qkv_f32.shape[0]
    check(lib().nuwa_rotary_to_bf16(ptr(qkv_f32), ptr(out), ptr(inv_freq), rows, n, H, dh, rot, stream()),
          "nuwa_rotary_to_bf16")
    return out


def cross_entropy_mean(logits, target):
    rows, V = logits.shape
    ws = torch.empty(rows, dtype=torch.float32, device=logits.device)
    out = torch.empty((), dtype=torch.float32, device=logits.device)
    check(lib().nuwa_cross_entropy_mean(ptr(logits), logits.stride(0), ptr(target), ptr(ws), ptr(out), rows, V,
                                        stream()), "nuwa_cross_entropy_mean")
    return out


def sample_topk_gumbel(cond, uncond, noise, k, cond_scale, temperature, want_guided=False):
    B, V = cond.shape
    out = torch.empty(B, dtype=torch.int64, device=cond.device)
    guided = torch.empty_like(cond) if want_guided else None
    check(lib().nuwa_sample_topk_gumbel(ptr(cond), ptr(uncond), ptr(noise), ptr(out), ptr(guided), B, V, k,
                                        float(cond_scale), float(temperature), stream()), "nuwa_sample_topk_gumbel")
    return (out, guided) if want_guided else out


# ------------------------------------------------------------------------------------------------
# VAE support
# ------------------------------------------------------------------------------------------------
def nchw_to_nhwc_bf16(x):
    B, C, H, W = x.shape
    out = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=x.device)
    check(lib().nuwa_nchw_f32_to_nhwc_bf16(ptr(x.contiguous().float()), ptr(out), B, C, H, W, stream()),
          "nuwa_nchw_f32_to_nhwc_bf16")
    return out


def nhwc_to_nchw_f32(x):
    B, H, W, C = x.shape
    out = torch.empty(B, C, H, W, dtype=torch.float32, device=x.device)
    check(lib().nuwa_nhwc_to_nchw_f32(ptr(x), int(x.dtype == torch.bfloat16), ptr(out), B, C, H, W, stream()),
          "nuwa_nhwc_to_nchw_f32")
    return out


def im2col(img, ks, kpad):
    B, C, H, W = img.shape
    out = torch.empty(B * H * W, kpad, dtype=torch.bfloat16, device=img.device)
    check(lib().nuwa_im2col_nchw_f32(ptr(img), ptr(out), B, C, H, W, ks, kpad, stream()), "nuwa_im2col_nchw_f32")
    return out


def pack_im2col_weight(w):
    """(Cout, C, KS, KS) -> bf16 (Cout, Kpad) with K index (kh*KS+kw)*C + c, Kpad = roundup(KS*KS*C, 64)."""
    cout, c, ks, _ = w.shape
    k = ks * ks * c
    kp = _round_up(k, 64)
    wt = w.permute(0, 2, 3, 1).reshape(cout, k)
    if kp != k:
        wt = torch.cat([wt, torch.zeros(cout, kp - k, dtype=w.dtype, device=w.device)], dim=1)
    return wt.to(torch.bfloat16).contiguous()


def groupnorm_nhwc(x, w, b, groups, leaky=False, want_bf16=True, want_f32=False):
    B, H, W, C = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    stats = torch.empty(B * groups * 2, dtype=torch.float32, device=x.device)
    o16 = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    o32 = torch.empty_like(x) if want_f32 else None
    check(lib().nuwa_groupnorm_nhwc(ptr(x), ptr(w), ptr(b), ptr(stats), ptr(o16), ptr(o32), B, H * W, C, groups,
                                    int(leaky), stream()), "nuwa_groupnorm_nhwc")
    return o16 if not want_f32 else (o16, o32)


def upsample2x(x):
    B, H, W, C = x.shape
    out = torch.empty(B, 2 * H, 2 * W, C, dtype=torch.bfloat16, device=x.device)
    check(lib().nuwa_upsample2x_nhwc_bf16(ptr(x), ptr(out), B, H, W, C, stream()), "nuwa_upsample2x_nhwc_bf16")
    return out


def vae_attn_prep(qkv_f32, B, n, inner):
    out = torch.empty(qkv_f32.shape, dtype=torch.bfloat16, device=qkv_f32.device)
    check(lib().nuwa_vae_attn_prep(ptr(qkv_f32), ptr(out), B, n, inner, stream()), "nuwa_vae_attn_prep")
    return out


def vq_argmax(x, code, code_sq=None, cosine=True, code16=None, emax=None, variant='auto'):
    """arg-max codebook index per row of x (M, D) fp32.  variant 'auto': the tensor-core path (bf16 similarity GEMM +
    exact fp32 re-score of the codes inside the bf16 error band: csrc/vae_ops.cu vq_argmax_tc) when a bf16 codebook copy
    `code16` and `emax` = max |code row| (a 1-element fp32 DEVICE tensor) are supplied and the problem is large enough to pay for the score matrix;
    'fp32' pins the CUDA-core kernel, 'tc' the tensor-core path."""
    M, D = x.shape
    Kc = code.shape[0]
    out = torch.empty(M, dtype=torch.int64, device=x.device)
    want_tc = variant == 'tc' or (variant == 'auto' and code16 is not None and emax is not None and M * Kc >= (1 << 20))
    if want_tc:
        if code16 is None:
            code16 = code.to(torch.bfloat16)
        if emax is None:
            emax = code.norm(dim=-1).max().reshape(1).float()
        nbytes = int(lib().nuwa_vq_argmax_tc_workspace(M, Kc, D))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        rc = lib().nuwa_vq_argmax_tc(ptr(x), ptr(code), ptr(code_sq), ptr(code16), ptr(emax), ptr(out), M, Kc, D,
                                     int(cosine), ptr(ws), nbytes, stream())
        if rc == 0:
            return out
        if variant == 'tc' or rc != _lib.NUWA_ERR_INVALID:
            check(rc, "nuwa_vq_argmax_tc")
    check(lib().nuwa_vq_argmax(ptr(x), ptr(code), ptr(code_sq), ptr(out), M, Kc, D, int(cosine), stream()),
          "nuwa_vq_argmax")
    return out


def split3(x):
    """fp32 (rows, K) -> bf16 (rows, 3K) = [x0 | x1 | x2] with x == x0 + x1 + x2 exactly (see nuwa_split3_f32_bf16)."""
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    rows, K = x.shape
    out = torch.empty(rows, 3 * K, dtype=torch.bfloat16, device=x.device)
    check(lib().nuwa_split3_f32_bf16(ptr(x), x.stride(0), ptr(out), rows, K, stream()), "nuwa_split3_f32_bf16")
    return out


def linear_f32x3(x, w3, bias=None, also_bf16=False):
    """fp32-faithful x @ W.T + bias on the tensor cores; w3 = split3(W fp32 (N, K)).  x: fp32 (M, K)."""
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1 and w3.dtype == torch.bfloat16
    M, K = x.shape
    N = w3.shape[0]
    assert w3.shape[1] == 3 * K
    out = torch.empty(M, N, dtype=torch.float32, device=x.device)
    nbytes = int(lib().nuwa_linear_f32x3_workspace(M, K))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    o16 = torch.empty(M, N, dtype=torch.bfloat16, device=x.device) if also_bf16 else None
    check(lib().nuwa_linear_f32x3(ptr(x), x.stride(0), ptr(w3), M, N, K, ptr(bias), ptr(out), ptr(o16), N, ptr(ws), nbytes,
                                  stream()), "nuwa_linear_f32x3")
    return (out, o16) if also_bf16 else out


def gather_rows(table, idx, want_bf16=True, want_f32=False):
    M, D = idx.numel(), table.shape[1]
    o16 = torch.empty(M, D, dtype=torch.bfloat16, device=table.device) if want_bf16 else None
    o32 = torch.empty(M, D, dtype=torch.float32, device=table.device) if want_f32 else None
    check(lib().nuwa_gather_rows(ptr(table), ptr(idx.contiguous()), ptr(o16), ptr(o32), M, D, stream()),
          "nuwa_gather_rows")
    if want_bf16 and want_f32:
        return o16, o32
    return o16 if want_bf16 else o32


def conv1x1_to_nchw(x, w, b):
    B, H, W, C = x.shape
    cout = w.shape[0]
    out = torch.empty(B, cout, H, W, dtype=torch.float32, device=x.device)
    check(lib().nuwa_conv1x1_nhwc_to_nchw(ptr(x), ptr(w), ptr(b), ptr(out), B, H * W, C, cout, stream()),
          "nuwa_conv1x1_nhwc_to_nchw")
    return out


def recon_loss(a, b, l2=False):
    """mean |a-b| (F.l1_loss) or mean (a-b)^2 (F.mse_loss) of two fp32 tensors -> 0-d fp32 device tensor."""
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.shape == b.shape
    a, b = a.contiguous(), b.contiguous()
    nparts = 4 * 148
    partials = torch.empty(nparts, dtype=torch.float32, device=a.device)
    out = torch.empty((), dtype=torch.float32, device=a.device)
    check(lib().nuwa_recon_loss_f32(ptr(a), ptr(b), a.numel(), int(bool(l2)), ptr(partials), nparts, ptr(out), stream()),
          "nuwa_recon_loss_f32")
    return out
