"""Python wrappers over the training (backward) entry points of the C-ABI (include/nuwa_b200.h, second half).

Same rules as ops.py: torch allocates, raw pointers cross the boundary, every status is checked; the arithmetic of
the backward pass happens inside libnuwa_b200.so (csrc/backward.cu, bgemm.cu, attention.cu, gemm_tcgen05.cu).
"""
import torch

from . import _lib, ops
from ._lib import check, lib, ptr, stream


def _round_up(x, m):
    return (x + m - 1) // m * m


# ------------------------------------------------------------------------------------------------
# dense layers
# ------------------------------------------------------------------------------------------------
def transpose(x):
    """bf16 (R, C) (row stride >= C) -> bf16 view (C, R) of a fresh buffer whose row pitch is a multiple of 8."""
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1
    R, C = x.shape
    ld = _round_up(R, 8)
    out = torch.empty(C, ld, dtype=torch.bfloat16, device=x.device)
    check(lib().nuwa_transpose_bf16(ptr(x), x.stride(0), ptr(out), ld, R, C, stream()), "nuwa_transpose_bf16")
    return out[:, :R]


def gemm_splitk(a, w, out, splits=64):
    """out (N_a, N_w) fp32 += a @ w.T ; a (N_a, K), w (N_w, K) bf16 with the LONG contraction K contiguous."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and out.dtype == torch.float32
    assert a.stride(1) == 1 and w.stride(1) == 1 and out.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and out.shape == (M, N)
    force_bn = 256 if N >= 256 else (128 if N >= 128 else 64)
    check(lib().nuwa_gemm_bf16_splitk(ptr(a), a.stride(0), ptr(w), w.stride(0), M, N, K, ptr(out), out.stride(0), splits,
                                      force_bn, stream()), "nuwa_gemm_bf16_splitk")


# weight gradients read dY and X contraction-major (no transposed copies); False = the transpose + K-major path, kept
# for the A/B test (tests/test_gemm_gpu.py) and measurements
WGRAD_TN = True


def gemm_splitk_tn(dy, a, out, splits=64):
    """out (N, K) fp32 += dy.T @ a ; dy (M, N), a (M, K) bf16 row-major (column slices allowed): dW = dY^T X."""
    assert dy.dtype == torch.bfloat16 and a.dtype == torch.bfloat16 and out.dtype == torch.float32
    assert dy.stride(1) == 1 and a.stride(1) == 1 and out.stride(1) == 1
    M, N = dy.shape
    K = a.shape[1]
    assert a.shape[0] == M and out.shape == (N, K)
    if not WGRAD_TN:
        return gemm_splitk(transpose(dy), transpose(a), out, splits)
    force_bn = 256 if K >= 256 else (128 if K >= 128 else 64)
    check(lib().nuwa_gemm_bf16_tn_splitk(ptr(dy), dy.stride(0), ptr(a), a.stride(0), N, K, M, ptr(out), out.stride(0),
                                         splits, force_bn, stream()), "nuwa_gemm_bf16_tn_splitk")


def linear_bwd(dy16, a16, w_t16, dW, want_da=True, da_residual=None):
    """y = a @ W.T  ->  da = dy @ W (fp32, optionally + da_residual), dW += dy.T @ a.
    dy16 (M, N) bf16, a16 (M, K) bf16, w_t16 = W.T as (K, N) bf16 contiguous, dW (N, K) fp32 accumulator (or None)."""
    da = None
    if want_da:
        da = ops.gemm(dy16, w_t16, residual=da_residual, out_dtype=torch.float32)
    if dW is not None:
        gemm_splitk_tn(dy16, a16, dW)
    return da


def bgemm(A, B, C, *, M, N, K, a_trans, b_trans, lda, ldb, ldc, batch1, batch2, a_s, b_s, c_s, alpha=1.0, accumulate=False):
    """A, B, C: (tensor-or-pointer).  a_s / b_s / c_s: (stride1, stride2) in elements."""
    p = _lib.BgemmParams()
    p.A = A if isinstance(A, int) else ptr(A)
    p.B = B if isinstance(B, int) else ptr(B)
    c_t = C[0] if isinstance(C, tuple) else C
    p.C = C[1] if isinstance(C, tuple) else ptr(C)
    p.M, p.N, p.K, p.a_trans, p.b_trans = M, N, K, int(a_trans), int(b_trans)
    p.lda, p.ldb, p.ldc = lda, ldb, ldc
    p.batch1, p.batch2 = batch1, batch2
    p.a_s1, p.a_s2 = a_s
    p.b_s1, p.b_s2 = b_s
    p.c_s1, p.c_s2 = c_s
    p.alpha = float(alpha)
    p.c_bf16 = int(c_t.dtype == torch.bfloat16)
    p.accumulate = int(bool(accumulate))
    check(lib().nuwa_bgemm(p, stream()), "nuwa_bgemm")


# ------------------------------------------------------------------------------------------------
# norms / elementwise
# ------------------------------------------------------------------------------------------------
def ln_bwd(dout, x, w, *, nt, dw=None, db=None, dcol=None, x2=None, stable=False, unshift=False, fmap=0,
           dx_bf16=False, dx_f32=None, dx2_f32=None, accumulate=False):
    """LayerNorm backward over rows of x (rows, D) fp32.  dout: fp32 or bf16 (rows, D).  Gradients of the affine
    parameters are ADDED to dw / db (and the column sum of dx to dcol).  Returns the bf16 dx when dx_bf16."""
    rows, D = x.shape
    p = _lib.LnBwdParams()
    p.rows, p.nt, p.D = rows, nt, D
    if dout.dtype == torch.float32:
        p.dout_f32 = ptr(dout)
    else:
        assert dout.dtype == torch.bfloat16
        p.dout_bf16 = ptr(dout)
    assert dout.is_contiguous() and x.is_contiguous()
    p.unshift, p.fmap = int(bool(unshift)), fmap or 0
    p.x, p.x2, p.stable, p.w, p.eps = ptr(x), ptr(x2), int(bool(stable)), ptr(w), 1e-5
    out16 = torch.empty(rows, D, dtype=torch.bfloat16, device=x.device) if dx_bf16 else None
    p.dx_bf16, p.dx_f32, p.dx2_f32, p.accumulate = ptr(out16), ptr(dx_f32), ptr(dx2_f32), int(bool(accumulate))
    p.dw, p.db, p.dcol = ptr(dw), ptr(db), ptr(dcol)
    check(lib().nuwa_ln_bwd(p, stream()), "nuwa_ln_bwd")
    return out16


def geglu_fwd(h):
    M, two_ip = h.shape
    g = torch.empty(M, two_ip // 2, dtype=torch.bfloat16, device=h.device)
    check(lib().nuwa_geglu_fwd(ptr(h), ptr(g), M, two_ip // 2, stream()), "nuwa_geglu_fwd")
    return g


def geglu_bwd(dg, h):
    M, two_ip = h.shape
    dh = torch.empty_like(h)
    check(lib().nuwa_geglu_bwd(ptr(dg), ptr(h), ptr(dh), M, two_ip // 2, stream()), "nuwa_geglu_bwd")
    return dh


def ce_bwd(logits, target, gscale=None):
    rows, V = logits.shape
    dl = torch.empty(rows, V, dtype=torch.bfloat16, device=logits.device)
    check(lib().nuwa_ce_bwd(ptr(logits), logits.stride(0), ptr(target), ptr(gscale), ptr(dl), V, rows, V, stream()),
          "nuwa_ce_bwd")
    return dl


def embed_bwd(dx, idx, dtable, *, nt, frac=1.0, dbos=None, daxials=(None, None, None), dims=(1, 1, 1)):
    B = idx.shape[0]
    D = dx.shape[-1]
    p = _lib.EmbedBwdParams()
    p.dx, p.idx, p.idx_bs, p.dtable, p.dbos = ptr(dx), ptr(idx), idx.stride(0), ptr(dtable), ptr(dbos)
    p.dax1, p.dax2, p.dax3 = (ptr(t) for t in daxials)
    p.d2, p.d3 = dims[1], dims[2]
    p.has_bos, p.B, p.nt, p.D, p.frac = int(dbos is not None), B, nt, D, float(frac)
    check(lib().nuwa_embed_bwd(p, stream()), "nuwa_embed_bwd")


def rotary_bwd_to_bf16(dqkv32, inv_freq, n, H, dh, rot):
    rows = dqkv32.shape[0]
    out = torch.empty(dqkv32.shape, dtype=torch.bfloat16, device=dqkv32.device)
    check(lib().nuwa_rotary_bwd_to_bf16(ptr(dqkv32), ptr(out), ptr(inv_freq), rows, n, H, dh, rot, stream()),
          "nuwa_rotary_bwd_to_bf16")
    return out


def add_rows(dst, src, row_map=None, cols=None, accumulate=True):
    """dst[row_map[r], :cols] (+)= src[r, :cols]  (fp32, 2-D, unit inner stride)."""
    rows = src.shape[0]
    cols = cols if cols is not None else src.shape[1]
    check(lib().nuwa_add_rows_f32(ptr(dst), dst.stride(0), ptr(src), src.stride(0), ptr(row_map), rows, cols,
                                  int(bool(accumulate)), stream()), "nuwa_add_rows_f32")


# ------------------------------------------------------------------------------------------------
# attention backward
# ------------------------------------------------------------------------------------------------
def _rows(S, dPp, talk, dtalk, B, H, nq, J, jp, out_scale, key_mask=None, has_null=0):
    Pp = torch.empty(B, H, nq, jp, dtype=torch.bfloat16, device=S.device)
    dS = torch.empty(B, H, nq, jp, dtype=torch.bfloat16, device=S.device)
    p = _lib.AttnRowsParams()
    p.S, p.dPp, p.Pp, p.dS, p.talk, p.dtalk = ptr(S), ptr(dPp), ptr(Pp), ptr(dS), ptr(talk), ptr(dtalk)
    p.B, p.H, p.nq, p.J, p.jp, p.out_scale = B, H, nq, J, jp, float(out_scale)
    if key_mask is not None:
        p.key_mask, p.mask_bs, p.has_null = ptr(key_mask), key_mask.stride(0), int(has_null)
    check(lib().nuwa_attn_bwd_rows(p, stream()), "nuwa_attn_bwd_rows")
    return Pp, dS


# scores of the Sparse3DNA backward: 'auto' = tcgen05 kernel when inside its envelope and the pass has enough 128-query
# tiles to occupy the machine, else the gather kernel; 'umma' / 'gather' force one (tests, tools)
SCORES_VARIANT = 'auto'


def attn_sparse3dna_bwd(qkv, do, *, B, n, H, dh, talk, dtalk, fmap, max_frames, kernel, dilation, causal):
    """Backward of ops.attn_sparse3dna over a full teacher-forced pass (positions 0..n-1, bos at 0).
    qkv: bf16 (B, n, 3*inner) saved by the forward; do: bf16 (B, n, inner) gradient of the attention output.
    Returns dqkv bf16 (B, n, 3*inner); adds the talking-heads gradient to dtalk (H, H) fp32."""
    inner = H * dh
    dev = qkv.device
    dqkv = torch.empty(B, n, 3 * inner, dtype=torch.bfloat16, device=dev)
    nq = n - 1
    base = qkv.data_ptr()
    J = 1 + kernel[0] * kernel[1] * kernel[2]
    jp = _round_up(J, 8)
    tmp = torch.zeros(2, B, inner, dtype=torch.float32, device=dev)
    if nq > 0:
        p = ops._attn_base(base + 3 * inner * 2, base + inner * 2, base + 2 * inner * 2, None, B, nq, 1, H, dh,
                           n * 3 * inner, n * 3 * inner, n * 3 * inner, 0, 3 * inner, 3 * inner, 3 * inner, inner, talk)
        p.fmap, p.max_frames, p.nv = fmap, max_frames, nq
        p.kt, p.kh, p.kw = kernel
        p.dt, p.dh_, p.dw = dilation
        p.causal = int(bool(causal))
        p.jmax = J
        S = torch.empty(B, H, nq, jp, dtype=torch.float32, device=dev)
        dPp = torch.empty(B, H, nq, jp, dtype=torch.float32, device=dev)
        do_q = do.data_ptr() + inner * 2  # gradient rows of the non-bos queries
        rc = _lib.NUWA_ERR_INVALID
        if SCORES_VARIANT == 'umma' or (SCORES_VARIANT == 'auto' and B * ((nq + 255) // 256) * 2 >= 64):
            rc = lib().nuwa_attn3dna_bwd_scores_umma(p, do_q, n * inner, inner, ptr(S), ptr(dPp), jp, stream())
            if rc != _lib.NUWA_ERR_INVALID or SCORES_VARIANT == 'umma':
                check(rc, "nuwa_attn3dna_bwd_scores_umma")
        if rc == _lib.NUWA_ERR_INVALID:
            check(lib().nuwa_attn3dna_bwd_scores(p, do_q, n * inner, inner, ptr(S), ptr(dPp), jp, stream()),
                  "nuwa_attn3dna_bwd_scores")
        # dS leaves the row kernel multiplied by the logit scale dh^-0.5, so dq = sum dS k and dk = sum dS q need no more
        Pp, dS = _rows(S, dPp, talk, dtalk, B, H, nq, J, jp, dh ** -0.5)
        rc = _lib.NUWA_ERR_INVALID
        if SCORES_VARIANT == 'umma' or (SCORES_VARIANT == 'auto' and B * ((nq + 255) // 256) * 2 >= 64):
            rc = lib().nuwa_attn3dna_bwd_dq_umma(p, ptr(dS), jp, dqkv.data_ptr() + 3 * inner * 2, n * 3 * inner, 3 * inner,
                                                 stream())
            if rc != _lib.NUWA_ERR_INVALID:
                check(rc, "nuwa_attn3dna_bwd_dq_umma")
        if rc == _lib.NUWA_ERR_INVALID:
            check(lib().nuwa_attn3dna_bwd_dq(p, ptr(dS), jp, dqkv.data_ptr() + 3 * inner * 2, n * 3 * inner, 3 * inner,
                                             stream()), "nuwa_attn3dna_bwd_dq")
        check(lib().nuwa_attn3dna_bwd_dkdv(p, do_q, n * inner, inner, ptr(dS), ptr(Pp), jp, dqkv.data_ptr() + inner * 2,
                                           dqkv.data_ptr() + 2 * inner * 2, n * 3 * inner, 3 * inner, stream()),
              "nuwa_attn3dna_bwd_dkdv")
        check(lib().nuwa_attn_bwd_first_key(base + 3 * inner * 2, n * 3 * inner, 3 * inner, do_q, n * inner, inner,
                                            ptr(dS), ptr(Pp), jp, B, H, dh, nq, ptr(tmp[0]), ptr(tmp[1]), inner,
                                            stream()), "nuwa_attn_bwd_first_key")
    check(lib().nuwa_attn3dna_bwd_first_key_finalize(ptr(tmp[0]), ptr(tmp[1]), ptr(do), n * inner, ptr(dqkv),
                                                     n * 3 * inner, inner, B, stream()),
          "nuwa_attn3dna_bwd_first_key_finalize")
    return dqkv


# single-query dense attention (bos query of SparseCross2DNA): dedicated kernels (csrc/attention_q1.cu); False = the generic
# chain, kept for the A/B test
Q1_KERNELS = True


def key_mask_ok(key_mask):
    return key_mask is None or (key_mask.dtype == torch.uint8 and key_mask.stride(1) == 1)


# dense attention backward: True = fused probability stage (csrc/attention_dense_bwd.cu: S and dP' recomputed per 16-query
# tile on the tensor cores, head mixes / dW on mma.sync from shared memory, nothing but P' and dS written to HBM) when the
# call is inside its envelope; False = the materialised-logits path (two batched GEMMs + the row kernel), kept for the
# A/B test.  Measured at the cfg-3 cross-attention shape (tools/dense_bwd_perf.py, whole attn_dense_bwd call per layer):
# materialised 0.62 ms, fused 0.57 ms (with the head mixes on the CUDA cores it was 0.69 ms).  Gradient error vs fp32
# autograd (materialised / fused): dq 2.4e-3 / 2.9e-3, dk|dv 2.4e-3 / 2.6e-3, dW_talk 1.6e-4 / 1.6e-3 (dP' is held as
# bf16 in shared memory).
DENSE_BWD_FUSED = True


def attn_dense_bwd(q_ptr, k_ptr, v_ptr, do, *, B, nq, nk, H, dh, q_bs, q_rs, kv_bs, kv_rs, talk, dtalk, null_k, null_v,
                   dnull_k, dnull_v, key_mask, dq_out, dq_bs, dq_rs, dk_ptr, dv_ptr, dkv_bs, dkv_rs, out_f32, side=None, keep=()):
    """Backward of ops.attn_dense (Attention core with the learned null key, key mask and talking heads).
    q/k/v: bf16 device pointers with element strides; do: bf16 (B, nq, inner) contiguous.
    dq_out: (dtype donor tensor, pointer) ; dk_ptr / dv_ptr: raw pointers of the same dtype (bf16, or fp32 when out_f32).
    Adds to dtalk (H,H), dnull_k / dnull_v (inner,) fp32."""
    inner = H * dh
    dev = do.device
    has_null = int(null_k is not None)
    if Q1_KERNELS and nq == 1 and talk is None and dh == 64 and out_f32 and key_mask_ok(key_mask):
        # one query per sample, no talking heads (the bos query of SparseCross2DNA): a single launch instead of the K/V
        # repack + five batched GEMMs + wide row kernel + split chain
        p = ops._attn_base(q_ptr, k_ptr, v_ptr, None, B, 1, 0, H, dh, q_bs, kv_bs, kv_bs, 0, q_rs, kv_rs, kv_rs, inner, None)
        p.null_k, p.null_v = ptr(null_k), ptr(null_v)
        if key_mask is not None:
            p.key_mask, p.mask_bs = ptr(key_mask), key_mask.stride(0)
        dq_ptr = dq_out[1] if isinstance(dq_out, tuple) else ptr(dq_out)
        rc = lib().nuwa_attn_dense_q1_bwd(p, nk, ptr(do), do.stride(0), dq_ptr, dq_bs, dk_ptr, dv_ptr, dkv_bs, dkv_rs,
                                          ptr(dnull_k), ptr(dnull_v), stream())
        if rc != _lib.NUWA_ERR_INVALID:
            check(rc, "nuwa_attn_dense_q1_bwd")
            return
    J = nk + has_null
    jp = _round_up(J, 8)
    scale = dh ** -0.5
    kfull = torch.empty(B, jp, inner, dtype=torch.bfloat16, device=dev)
    vfull = torch.empty(B, jp, inner, dtype=torch.bfloat16, device=dev)
    check(lib().nuwa_kv_full_build(k_ptr, v_ptr, kv_bs, kv_rs, ptr(null_k), ptr(null_v), ptr(kfull), ptr(vfull), B, nk, jp,
                                   inner, stream()), "nuwa_kv_full_build")
    sc = (H * nq * jp, nq * jp)
    full_s = (jp * inner, dh)
    Pp = dS = None
    if DENSE_BWD_FUSED and H == 8 and dh == 64 and nk <= 256 and nq >= 16:
        # probability stage fused: S and dP' are recomputed on the tensor cores inside the kernel and never reach HBM
        p = ops._attn_base(q_ptr, k_ptr, v_ptr, None, B, nq, 0, H, dh, q_bs, kv_bs, kv_bs, 0, q_rs, kv_rs, kv_rs, inner, talk)
        p.null_k, p.null_v = ptr(null_k), ptr(null_v)
        if key_mask is not None:
            p.key_mask, p.mask_bs = ptr(key_mask), key_mask.stride(0)
        Pp = torch.empty(B, H, nq, jp, dtype=torch.bfloat16, device=dev)
        dS = torch.empty(B, H, nq, jp, dtype=torch.bfloat16, device=dev)
        rc = lib().nuwa_attn_dense_bwd_fused(p, nk, ptr(do), nq * inner, inner, ptr(Pp), ptr(dS), jp, ptr(dtalk), scale, stream())
        if rc == _lib.NUWA_ERR_INVALID:
            Pp = dS = None
        else:
            check(rc, "nuwa_attn_dense_bwd_fused")
    if Pp is None:
        S = torch.empty(B, H, nq, jp, dtype=torch.float32, device=dev)
        dPp = torch.empty(B, H, nq, jp, dtype=torch.float32, device=dev)
        bgemm(q_ptr, kfull, S, M=nq, N=jp, K=dh, a_trans=0, b_trans=0, lda=q_rs, ldb=inner, ldc=jp, batch1=B, batch2=H,
              a_s=(q_bs, dh), b_s=full_s, c_s=sc, alpha=scale)
        bgemm(do, vfull, dPp, M=nq, N=jp, K=dh, a_trans=0, b_trans=0, lda=inner, ldb=inner, ldc=jp, batch1=B, batch2=H,
              a_s=(nq * inner, dh), b_s=full_s, c_s=sc)
        fuse_mask = J <= 288  # the register-resident row kernel applies the key mask itself; the wide fallback wants it in S
        if key_mask is not None and not fuse_mask:
            check(lib().nuwa_mask_scores(ptr(S), ptr(key_mask), key_mask.stride(0), B, H, nq, jp, nk, has_null, stream()),
                  "nuwa_mask_scores")
        Pp, dS = _rows(S, dPp, talk, dtalk, B, H, nq, J, jp, scale, key_mask=key_mask if fuse_mask else None,
                       has_null=has_null)
        del S, dPp
    # dQ = dS K
    bgemm(dS, kfull, dq_out, M=nq, N=dh, K=jp, a_trans=0, b_trans=1, lda=jp, ldb=inner, ldc=dq_rs, batch1=B, batch2=H,
          a_s=sc, b_s=full_s, c_s=(dq_bs, dh))
    # dK = dS^T Q ; dV = P'^T dO     (fp32, including the null slot and the zero pad rows).  The key / value branch feeds
    # nothing the query branch needs: with `side` (an object with run(fn, *keep), train._SideWork) it goes to the second
    # stream; `keep` = the tensors behind the raw q / k / v pointers
    def key_value_branch():
        dkfull = torch.empty(B, jp, inner, dtype=torch.float32, device=dev)
        dvfull = torch.empty(B, jp, inner, dtype=torch.float32, device=dev)
        bgemm(dS, q_ptr, dkfull, M=jp, N=dh, K=nq, a_trans=1, b_trans=1, lda=jp, ldb=q_rs, ldc=inner, batch1=B, batch2=H,
              a_s=sc, b_s=(q_bs, dh), c_s=full_s)
        bgemm(Pp, do, dvfull, M=jp, N=dh, K=nq, a_trans=1, b_trans=1, lda=jp, ldb=inner, ldc=inner, batch1=B, batch2=H,
              a_s=sc, b_s=(nq * inner, dh), c_s=full_s)
        k16, v16, k32, v32 = (None, None, dk_ptr, dv_ptr) if out_f32 else (dk_ptr, dv_ptr, None, None)
        check(lib().nuwa_kv_full_split(ptr(dkfull), ptr(dvfull), ptr(dnull_k), ptr(dnull_v), k16, v16, k32, v32, dkv_bs, dkv_rs,
                                       B, nk, jp, inner, stream()), "nuwa_kv_full_split")
    if side is not None:
        side.run(key_value_branch, dS, Pp, do, dnull_k, dnull_v, *keep)
    else:
        key_value_branch()


def attn_cross2dna_bwd(q, kv, do, *, B, n, nk, H, dh, talk, dtalk, null_k, null_v, dnull_k, dnull_v, key_mask, fmap, ck, cdil,
                       side=None):
    """Backward of the SparseCross2DNA core over a full teacher-forced pass (nuwa_pytorch.py:794-901): position 0 (bos) is a
    dense query over [null] + every context token WITHOUT talking heads (:828-844), positions 1.. attend to the null key
    and the ck x ck window at their own (y, x) in every context frame (:851-895).
    q: bf16 (B, n, inner); kv: bf16 (B, nk, 2*inner); do: bf16 (B, n, inner).  Returns dq (B, n, inner), dkv (B, nk, 2*inner)
    bf16; adds to dtalk, dnull_k, dnull_v."""
    inner = H * dh
    dev = q.device
    dq = torch.empty(B, n, inner, dtype=torch.bfloat16, device=dev)
    dkv = torch.empty(B, nk, 2 * inner, dtype=torch.bfloat16, device=dev)
    kb = kv.data_ptr()
    # ---- bos query: dense, fp32 key gradients kept as the base of the key-centric pass ----
    base = torch.empty(B, nk, 2 * inner, dtype=torch.float32, device=dev)
    do0 = do[:, 0].contiguous().view(B, 1, inner)
    attn_dense_bwd(q.data_ptr(), kb, kb + inner * 2, do0, B=B, nq=1, nk=nk, H=H, dh=dh, q_bs=n * inner, q_rs=inner,
                   kv_bs=nk * 2 * inner, kv_rs=2 * inner, talk=None, dtalk=None, null_k=null_k, null_v=null_v,
                   dnull_k=dnull_k, dnull_v=dnull_v, key_mask=key_mask, dq_out=(dq, dq.data_ptr()), dq_bs=n * inner,
                   dq_rs=inner, dk_ptr=base.data_ptr(), dv_ptr=base.data_ptr() + inner * 4, dkv_bs=nk * 2 * inner,
                   dkv_rs=2 * inner, out_f32=True)
    nq = n - 1
    if nq <= 0:
        dkv.copy_(base)
        return dq, dkv
    frames = nk // (fmap * fmap)
    J = 1 + frames * ck * ck
    jp = _round_up(J, 8)
    p = ops._attn_base(q.data_ptr() + inner * 2, kb, kb + inner * 2, None, B, nq, 1, H, dh, n * inner, nk * 2 * inner,
                       nk * 2 * inner, 0, inner, 2 * inner, 2 * inner, inner, talk)
    p.null_k, p.null_v = ptr(null_k), ptr(null_v)
    if key_mask is not None:
        p.key_mask, p.mask_bs = ptr(key_mask), key_mask.shape[1]
    p.fmap, p.ck, p.cdil, p.jmax = fmap, ck, cdil, J
    S = torch.empty(B, H, nq, jp, dtype=torch.float32, device=dev)
    dPp = torch.empty(B, H, nq, jp, dtype=torch.float32, device=dev)
    do_q = do.data_ptr() + inner * 2
    rc = _lib.NUWA_ERR_INVALID
    if SCORES_VARIANT == 'umma' or (SCORES_VARIANT == 'auto' and B * ((nq + 255) // 256) * 2 >= 64):
        rc = lib().nuwa_attnx2_bwd_scores_umma(p, do_q, n * inner, inner, ptr(S), ptr(dPp), jp, stream())
        if rc != _lib.NUWA_ERR_INVALID or SCORES_VARIANT == 'umma':
            check(rc, "nuwa_attnx2_bwd_scores_umma")
    if rc == _lib.NUWA_ERR_INVALID:
        check(lib().nuwa_attnx2_bwd_scores(p, do_q, n * inner, inner, ptr(S), ptr(dPp), jp, stream()), "nuwa_attnx2_bwd_scores")
    Pp, dS = _rows(S, dPp, talk, dtalk, B, H, nq, J, jp, dh ** -0.5)  # masks already folded into S by the gather
    rc = _lib.NUWA_ERR_INVALID
    if SCORES_VARIANT == 'umma' or (SCORES_VARIANT == 'auto' and B * ((nq + 255) // 256) * 2 >= 64):
        rc = lib().nuwa_attnx2_bwd_dq_umma(p, ptr(dS), jp, dq.data_ptr() + inner * 2, n * inner, inner, stream())
        if rc != _lib.NUWA_ERR_INVALID:
            check(rc, "nuwa_attnx2_bwd_dq_umma")
    if rc == _lib.NUWA_ERR_INVALID:
        check(lib().nuwa_attnx2_bwd_dq(p, ptr(dS), jp, dq.data_ptr() + inner * 2, n * inner, inner, stream()),
              "nuwa_attnx2_bwd_dq")
    # key / value branch (key-centric pass + null key / value column reduction): feeds only the to_kv weight gradient and the
    # context gradient -> second stream when `side` is given (train._SideWork)
    def key_value_branch():
        check(lib().nuwa_attnx2_bwd_dkdv(p, nk, do_q, n * inner, inner, ptr(dS), ptr(Pp), jp, ptr(base),
                                         base.data_ptr() + inner * 4, nk * 2 * inner, 2 * inner, dkv.data_ptr(),
                                         dkv.data_ptr() + inner * 2, nk * 2 * inner, 2 * inner, stream()),
              "nuwa_attnx2_bwd_dkdv")
        check(lib().nuwa_attn_bwd_first_key(q.data_ptr() + inner * 2, n * inner, inner, do_q, n * inner, inner, ptr(dS),
                                            ptr(Pp), jp, B, H, dh, nq, ptr(dnull_k), ptr(dnull_v), 0, stream()),
              "nuwa_attn_bwd_first_key")
    if side is not None:
        side.run(key_value_branch, q, kv, do, dS, Pp, base, dkv, key_mask, talk, dnull_k, dnull_v)
    else:
        key_value_branch()
    return dq, dkv
