"""Unit parity of the training (backward) kernels, through the C-ABI, against torch autograd on the CPU oracle math."""
import pytest
import torch
import torch.nn.functional as F

from oracle import nuwa_oracle as O
from tests.helpers import gen, rel

pytestmark = pytest.mark.gpu


def _pad8(t):
    """(R, C) -> same values inside a buffer whose row pitch is a multiple of 8 (garbage-free zero pad)."""
    R, C = t.shape
    buf = torch.zeros(R, (C + 7) // 8 * 8, dtype=t.dtype, device=t.device)
    buf[:, :C] = t
    return buf


@pytest.mark.parametrize("at,bt", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_bgemm_all_orientations(cuda_device, at, bt):
    from nuwa_pytorch_b200 import ops_bwd
    g = gen(10 + at * 2 + bt)
    b1, b2, M, N, K = 2, 3, 70, 50, 45
    A = torch.randn(b1, b2, M, K, generator=g).bfloat16()
    Bm = torch.randn(b1, b2, K, N, generator=g).bfloat16()
    ref = 0.5 * (A.float() @ Bm.float())
    a_store = (A.transpose(-1, -2) if at else A).contiguous()           # (.., K, M) or (.., M, K)
    b_store = (Bm if bt else Bm.transpose(-1, -2)).contiguous()          # (.., K, N) or (.., N, K)
    a_dev = _pad8(a_store.reshape(-1, a_store.shape[-1])).to(cuda_device)
    b_dev = _pad8(b_store.reshape(-1, b_store.shape[-1])).to(cuda_device)
    lda, ldb = a_dev.shape[1], b_dev.shape[1]
    ra, rb = a_store.shape[-2], b_store.shape[-2]
    for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 4e-3)):
        C = torch.full((b1, b2, M, N), 7.0, dtype=dtype, device=cuda_device)
        ops_bwd.bgemm(a_dev, b_dev, C, M=M, N=N, K=K, a_trans=at, b_trans=bt, lda=lda, ldb=ldb, ldc=N, batch1=b1,
                      batch2=b2, a_s=(b2 * ra * lda, ra * lda), b_s=(b2 * rb * ldb, rb * ldb), c_s=(b2 * M * N, M * N),
                      alpha=0.5)
        assert rel(C.float(), ref) < tol, (at, bt, dtype)
    C = torch.ones(b1, b2, M, N, device=cuda_device)
    ops_bwd.bgemm(a_dev, b_dev, C, M=M, N=N, K=K, a_trans=at, b_trans=bt, lda=lda, ldb=ldb, ldc=N, batch1=b1, batch2=b2,
                  a_s=(b2 * ra * lda, ra * lda), b_s=(b2 * rb * ldb, rb * ldb), c_s=(b2 * M * N, M * N), alpha=0.5,
                  accumulate=True)
    assert rel(C, ref + 1) < 1e-5


def _round_up8(x):
    return (x + 7) // 8 * 8


def test_transpose_and_splitk_gemm(cuda_device):
    from nuwa_pytorch_b200 import ops_bwd
    g = gen(20)
    x = torch.randn(203, 77, generator=g).bfloat16()
    xt = ops_bwd.transpose(x.to(cuda_device))
    assert xt.shape == (77, 203) and torch.equal(xt.cpu(), x.t())
    # 16-byte-aligned pitches take the bf16-pair kernel: ragged rows / columns, a strided view, exact tiles
    for (R, C, pitch) in ((203, 88, 88), (256, 128, 128), (130, 72, 96), (2048, 512, 1536), (7, 8, 8)):
        base = torch.randn(R, pitch, generator=g).bfloat16().to(cuda_device)
        xv = base[:, :C]
        xt = ops_bwd.transpose(xv)
        assert xt.shape == (C, R) and torch.equal(xt, xv.t()), (R, C, pitch)
    for (M, N, K) in ((512, 192, 5000), (96, 40, 333), (1536, 512, 20480 // 4)):
        a = torch.randn(M, K, generator=g).bfloat16()
        w = torch.randn(N, K, generator=g).bfloat16()
        a_d, w_d = _pad8(a).to(cuda_device)[:, :K], _pad8(w).to(cuda_device)[:, :K]
        out = torch.ones(M, N, device=cuda_device)
        ops_bwd.gemm_splitk(a_d, w_d, out)
        ref = a.float() @ w.float().t() + 1
        assert rel(out, ref) < 2e-5, (M, N, K)


@pytest.mark.parametrize("tokens,n_out,k_in", [(5000, 512, 192), (333, 96, 40), (20480 // 4, 1536, 512), (2048, 64, 4096),
                                               (777, 200, 520), (64, 128, 256), (4100, 2816, 512)])
def test_weight_gradient_gemm_contraction_major(cuda_device, tokens, n_out, k_in):
    """dW = dY^T X on the tcgen05 kernel with BOTH operands read MN-major straight from dY [tokens, n_out] and
    X [tokens, k_in] (column-sliced views included) vs fp32 matmul, and vs the transpose + K-major path (same products,
    same split-K reduce: equal up to the order of the fp32 reduce-adds).  Ragged tokens / n_out / k_in exercise the TMA
    zero fill of partial 64-wide blocks and contraction rows."""
    from nuwa_pytorch_b200 import ops_bwd
    g = gen(tokens + n_out)
    pitch_y, pitch_x = _round_up8(n_out) + 64, _round_up8(k_in) + 8
    dy = torch.randn(tokens, pitch_y, generator=g).bfloat16().to(cuda_device)[:, 64:64 + n_out]   # a column slice
    x = torch.randn(tokens, pitch_x, generator=g).bfloat16().to(cuda_device)[:, :k_in]
    out = torch.ones(n_out, k_in, device=cuda_device)
    ops_bwd.gemm_splitk_tn(dy, x, out)
    ref = dy.float().t() @ x.float() + 1
    r = rel(out, ref)
    out2 = torch.ones(n_out, k_in, device=cuda_device)
    ops_bwd.WGRAD_TN = False
    try:
        ops_bwd.gemm_splitk_tn(dy, x, out2)
    finally:
        ops_bwd.WGRAD_TN = True
    print(f"  wgrad tn tokens={tokens} n_out={n_out} k_in={k_in}: rel vs fp32 {r:.2e}, vs transposed path {rel(out, out2):.2e}")
    assert r < 2e-5
    assert rel(out, out2) < 2e-6


@pytest.mark.parametrize("D", [64, 512, 1024])
def test_ln_bwd_plain_shift_stable(cuda_device, D):
    from nuwa_pytorch_b200 import ops_bwd
    g = gen(30 + D)
    B, fmap = 2, 4
    n = 1 + 2 * 16 + 5
    dv = lambda t: t.to(cuda_device).contiguous()
    x = torch.randn(B, n, D, generator=g) * 1.5 + 0.3
    w, b = torch.randn(D, generator=g), torch.randn(D, generator=g)
    dout = torch.randn(B, n, D, generator=g)
    # ---- plain LayerNorm, fp32 upstream gradient, accumulate into an existing gradient ----
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    F.layer_norm(xr, (D,), wr, br).backward(dout)
    dw, db, dcol = (torch.ones(D, device=cuda_device) for _ in range(3))
    acc = torch.ones(B * n, D, device=cuda_device)
    d16 = ops_bwd.ln_bwd(dv(dout.view(-1, D)), dv(x.view(-1, D)), dv(w), nt=n, dw=dw, db=db, dcol=dcol, dx_bf16=True,
                         dx_f32=acc, accumulate=True)
    assert rel(acc - 1, xr.grad.view(-1, D)) < 1e-5 and rel(d16.float(), xr.grad.view(-1, D)) < 4e-3
    assert rel(dw - 1, wr.grad) < 1e-5 and rel(db - 1, br.grad) < 1e-5
    assert rel(dcol - 1, xr.grad.sum((0, 1))) < 1e-4
    # ---- pre-norm of a ShiftVideoTokens block: bf16 upstream gradient of the shifted operand ----
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    da = dout.bfloat16()
    O.shift_video_tokens(F.layer_norm(xr, (D,), wr, br), fmap).backward(da.float())
    dw, db = torch.zeros(D, device=cuda_device), torch.zeros(D, device=cuda_device)
    dx = torch.zeros(B * n, D, device=cuda_device)
    ops_bwd.ln_bwd(dv(da.view(-1, D)), dv(x.view(-1, D)), dv(w), nt=n, dw=dw, db=db, unshift=True, fmap=fmap, dx_f32=dx)
    assert rel(dx, xr.grad.view(-1, D)) < 1e-5 and rel(dw, wr.grad) < 1e-5 and rel(db, br.grad) < 1e-5
    # ---- StableLayerNorm of the sum of two streams ----
    x2 = torch.randn(B, n, D, generator=g)
    xr, x2r, wr, br = (t.clone().requires_grad_() for t in (x, x2, w, b))
    O.stable_layer_norm(xr + x2r, wr, br).backward(dout)
    dw, db = torch.zeros(D, device=cuda_device), torch.zeros(D, device=cuda_device)
    d1, d2 = torch.zeros(B * n, D, device=cuda_device), torch.zeros(B * n, D, device=cuda_device)
    ops_bwd.ln_bwd(dv(dout.view(-1, D)), dv(x.view(-1, D)), dv(w), nt=n, dw=dw, db=db, x2=dv(x2.view(-1, D)), stable=True,
                   dx_f32=d1, dx2_f32=d2)
    assert rel(d1, xr.grad.view(-1, D)) < 1e-4 and torch.equal(d1, d2)  # x / amax(x) amplifies fp32 rounding
    assert rel(dw, wr.grad) < 1e-4 and rel(db, br.grad) < 1e-4


def test_geglu_ce_embed_rotary_bwd(cuda_device):
    from nuwa_pytorch_b200 import ops, ops_bwd
    g = gen(40)
    dv = lambda t: t.to(cuda_device).contiguous()
    # ---- GEGLU on the pair-packed layout ----
    M, inner = 37, 170
    h = torch.randn(M, 2 * inner, generator=g)
    hp = ops.pack_pairs(h.t().contiguous()).t().contiguous().bfloat16()    # (M, 2*ip) packed columns
    ip = hp.shape[1] // 2
    gk = ops_bwd.geglu_fwd(dv(hp))
    hq = hp.float()
    unpack = ops.pack_pairs(torch.arange(2 * inner, dtype=torch.float32)[:, None]).view(-1).long()  # packed col -> source col
    hsrc = torch.zeros(M, 2 * inner)
    valid = torch.ones(2 * ip, dtype=torch.bool)
    pad_cols = ops.pack_pairs(torch.ones(2 * inner, 1)).view(-1) == 0
    valid[pad_cols] = False
    hsrc[:, unpack[valid]] = hq[:, valid]
    hr = hsrc.clone().requires_grad_()
    a, gt = hr.chunk(2, -1)
    ref = a * F.gelu(gt)
    assert rel(gk.float()[:, :inner], ref) < 4e-3 and gk[:, inner:].float().abs().max().item() == 0
    dg = torch.zeros(M, ip)
    dg[:, :inner] = torch.randn(M, inner, generator=g)
    dg = dg.bfloat16()
    ref.backward(dg.float()[:, :inner])
    dh = ops_bwd.geglu_bwd(dv(dg), dv(hp)).float().cpu()
    got = torch.zeros(M, 2 * inner)
    got[:, unpack[valid]] = dh[:, valid]
    assert rel(got, hr.grad) < 6e-3
    # ---- cross entropy ----
    rows, V = 19, 300
    logits = (torch.randn(rows, V, generator=g) * 3).requires_grad_()
    tgt = torch.randint(0, V, (rows,), generator=g)
    (F.cross_entropy(logits, tgt) * 0.25).backward()
    dl = ops_bwd.ce_bwd(dv(logits.detach()), dv(tgt), gscale=torch.tensor([0.25], device=cuda_device))
    assert rel(dl.float(), logits.grad) < 4e-3
    # ---- embedding / axial positions / bos ----
    D, Fr, hh = 32, 3, 4
    table, bos = torch.randn(50, D, generator=g).requires_grad_(), torch.randn(D, generator=g).requires_grad_()
    a1, a2, a3 = (torch.randn(s, D, generator=g).requires_grad_() for s in (Fr, hh, hh))
    idx = torch.randint(0, 50, (2, 40), generator=g)
    pos = (a1[:, None, None] + a2[None, :, None] + a3[None, None, :]).reshape(-1, D)
    x = torch.cat([bos[None, None].expand(2, 1, D), O.frac_gradient(table[idx], 0.2) + pos[:40]], dim=1)
    dx = torch.randn(2, 41, D, generator=g)
    x.backward(dx)
    dt, dbos = torch.zeros(50, D, device=cuda_device), torch.zeros(D, device=cuda_device)
    dax = tuple(torch.zeros(s, D, device=cuda_device) for s in (Fr, hh, hh))
    ops_bwd.embed_bwd(dv(dx), dv(idx), dt, nt=41, frac=0.2, dbos=dbos, daxials=dax, dims=(Fr, hh, hh))
    assert rel(dt, table.grad) < 1e-5 and rel(dbos, bos.grad) < 1e-5
    for got_, want in zip(dax, (a1, a2, a3)):
        assert rel(got_, want.grad) < 1e-5
    # ---- rotary ----
    B, n, H, dh, rot = 2, 9, 2, 32, 32
    inv = 1. / (10000 ** (torch.arange(0, rot, 2).float() / rot))
    t = torch.randn(3, B, H, n, dh, generator=g).requires_grad_()
    y = O.apply_rotary(O.rotary_freqs(inv, n), t)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    to_rows = lambda u: u.permute(1, 3, 0, 2, 4).reshape(B * n, -1)  # (B*n, 3*H*dh)
    got = ops_bwd.rotary_bwd_to_bf16(dv(to_rows(dy)), dv(inv), n, H, dh, rot)
    assert rel(got.float(), to_rows(t.grad)) < 4e-3


@pytest.mark.parametrize("causal,kernel,dil,n,H,dh", [(True, (5, 3, 3), 1, 49, 2, 32), (True, (5, 3, 3), 2, 40, 2, 32),
                                                       (True, (3, 3, 3), 4, 49, 8, 64), (False, (3, 3, 3), 1, 33, 2, 32),
                                                       (False, (3, 5, 3), 2, 49, 4, 16)])
def test_attn_sparse3dna_bwd(cuda_device, causal, kernel, dil, n, H, dh):
    from nuwa_pytorch_b200 import ops, ops_bwd
    g = gen(50 + n + dil)
    B, fmap, maxf = 2, 4, 3
    inner = H * dh
    qkv = torch.randn(B, n, 3 * inner, generator=g).bfloat16()
    talk = torch.randn(H, H, generator=g) / 2
    do = torch.randn(B, n, inner, generator=g).bfloat16()
    eye, z = torch.eye(inner), torch.zeros(inner, inner)
    x = qkv.float().clone().requires_grad_()
    tk = talk.clone().requires_grad_()
    p = {'to_q.weight': torch.cat([eye, z, z], 1), 'to_kv.weight': torch.cat([torch.cat([z, eye, z], 1), torch.cat([z, z, eye], 1)]),
         'talking_heads.weight': tk[:, :, None, None], 'to_out.weight': eye, 'to_out.bias': torch.zeros(inner)}
    out = O.sparse3dna(x, p, H, (maxf, fmap, fmap), kernel, (dil,) * 3, causal)
    out.backward(do.float())
    # forward consistency of the saved q|k|v with the product kernel, then the backward
    o = torch.empty(B, n, inner, dtype=torch.bfloat16, device=cuda_device)
    geom = dict(H=H, dh=dh, fmap=fmap, max_frames=maxf, kernel=kernel, dilation=(dil,) * 3, causal=causal)
    ops.attn_sparse3dna(qkv.to(cuda_device), o, B=B, nq=n, t0=0, npos=n, nv=n - 1, talk=talk.to(cuda_device), **geom)
    assert rel(o.float(), out) < 1e-2
    dtalk = torch.zeros(H, H, device=cuda_device)
    dqkv = ops_bwd.attn_sparse3dna_bwd(qkv.to(cuda_device), do.to(cuda_device), B=B, n=n, talk=talk.to(cuda_device),
                                       dtalk=dtalk, **geom)
    r = rel(dqkv.float(), x.grad)
    print(f"  3dna bwd causal={causal} k={kernel} d={dil} n={n}: dqkv rel {r:.2e}, dtalk rel {rel(dtalk, tk.grad):.2e}")
    assert r < 1.5e-2 and rel(dtalk, tk.grad) < 1.5e-2
    for part, name in ((slice(0, inner), 'dq'), (slice(inner, 2 * inner), 'dk'), (slice(2 * inner, 3 * inner), 'dv')):
        assert rel(dqkv.float()[..., part], x.grad[..., part]) < 2e-2, name


@pytest.mark.parametrize("B,nq,nk,null,masked", [(2, 2560, 256, True, True), (1, 300, 256, True, False), (3, 77, 100, True, True),
                                                 (2, 256, 256, False, True), (1, 16, 1, True, False), (2, 50, 33, False, False)])
def test_dense_bwd_fused_probability_stage_matches_materialised_path(cuda_device, B, nq, nk, null, masked):
    """attention_dense_bwd.cu (S and dP' recomputed per 16-query tile on the tensor cores, softmax / talking-heads backward
    from shared memory) vs the materialised-logits path (two batched GEMMs + row kernel) on identical bf16 operands:
    dq, dk, dv, talking-heads and null key / value gradients.  Cfg-3 text-context shape, ragged query tiles, partial key
    chunks, a single key, fully masked samples, with / without the null key."""
    from nuwa_pytorch_b200 import ops_bwd
    H, dh = 8, 64
    inner = H * dh
    g = gen(nq + nk + int(null))
    dv_ = lambda t: None if t is None else t.to(cuda_device).contiguous()
    q = dv_(torch.randn(B, nq, inner, generator=g).bfloat16())
    kv = dv_(torch.randn(B, nk, 2 * inner, generator=g).bfloat16())
    do = dv_((torch.randn(B, nq, inner, generator=g) / 8).bfloat16())
    talk = dv_(torch.randn(H, H, generator=g) / 2)
    null_k = dv_(torch.randn(inner, generator=g)) if null else None
    null_v = dv_(torch.randn(inner, generator=g)) if null else None
    mask = None
    if masked:
        mask = torch.rand(B, nk, generator=g) > 0.3
        if null:
            mask[0] = False          # a sample that sees only the null key
        else:
            mask[:, 0] = True        # without a null key every row needs one visible key
        mask = dv_(mask.to(torch.uint8))
    outs = {}
    for fused in (False, True):
        ops_bwd.DENSE_BWD_FUSED = fused
        try:
            dtalk = torch.zeros(H, H, device=cuda_device)
            dnk, dnv = torch.zeros(inner, device=cuda_device), torch.zeros(inner, device=cuda_device)
            dq = torch.empty(B, nq, inner, dtype=torch.bfloat16, device=cuda_device)
            dkv = torch.empty(B, nk, 2 * inner, dtype=torch.bfloat16, device=cuda_device)
            ops_bwd.attn_dense_bwd(q.data_ptr(), kv.data_ptr(), kv.data_ptr() + inner * 2, do, B=B, nq=nq, nk=nk, H=H, dh=dh,
                                   q_bs=nq * inner, q_rs=inner, kv_bs=nk * 2 * inner, kv_rs=2 * inner, talk=talk, dtalk=dtalk,
                                   null_k=null_k, null_v=null_v, dnull_k=dnk if null else None, dnull_v=dnv if null else None,
                                   key_mask=mask, dq_out=dq, dq_bs=nq * inner, dq_rs=inner, dk_ptr=dkv.data_ptr(),
                                   dv_ptr=dkv.data_ptr() + inner * 2, dkv_bs=nk * 2 * inner, dkv_rs=2 * inner, out_f32=False)
        finally:
            ops_bwd.DENSE_BWD_FUSED = True
        torch.cuda.synchronize()
        outs[fused] = [t.float().clone() for t in (dq, dkv, dtalk, dnk, dnv)]
    rs = [rel(a, b) for a, b in zip(outs[True], outs[False])]
    print(f"  dense bwd fused vs materialised B={B} nq={nq} nk={nk} null={null} masked={masked}: dq {rs[0]:.2e} dkv {rs[1]:.2e} "
          f"dtalk {rs[2]:.2e} dnull {rs[3]:.2e}/{rs[4]:.2e}")
    assert all(torch.isfinite(t).all() for t in outs[True])
    # the fused kernel keeps dP' as bf16 and P as fp16 x fp32 factor in shared memory (the other path: fp32 in HBM) and uses
    # the exact fp32 null-key logit like the forward kernel (the other path rounds the null key to bf16)
    assert rs[0] < 6e-3 and rs[1] < 6e-3 and rs[2] < 6e-3
    if null:
        assert rs[3] < 6e-3 and rs[4] < 6e-3


@pytest.mark.parametrize("B,nk,H,null,masked", [(4, 768, 8, True, True), (2, 256, 8, True, False), (3, 100, 4, False, True)])
def test_single_query_dense_attention_kernels_match_the_generic_chain(cuda_device, B, nk, H, null, masked):
    """attention_q1.cu (one CTA per (sample, head) for the single bos query of SparseCross2DNA, forward and backward) vs the
    generic paths (decode kernel; K/V repack + batched GEMMs + row kernel + split) on identical bf16 operands."""
    from nuwa_pytorch_b200 import ops, ops_bwd
    dh = 64
    inner = H * dh
    g = gen(nk + H)
    dv_ = lambda t: None if t is None else t.to(cuda_device).contiguous()
    q = dv_(torch.randn(B, 5, inner, generator=g).bfloat16())          # the query is row 0 of a longer sequence
    kv = dv_(torch.randn(B, nk, 2 * inner, generator=g).bfloat16())
    do = dv_((torch.randn(B, 1, inner, generator=g) / 8).bfloat16())
    null_k = dv_(torch.randn(inner, generator=g)) if null else None
    null_v = dv_(torch.randn(inner, generator=g)) if null else None
    mask = None
    if masked:
        mask = torch.rand(B, nk, generator=g) > 0.4
        mask[:, 0] = True
        if null:
            mask[0] = False
        mask = dv_(mask.to(torch.uint8))
    common = dict(B=B, nq=1, nk=nk, H=H, dh=dh, q_bs=5 * inner, q_rs=inner, k_bs=nk * 2 * inner, k_rs=2 * inner,
                  v_bs=nk * 2 * inner, v_rs=2 * inner, o_bs=inner, o_rs=inner, talk=None, null_k=null_k, null_v=null_v, key_mask=mask)
    o_new = torch.full((B, 1, inner), float('nan'), dtype=torch.bfloat16, device=cuda_device)
    o_ref = torch.zeros_like(o_new)
    ops.attn_dense(q.data_ptr(), kv.data_ptr(), kv.data_ptr() + inner * 2, o_new, **common)
    ops.attn_dense(q.data_ptr(), kv.data_ptr(), kv.data_ptr() + inner * 2, o_ref, variant='generic', **common)
    torch.cuda.synchronize()
    rf = rel(o_new.float(), o_ref.float())
    outs = {}
    for q1 in (False, True):
        ops_bwd.Q1_KERNELS = q1
        try:
            dnk, dnv = torch.zeros(inner, device=cuda_device), torch.zeros(inner, device=cuda_device)
            dq = torch.zeros(B, 5, inner, dtype=torch.bfloat16, device=cuda_device)
            base = torch.full((B, nk, 2 * inner), float('nan'), dtype=torch.float32, device=cuda_device)
            ops_bwd.attn_dense_bwd(q.data_ptr(), kv.data_ptr(), kv.data_ptr() + inner * 2, do, B=B, nq=1, nk=nk, H=H, dh=dh,
                                   q_bs=5 * inner, q_rs=inner, kv_bs=nk * 2 * inner, kv_rs=2 * inner, talk=None, dtalk=None,
                                   null_k=null_k, null_v=null_v, dnull_k=dnk if null else None, dnull_v=dnv if null else None,
                                   key_mask=mask, dq_out=(dq, dq.data_ptr()), dq_bs=5 * inner, dq_rs=inner,
                                   dk_ptr=base.data_ptr(), dv_ptr=base.data_ptr() + inner * 4, dkv_bs=nk * 2 * inner,
                                   dkv_rs=2 * inner, out_f32=True)
        finally:
            ops_bwd.Q1_KERNELS = True
        torch.cuda.synchronize()
        outs[q1] = [t.float().clone() for t in (dq[:, 0], base, dnk, dnv)]
    rs = [rel(a, b) for a, b in zip(outs[True], outs[False])]
    print(f"  single-query dense B={B} nk={nk} H={H} null={null} masked={masked}: fwd {rf:.2e}  dq {rs[0]:.2e} dk|dv {rs[1]:.2e} "
          f"dnull {rs[2]:.2e}/{rs[3]:.2e}")
    assert torch.isfinite(o_new.float()).all() and all(torch.isfinite(t).all() for t in outs[True])
    assert rf < 4e-3 and rs[0] < 6e-3 and rs[1] < 6e-3     # bf16 outputs; the generic chain rounds P' / dS / null key to bf16
    if null:
        assert rs[2] < 6e-3 and rs[3] < 6e-3


@pytest.mark.parametrize("n,frames,cdil,B,masked", [(2561, 3, 1, 2, True), (2561, 3, 2, 2, True), (1281, 3, 4, 2, False),
                                                    (1000, 1, 2, 3, True), (258, 2, 1, 2, True),
                                                    (2561, 3, 1, 9, True)])   # 180 tiles: two tiles per CTA
def test_cross2dna_bwd_scores_tcgen05_matches_gather(cuda_device, n, frames, cdil, B, masked):
    """Backward of the SparseCross2DNA core with S / dP' from the tcgen05 kernel in scores mode and dq in PV mode vs the
    gather kernels:
    dq, dk|dv, talking-heads and null key / value gradients; context mask incl. a fully masked frame."""
    from nuwa_pytorch_b200 import ops_bwd
    H, dh, fmap, ck = 8, 64, 16, 3
    inner, nk = H * dh, frames * fmap * fmap
    g = gen(n + 7 * frames + cdil)
    dv_ = lambda t: None if t is None else t.to(cuda_device).contiguous()
    q = dv_(torch.randn(B, n, inner, generator=g).bfloat16())
    kv = dv_(torch.randn(B, nk, 2 * inner, generator=g).bfloat16())
    do = dv_((torch.randn(B, n, inner, generator=g) / 8).bfloat16())
    talk = dv_(torch.randn(H, H, generator=g) / 2)
    null_k, null_v = dv_(torch.randn(inner, generator=g)), dv_(torch.randn(inner, generator=g))
    mask = None
    if masked:
        mask = torch.rand(B, nk, generator=g) > 0.3
        mask[0, :fmap * fmap] = False
        mask = dv_(mask.to(torch.uint8))
    outs = {}
    for variant in ('gather', 'umma'):
        ops_bwd.SCORES_VARIANT = variant
        try:
            dtalk = torch.zeros(H, H, device=cuda_device)
            dnk, dnv = torch.zeros(inner, device=cuda_device), torch.zeros(inner, device=cuda_device)
            dq, dkv = ops_bwd.attn_cross2dna_bwd(q, kv, do, B=B, n=n, nk=nk, H=H, dh=dh, talk=talk, dtalk=dtalk, null_k=null_k,
                                                 null_v=null_v, dnull_k=dnk, dnull_v=dnv, key_mask=mask, fmap=fmap, ck=ck,
                                                 cdil=cdil)
        finally:
            ops_bwd.SCORES_VARIANT = 'auto'
        torch.cuda.synchronize()
        outs[variant] = [t.float().clone() for t in (dq, dkv, dtalk, dnk, dnv)]
    rs = [rel(a, b) for a, b in zip(outs['umma'], outs['gather'])]
    print(f"  x2dna bwd scores tcgen05 vs gather n={n} frames={frames} d={cdil} masked={masked}: dq {rs[0]:.2e} dkv {rs[1]:.2e} "
          f"dtalk {rs[2]:.2e} dnull {rs[3]:.2e}/{rs[4]:.2e}")
    assert all(torch.isfinite(t).all() for t in outs['umma'])
    assert rs[0] < 3e-3 and rs[1] < 3e-3 and rs[2] < 1e-3 and rs[3] < 1e-3 and rs[4] < 1e-3


@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("kernel,dil,nv,B,maxf", [((5, 3, 3), (1, 1, 1), 768, 2, 10), ((5, 3, 3), (2, 2, 2), 1279, 2, 10),
                                                  ((5, 3, 3), (4, 4, 4), 2559, 3, 10), ((5, 3, 3), (1, 2, 4), 601, 1, 10),
                                                  ((3, 3, 3), (2, 4, 2), 530, 2, 10), ((3, 1, 3), (1, 1, 4), 256, 2, 10),
                                                  ((5, 3, 3), (2, 2, 2), 767, 2, 3), ((5, 3, 3), (1, 1, 1), 17, 1, 10),
                                                  ((5, 3, 3), (1, 1, 1), 2560, 9, 10)])   # 180 tiles: two tiles per CTA
def test_sparse3dna_bwd_scores_tcgen05_matches_gather(cuda_device, kernel, dil, nv, B, maxf, causal):
    """Backward of the Sparse3DNA core with the logits S and dP' = dO V^T produced by the tcgen05 kernel in scores mode and
    dq by the same kernel in PV mode (dS in the place of the probabilities, V := K; kernel height 3) vs the gather kernels
    (same bf16 operands): dq|dk|dv (bf16) and the talking-heads gradient; causal and centred
    windows (incl. visible zero keys past the sequence), kernel heights 1 and 3, ragged last frame."""
    from nuwa_pytorch_b200 import ops_bwd
    H, dh, fmap = 8, 64, 16
    inner, n = H * dh, nv + 1
    g = gen(nv + 13 * dil[1] + int(causal))
    qkv = torch.randn(B, n, 3 * inner, generator=g).bfloat16().to(cuda_device)
    do = (torch.randn(B, n, inner, generator=g) / 8).bfloat16().to(cuda_device)
    talk = (torch.randn(H, H, generator=g) / 2).to(cuda_device)
    outs = {}
    for variant in ('gather', 'umma'):
        ops_bwd.SCORES_VARIANT = variant
        try:
            dtalk = torch.zeros(H, H, device=cuda_device)
            dqkv = ops_bwd.attn_sparse3dna_bwd(qkv, do, B=B, n=n, H=H, dh=dh, talk=talk, dtalk=dtalk, fmap=fmap,
                                               max_frames=maxf, kernel=kernel, dilation=dil, causal=causal)
        finally:
            ops_bwd.SCORES_VARIANT = 'auto'
        torch.cuda.synchronize()
        outs[variant] = (dqkv.float(), dtalk.clone())
    r, rt = rel(outs['umma'][0], outs['gather'][0]), rel(outs['umma'][1], outs['gather'][1])
    print(f"  3dna bwd scores tcgen05 vs gather causal={causal} k={kernel} d={dil} nv={nv}: dqkv {r:.2e} dtalk {rt:.2e}")
    assert torch.isfinite(outs['umma'][0]).all()
    assert r < 3e-3 and rt < 1e-3   # both paths accumulate the same bf16 products in fp32; outputs are bf16-rounded


@pytest.mark.parametrize("B,nq,nk,H,dh,null,masked", [(2, 37, 12, 2, 32, True, True), (3, 70, 50, 8, 64, True, False),
                                                       (2, 16, 16, 4, 16, False, True),
                                                       (2, 45, 256, 8, 64, True, True),     # cfg-3 text context: 257 slots,
                                                       (1, 33, 200, 8, 64, False, False),   # two warps per row kernel
                                                       (1, 20, 300, 4, 32, True, True)])    # 301 slots
def test_attn_dense_bwd(cuda_device, B, nq, nk, H, dh, null, masked):
    from nuwa_pytorch_b200 import ops_bwd
    g = gen(60 + nq)
    inner = H * dh
    q = torch.randn(B, nq, inner, generator=g).bfloat16()
    kv = torch.randn(B, nk, 2 * inner, generator=g).bfloat16()
    do = torch.randn(B, nq, inner, generator=g).bfloat16()
    talk = torch.randn(H, H, generator=g) / 2
    null_k, null_v = torch.randn(inner, generator=g), torch.randn(inner, generator=g)
    mask = torch.rand(B, nk, generator=g) > 0.3 if masked else None
    if masked:
        mask[0] = False
        if not null:
            mask[:, 0] = True
    qr, kvr, tk, nkr, nvr = (t.float().clone().requires_grad_() for t in (q, kv, talk, null_k.bfloat16(), null_v.bfloat16()))
    qh = O._heads(qr, H) * dh ** -0.5
    k, v = kvr.chunk(2, -1)
    kh, vh = O._heads(k, H), O._heads(v, H)
    if null:
        kh = torch.cat([nkr.view(1, H, 1, dh).expand(B, -1, -1, -1), kh], 2)
        vh = torch.cat([nvr.view(1, H, 1, dh).expand(B, -1, -1, -1), vh], 2)
    sim = qh @ kh.transpose(-1, -2)
    if masked:
        m = F.pad(mask, (1, 0), value=True) if null else mask
        sim = sim.masked_fill(~m[:, None, None], O.NEG)
    attn = O._talking_heads(sim.softmax(-1), tk[:, :, None, None])
    O._merge(attn @ vh).backward(do.float())
    dv_ = lambda t: t.to(cuda_device).contiguous()
    qd, kvd, dod = dv_(q), dv_(kv), dv_(do)
    dq = torch.empty(B, nq, inner, dtype=torch.bfloat16, device=cuda_device)
    dkv = torch.empty(B, nk, 2 * inner, dtype=torch.bfloat16, device=cuda_device)
    dtalk = torch.zeros(H, H, device=cuda_device)
    dnk, dnv = (torch.zeros(inner, device=cuda_device), torch.zeros(inner, device=cuda_device)) if null else (None, None)
    ops_bwd.attn_dense_bwd(qd.data_ptr(), kvd.data_ptr(), kvd.data_ptr() + inner * 2, dod, B=B, nq=nq, nk=nk, H=H, dh=dh,
                           q_bs=nq * inner, q_rs=inner, kv_bs=nk * 2 * inner, kv_rs=2 * inner, talk=dv_(talk), dtalk=dtalk,
                           null_k=dv_(null_k) if null else None, null_v=dv_(null_v) if null else None, dnull_k=dnk,
                           dnull_v=dnv, key_mask=dv_(mask.to(torch.uint8)) if masked else None, dq_out=dq,
                           dq_bs=nq * inner, dq_rs=inner, dk_ptr=dkv.data_ptr(), dv_ptr=dkv.data_ptr() + inner * 2,
                           dkv_bs=nk * 2 * inner, dkv_rs=2 * inner, out_f32=False)
    assert rel(dq.float(), qr.grad) < 1.5e-2
    assert rel(dkv.float(), kvr.grad) < 1.5e-2
    assert rel(dtalk, tk.grad) < 1.5e-2
    if null:
        assert rel(dnk, nkr.grad) < 1.5e-2 and rel(dnv, nvr.grad) < 1.5e-2


@pytest.mark.parametrize("n,frames,ck,cdil,H,dh,masked", [(49, 3, 3, 1, 2, 32, True), (40, 2, 3, 2, 2, 32, False),
                                                           (33, 1, 5, 1, 8, 64, True), (1, 2, 3, 1, 2, 32, True)])
def test_attn_cross2dna_bwd(cuda_device, n, frames, ck, cdil, H, dh, masked):
    """SparseCross2DNA core backward (dense bos query + windowed queries + null key) vs autograd on the oracle."""
    from nuwa_pytorch_b200 import ops_bwd
    g = gen(70 + n + ck)
    B, fmap = 2, 4
    inner, nk = H * dh, frames * 16
    q = torch.randn(B, n, inner, generator=g).bfloat16()
    kv = torch.randn(B, nk, 2 * inner, generator=g).bfloat16()
    do = torch.randn(B, n, inner, generator=g).bfloat16()
    talk = torch.randn(H, H, generator=g) / 2
    null_k, null_v = torch.randn(inner, generator=g).bfloat16().float(), torch.randn(inner, generator=g).bfloat16().float()
    mask = torch.rand(B, nk, generator=g) > 0.3 if masked else None
    if masked:
        mask[0, :16] = False
    eye = torch.eye(inner)
    qr, kvr, tk, nkr, nvr = (t.float().clone().requires_grad_() for t in (q, kv, talk, null_k, null_v))
    p = {'to_q.weight': eye, 'to_kv.weight': torch.eye(2 * inner), 'to_out.weight': eye,
         'talking_heads.weight': tk[:, :, None, None, None], 'null_k': nkr.view(H, 1, dh), 'null_v': nvr.view(H, 1, dh)}
    out = O.sparse_cross2dna(qr, p, H, kvr, mask, fmap, ck, cdil)
    out.backward(do.float())
    dv_ = lambda t: t.to(cuda_device).contiguous()
    dtalk = torch.zeros(H, H, device=cuda_device)
    dnk, dnv = torch.zeros(inner, device=cuda_device), torch.zeros(inner, device=cuda_device)
    dq, dkv = ops_bwd.attn_cross2dna_bwd(dv_(q), dv_(kv), dv_(do), B=B, n=n, nk=nk, H=H, dh=dh, talk=dv_(talk), dtalk=dtalk,
                                         null_k=dv_(null_k), null_v=dv_(null_v), dnull_k=dnk, dnull_v=dnv,
                                         key_mask=dv_(mask.to(torch.uint8)) if masked else None, fmap=fmap, ck=ck, cdil=cdil)
    print(f"  x2dna bwd n={n} frames={frames} ck={ck} d={cdil}: dq {rel(dq.float(), qr.grad):.2e} dkv "
          f"{rel(dkv.float(), kvr.grad):.2e} dnull {rel(dnk, nkr.grad):.2e}/{rel(dnv, nvr.grad):.2e}")
    assert rel(dq.float(), qr.grad) < 1.5e-2 and rel(dkv.float(), kvr.grad) < 1.5e-2
    assert rel(dnk, nkr.grad) < 1.5e-2 and rel(dnv, nvr.grad) < 1.5e-2
    if n > 1:
        assert rel(dtalk, tk.grad) < 1.5e-2
