"""Unit parity of the individual CUDA kernels (through the C-ABI) against the CPU oracle."""
import pytest
import torch
import torch.nn.functional as F

from oracle import nuwa_oracle as O
from tests.helpers import gen, rel

pytestmark = pytest.mark.gpu


def test_groupnorm_nhwc(cuda_device):
    from nuwa_pytorch_b200 import ops
    g = gen(1)
    x = torch.randn(3, 64, 6, 5, generator=g) * 2 + 0.5
    w, b = torch.randn(64, generator=g), torch.randn(64, generator=g)
    for leaky in (False, True):
        ref = F.group_norm(x, 16, w, b)
        if leaky:
            ref = F.leaky_relu(ref, 0.1)
        o16, o32 = ops.groupnorm_nhwc(x.permute(0, 2, 3, 1).contiguous().to(cuda_device), w.to(cuda_device),
                                      b.to(cuda_device), 16, leaky=leaky, want_f32=True)
        assert rel(o32.permute(0, 3, 1, 2), ref) < 1e-5
        assert rel(o16.float().permute(0, 3, 1, 2), ref) < 4e-3


def test_upsample2x(cuda_device):
    from nuwa_pytorch_b200 import ops
    x = torch.randn(2, 16, 5, 7, generator=gen(2)).bfloat16()
    ref = F.interpolate(x.float(), scale_factor=2, mode='bilinear', align_corners=False)
    out = ops.upsample2x(x.permute(0, 2, 3, 1).contiguous().to(cuda_device))
    assert rel(out.float().permute(0, 3, 1, 2), ref) < 4e-3


def test_layout_and_first_conv(cuda_device):
    from nuwa_pytorch_b200 import ops
    g = gen(3)
    x = torch.randn(2, 5, 12, 9, generator=g)
    nhwc = ops.nchw_to_nhwc_bf16(x.to(cuda_device))
    assert torch.equal(nhwc.cpu(), x.permute(0, 2, 3, 1).bfloat16())
    back = ops.nhwc_to_nchw_f32(nhwc)
    assert torch.equal(back.cpu(), x.bfloat16().float())
    w = torch.randn(32, 5, 5, 5, generator=g) / 11
    bias = torch.randn(32, generator=g)
    wp = ops.pack_im2col_weight(w.to(cuda_device))
    a = ops.im2col(x.to(cuda_device), 5, wp.shape[1])
    out = ops.gemm(a, wp, bias=bias.to(cuda_device), out_dtype=torch.float32).view(2, 12, 9, 32)
    ref = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), bias, padding=2)
    assert rel(out.permute(0, 3, 1, 2), ref) < 1e-5


def test_conv1x1_to_nchw(cuda_device):
    from nuwa_pytorch_b200 import ops
    g = gen(4)
    # C % 64 == 0 -> tensor-core kernel (weights rounded to bf16); otherwise the fp32-weight warp-per-pixel kernel
    for C, cout in ((64, 3), (128, 3), (512, 8), (40, 3), (8, 1)):
        x = torch.randn(2, 7, 6, C, generator=g).bfloat16()
        w, b = torch.randn(cout, C, generator=g) / 8, torch.randn(cout, generator=g)
        out = ops.conv1x1_to_nchw(x.to(cuda_device), w.to(cuda_device), b.to(cuda_device))
        w_eff = w.bfloat16().float() if C % 64 == 0 else w
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w_eff[:, :, None, None], b)
        assert rel(out, ref) < 1e-5, (C, cout)


def test_sandwich_ln_shift_scatter(cuda_device):
    from nuwa_pytorch_b200 import ops
    g = gen(5)
    B, n, D, fmap = 2, 1 + 2 * 16 + 5, 64, 4
    y, xres = torch.randn(B, n, D, generator=g), torch.randn(B, n, D, generator=g)
    pw, pb, qw, qb = (torch.randn(D, generator=g) for _ in range(4))
    x_ref = xres + F.layer_norm(y, (D,), pw, pb)
    a_ref = O.shift_video_tokens(F.layer_norm(x_ref, (D,), qw, qb), fmap)
    dv = lambda t: t.to(cuda_device).contiguous()
    x_out = torch.empty(B, n, D, device=cuda_device)
    a_out = torch.full((B, n, D), 7.0, dtype=torch.bfloat16, device=cuda_device)
    ops.sandwich_ln(B, n, D, y=dv(y), post=(dv(pw), dv(pb)), res_in=dv(xres), x_out=x_out, pre=(dv(qw), dv(qb)),
                    a_out=a_out, a_bs=n * D, a_rs=D, a_t0=0, a_npos=n, shift=True, fmap=fmap)
    assert rel(x_out, x_ref) < 1e-6
    assert rel(a_out.float(), a_ref) < 4e-3
    # incremental: feeding tokens one by one into a persistent buffer gives the same operand rows
    a_inc = torch.zeros(B, n, D, dtype=torch.bfloat16, device=cuda_device)
    for t in range(n):
        ops.sandwich_ln(B, 1, D, res_in=dv(x_ref[:, t:t + 1]), pre=(dv(qw), dv(qb)), a_out=a_inc, a_bs=n * D, a_rs=D,
                        a_t0=0, a_npos=n, shift=True, fmap=fmap, t0=t)
    assert torch.equal(a_inc, a_out)


def test_stable_ln(cuda_device):
    from nuwa_pytorch_b200 import ops
    g = gen(6)
    a, b2 = torch.randn(37, 64, generator=g), torch.randn(37, 64, generator=g)
    w, b = torch.randn(64, generator=g), torch.randn(64, generator=g)
    o32, o16 = ops.stable_ln(a.to(cuda_device), w.to(cuda_device), b.to(cuda_device), b2=b2.to(cuda_device),
                             want_bf16=True)
    ref = O.stable_layer_norm(a + b2, w, b)
    assert rel(o32, ref) < 1e-5 and rel(o16.float(), ref) < 4e-3


def test_embed_axial(cuda_device):
    from nuwa_pytorch_b200 import ops
    g = gen(7)
    D, F_, h = 32, 3, 4
    table = torch.randn(50, D, generator=g)
    bos = torch.randn(D, generator=g)
    a1, a2, a3 = torch.randn(F_, D, generator=g), torch.randn(h, D, generator=g), torch.randn(h, D, generator=g)
    idx = torch.randint(0, 50, (2, 40), generator=g)
    pos = (a1[:, None, None] + a2[None, :, None] + a3[None, None, :]).reshape(-1, D)
    ref = torch.cat([bos[None, None].expand(2, 1, D), table[idx] + pos[:40]], dim=1)
    dv = lambda t: t.to(cuda_device)
    out = ops.embed_tokens(dv(idx), dv(table), nt=41, bos=dv(bos), axials=(dv(a1), dv(a2), dv(a3)), dims=(F_, h, h))
    assert rel(out, ref) < 1e-6
    part = ops.embed_tokens(dv(idx), dv(table), nt=3, t0=20, bos=dv(bos), axials=(dv(a1), dv(a2), dv(a3)), dims=(F_, h, h))
    assert torch.equal(part, out[:, 20:23])
    txt = ops.embed_tokens(dv(idx), dv(table), nt=40)
    assert torch.equal(txt.cpu(), table[idx])


def test_rotary(cuda_device):
    from nuwa_pytorch_b200 import ops
    g = gen(8)
    B, n, H, dh, rot = 2, 12, 2, 32, 32
    qkv = torch.randn(B * n, 3 * H * dh, generator=g)
    inv = 1. / (10000 ** (torch.arange(0, rot, 2).float() / rot))
    out = ops.rotary_to_bf16(qkv.to(cuda_device), inv.to(cuda_device), n, H, dh, rot)
    fr = O.rotary_freqs(inv, n)
    t = qkv.view(B, n, 3, H, dh).permute(0, 2, 3, 1, 4)  # b 3 h n d
    ref = O.apply_rotary(fr, t).permute(0, 3, 1, 2, 4).reshape(B * n, -1)
    assert rel(out.float(), ref) < 4e-3


def test_cross_entropy_and_sampling(cuda_device):
    from nuwa_pytorch_b200 import ops
    g = gen(9)
    logits = torch.randn(300, 1000, generator=g) * 3
    tgt = torch.randint(0, 1000, (300,), generator=g)
    loss = ops.cross_entropy_mean(logits.to(cuda_device), tgt.to(cuda_device))
    assert abs(loss.item() - F.cross_entropy(logits, tgt).item()) < 1e-5
    c, u = torch.randn(5, 8192, generator=g), torch.randn(5, 8192, generator=g)
    noise = torch.rand(5, 8192, generator=g)
    k = max(int((1 - 0.9) * 8192), 1)
    got, guided = ops.sample_topk_gumbel(c.to(cuda_device), u.to(cuda_device), noise.to(cuda_device), k, 2.0, 1.0,
                                         want_guided=True)
    mixed = u + (c - u) * 2.0
    want = O.gumbel_argmax(O.top_k_filter(mixed, 0.9), noise, 1.0)
    assert torch.equal(guided.cpu(), mixed)
    assert torch.equal(got.cpu(), want)


def _qkv_from(x, p, heads):
    q = x @ p['to_q.weight'].t()
    kv = x @ p['to_kv.weight'].t()
    return torch.cat([q, kv], dim=-1).bfloat16()


@pytest.mark.parametrize("causal,kernel,dil,n,geom", [
    (True, (5, 3, 3), 1, 49, None), (True, (5, 3, 3), 2, 30, None), (True, (3, 3, 3), 4, 21, None),
    (False, (3, 3, 3), 1, 23, None), (False, (5, 3, 3), 2, 48, None),
    # the model geometry (8 x 64 heads, 16-wide grid): causal full passes run attention_3dna_halo.cu -- checked here
    # directly against the oracle math (tests/test_decode_kernels_gpu.py compares it with the gather kernel)
    (True, (5, 3, 3), 1, 601, (1, 8, 64, 16, 3)), (True, (5, 3, 3), 2, 769, (1, 8, 64, 16, 3)),
    (True, (3, 3, 3), 4, 1025, (1, 8, 64, 16, 5))])
def test_attn_sparse3dna_core(cuda_device, causal, kernel, dil, n, geom):
    """Attention core only: identical bf16 q/k/v in, compare with the oracle run on those same bf16 values."""
    from nuwa_pytorch_b200 import ops
    g = gen(10 + n)
    B, H, dh, fmap, maxf = geom if geom is not None else (2, 2, 32, 4, 3)
    inner = H * dh
    qkv = (torch.randn(B, n, 3 * inner, generator=g)).bfloat16()
    talk = torch.randn(H, H, generator=g) / 2
    o = torch.empty(B, n, inner, dtype=torch.bfloat16, device=cuda_device)
    ops.attn_sparse3dna(qkv.to(cuda_device), o, B=B, nq=n, t0=0, npos=n, H=H, dh=dh, talk=talk.to(cuda_device),
                        fmap=fmap, max_frames=maxf, nv=n - 1, kernel=kernel, dilation=(dil,) * 3, causal=causal)
    # oracle with identity projections: feed q/k/v through to_q = I etc. by calling the core pieces directly
    q, k, v = qkv.float().split(inner, dim=-1)
    eye = torch.eye(inner)
    p = {'to_q.weight': eye, 'to_kv.weight': torch.cat([eye, eye]), 'talking_heads.weight': talk[:, :, None, None],
         'to_out.weight': eye, 'to_out.bias': torch.zeros(inner)}
    # the oracle derives k and v from ONE input; run it twice (x=k for keys, x=v for values) is not possible, so
    # restate the core with explicit q/k/v using its neighbour function
    T = fmap * fmap
    pad = (-(n - 1)) % T
    cur = (n + pad) // T
    idx, in_cur, masked = O.sparse3dna_neighbours(n - 1, (maxf, fmap, fmap), kernel, (dil,) * 3, causal, cur)
    qh, kh, vh = (O._heads(t, H) for t in (q, k, v))
    kh_p = F.pad(kh[:, :, 1:], (0, 0, 0, pad))
    vh_p = F.pad(vh[:, :, 1:], (0, 0, 0, pad))
    kg = kh_p[:, :, idx] * in_cur[None, None, :, :, None]
    vg = vh_p[:, :, idx] * in_cur[None, None, :, :, None]
    kg = torch.cat([kh[:, :, :1, None].expand(-1, -1, n - 1, -1, -1), kg], dim=3)
    vg = torch.cat([vh[:, :, :1, None].expand(-1, -1, n - 1, -1, -1), vg], dim=3)
    sim = torch.einsum('bhid,bhijd->bhij', qh[:, :, 1:] * dh ** -0.5, kg)
    sim = sim.masked_fill(F.pad(masked, (1, 0), value=False)[None, None], O.NEG)
    attn = O._talking_heads(sim.softmax(-1), talk[:, :, None, None])
    out = torch.cat([vh[:, :, :1], torch.einsum('bhij,bhijd->bhid', attn, vg)], dim=2)
    ref = O._merge(out)
    # output rounded to bf16 once; the tensor-core kernels also round the mixed probabilities to bf16 for P'V
    r = rel(o.float(), ref)
    print(f"  3dna core causal={causal} k={kernel} d={dil} n={n}: rel {r:.2e} (auto variant)")
    assert r < (4e-3 if geom is None else 3e-3)
    if geom is not None:   # the same case on each tensor-core kernel explicitly
        for variant in ('umma', 'halo'):
            o2 = torch.empty_like(o)
            ops.attn_sparse3dna(qkv.to(cuda_device), o2, B=B, nq=n, t0=0, npos=n, H=H, dh=dh, talk=talk.to(cuda_device),
                                fmap=fmap, max_frames=maxf, nv=n - 1, kernel=kernel, dilation=(dil,) * 3, causal=causal,
                                variant=variant)
            r2 = rel(o2.float(), ref)
            print(f"    {variant}: rel {r2:.2e}")
            assert r2 < 3e-3


def test_attn_dense_core(cuda_device):
    from nuwa_pytorch_b200 import ops
    g = gen(20)
    B, nq, nk, H, dh = 2, 13, 9, 2, 32
    inner = H * dh
    q = torch.randn(B, nq, inner, generator=g).bfloat16()
    kv = torch.randn(B, nk, 2 * inner, generator=g).bfloat16()
    null_k, null_v = torch.randn(H, dh, generator=g), torch.randn(H, dh, generator=g)
    talk = torch.randn(H, H, generator=g) / 2
    mask = torch.ones(B, nk, dtype=torch.bool)
    mask[1, 5:] = False
    mask[0, :] = False  # sample 0: unconditional sweep (only the null key visible)
    o = torch.empty(B, nq, inner, dtype=torch.bfloat16, device=cuda_device)
    qd, kvd = q.to(cuda_device), kv.to(cuda_device)
    ops.attn_dense(qd.data_ptr(), kvd.data_ptr(), kvd.data_ptr() + inner * 2, o, B=B, nq=nq, nk=nk, H=H, dh=dh,
                   q_bs=nq * inner, q_rs=inner, k_bs=nk * 2 * inner, k_rs=2 * inner, v_bs=nk * 2 * inner, v_rs=2 * inner,
                   o_bs=nq * inner, o_rs=inner, talk=talk.to(cuda_device), null_k=null_k.to(cuda_device).contiguous(),
                   null_v=null_v.to(cuda_device).contiguous(), key_mask=mask.to(torch.uint8).to(cuda_device))
    qh = O._heads(q.float(), H) * dh ** -0.5
    k, v = kv.float().chunk(2, -1)
    kh = torch.cat([null_k[None, :, None].expand(B, -1, -1, -1), O._heads(k, H)], 2)
    vh = torch.cat([null_v[None, :, None].expand(B, -1, -1, -1), O._heads(v, H)], 2)
    sim = qh @ kh.transpose(-1, -2)
    sim = sim.masked_fill(~F.pad(mask, (1, 0), value=True)[:, None, None], O.NEG)
    attn = O._talking_heads(sim.softmax(-1), talk[:, :, None, None])
    ref = O._merge(attn @ vh)
    assert rel(o.float(), ref) < 4e-3


@pytest.mark.parametrize("B,nq,nk,H,dh,null,talk,bias", [(2, 37, 256, 8, 64, True, True, False), (3, 256, 256, 8, 64, False, False, True),
                                                         (2, 50, 100, 4, 32, True, True, False)])
def test_attn_dense_tensor_core_variant(cuda_device, B, nq, nk, H, dh, null, talk, bias):
    """mma.sync variant vs the fp32 evaluation on identical bf16 operands, and vs the generic CUDA-core kernel."""
    from nuwa_pytorch_b200 import ops
    g = gen(B * 1000 + nq)
    inner = H * dh
    q = torch.randn(B, nq, inner, generator=g).bfloat16()
    kv = torch.randn(B, nk, 2 * inner, generator=g).bfloat16()
    null_k = torch.randn(H, dh, generator=g) if null else None
    null_v = torch.randn(H, dh, generator=g) if null else None
    tk = torch.randn(H, H, generator=g) / 2 if talk else None
    bs = torch.randn(H, nq, nk, generator=g) if bias else None
    hscale = torch.rand(H, generator=g) + 0.5 if bias else None
    mask = None
    if null:
        mask = torch.rand(B, nk, generator=g) > 0.3
        mask[0] = False
    dv = lambda t: None if t is None else t.to(cuda_device).contiguous()
    qd, kvd = dv(q), dv(kv)
    outs = []
    for use_mma in (True, False):
        o = torch.zeros(B, nq, inner, dtype=torch.bfloat16, device=cuda_device)
        ops.attn_dense(qd.data_ptr(), kvd.data_ptr(), kvd.data_ptr() + inner * 2, o, B=B, nq=nq, nk=nk, H=H, dh=dh,
                       q_bs=nq * inner, q_rs=inner, k_bs=nk * 2 * inner, k_rs=2 * inner, v_bs=nk * 2 * inner,
                       v_rs=2 * inner, o_bs=nq * inner, o_rs=inner, talk=dv(tk), null_k=dv(null_k), null_v=dv(null_v),
                       key_mask=None if mask is None else dv(mask.to(torch.uint8)), head_scale=dv(hscale), bias=dv(bs),
                       use_mma=use_mma)
        outs.append(o.float().cpu())
    qh = O._heads(q.float(), H) * dh ** -0.5
    k, v = kv.float().chunk(2, -1)
    kh, vh = O._heads(k, H), O._heads(v, H)
    if null:
        kh = torch.cat([null_k[None, :, None].expand(B, -1, -1, -1), kh], 2)
        vh = torch.cat([null_v[None, :, None].expand(B, -1, -1, -1), vh], 2)
    sim = qh @ kh.transpose(-1, -2)
    if bias:
        sim = sim * hscale[None, :, None, None] + bs[None]
    if mask is not None:
        sim = sim.masked_fill(~F.pad(mask, (1, 0), value=True)[:, None, None], O.NEG)
    attn = sim.softmax(-1)
    if talk:
        attn = O._talking_heads(attn, tk[:, :, None, None])
    ref = O._merge(attn @ vh)
    r_mma, r_gen = rel(outs[0], ref), rel(outs[1], ref)
    print(f"  dense attn B={B} nq={nq} nk={nk}: mma rel {r_mma:.2e}, generic rel {r_gen:.2e}")
    assert r_gen < 4e-3      # fp32 math, one bf16 rounding of the output
    assert r_mma < 8e-3      # additionally rounds the probabilities to bf16 for the PV tensor-core product


@pytest.mark.parametrize("variant", ["lib", "pres"])
@pytest.mark.parametrize("B,nq,nk,null,talk,masked", [(2, 100, 256, True, True, True), (1, 64, 50, True, True, False),
                                                       (3, 16, 12, False, True, True), (2, 257, 77, True, False, True),
                                                       (1, 2560, 256, True, True, True), (2, 300, 256, True, True, False),
                                                       (2, 33, 224, False, False, False)])
def test_attn_dense_x64_two_pass_kernel(cuda_device, B, nq, nk, null, talk, masked, variant):
    """attention_x64.cu (8 heads x 64, 64-query tiles, statistics pass + register talking-heads mix; variant 'lib') and
    attention_dense_pres.cu (probability slab resident in shared memory, tensor-core mix; 'pres') vs the oracle math."""
    from nuwa_pytorch_b200 import ops
    g = gen(300 + nq + nk)
    H, dh = 8, 64
    inner = H * dh
    q = torch.randn(B, nq, inner, generator=g).bfloat16()
    kv = torch.randn(B, nk, 2 * inner, generator=g).bfloat16()
    tk = torch.randn(H, H, generator=g) / 2 if talk else None
    null_k, null_v = (torch.randn(inner, generator=g), torch.randn(inner, generator=g)) if null else (None, None)
    mask = None
    if masked:
        mask = torch.rand(B, nk, generator=g) > 0.3
        mask[0, :nk // 2] = False
        if not null:
            mask[:, -1] = True
    qh = O._heads(q.float(), H) * dh ** -0.5
    k, v = kv.float().chunk(2, -1)
    kh, vh = O._heads(k, H), O._heads(v, H)
    if null:
        kh = torch.cat([null_k.view(1, H, 1, dh).expand(B, -1, -1, -1), kh], 2)
        vh = torch.cat([null_v.view(1, H, 1, dh).expand(B, -1, -1, -1), vh], 2)
    sim = qh @ kh.transpose(-1, -2)
    if masked:
        mm = F.pad(mask, (1, 0), value=True) if null else mask
        sim = sim.masked_fill(~mm[:, None, None], O.NEG)
    attn = sim.softmax(-1)
    if talk:
        attn = O._talking_heads(attn, tk[:, :, None, None])
    ref = O._merge(attn @ vh)
    dv = lambda t_: t_.to(cuda_device).contiguous() if t_ is not None else None
    qd, kvd = dv(q), dv(kv)
    o = torch.zeros(B, nq, inner, dtype=torch.bfloat16, device=cuda_device)
    ops.attn_dense(qd.data_ptr(), kvd.data_ptr(), kvd.data_ptr() + inner * 2, o, B=B, nq=nq, nk=nk, H=H, dh=dh,
                   q_bs=nq * inner, q_rs=inner, k_bs=nk * 2 * inner, k_rs=2 * inner, v_bs=nk * 2 * inner, v_rs=2 * inner,
                   o_bs=nq * inner, o_rs=inner, talk=dv(tk), null_k=dv(null_k), null_v=dv(null_v),
                   key_mask=dv(mask.to(torch.uint8)) if masked else None, variant=variant)
    r = rel(o.float(), ref)
    print(f"  {variant} dense attention B={B} nq={nq} nk={nk}: rel {r:.2e}")
    assert r < 6e-3  # probabilities are rounded to bf16 for the tensor-core P'V, output stored as bf16
