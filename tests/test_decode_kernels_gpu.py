"""Decode-step kernels (skinny-M GEMM, lane-per-key attention): same results as the batch kernels / fp32 math."""
import pytest
import torch
import torch.nn.functional as F

from oracle import nuwa_oracle as O
from tests.helpers import gen, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(8, 512, 512), (8, 1536, 512), (3, 8192, 512), (32, 512, 1376), (1, 64, 136)])
def test_skinny_gemm_plain(cuda_device, M, N, K):
    from nuwa_pytorch_b200 import ops
    g = gen(M + N + K)
    a = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
    bias, res = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    ref = a.float() @ w.float().t() + bias + res
    out = ops.gemm(a.to(cuda_device), w.to(cuda_device), bias=bias.to(cuda_device), residual=res.to(cuda_device),
                   out_dtype=torch.float32)
    assert rel(out, ref) < 1e-5
    # the tensor-core kernel (forced) agrees
    out_tc = ops.gemm(a.to(cuda_device), w.to(cuda_device), bias=bias.to(cuda_device), residual=res.to(cuda_device),
                      out_dtype=torch.float32, force_bn=64)
    assert rel(out_tc, ref) < 1e-5


def test_skinny_gemm_geglu_and_strided_rows(cuda_device):
    from nuwa_pytorch_b200 import ops
    g = gen(5)
    M, K, inner = 8, 512, 1365
    buf = torch.randn(M, 7, K, generator=g).bfloat16().to(cuda_device)
    a = buf[:, 3, :]  # rows strided by 7*K, as the decode path reads cache rows
    w = (torch.randn(2 * inner, K, generator=g) / K ** 0.5).bfloat16()
    wp = ops.pack_pairs(w.to(cuda_device))
    out = ops.gemm(a, wp, act='geglu', out_dtype=torch.bfloat16)
    h = a.float().cpu() @ w.float().t()
    ref = h[:, :inner] * F.gelu(h[:, inner:])
    assert rel(out.float()[:, :inner], ref) < 4e-3
    assert out[:, inner:].float().abs().max().item() == 0.0  # zero padded tail columns


def test_decode_attention_matches_batch_kernels(cuda_device):
    """nq == 1 dispatches to the lane-per-key kernel; it must agree with the warp-per-query kernel on the same
    position (Sparse3DNA causal window and dense null-key attention)."""
    from nuwa_pytorch_b200 import ops
    g = gen(6)
    B, H, dh, fmap, maxf = 3, 8, 64, 4, 3
    inner = H * dh
    n = 1 + 2 * 16 + 7
    qkv = torch.randn(B, n, 3 * inner, generator=g).bfloat16().to(cuda_device)
    talk = (torch.randn(H, H, generator=g) / 2).to(cuda_device)
    full = torch.empty(B, n, inner, dtype=torch.bfloat16, device=cuda_device)
    geom = dict(H=H, dh=dh, talk=talk, fmap=fmap, max_frames=maxf, kernel=(5, 3, 3), dilation=(2, 2, 2), causal=True)
    ops.attn_sparse3dna(qkv, full, B=B, nq=n, t0=0, npos=n, nv=n - 1, **geom)
    for t in (0, 1, 17, n - 1):
        o = torch.empty(B, 1, inner, dtype=torch.bfloat16, device=cuda_device)
        t_dev = torch.tensor([t], dtype=torch.int32, device=cuda_device)
        ops.attn_sparse3dna_decode(qkv[:, t, :].contiguous(), qkv, o, t_dev, B=B, npos=n, **geom)
        assert rel(o[:, 0].float(), full[:, t].float()) < 1e-2, t
    # dense with null key, mask, talking heads
    nk = 50
    q = torch.randn(B, 1, inner, generator=g).bfloat16().to(cuda_device)
    kv = torch.randn(B, nk, 2 * inner, generator=g).bfloat16().to(cuda_device)
    null_k, null_v = torch.randn(H * dh, generator=g).to(cuda_device), torch.randn(H * dh, generator=g).to(cuda_device)
    mask = (torch.rand(B, nk, generator=g) > 0.3)
    mask[0] = False
    o1 = torch.empty(B, 1, inner, dtype=torch.bfloat16, device=cuda_device)
    ops.attn_dense(q.data_ptr(), kv.data_ptr(), kv.data_ptr() + inner * 2, o1, B=B, nq=1, nk=nk, H=H, dh=dh, q_bs=inner,
                   q_rs=inner, k_bs=nk * 2 * inner, k_rs=2 * inner, v_bs=nk * 2 * inner, v_rs=2 * inner, o_bs=inner,
                   o_rs=inner, talk=talk, null_k=null_k, null_v=null_v, key_mask=mask.to(torch.uint8).to(cuda_device))
    qh = O._heads(q.float().cpu(), H) * dh ** -0.5
    k, v = kv.float().cpu().chunk(2, -1)
    kh = torch.cat([null_k.cpu().view(1, H, 1, dh).expand(B, -1, -1, -1), O._heads(k, H)], 2)
    vh = torch.cat([null_v.cpu().view(1, H, 1, dh).expand(B, -1, -1, -1), O._heads(v, H)], 2)
    sim = (qh @ kh.transpose(-1, -2)).masked_fill(~F.pad(mask, (1, 0), value=True)[:, None, None], O.NEG)
    attn = O._talking_heads(sim.softmax(-1), talk.cpu()[:, :, None, None])
    assert rel(o1.float(), O._merge(attn @ vh)) < 4e-3


@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("kernel,dil,nv,B,talk_on,maxf", [
    ((5, 3, 3), (1, 1, 1), 768, 2, True, 10), ((5, 3, 3), (2, 2, 2), 1279, 2, True, 10), ((5, 3, 3), (4, 4, 4), 2559, 3, True, 10),
    ((5, 3, 3), (1, 2, 4), 601, 1, True, 10), ((3, 3, 3), (2, 4, 2), 530, 2, False, 10), ((3, 1, 3), (1, 1, 4), 256, 2, True, 10),
    ((5, 3, 3), (3, 5, 2), 1024, 1, True, 10), ((1, 3, 3), (1, 2, 1), 300, 2, True, 10), ((5, 3, 3), (1, 1, 1), 1, 2, True, 10),
    ((5, 3, 3), (1, 2, 4), 5, 1, True, 10), ((5, 3, 3), (2, 2, 2), 16, 2, True, 10), ((5, 3, 3), (1, 1, 1), 17, 1, True, 10),
    ((3, 3, 3), (1, 1, 1), 255, 2, True, 10), ((5, 3, 3), (4, 4, 4), 257, 3, True, 10), ((5, 3, 3), (9, 9, 4), 2560, 1, True, 10),
    ((5, 3, 3), (1, 1, 1), 767, 2, True, 3), ((5, 3, 3), (2, 2, 2), 767, 2, True, 3), ((5, 3, 3), (4, 4, 4), 767, 2, True, 3),
    ((5, 3, 3), (1, 8, 2), 700, 1, True, 5), ((3, 3, 3), (2, 16, 1), 512, 1, True, 4), ((5, 3, 3), (2, 2, 2), 300, 9, True, 10)])
def test_sparse3dna_umma_kernel_matches_gather_kernel(cuda_device, kernel, dil, nv, B, talk_on, maxf, causal):
    """attention_3dna_umma.cu (tcgen05 / TMEM: one UMMA per (head, frame offset) unit over a 128-query tile) vs the
    generic gather kernel on identical bf16 q|k|v: causal (decoder) and centred (sketch encoder, incl. the visible zero
    keys of frames beyond the sequence) windows, partial last rows / frames, mixed per-axis dilations, one-
    and two-class tiles (row dilation >= 4), more tiles than SMs."""
    from nuwa_pytorch_b200 import ops
    H, dh, fmap = 8, 64, 16
    g = gen(nv * 7 + dil[0] + 31 * kernel[2] + int(causal))
    inner = H * dh
    n = nv + 1
    qkv = torch.randn(B, n, 3 * inner, generator=g).bfloat16().to(cuda_device)
    talk = (torch.randn(H, H, generator=g) / 2).to(cuda_device) if talk_on else None
    geom = dict(B=B, nq=n, t0=0, npos=n, H=H, dh=dh, talk=talk, fmap=fmap, max_frames=maxf, nv=nv, kernel=kernel,
                dilation=dil, causal=causal)
    o_ref = torch.zeros(B, n, inner, dtype=torch.bfloat16, device=cuda_device)
    o_new = torch.full((B, n, inner), float('nan'), dtype=torch.bfloat16, device=cuda_device)
    ops.attn_sparse3dna(qkv, o_ref, variant='gather', **geom)
    ops.attn_sparse3dna(qkv, o_new, variant='umma', **geom)
    torch.cuda.synchronize()
    assert torch.isfinite(o_new.float()).all()  # every output row was written
    r = rel(o_new.float(), o_ref.float())
    worst = (o_new.float() - o_ref.float()).abs().max().item()
    print(f"  3dna umma vs gather causal={causal} k={kernel} d={dil} nv={nv} B={B}: rel {r:.2e} max abs {worst:.2e}")
    assert torch.equal(o_new[:, 0], o_ref[:, 0])  # bos row: its own value
    assert r < 4e-3  # P' is rounded to bf16 for the tensor-core PV (both outputs are bf16)


@pytest.mark.gpu
@pytest.mark.parametrize("n,frames,cdil,B,masked,talk_on", [
    (2561, 3, 1, 2, True, True), (2561, 3, 2, 2, True, True), (2561, 3, 4, 2, False, True), (1281, 5, 1, 1, True, True),
    (1000, 1, 2, 3, True, False), (258, 2, 4, 2, True, True), (2, 3, 1, 2, True, True), (17, 4, 2, 1, False, True),
    (2561, 3, 1, 9, True, True)])
def test_cross2dna_umma_kernel_matches_gather_kernel_and_oracle(cuda_device, n, frames, cdil, B, masked, talk_on):
    """SparseCross2DNA (nuwa_pytorch.py:851-895) on the tcgen05 / TMEM kernel (separate query and context buffers, unit =
    context frame, learned null key / value in slot 0, context mask on the gathered scores) vs the gather kernel on
    identical bf16 operands, and vs the fp32 oracle of the layer core; ragged last frame, fully masked context frame,
    one- and two-class tiles, more tiles than SMs."""
    from nuwa_pytorch_b200 import ops
    H, dh, fmap, ck = 8, 64, 16, 3
    inner, nk = H * dh, frames * fmap * fmap
    g = gen(n * 3 + frames * 17 + cdil)
    q = torch.randn(B, n, inner, generator=g).bfloat16()
    kv = torch.randn(B, nk, 2 * inner, generator=g).bfloat16()
    talk = torch.randn(H, H, generator=g) / 2 if talk_on else None
    null_k, null_v = torch.randn(inner, generator=g), torch.randn(inner, generator=g)
    mask = None
    if masked:
        mask = torch.rand(B, nk, generator=g) > 0.3
        mask[0, :fmap * fmap] = False          # a whole context frame masked
        mask[-1, -40:] = False
    dv_ = lambda t: None if t is None else t.to(cuda_device).contiguous()
    qd, kvd = dv_(q), dv_(kv)
    common = dict(B=B, nq=n - 1, t0=1, H=H, dh=dh, q_bs=n * inner, q_rs=inner, k_bs=nk * 2 * inner, k_rs=2 * inner,
                  v_bs=nk * 2 * inner, v_rs=2 * inner, o_bs=n * inner, o_rs=inner, talk=dv_(talk), null_k=dv_(null_k),
                  null_v=dv_(null_v), key_mask=dv_(mask.to(torch.uint8)) if masked else None, fmap=fmap, frames=frames, ck=ck,
                  cdil=cdil)
    outs = {}
    for variant in ('gather', 'umma'):
        o = torch.full((B, n, inner), float('nan'), dtype=torch.bfloat16, device=cuda_device)
        o[:, 0] = 0  # the bos query belongs to the dense kernel
        ops.attn_cross2dna(qd.data_ptr() + inner * 2, kvd.data_ptr(), kvd.data_ptr() + inner * 2, o.data_ptr() + inner * 2,
                           variant=variant, **common)
        torch.cuda.synchronize()
        assert torch.isfinite(o.float()).all(), variant
        outs[variant] = o.float().cpu()
    r = rel(outs['umma'], outs['gather'])
    print(f"  x2dna umma vs gather n={n} frames={frames} d={cdil} B={B} masked={masked}: rel {r:.2e}")
    assert r < 4e-3  # P' is rounded to bf16 for the tensor-core PV (both outputs are bf16)
    if B <= 3:
        eye = torch.eye(inner)
        tk = talk if talk_on else torch.eye(H)
        p = {'to_q.weight': eye, 'to_kv.weight': torch.eye(2 * inner), 'to_out.weight': eye,
             'talking_heads.weight': tk[:, :, None, None, None], 'null_k': null_k.view(H, 1, dh), 'null_v': null_v.view(H, 1, dh)}
        ref = O.sparse_cross2dna(q.float(), p, H, kv.float(), mask, fmap, ck, cdil)
        ro = rel(outs['umma'][:, 1:], ref[:, 1:])
        print(f"     vs oracle: rel {ro:.2e}")
        assert ro < 6e-3


@pytest.mark.gpu
@pytest.mark.parametrize("kernel,dil,nv,B,talk_on", [((5, 3, 3), (1, 1, 1), 768, 2, True), ((5, 3, 3), (2, 2, 2), 1279, 2, True),
                                                     ((5, 3, 3), (4, 4, 4), 2559, 3, True), ((5, 3, 3), (1, 2, 4), 601, 1, True),
                                                     ((3, 3, 5), (2, 4, 2), 530, 2, False), ((3, 1, 3), (1, 1, 3), 256, 2, True),
                                                     ((5, 3, 3), (3, 5, 2), 1024, 1, True), ((1, 3, 15), (1, 2, 1), 300, 2, True),
                                                     ((5, 3, 3), (1, 1, 1), 1, 2, True), ((5, 3, 3), (1, 2, 4), 5, 1, True),
                                                     ((5, 3, 3), (2, 2, 2), 16, 2, True), ((5, 3, 3), (1, 1, 1), 17, 1, True),
                                                     ((3, 3, 3), (1, 1, 1), 255, 2, True), ((5, 3, 3), (4, 4, 4), 257, 3, True),
                                                     ((5, 3, 3), (9, 9, 9), 2560, 1, True)])
def test_sparse3dna_halo_kernel_matches_gather_kernel(cuda_device, kernel, dil, nv, B, talk_on):
    """attention_3dna_halo.cu (TMA-staged key rows shared by 4 query rows, banded mma.sync blocks) vs the generic
    gather kernel on identical bf16 q|k|v, incl. partial last rows / frames and mixed per-axis dilations."""
    from nuwa_pytorch_b200 import ops
    H, dh, fmap, maxf = 8, 64, 16, 10
    g = gen(nv * 7 + dil[0] + 31 * kernel[2])
    inner = H * dh
    n = nv + 1
    qkv = torch.randn(B, n, 3 * inner, generator=g).bfloat16().to(cuda_device)
    talk = (torch.randn(H, H, generator=g) / 2).to(cuda_device) if talk_on else None
    geom = dict(B=B, nq=n, t0=0, npos=n, H=H, dh=dh, talk=talk, fmap=fmap, max_frames=maxf, nv=nv, kernel=kernel,
                dilation=dil, causal=True)
    o_ref = torch.zeros(B, n, inner, dtype=torch.bfloat16, device=cuda_device)
    o_halo = torch.full((B, n, inner), float('nan'), dtype=torch.bfloat16, device=cuda_device)
    ops.attn_sparse3dna(qkv, o_ref, variant='gather', **geom)
    ops.attn_sparse3dna(qkv, o_halo, variant='halo', **geom)
    torch.cuda.synchronize()
    assert torch.isfinite(o_halo.float()).all()  # every output row was written
    r = rel(o_halo.float(), o_ref.float())
    worst = (o_halo.float() - o_ref.float()).abs().max().item()
    print(f"  3dna halo vs gather k={kernel} d={dil} nv={nv}: rel {r:.2e} max abs {worst:.2e}")
    assert torch.equal(o_halo[:, 0], o_ref[:, 0])  # bos row: its own value
    assert r < 8e-3  # probabilities are rounded to bf16 for the tensor-core PV


@pytest.mark.gpu
def test_sparse3dna_halo_envelope(cuda_device):
    """Outside its envelope the halo kernel launches nothing and reports NUWA_ERR_INVALID ('auto' then falls back)."""
    from nuwa_pytorch_b200 import ops
    from nuwa_pytorch_b200._lib import NuwaB200Error
    H, dh, nv = 4, 32, 64
    qkv = torch.randn(1, nv + 1, 3 * H * dh, device=cuda_device).bfloat16()
    o = torch.zeros(1, nv + 1, H * dh, dtype=torch.bfloat16, device=cuda_device)
    geom = dict(B=1, nq=nv + 1, t0=0, npos=nv + 1, H=H, dh=dh, talk=None, fmap=8, max_frames=2, nv=nv, kernel=(3, 3, 3),
                dilation=(1, 1, 1), causal=True)
    with pytest.raises(NuwaB200Error):
        ops.attn_sparse3dna(qkv, o, variant='halo', **geom)
    ops.attn_sparse3dna(qkv, o, variant='auto', **geom)  # gather kernel
    torch.cuda.synchronize()
    assert o.float().abs().sum() > 0
