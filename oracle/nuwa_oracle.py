"""TEST INFRASTRUCTURE ONLY -- CPU fp32 oracle for the NUWA hot paths.

A functional restatement (plain PyTorch CPU ops, explicit index arithmetic, no nn.Module tree) of the
algorithm in lucidrains/nuwa-pytorch @ a3e3a6d for the in-scope path:
  VQGanVAE encode / decode / VQ lookup          (vqgan_vae.py)
  Sparse3DNA, Attention, SparseCross2DNA, FeedForward, SandwichNorm, ShiftVideoTokens,
  Transformer / ReversibleTransformer forward, NUWA / NUWASketch forward + generate steps (nuwa_pytorch.py)
Each function cites the reference lines it follows.  All functions take the reference's state_dict
(same key names) so a checkpoint of the reference drives the oracle directly.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
module; the product (nuwa_pytorch_b200) never does.

Parity pin: the reference ships no tests / golden vectors, so this oracle is pinned
against outputs of the UNMODIFIED reference executed in the build container (oracle/make_golden.py,
fixtures under tests/golden/).  The two third-party dependencies the reference needs (unfoldNd,
vector-quantize-pytorch) are absent from the image; they are restated in oracle/standins/ from their
published behaviour -- that part of the parity is UNPINNED.
"""
import math

import torch
import torch.nn.functional as F

NEG = -torch.finfo(torch.float32).max


# --------------------------------------------------------------------------------------------------
# small pieces
# --------------------------------------------------------------------------------------------------
def layer_norm(x, w, b, eps=1e-5):
    """nn.LayerNorm(dim) -- nuwa_pytorch.py:91,120-121."""
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def stable_layer_norm(x, w, b):
    """StableLayerNorm: LN(x / amax(x)) -- nuwa_pytorch.py:88-95."""
    return layer_norm(x / x.amax(dim=-1, keepdim=True), w, b)


def shift_video_tokens(x, fmap):
    """ShiftVideoTokens(shift_space=True, shift_time=False) -- nuwa_pytorch.py:200-253.
    Token 0 is bos.  Channels [0,d/4) come from the token one row up, [d/4,d/2) from one column left
    (zero at the frame border); the rest is untouched."""
    b, n, d = x.shape
    nv = n - 1
    if nv <= 0:
        return x
    q = -(-d // 4)  # torch.chunk(4) chunk size
    T = fmap * fmap
    v = torch.arange(nv)
    pos = v % T
    row, col = pos // fmap, pos % fmap
    vid = x[:, 1:]
    out = vid.clone()
    # chunk 0 : shifted along height
    src_h = (v - fmap).clamp(min=0)
    ch0 = vid[:, src_h, :q] * (row > 0).to(x.dtype)[None, :, None]
    out[:, :, :q] = ch0
    # chunk 1 : shifted along width
    src_w = (v - 1).clamp(min=0)
    ch1 = vid[:, src_w, q:2 * q] * (col > 0).to(x.dtype)[None, :, None]
    out[:, :, q:2 * q] = ch1
    return torch.cat([x[:, :1], out], dim=1)


def feed_forward(x, w1, w2):
    """FeedForward / GEGLU (bias-free, exact-erf GELU) -- nuwa_pytorch.py:255-278."""
    h = x @ w1.t()
    a, g = h.chunk(2, dim=-1)
    return (a * F.gelu(g)) @ w2.t()


def rotary_freqs(inv_freq, n):
    """RotaryEmbedding.forward -- nuwa_pytorch.py:138-142."""
    t = torch.arange(n).to(inv_freq.dtype)
    fr = t[:, None] * inv_freq[None, :]
    return torch.cat([fr, fr], dim=-1)


def apply_rotary(freqs, t):
    """apply_rotary_pos_emb / rotate_half -- nuwa_pytorch.py:144-153."""
    rd = freqs.shape[-1]
    tr, tp = t[..., :rd], t[..., rd:]
    h = rd // 2
    rot = torch.cat([-tr[..., h:], tr[..., :h]], dim=-1)
    tr = tr * freqs.cos() + rot * freqs.sin()
    return torch.cat([tr, tp], dim=-1)


def _heads(t, h):
    b, n, _ = t.shape
    return t.reshape(b, n, h, -1).permute(0, 2, 1, 3)  # b h n d


def _merge(t):
    b, h, n, d = t.shape
    return t.permute(0, 2, 1, 3).reshape(b, n, h * d)


def _talking_heads(attn, w):
    """Conv2d/Conv3d(H,H,1,bias=False) over the head axis of (b,h,...) -- nuwa_pytorch.py:309,372,404,556-558,781,889."""
    W = w.reshape(w.shape[0], w.shape[1])
    return torch.einsum('gh,bh...->bg...', W, attn)


# --------------------------------------------------------------------------------------------------
# attention variants
# --------------------------------------------------------------------------------------------------
def dense_attention(x, p, heads, context=None, key_mask=None, rotary=None, causal=False):
    """Attention.forward -- nuwa_pytorch.py:315-379.  p: to_q/to_kv/to_out/talking_heads weights, null_k/null_v."""
    b = x.shape[0]
    src = context if context is not None else x
    q = _heads(x @ p['to_q.weight'].t(), heads)
    k, v = (src @ p['to_kv.weight'].t()).chunk(2, dim=-1)
    k, v = _heads(k, heads), _heads(v, heads)
    if context is None and rotary is not None:
        q, k, v = (apply_rotary(rotary, t) for t in (q, k, v))  # values too (nuwa_pytorch.py:333-335)
    nk = p['null_k'][None].expand(b, -1, -1, -1)
    nv = p['null_v'][None].expand(b, -1, -1, -1)
    k, v = torch.cat([nk, k], dim=2), torch.cat([nv, v], dim=2)
    dh = q.shape[-1]
    sim = (q * dh ** -0.5) @ k.transpose(-1, -2)
    if key_mask is not None:
        km = F.pad(key_mask, (1, 0), value=True)
        sim = sim.masked_fill(~km[:, None, None, :], NEG)
    if causal:
        i, j = sim.shape[-2:]
        cm = torch.ones(i, j, dtype=torch.bool).triu_(j - i + 1)
        sim = sim.masked_fill(cm, NEG)
    attn = sim.softmax(dim=-1, dtype=torch.float32)
    attn = _talking_heads(attn, p['talking_heads.weight'])
    out = _merge(attn @ v)
    return out @ p['to_out.weight'].t()


def sparse3dna_neighbours(nq, video_shape, kernel, dilation, causal, cur_frames):
    """Key coordinates of Sparse3DNA -- nuwa_pytorch.py:420-429 (padding), :444-457 (mask), :507-527 (unfold).
    Returns idx (nq, J) into the (cur_frames*h*w) key grid, in_cur (nq,J) bool (key exists, else a zero
    vector) and masked (nq,J) bool (outside the max_frames grid => masked)."""
    maxf, hh, ww = video_shape
    kt, kh, kw = kernel
    dt, dh_, dw = dilation
    pads = [dt * (kt - 1) // 2, dh_ * (kh - 1) // 2, dw * (kw - 1) // 2]
    P = [2 * p if causal else p for p in pads]
    v = torch.arange(nq)
    T = hh * ww
    f, y, x = v // T, (v % T) // ww, v % ww
    a = torch.arange(kt)[:, None, None].expand(kt, kh, kw).reshape(-1)
    bq = torch.arange(kh)[None, :, None].expand(kt, kh, kw).reshape(-1)
    c = torch.arange(kw)[None, None, :].expand(kt, kh, kw).reshape(-1)
    ff = f[:, None] + a[None] * dt - P[0]
    yy = y[:, None] + bq[None] * dh_ - P[1]
    xx = x[:, None] + c[None] * dw - P[2]
    in_hw = (yy >= 0) & (yy < hh) & (xx >= 0) & (xx < ww)
    in_max = in_hw & (ff >= 0) & (ff < maxf)
    in_cur = in_hw & (ff >= 0) & (ff < cur_frames)
    idx = (ff.clamp(0, max(cur_frames - 1, 0)) * hh + yy.clamp(0, hh - 1)) * ww + xx.clamp(0, ww - 1)
    return idx, in_cur, ~in_max


def sparse3dna(x, p, heads, video_shape, kernel, dilation, causal, query_chunk=1024):
    """Sparse3DNA.forward -- nuwa_pytorch.py:459-613.  kernel / dilation are 3-tuples."""
    b, n, _ = x.shape
    hh = video_shape[1]
    T = hh * hh
    pad = (-(n - 1)) % T
    cur_frames = (n + pad) // T
    xp = F.pad(x, (0, 0, 0, pad)) if pad > 0 else x
    q = x @ p['to_q.weight'].t()
    k, v = (xp @ p['to_kv.weight'].t()).chunk(2, dim=-1)
    if n == 1:
        return v @ p['to_out.weight'].t() + p['to_out.bias']
    q, k, v = _heads(q, heads), _heads(k, heads), _heads(v, heads)
    dh = q.shape[-1]
    q = q * dh ** -0.5
    q = q[:, :, 1:]
    k_bos, v_bos = k[:, :, :1], v[:, :, :1]
    k, v = k[:, :, 1:], v[:, :, 1:]  # (b,h,cur_frames*T,dh), zero rows for the padded tail
    nq = n - 1
    assert nq <= video_shape[0] * T, 'more tokens than the precomputed mask covers (nuwa_pytorch.py:573)'
    idx, in_cur, masked = sparse3dna_neighbours(nq, video_shape, kernel, dilation, causal, cur_frames)
    W = p['talking_heads.weight']
    outs = []
    for s in range(0, nq, query_chunk):
        e = min(nq, s + query_chunk)
        ii, ic, mk = idx[s:e], in_cur[s:e], masked[s:e]
        kg = k[:, :, ii] * ic[None, None, :, :, None].to(k.dtype)  # (b,h,c,J,dh)
        vg = v[:, :, ii] * ic[None, None, :, :, None].to(v.dtype)
        kg = torch.cat([k_bos[:, :, None].expand(-1, -1, e - s, -1, -1), kg], dim=3)
        vg = torch.cat([v_bos[:, :, None].expand(-1, -1, e - s, -1, -1), vg], dim=3)
        sim = torch.einsum('bhid,bhijd->bhij', q[:, :, s:e], kg)
        mfull = F.pad(mk, (1, 0), value=False)  # bos never masked (nuwa_pytorch.py:456)
        sim = sim.masked_fill(mfull[None, None], NEG)
        attn = sim.softmax(dim=-1, dtype=torch.float32)
        attn = _talking_heads(attn, W)
        outs.append(torch.einsum('bhij,bhijd->bhid', attn, vg))
    out = torch.cat([v_bos] + outs, dim=2)  # bos output = its own value (nuwa_pytorch.py:608)
    return _merge(out) @ p['to_out.weight'].t() + p['to_out.bias']


def sparse_cross2dna(x, p, heads, context, context_mask, fmap, kernel, dilation):
    """SparseCross2DNA.forward -- nuwa_pytorch.py:794-901."""
    b, n, _ = x.shape
    T = fmap * fmap
    J = kernel * kernel
    pad = dilation * (kernel - 1) // 2
    clen = context.shape[1]
    fs = clen // T
    if context_mask is None:
        context_mask = torch.ones(b, clen, dtype=torch.bool)
    q = _heads(x @ p['to_q.weight'].t(), heads)
    k, v = (context @ p['to_kv.weight'].t()).chunk(2, dim=-1)
    k, v = _heads(k, heads), _heads(v, heads)
    dh = q.shape[-1]
    q = q * dh ** -0.5
    nk = p['null_k'][None].expand(b, -1, -1, -1)
    nv = p['null_v'][None].expand(b, -1, -1, -1)
    # bos query: dense over [null] + all context, no talking heads (nuwa_pytorch.py:828-844)
    kb, vb = torch.cat([nk, k], dim=2), torch.cat([nv, v], dim=2)
    sim_b = q[:, :, :1] @ kb.transpose(-1, -2)
    mb = F.pad(context_mask, (1, 0), value=True)
    sim_b = sim_b.masked_fill(~mb[:, None, None, :], NEG)
    out_b = sim_b.softmax(dim=-1, dtype=torch.float32) @ vb  # (b,h,1,dh)
    if n == 1:
        return _merge(out_b) @ p['to_out.weight'].t()
    # other queries: position i of ANY frame attends the k x k neighbourhood of i in every context frame
    i = torch.arange(T)
    y, xx = i // fmap, i % fmap
    a = torch.arange(kernel)[:, None].expand(kernel, kernel).reshape(-1)
    c = torch.arange(kernel)[None, :].expand(kernel, kernel).reshape(-1)
    ny = y[:, None] + a[None] * dilation - pad
    nx = xx[:, None] + c[None] * dilation - pad
    inb = (ny >= 0) & (ny < fmap) & (nx >= 0) & (nx < fmap)  # (T,J)
    sp = ny.clamp(0, fmap - 1) * fmap + nx.clamp(0, fmap - 1)
    fr = torch.arange(fs)
    idx = (fr[None, :, None] * T + sp[:, None, :]).reshape(T, fs * J)  # ordered (f, j) -- nuwa_pytorch.py:855
    inb_f = inb[:, None, :].expand(T, fs, J).reshape(T, fs * J)
    kg = k[:, :, idx] * inb_f[None, None, :, :, None].to(k.dtype)  # (b,h,T,fs*J,dh)
    vg = v[:, :, idx] * inb_f[None, None, :, :, None].to(v.dtype)
    kg = torch.cat([nk[:, :, None].expand(-1, -1, T, -1, -1), kg], dim=3)
    vg = torch.cat([nv[:, :, None].expand(-1, -1, T, -1, -1), vg], dim=3)
    qv = q[:, :, 1:]
    nqv = qv.shape[2]
    qpad = (-nqv) % T
    qv = F.pad(qv, (0, 0, 0, qpad))
    nf = qv.shape[2] // T
    qv = qv.reshape(b, heads, nf, T, dh)
    sim = torch.einsum('bhfid,bhijd->bhfij', qv, kg)
    cm = context_mask[:, idx] & inb_f[None]  # unfolded mask: zero padding => masked (nuwa_pytorch.py:878-881)
    cm = F.pad(cm, (1, 0), value=True)
    sim = sim.masked_fill(~cm[:, None, None], NEG)
    attn = sim.softmax(dim=-1, dtype=torch.float32)
    attn = _talking_heads(attn, p['talking_heads.weight'])
    out = torch.einsum('bhfij,bhijd->bhfid', attn, vg).reshape(b, heads, nf * T, dh)
    out = torch.cat([out_b, out], dim=2)
    return (_merge(out) @ p['to_out.weight'].t())[:, :n]


# --------------------------------------------------------------------------------------------------
# transformer stacks
# --------------------------------------------------------------------------------------------------
def _sub(sd, prefix):
    """state-dict view below `prefix.`"""
    pl = len(prefix) + 1
    return {k[pl:]: v for k, v in sd.items() if k.startswith(prefix + '.')}


def _inner(p):
    """strip SandwichNorm.fn / ShiftVideoTokens.fn nesting: returns the wrapped module's params."""
    q = _sub(p, 'fn')
    if any(k.startswith('fn.') for k in q):
        q = _sub(q, 'fn')
    return q


class StackSpec:
    """Hyper-parameters of one Transformer / ReversibleTransformer (nuwa_pytorch.py:1071-1295)."""

    def __init__(self, depth, heads=8, causal=False, cross=None, reversible=False, sparse3dna=False,
                 video_shape=None, kernel=(3, 3, 3), dilations=(1,), shift=False, cross_fmap=None,
                 cross_kernel=3, cross_dilations=(1,)):
        self.depth, self.heads, self.causal, self.cross = depth, heads, causal, cross  # cross: None|'dense'|'2dna'
        self.reversible, self.sparse3dna, self.video_shape = reversible, sparse3dna, video_shape
        self.kernel = tuple(kernel) if isinstance(kernel, (tuple, list)) else (kernel,) * 3
        self.dilations, self.shift = tuple(dilations), shift
        self.cross_fmap, self.cross_kernel, self.cross_dilations = cross_fmap, cross_kernel, tuple(cross_dilations)


def _sandwich(x, p, fn):
    """SandwichNorm -- nuwa_pytorch.py:124-128."""
    y = layer_norm(x, p['prenorm.weight'], p['prenorm.bias'])
    y = fn(y)
    return layer_norm(y, p['postnorm.weight'], p['postnorm.bias'])


def _self_attn_block(x, p, spec, layer, key_mask, rotary):
    fmap = spec.video_shape[1] if spec.sparse3dna else None
    inner = _inner(p)

    def fn(y):
        if spec.sparse3dna:
            if spec.shift:
                y = shift_video_tokens(y, fmap)
            d = spec.dilations[layer % len(spec.dilations)]
            d3 = tuple(d) if isinstance(d, (tuple, list)) else (d,) * 3
            return sparse3dna(y, inner, spec.heads, spec.video_shape, spec.kernel, d3, spec.causal)
        return dense_attention(y, inner, spec.heads, key_mask=key_mask, rotary=rotary, causal=spec.causal)

    return _sandwich(x, p, fn)


def _cross_block(x, p, spec, layer, context, context_mask):
    inner = _inner(p)

    def fn(y):
        if spec.cross == '2dna':
            d = spec.cross_dilations[layer % len(spec.cross_dilations)]
            return sparse_cross2dna(y, inner, spec.heads, context, context_mask, spec.cross_fmap, spec.cross_kernel, d)
        return dense_attention(y, inner, spec.heads, context=context, key_mask=context_mask)

    return _sandwich(x, p, fn)


def _ff_block(x, p, spec):
    inner = _inner(p)
    fmap = spec.video_shape[1] if spec.sparse3dna else None

    def fn(y):
        if spec.sparse3dna and spec.shift:
            y = shift_video_tokens(y, fmap)
        return feed_forward(y, inner['net.0.weight'], inner['net.3.weight'])

    return _sandwich(x, p, fn)


def transformer(x, sd, spec, mask=None, context=None, context_mask=None, rotary=None, final_norm=True):
    """Transformer.forward (nuwa_pytorch.py:1167-1182) or ReversibleTransformer.forward
    (:1289-1295 + reversible.py:60-68,132-142, inference form).  sd: state-dict below the stack prefix."""
    if not spec.reversible:
        for i in range(spec.depth):
            x = _self_attn_block(x, _sub(sd, f'layers.{i}.0'), spec, i, mask, rotary) + x
            if spec.cross:
                x = _cross_block(x, _sub(sd, f'layers.{i}.1'), spec, i, context, context_mask) + x
            x = _ff_block(x, _sub(sd, f'layers.{i}.2'), spec) + x
    else:
        x1, x2 = x, x
        li = 0
        for i in range(spec.depth):
            y1 = x1 + _self_attn_block(x2, _sub(sd, f'layers.{li}.0'), spec, i, mask, rotary)
            y2 = x2 + _ff_block(y1, _sub(sd, f'layers.{li}.1'), spec)
            x1, x2 = y1, y2
            li += 1
            if spec.cross:
                y1 = x1 + _cross_block(x2, _sub(sd, f'layers.{li}.0'), spec, i, context, context_mask)
                y2 = x2 + _ff_block(y1, _sub(sd, f'layers.{li}.1'), spec)
                x1, x2 = y1, y2
                li += 1
        x = x1 + x2
    if final_norm:
        x = stable_layer_norm(x, sd['norm.norm.weight'], sd['norm.norm.bias'])
    return x


# --------------------------------------------------------------------------------------------------
# VQGanVAE
# --------------------------------------------------------------------------------------------------
class VAESpec:
    """VQGanVAE hyper-parameters (vqgan_vae.py:289-315) that shape the forward path."""

    def __init__(self, dim, image_size, channels=3, num_layers=4, num_resnet_blocks=1, codebook_dim=256,
                 codebook_size=512, use_cosine_sim=True, attn_heads=8, attn_dim_head=64, groups=16,
                 first_conv_kernel_size=5, use_attn=True):
        self.dim, self.image_size, self.channels, self.num_layers = dim, image_size, channels, num_layers
        self.num_resnet_blocks, self.codebook_dim, self.codebook_size = num_resnet_blocks, codebook_dim, codebook_size
        self.use_cosine_sim, self.attn_heads, self.attn_dim_head, self.groups = use_cosine_sim, attn_heads, attn_dim_head, groups
        self.first_conv_kernel_size, self.use_attn = first_conv_kernel_size, use_attn
        self.layer_dims = [dim * 2 ** i for i in range(num_layers)]

    # module lists as the reference assembles them (vqgan_vae.py:351-366): list of (kind, index)
    def encoder_modules(self):
        mods = ['conv_in'] + ['down'] * self.num_layers + ['res'] * self.num_resnet_blocks
        if self.use_attn:
            mods.append('attn')
        return mods

    def decoder_modules(self):
        mods = ['glures'] * self.num_resnet_blocks
        if self.use_attn:
            mods.append('attn')
        return mods + ['up'] * self.num_layers + ['conv_out']


def _leaky(x):
    return F.leaky_relu(x, 0.1)  # slope is always 0.1 (vqgan_vae.py:94-95)


def vae_res_block(x, p, groups):
    """ResBlock -- vqgan_vae.py:228-242."""
    h = F.conv2d(x, p['net.0.weight'], p['net.0.bias'], padding=1)
    h = _leaky(F.group_norm(h, groups, p['net.1.weight'], p['net.1.bias']))
    h = F.conv2d(h, p['net.3.weight'], p['net.3.bias'], padding=1)
    h = _leaky(F.group_norm(h, groups, p['net.4.weight'], p['net.4.bias']))
    return F.conv2d(h, p['net.6.weight'], p['net.6.bias']) + x


def vae_glu_res_block(x, p, groups):
    """GLUResBlock -- vqgan_vae.py:212-226."""
    h = F.glu(F.conv2d(x, p['net.0.weight'], p['net.0.bias'], padding=1), dim=1)
    h = F.group_norm(h, groups, p['net.2.weight'], p['net.2.bias'])
    h = F.glu(F.conv2d(h, p['net.3.weight'], p['net.3.bias'], padding=1), dim=1)
    h = F.group_norm(h, groups, p['net.5.weight'], p['net.5.bias'])
    return F.conv2d(h, p['net.6.weight'], p['net.6.bias']) + x


def vae_cpb_bias(p, fmap):
    """ContinuousPositionBias -- vqgan_vae.py:192-210: MLP on sign(d)*log(|d|+1) of all grid offsets -> (H,n,n)."""
    pos = torch.arange(fmap)
    gy, gx = torch.meshgrid(pos, pos, indexing='ij')
    grid = torch.stack([gy, gx]).reshape(2, -1).t()
    rel = grid[:, None, :] - grid[None, :, :]
    rel = (torch.sign(rel) * torch.log(rel.abs() + 1)).float()
    n_lin = len({k.split('.')[1] for k in p if k.startswith('net.')})
    h = rel
    for li in range(n_lin - 1):
        h = _leaky(F.linear(h, p[f'net.{li}.0.weight'], p[f'net.{li}.0.bias']))
    h = F.linear(h, p[f'net.{n_lin - 1}.weight'], p[f'net.{n_lin - 1}.bias'])
    return h.permute(2, 0, 1)


def vae_attention(x, p, heads):
    """VQGanAttention -- vqgan_vae.py:265-286 (q,k l2-normalised over the SPATIAL axis, D9)."""
    b, c, hh, ww = x.shape
    qkv = F.conv2d(x, p['to_qkv.weight'])
    q, k, v = (t.reshape(b, heads, -1, hh * ww) for t in qkv.chunk(3, dim=1))  # b h c n
    q, k = F.normalize(q, dim=-1), F.normalize(k, dim=-1)
    sim = torch.einsum('bhci,bhcj->bhij', q, k) * p['scale'].exp()
    sim = sim + vae_cpb_bias(_sub(p, 'cpb'), hh)
    alpha = 32 ** 2  # stable_softmax, vqgan_vae.py:97-100
    t = sim / alpha
    t = t - t.amax(dim=-1, keepdim=True)
    attn = (t * alpha).softmax(dim=-1)
    out = torch.einsum('bhij,bhcj->bhci', attn, v).reshape(b, -1, hh, ww)
    out = F.conv2d(out, p['to_out.weight'], p['to_out.bias'])
    var = out.var(dim=1, unbiased=False, keepdim=True)
    mean = out.mean(dim=1, keepdim=True)
    out = (out - mean) / (var + 1e-5).sqrt() * p['post_norm.g'] + p['post_norm.b']  # LayerNormChan :140-143
    return out + x


def vq_lookup(flat, embed, use_cosine_sim=True):
    """Codebook arg-max of vector-quantize-pytorch (contract in oracle/standins/vector_quantize_pytorch).
    flat (M, D) fp32 after project_in.  Returns int64 indices (first maximum wins)."""
    if use_cosine_sim:
        dist = F.normalize(flat, dim=-1) @ F.normalize(embed, dim=-1).t()
    else:
        e = embed.t()
        dist = -(flat.pow(2).sum(1, keepdim=True) - 2 * flat @ e + e.pow(2).sum(0, keepdim=True))
    return dist.max(dim=-1).indices


def vae_encode_fmap(img, sd, spec):
    """encoders of VQGanVAE.encode up to (not including) the VQ -- vqgan_vae.py:431-433."""
    x = img
    for i, kind in enumerate(spec.encoder_modules()):
        p = _sub(sd, f'encoders.{i}')
        if kind == 'conv_in':
            x = F.conv2d(x, p['weight'], p['bias'], padding=spec.first_conv_kernel_size // 2)
        elif kind == 'down':
            x = _leaky(F.conv2d(x, p['0.weight'], p['0.bias'], stride=2, padding=1))
        elif kind == 'res':
            x = vae_res_block(x, p, spec.groups)
        elif kind == 'attn':
            x = vae_attention(x, p, spec.attn_heads)
    return x


def vae_quantize(fmap, sd, spec):
    """self.vq(fmap) in eval mode -- vqgan_vae.py:435 (+ third-party contract).  Returns (quantized NCHW, indices)."""
    b, c, hh, ww = fmap.shape
    flat = fmap.permute(0, 2, 3, 1).reshape(-1, c)
    if 'vq.project_in.weight' in sd:
        flat = F.linear(flat, sd['vq.project_in.weight'], sd['vq.project_in.bias'])
    embed = sd['vq._codebook.embed']
    ind = vq_lookup(flat, embed, spec.use_cosine_sim)
    quant = embed[ind]
    if 'vq.project_out.weight' in sd:
        quant = F.linear(quant, sd['vq.project_out.weight'], sd['vq.project_out.bias'])
    quant = quant.reshape(b, hh, ww, -1).permute(0, 3, 1, 2)
    return quant, ind.reshape(b, hh, ww)


def vae_encode(img, sd, spec):
    """VQGanVAE.encode -- vqgan_vae.py:431-435 (eval: loss = 0)."""
    quant, ind = vae_quantize(vae_encode_fmap(img, sd, spec), sd, spec)
    return quant, ind, torch.zeros(1)


def vae_decode(fmap, sd, spec):
    """VQGanVAE.decode -- vqgan_vae.py:437-441."""
    x = fmap
    for i, kind in enumerate(spec.decoder_modules()):
        p = _sub(sd, f'decoders.{i}')
        if kind == 'glures':
            x = vae_glu_res_block(x, p, spec.groups)
        elif kind == 'attn':
            x = vae_attention(x, p, spec.attn_heads)
        elif kind == 'up':
            x = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)
            x = _leaky(F.conv2d(x, p['1.weight'], p['1.bias'], padding=1))
        elif kind == 'conv_out':
            x = F.conv2d(x, p['weight'], p['bias'])
    return x


def vae_forward(img, sd, spec):
    """VQGanVAE.forward(img) reconstruction -- vqgan_vae.py:460-477."""
    quant, _, _ = vae_encode(img, sd, spec)
    return vae_decode(quant, sd, spec)


def vae_get_video_indices(video, sd, spec):
    """get_video_indices -- vqgan_vae.py:452-458."""
    b, f = video.shape[:2]
    _, ind, _ = vae_encode(video.reshape(b * f, *video.shape[2:]), sd, spec)
    return ind.reshape(b, f, *ind.shape[1:])


def vae_codebook_indices_to_video(indices, sd, spec, fmap_size):
    """codebook_indices_to_video -- vqgan_vae.py:443-450 (raw codebook vectors straight into decode, D4)."""
    b = indices.shape[0]
    codes = sd['vq._codebook.embed'][indices]  # (b, n, d)
    codes = codes.reshape(b, -1, fmap_size, fmap_size, codes.shape[-1]).permute(0, 1, 4, 2, 3)
    f = codes.shape[1]
    video = vae_decode(codes.reshape(b * f, -1, fmap_size, fmap_size), sd, spec)
    return video.reshape(b, f, *video.shape[1:])


# --------------------------------------------------------------------------------------------------
# NUWA / NUWASketch
# --------------------------------------------------------------------------------------------------
def axial_pos_emb(sd, prefix, shape):
    """AxialPositionalEmbedding.forward -- nuwa_pytorch.py:1693-1709 (axes of length 1 dropped, :1683)."""
    pos = None
    ax = 1
    for length in shape:
        if length <= 1:
            continue
        a = sd[f'{prefix}.axial{ax}']
        ax += 1
        pos = a if pos is None else pos[..., None, :] + a
    return pos.reshape(-1, pos.shape[-1])


def frac_gradient(x, frac=0.2):
    """Embedding.forward's gradient scaling -- nuwa_pytorch.py:1666-1670 (value unchanged, gradient x frac;
    embed_gradient_frac defaults to 0.2, :1741 / :2314)."""
    return x * frac + x.detach() * (1 - frac) if frac != 1. else x


class NUWASpec:
    def __init__(self, dim, fmap, max_video_frames, num_image_tokens, text_enc_depth=6, text_enc_heads=8,
                 dec_depth=6, dec_heads=8, dec_reversible=False, enc_reversible=True, kernel=3, dilation=1,
                 shift_video_tokens=True, text_enc_dim_head=64):
        self.text_dim_head = text_enc_dim_head
        self.dim, self.fmap, self.max_video_frames, self.num_image_tokens = dim, fmap, max_video_frames, num_image_tokens
        dil = tuple(range(1, dilation + 1)) if not isinstance(dilation, (list, tuple)) else tuple(dilation)
        self.video_shape = (max_video_frames, fmap, fmap)
        self.text = StackSpec(text_enc_depth, text_enc_heads, reversible=enc_reversible)
        self.dec = StackSpec(dec_depth, dec_heads, causal=True, cross='dense', reversible=dec_reversible,
                             sparse3dna=True, video_shape=self.video_shape, kernel=kernel, dilations=dil,
                             shift=shift_video_tokens)


def nuwa_embed_text(text, sd, spec):
    """NUWA.embed_text -- nuwa_pytorch.py:1821-1839 (rotary position embedding, text_rotary_pos_emb=True)."""
    mask = text != 0
    tok = frac_gradient(sd['text_embedding.embed.weight'][text])
    if 'text_abs_pos_emb.embed.weight' in sd:  # text_rotary_pos_emb=False: learned absolute positions -- :1829-1831
        tok = tok + sd['text_abs_pos_emb.embed.weight'][:text.shape[1]]
        rot = None
    else:
        inv = sd.get('text_rotary_pos_emb.inv_freq')
        if inv is None:  # buffer of RotaryEmbedding(dim=min(32, dim_head)) -- nuwa_pytorch.py:135,1769
            rd = min(32, spec.text_dim_head)
            inv = 1. / (10000 ** (torch.arange(0, rd, 2).float() / rd))
        rot = rotary_freqs(inv, text.shape[1])
    emb = transformer(tok, _sub(sd, 'text_transformer'), spec.text, mask=mask, rotary=rot)
    return emb, mask


def nuwa_decoder_input(indices_in, sd):
    """bos + image_embedding + axial positions -- nuwa_pytorch.py:1940-1944 / :1879-1881."""
    b, n = indices_in.shape
    emb = sd['image_embedding.embed.weight'][indices_in]
    return emb, b, n


def nuwa_logits(text, frame_indices, sd, spec, return_loss=True, context_mask_override=None):
    """NUWA.forward in eval mode (no cond-dropout) -- nuwa_pytorch.py:1917-1964.  frame_indices (b, F*h*w) int64."""
    text_emb, text_mask = nuwa_embed_text(text, sd, spec)
    if context_mask_override is not None:
        text_mask = context_mask_override
    idx_in = frame_indices[:, :-1] if return_loss else frame_indices
    b, n = idx_in.shape
    pos = axial_pos_emb(sd, 'video_pos_emb', spec.video_shape)
    x = frac_gradient(sd['image_embedding.embed.weight'][idx_in]) + pos[:n]
    x = torch.cat([sd['video_bos'][None, None].expand(b, 1, -1), x], dim=1)
    x = transformer(x, _sub(sd, 'video_transformer'), spec.dec, context=text_emb, context_mask=text_mask)
    logits = x @ sd['to_logits.weight'].t()
    if not return_loss:
        return logits
    loss = F.cross_entropy(logits.transpose(1, 2), frame_indices)
    return logits, loss


def top_k_filter(logits, thres=0.9):
    """top_k -- nuwa_pytorch.py:1713-1719."""
    k = max(int((1 - thres) * logits.shape[-1]), 1)
    val, ind = torch.topk(logits, k)
    out = torch.full_like(logits, float('-inf'))
    out.scatter_(1, ind, val)
    return out


def gumbel_argmax(logits, uniform, temperature=1.):
    """gumbel_sample with injected U(0,1) noise -- nuwa_pytorch.py:55-66."""
    def lg(t):
        return torch.log(t.clamp(min=1e-20))
    return ((logits / temperature) + (-lg(-lg(uniform)))).argmax(dim=-1)


def nuwa_generate_step_logits(text_emb, text_mask, prefix_indices, sd, spec, cond_scale=2.):
    """One iteration of NUWA.generate's loop body -- nuwa_pytorch.py:1879-1903, INCLUDING the reference's D8
    behaviour: the 'unconditional' sweep consumes the OUTPUT of the conditional sweep.  Returns the
    guided logits of the last position (b, V)."""
    b, n = prefix_indices.shape
    pos = axial_pos_emb(sd, 'video_pos_emb', spec.video_shape)
    x = sd['image_embedding.embed.weight'][prefix_indices] + pos[:n]
    x = torch.cat([sd['video_bos'][None, None].expand(b, 1, -1), x], dim=1)
    dec = _sub(sd, 'video_transformer')
    y = transformer(x, dec, spec.dec, context=text_emb, context_mask=text_mask)
    logits = y @ sd['to_logits.weight'].t()
    if cond_scale != 1:
        y2 = transformer(y, dec, spec.dec, context=text_emb, context_mask=torch.zeros_like(text_mask))
        ul = y2 @ sd['to_logits.weight'].t()
        logits = ul + (logits - ul) * cond_scale
    return logits[:, -1]


class SketchSpec:
    def __init__(self, dim, fmap, max_video_frames, sketch_max_video_frames, num_image_tokens,
                 sketch_enc_depth=6, sketch_enc_heads=8, sketch_enc_use_sparse_3dna=False, enc_reversible=False,
                 dec_depth=6, dec_heads=8, dec_reversible=False, kernel=3, dilation=1, cross_kernel=3,
                 cross_dilation=1, shift_video_tokens=True):
        dil = tuple(range(1, dilation + 1)) if not isinstance(dilation, (list, tuple)) else tuple(dilation)
        cdil = tuple(range(1, cross_dilation + 1)) if not isinstance(cross_dilation, (list, tuple)) else tuple(cross_dilation)
        self.dim, self.fmap = dim, fmap
        self.video_shape = (max_video_frames, fmap, fmap)
        self.sketch_shape = (sketch_max_video_frames, fmap, fmap)
        self.enc = StackSpec(sketch_enc_depth, sketch_enc_heads, reversible=enc_reversible,
                             sparse3dna=sketch_enc_use_sparse_3dna, video_shape=self.sketch_shape, kernel=kernel,
                             dilations=dil, shift=shift_video_tokens)
        self.dec = StackSpec(dec_depth, dec_heads, causal=True, cross='2dna', reversible=dec_reversible,
                             sparse3dna=True, video_shape=self.video_shape, kernel=kernel, dilations=dil,
                             shift=shift_video_tokens, cross_fmap=fmap, cross_kernel=cross_kernel,
                             cross_dilations=cdil)


def sketch_embed(sketch_indices, sketch_mask_frames, sd, spec):
    """NUWASketch.embed_sketch after the VAE tokenisation -- nuwa_pytorch.py:2420-2436.
    sketch_indices (b, f, h, w) int64 ; sketch_mask_frames (b, f) bool or None."""
    b, f = sketch_indices.shape[:2]
    idx = sketch_indices.reshape(b, -1)
    n = idx.shape[1]
    tok = frac_gradient(sd['sketch_embedding.embed.weight'][idx]) + axial_pos_emb(sd, 'sketch_pos_emb', spec.sketch_shape)[:n]
    if sketch_mask_frames is not None:
        mask = sketch_mask_frames[:, :, None].expand(b, f, n // f).reshape(b, n)
    else:
        mask = torch.ones(b, n, dtype=torch.bool)
    emb = transformer(tok, _sub(sd, 'sketch_transformer'), spec.enc, mask=mask)
    return emb, mask


def sketch_logits(sketch_indices, sketch_mask_frames, frame_indices, sd, spec, return_loss=True):
    """NUWASketch.forward in eval mode -- nuwa_pytorch.py:2514-2571."""
    ctx, cmask = sketch_embed(sketch_indices, sketch_mask_frames, sd, spec)
    idx_in = frame_indices[:, :-1] if return_loss else frame_indices
    b, n = idx_in.shape
    pos = axial_pos_emb(sd, 'video_pos_emb', spec.video_shape)
    x = frac_gradient(sd['image_embedding.embed.weight'][idx_in]) + pos[:n]
    x = torch.cat([sd['video_bos'][None, None].expand(b, 1, -1), x], dim=1)
    x = transformer(x, _sub(sd, 'video_transformer'), spec.dec, context=ctx, context_mask=cmask)
    logits = x @ sd['to_logits.weight'].t()
    if not return_loss:
        return logits
    return logits, F.cross_entropy(logits.transpose(1, 2), frame_indices)
