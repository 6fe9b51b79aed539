"""NUWA / NUWASketch / Sparse3DNA on sm_100a kernels -- drop-in for the in-scope classes of
nuwa_pytorch/nuwa_pytorch.py (NUWA :1723-1964, NUWASketch :2297-2571, Sparse3DNA :381-613).

The nn.Module tree is a parameter container with the reference's attribute names, so state_dict keys are
identical (including the duplicated `net.blocks.*` aliases of reversible stacks and the persistent
`mask` buffer of Sparse3DNA).  No contained torch module is ever called: forward()/generate() go through
nuwa_pytorch_b200.engine -> libnuwa_b200.so.  generate() decodes incrementally with per-layer KV / shift
caches (mathematically identical to the reference's full recompute because every decoder op is causal,
SURVEY.md §3.3), including the reference's behaviour of feeding the conditional sweep's OUTPUT to the
'unconditional' sweep.
"""
from functools import wraps

import torch
from torch import nn

from . import engine, ops
from .vqgan_vae import VQGanVAE  # noqa: F401  (re-export, as the reference package does)


def _exists(v):
    return v is not None


def _cast_tuple(val, size=1):
    return val if isinstance(val, tuple) else (val,) * size


def _eval_decorator(fn):
    @wraps(fn)
    def inner(model, *args, **kwargs):
        was_training = model.training
        model.eval()
        out = fn(model, *args, **kwargs)
        model.train(was_training)
        return out
    return inner


# ------------------------------------------------------------------------------------------------
# parameter containers (attribute names == reference, nuwa_pytorch.py)
# ------------------------------------------------------------------------------------------------
class StableLayerNorm(nn.Module):  # :88-95
    def __init__(self, dim):
        super().__init__()
        self.norm = nn.LayerNorm(dim)


class SandwichNorm(nn.Module):  # :112-128
    def __init__(self, *, dim, fn):
        super().__init__()
        self.prenorm = nn.LayerNorm(dim)
        self.postnorm = nn.LayerNorm(dim)
        self.fn = fn


class RotaryEmbedding(nn.Module):  # :132-142
    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.register_buffer('inv_freq', 1. / (10000 ** (torch.arange(0, dim, 2).float() / dim)))


class ShiftVideoTokens(nn.Module):  # :185-253 (the shift itself is fused into the pre-norm kernel)
    def __init__(self, fn, image_size, shift_space=True, shift_time=False):
        super().__init__()
        self.fn = fn
        self.image_size = image_size
        self.shift_time = shift_time
        self.shift_space = shift_space


class GEGLU(nn.Module):  # :255-258 (fused into the first FF GEMM's epilogue)
    pass


class FeedForward(nn.Module):  # :260-286
    def __init__(self, *, dim, mult=4, dropout=0., chunk_size=None):
        super().__init__()
        inner_dim = (dim * mult * 2) // 3
        self.chunk_size = chunk_size  # output-neutral memory work-around of the reference; ignored
        self.net = nn.Sequential(nn.Linear(dim, inner_dim * 2, bias=False), GEGLU(), nn.Dropout(dropout),
                                 nn.Linear(inner_dim, dim, bias=False))


class Attention(nn.Module):  # :290-379
    def __init__(self, *, dim, heads=8, dim_head=64, causal=False, dropout=0.):
        super().__init__()
        inner_dim = heads * dim_head
        self.heads, self.causal, self.scale = heads, causal, dim_head ** -0.5
        self.null_k = nn.Parameter(torch.randn(heads, 1, dim_head))
        self.null_v = nn.Parameter(torch.randn(heads, 1, dim_head))
        self.talking_heads = nn.Conv2d(heads, heads, 1, bias=False)
        self.dropout = nn.Dropout(dropout)
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_kv = nn.Linear(dim, inner_dim * 2, bias=False)
        self.to_out = nn.Linear(inner_dim, dim, bias=False)


def _sparse3dna_mask(video_shape, kernel, dilation, causal):
    """The persistent `mask` buffer (N, J+1) of Sparse3DNA (:442-457): True where the key slot falls outside
    the (max_frames, h, w) grid; column 0 (bos) is never masked."""
    maxf, hh, ww = video_shape
    kt, kh, kw = kernel
    pads = [dilation[0] * (kt - 1) // 2, dilation[1] * (kh - 1) // 2, dilation[2] * (kw - 1) // 2]
    P = [2 * p if causal else p for p in pads]
    n = maxf * hh * ww
    v = torch.arange(n)
    f, y, x = v // (hh * ww), (v % (hh * ww)) // ww, v % ww
    j = torch.arange(kt * kh * kw)
    a, b, c = j // (kh * kw), (j // kw) % kh, j % kw
    ff = f[:, None] + a[None] * dilation[0] - P[0]
    yy = y[:, None] + b[None] * dilation[1] - P[1]
    xx = x[:, None] + c[None] * dilation[2] - P[2]
    outside = (ff < 0) | (ff >= maxf) | (yy < 0) | (yy >= hh) | (xx < 0) | (xx >= ww)
    return torch.cat([torch.zeros(n, 1, dtype=torch.bool), outside], dim=1)


class Sparse3DNA(nn.Module):  # :381-613
    def __init__(self, dim, video_shape, kernel_size=3, dilation=1, heads=8, dim_head=64, dropout=0., causal=False,
                 query_num_frames_chunk=None, rel_pos_bias=False):
        super().__init__()
        inner_dim = dim_head * heads
        self.heads, self.scale, self.causal = heads, dim_head ** -0.5, causal
        self.dropout = nn.Dropout(dropout)
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_kv = nn.Linear(dim, inner_dim * 2, bias=False)
        self.talking_heads = nn.Conv2d(heads, heads, 1, bias=False)
        self.to_out = nn.Linear(inner_dim, dim)
        self.dilation = _cast_tuple(dilation, size=3)
        self.kernel_size = _cast_tuple(kernel_size, size=3)
        assert all(map(lambda n: n % 2 == 1, self.kernel_size)), 'kernel size must be odd'
        self.kernel_numel = self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        if rel_pos_bias:
            raise NotImplementedError('sparse_3dna_rel_pos_bias only broadcasts for batch 1 in the reference (SURVEY D11)')
        self.rel_pos_bias = None
        self.video_shape = video_shape
        max_frames, fmap_size, _ = video_shape
        self.max_num_tokens = max_frames * fmap_size * fmap_size
        # accepted for signature compatibility; the fused kernel never materialises the unfolded keys
        self.query_num_frames_chunk = query_num_frames_chunk if _exists(query_num_frames_chunk) else max_frames
        self.register_buffer('mask', _sparse3dna_mask(video_shape, self.kernel_size, self.dilation, causal))

    def forward(self, x, **kwargs):
        """Standalone use of the exported class: x (b, n, dim) fp32 with position 0 = bos -> (b, n, dim)."""
        b, n, d = x.shape
        assert n - 1 <= self.max_num_tokens
        with torch.no_grad():
            a = x.detach().to(torch.bfloat16).contiguous().view(b * n, d)
            w_qkv = torch.cat([self.to_q.weight, self.to_kv.weight]).detach().to(torch.bfloat16).contiguous()
            inner = self.to_q.weight.shape[0]
            if n == 1:  # bos only: to_out(v)  (:485-486)
                kv = ops.gemm(a, w_qkv, out_dtype=torch.bfloat16)
                o = kv[:, 2 * inner:].contiguous()
            else:
                qkv = ops.gemm(a, w_qkv, out_dtype=torch.bfloat16)
                o = torch.empty(b * n, inner, dtype=torch.bfloat16, device=x.device)
                ops.attn_sparse3dna(qkv, o, B=b, nq=n, t0=0, npos=n, H=self.heads, dh=inner // self.heads,
                                    talk=self.talking_heads.weight.detach().float().reshape(self.heads, self.heads).contiguous(),
                                    fmap=self.video_shape[1], max_frames=self.video_shape[0], nv=n - 1,
                                    kernel=self.kernel_size, dilation=self.dilation, causal=self.causal)
            y = ops.gemm(o, self.to_out.weight.detach().to(torch.bfloat16).contiguous(),
                         bias=self.to_out.bias.detach().float().contiguous(), out_dtype=torch.float32)
        return y.view(b, n, d)


class SparseCross2DNA(nn.Module):  # :761-901
    def __init__(self, *, dim, image_size, heads=8, dim_head=64, dropout=0., kernel_size=3, dilation=1):
        super().__init__()
        inner_dim = heads * dim_head
        self.heads, self.scale = heads, dim_head ** -0.5
        self.null_k = nn.Parameter(torch.randn(heads, 1, dim_head))
        self.null_v = nn.Parameter(torch.randn(heads, 1, dim_head))
        self.talking_heads = nn.Conv3d(heads, heads, 1, bias=False)
        self.dropout = nn.Dropout(dropout)
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_kv = nn.Linear(dim, inner_dim * 2, bias=False)
        self.to_out = nn.Linear(inner_dim, dim, bias=False)
        self.image_size, self.kernel_size, self.dilation = image_size, kernel_size, dilation
        self.padding = dilation * (kernel_size - 1) // 2


def _build_layers(*, dim, depth, causal, heads, dim_head, ff_mult, cross_attend, attn_dropout, ff_dropout,
                  ff_chunk_size, cross_2dna_attn, cross_2dna_image_size, cross_2dna_kernel_size, cross_2dna_dilations,
                  sparse_3dna_attn, sparse_3dna_kernel_size, sparse_3dna_video_shape, sparse_3dna_query_num_frames_chunk,
                  sparse_3dna_dilations, sparse_3dna_rel_pos_bias, shift_video_tokens, reversible):
    """Shared layer assembly of Transformer (:1102-1163) and ReversibleTransformer (:1215-1277)."""
    assert not (sparse_3dna_attn and not _exists(sparse_3dna_video_shape)), 'sparse_3dna_video_shape must be defined if turned on'
    assert not (cross_2dna_attn and not _exists(cross_2dna_image_size)), 'cross_2dna_image_size must be defined'
    layers = nn.ModuleList([])
    for ind in range(depth):
        if sparse_3dna_attn:
            self_attn = Sparse3DNA(dim=dim, heads=heads, dim_head=dim_head, causal=causal,
                                   kernel_size=sparse_3dna_kernel_size,
                                   dilation=sparse_3dna_dilations[ind % len(sparse_3dna_dilations)],
                                   video_shape=sparse_3dna_video_shape,
                                   query_num_frames_chunk=sparse_3dna_query_num_frames_chunk,
                                   rel_pos_bias=sparse_3dna_rel_pos_bias)
            image_size = sparse_3dna_video_shape[-1]
        else:
            self_attn = Attention(dim=dim, heads=heads, dim_head=dim_head, causal=causal, dropout=attn_dropout)
            image_size = None
        cross_attn = None
        if cross_attend:
            if cross_2dna_attn:
                cross_attn = SparseCross2DNA(dim=dim, heads=heads, dim_head=dim_head, dropout=attn_dropout,
                                             image_size=cross_2dna_image_size, kernel_size=cross_2dna_kernel_size,
                                             dilation=cross_2dna_dilations[ind % len(cross_2dna_dilations)])
            else:
                cross_attn = Attention(dim=dim, heads=heads, dim_head=dim_head, dropout=attn_dropout)

        def make_ff():
            return FeedForward(dim=dim, mult=ff_mult, dropout=ff_dropout, chunk_size=ff_chunk_size)

        def mark_cross(s):
            s._is_cross = True
            return s

        if not reversible:
            ff = make_ff()
            if sparse_3dna_attn and shift_video_tokens:
                self_attn = ShiftVideoTokens(self_attn, image_size=image_size)
                ff = ShiftVideoTokens(ff, image_size=image_size)
            layers.append(nn.ModuleList([SandwichNorm(dim=dim, fn=self_attn),
                                         mark_cross(SandwichNorm(dim=dim, fn=cross_attn)) if cross_attend else None,
                                         SandwichNorm(dim=dim, fn=ff)]))
        else:
            def wrap(fn):  # the reversible stack always wraps (shift disabled when not 3dna / not requested), :1244
                return ShiftVideoTokens(fn, image_size=image_size, shift_space=bool(sparse_3dna_attn and shift_video_tokens))
            layers.append(nn.ModuleList([SandwichNorm(dim=dim, fn=wrap(self_attn)), SandwichNorm(dim=dim, fn=wrap(make_ff()))]))
            if cross_attend:
                layers.append(nn.ModuleList([mark_cross(SandwichNorm(dim=dim, fn=cross_attn)),
                                             SandwichNorm(dim=dim, fn=wrap(make_ff()))]))
    return layers


_STACK_DEFAULTS = dict(causal=False, heads=8, dim_head=64, ff_mult=4, cross_attend=False, attn_dropout=0., ff_dropout=0.,
                       ff_chunk_size=None, cross_2dna_attn=False, cross_2dna_image_size=None, cross_2dna_kernel_size=3,
                       cross_2dna_dilations=(1,), sparse_3dna_attn=False, sparse_3dna_kernel_size=3,
                       sparse_3dna_video_shape=None, sparse_3dna_query_num_frames_chunk=None, sparse_3dna_dilations=(1,),
                       sparse_3dna_rel_pos_bias=False, shift_video_tokens=False)


class Transformer(nn.Module):  # :1071-1182
    def __init__(self, *, dim, depth, rotary_pos_emb=False, **kw):
        super().__init__()
        cfg = {**_STACK_DEFAULTS, **kw}
        self.layers = _build_layers(dim=dim, depth=depth, reversible=False, **cfg)
        self.norm = StableLayerNorm(dim)

    def forward(self, x, mask=None, context=None, context_mask=None):
        raise TypeError('Transformer is executed by nuwa_pytorch_b200.engine.run_stack')


class Deterministic(nn.Module):  # reversible.py:20-50 (RNG replay only matters for dropout in training)
    def __init__(self, net):
        super().__init__()
        self.net = net


class ReversibleBlock(nn.Module):  # reversible.py:54-58
    def __init__(self, f, g):
        super().__init__()
        self.f = Deterministic(f)
        self.g = Deterministic(g)


class ReversibleSequence(nn.Module):  # reversible.py:126-130
    def __init__(self, blocks, args_route={}):
        super().__init__()
        self.args_route = args_route
        self.blocks = nn.ModuleList([ReversibleBlock(f=f, g=g) for f, g in blocks])


class ReversibleTransformer(nn.Module):  # :1184-1295
    def __init__(self, *, dim, depth, rotary_pos_emb=False, **kw):
        super().__init__()
        cfg = {**_STACK_DEFAULTS, **kw}
        self.layers = _build_layers(dim=dim, depth=depth, reversible=True, **cfg)
        self.net = ReversibleSequence(self.layers)  # same modules => aliased `net.blocks.*` state-dict keys
        self.norm = StableLayerNorm(dim)


class Embedding(nn.Module):  # :1659-1671 (frac_gradient only changes gradients)
    def __init__(self, *shape, frac_gradient=1.):
        super().__init__()
        self.frac_gradient = frac_gradient
        self.embed = nn.Embedding(*shape)


class AxialPositionalEmbedding(nn.Module):  # :1675-1709
    def __init__(self, dim, *, shape):
        super().__init__()
        self.full_shape = tuple(shape)
        shape = tuple(filter(lambda t: t > 1, shape))
        self.dim, self.shape, self.num_axials = dim, shape, len(shape)
        for axial_ind, axial_len in enumerate(shape):
            setattr(self, f'axial{axial_ind + 1}', nn.Parameter(torch.randn(axial_len, dim)))

    def tables(self):
        """axial tables aligned with the (f, h, w) axes of full_shape; None where an axis of length 1 was dropped."""
        out, ax = [], 1
        for length in self.full_shape:
            if length > 1:
                out.append(getattr(self, f'axial{ax}').detach().float().contiguous())
                ax += 1
            else:
                out.append(None)
        return tuple(out)


def _top_k_count(num_logits, thres):
    return max(int((1 - thres) * num_logits), 1)  # :1715


# ------------------------------------------------------------------------------------------------
# shared decoder logic of NUWA and NUWASketch
# ------------------------------------------------------------------------------------------------
class _VideoDecoderMixin:
    def _embed_video(self, indices, nt, t0=0, t_dev=None):
        """bos + image_embedding + axial positions for positions [t0, t0+nt)  (:1940-1944 / :1879-1881)."""
        return ops.embed_tokens(indices.contiguous(), self.image_embedding.embed.weight.detach().float().contiguous(),
                                nt=nt, t0=t0, t_dev=t_dev, bos=self.video_bos.detach().float().contiguous(),
                                axials=self.video_pos_emb.tables(), dims=self.video_pos_emb.full_shape)

    def _logits_weight(self):
        w = self.to_logits.weight
        key = (w.data_ptr(), w._version, ops._lib.WEIGHTS_EPOCH[0])
        cache = getattr(self, '_logits_cache', None)
        if cache is None or cache[0][0] != key[0] or cache[1].device != w.device:
            self._logits_cache = (key, w.detach().to(torch.bfloat16).contiguous())
        elif cache[0] != key:  # same storage, new values: refresh the bf16 copy in place (address stays graph-stable)
            with torch.no_grad():
                cache[1].copy_(w.detach())
            self._logits_cache = (key, cache[1])
        return self._logits_cache[1]

    @torch.no_grad()
    def refresh_packed_weights(self):
        """Re-derive every packed bf16 GEMM operand (stack weights, their dgrad transposes, the logits weight) from
        the fp32 parameters IN PLACE.  Device copies only, so it can be captured at the head of a CUDA graph
        (graphs.GraphedTrainStep): a replayed step then always computes with the current weights."""
        for name in ('text_transformer', 'sketch_transformer', 'video_transformer'):
            stack = getattr(self, name, None)
            if stack is not None:
                engine.pack_stack(stack).refresh()
        w = self._logits_weight()
        w.copy_(self.to_logits.weight.detach())

    def _decoder_logits(self, frame_indices, context, return_loss):
        """Teacher-forced pass: logits for every position, optionally the mean cross entropy (:1937-1964)."""
        b, N = frame_indices.shape
        n = N if return_loss else N + 1
        x = self._embed_video(frame_indices, n)
        _, y16 = engine.run_stack(self.video_transformer, x, context=context, want_bf16=True)
        logits = ops.gemm(y16.view(b * n, -1), self._logits_weight(), out_dtype=torch.float32)
        if not return_loss:
            return logits.view(b, n, -1)
        return ops.cross_entropy_mean(logits, frame_indices.reshape(-1).contiguous())

    @torch.no_grad()
    def _generate_indices(self, context, batch, *, num_frames, filter_thres, temperature, cond_scale, noise=None,
                          return_step_logits=False, use_graph=True, use_fused=True):
        """Autoregressive loop (:1858-1908 / :2455-2505), one token per step, KV-cached.

        Incremental mode keeps the position in a device scalar, so ONE decode step (both guidance sweeps, sampling,
        position increment) is captured in a CUDA graph and replayed for every token (dense-context decoders)."""
        dev = self.video_bos.device
        T = self.video_fmap_size ** 2
        total = T * num_frames
        max_tokens = T * self.max_video_frames
        V = self.to_logits.weight.shape[0]
        k = _top_k_count(V, filter_thres)
        video_indices = torch.zeros(batch, total, dtype=torch.int64, device=dev)
        pack = engine.pack_stack(self.video_transformer)
        uncond_ctx = context.with_mask(torch.zeros_like(context.mask)) if cond_scale != 1 else None
        w_log = self._logits_weight()
        step_logits = []
        if total <= max_tokens:
            t_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            st_c = engine.DecodeState(pack, batch, total, dev, t_dev)
            st_u = engine.DecodeState(pack, batch, total, dev, t_dev) if cond_scale != 1 else None
            engine.prime_context(self.video_transformer, context)
            noise_all = noise if noise is not None else torch.rand(total, batch, V, device=dev)
            noise_all = noise_all.contiguous()

            fused = use_fused and engine.FusedDecode.supported(pack, batch, context)
            if fused:
                # whole stack + to_logits of one token step in ONE persistent kernel per sweep (csrc/decode_stack.cu)
                plan_c = engine.FusedDecode(pack, st_c, context, w_log)
                plan_u = engine.FusedDecode(pack, st_u, uncond_ctx, w_log) if cond_scale != 1 else None

            def step(ind):
                x = self._embed_video(video_indices, 1, t0=ind, t_dev=t_dev)
                ulogits = None
                if fused:
                    y32, logits = plan_c.run(x)
                    if cond_scale != 1:
                        _, ulogits = plan_u.run(y32)  # second sweep consumes the first sweep's OUTPUT
                else:
                    y32, y16 = engine.run_stack(self.video_transformer, x, context=context, state=st_c, t0=ind,
                                                want_bf16=True)
                    logits = ops.gemm(y16.view(batch, -1), w_log, out_dtype=torch.float32)
                    if cond_scale != 1:
                        # the reference feeds the conditional sweep's OUTPUT to the second sweep
                        _, u16 = engine.run_stack(self.video_transformer, y32, context=uncond_ctx, state=st_u, t0=ind,
                                                  want_bf16=True)
                        ulogits = ops.gemm(u16.view(batch, -1), w_log, out_dtype=torch.float32)
                if return_step_logits:
                    step_logits.append(logits.clone() if ulogits is None else ulogits + (logits - ulogits) * cond_scale)
                ops.sample_topk_gumbel_at(logits, ulogits, noise_all, video_indices, t_dev, k, cond_scale, temperature)
                ops.step_increment(t_dev)

            graphable = use_graph and not return_step_logits and all(s.kind != 'x2dna' for s in pack.subs)
            if graphable and total > 2:
                step(0)  # eager warm-up: sizes every kernel's attributes / workspaces; its writes are redone by replay 0
                t_dev.zero_()
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    step(0)
                for _ in range(total):
                    graph.replay()
            else:
                for ind in range(total):
                    step(ind)
            return (video_indices, step_logits) if return_step_logits else video_indices

        for ind in range(total):
            # look-back window beyond max_video_frames (:1873-1877): positions shift every frame, so the window is
            # re-evaluated in full each step (reference semantics; rare path, only when num_frames > max_video_frames)
            window = video_indices[:, :ind]
            if ind > max_tokens:
                cur = ind % T
                lookback = (self.max_video_frames - (0 if cur == 0 else 1)) * T + cur
                window = window[:, -lookback:]
            n = window.shape[1] + 1
            x = self._embed_video(window, n)
            y32, y16 = engine.run_stack(self.video_transformer, x, context=context, want_bf16=True)
            logits = ops.gemm(y16[:, -1].contiguous(), w_log, out_dtype=torch.float32)
            ulogits = None
            if cond_scale != 1:
                _, u16 = engine.run_stack(self.video_transformer, y32, context=uncond_ctx, want_bf16=True)
                ulogits = ops.gemm(u16[:, -1].contiguous(), w_log, out_dtype=torch.float32)
            u = noise[ind] if noise is not None else torch.rand(batch, V, device=dev)
            res = ops.sample_topk_gumbel(logits, ulogits, u, k, cond_scale, temperature, want_guided=return_step_logits)
            if return_step_logits:
                res, guided = res
                step_logits.append(guided)
            video_indices[:, ind] = res
        return (video_indices, step_logits) if return_step_logits else video_indices

    @torch.no_grad()
    def _indices_to_video(self, video_indices, decode_max_batchsize):
        """vae.codebook[indices] -> decode in `decode_max_batchsize` chunks (:1910-1915; torch.chunk semantics)."""
        b = video_indices.shape[0]
        fm = self.video_fmap_size
        vae = self.vae
        c16, c32 = ops.gather_rows(vae._pack()['vq']['embed'], video_indices.reshape(-1).contiguous(), want_bf16=True,
                                   want_f32=True)
        d = c16.shape[1]
        c16, c32 = c16.view(-1, fm, fm, d), c32.view(-1, fm, fm, d)
        outs = [vae._decode_nhwc(a, bq) for a, bq in zip(c16.chunk(decode_max_batchsize, dim=0),
                                                          c32.chunk(decode_max_batchsize, dim=0))]
        imgs = torch.cat(outs, dim=0)
        return imgs.view(b, -1, *imgs.shape[1:])


# ------------------------------------------------------------------------------------------------
# NUWA
# ------------------------------------------------------------------------------------------------
class NUWA(nn.Module, _VideoDecoderMixin):
    def __init__(self, *, dim, vae=None, image_size=None, max_video_frames=5, text_num_tokens=49408,
                 text_max_seq_len=256, text_enc_depth=6, text_enc_dim_head=64, text_enc_heads=8, text_rotary_pos_emb=True,
                 enc_reversible=False, dec_depth=6, dec_dim_head=64, dec_heads=8, dec_reversible=False, attn_dropout=0.,
                 ff_dropout=0., ff_chunk_size=None, embed_gradient_frac=0.2, shift_video_tokens=True,
                 sparse_3dna_kernel_size=3, sparse_3dna_query_num_frames_chunk=None, sparse_3dna_dilation=1,
                 sparse_3dna_rel_pos_bias=False):
        super().__init__()
        assert _exists(vae) ^ _exists(image_size), 'either VAE or image size must be specified'
        if not _exists(vae):
            raise AttributeError("NUWA(image_size=...) without a vae fails in the reference too ('NoneType' object has "
                                 "no attribute 'num_layers', nuwa_pytorch.py:1760); pass vae=")
        if not enc_reversible:
            raise TypeError("NUWA with enc_reversible=False cannot run in the reference (Transformer.forward() got an "
                            "unexpected keyword argument 'rotary_pos_emb', nuwa_pytorch.py:1835-1839); use enc_reversible=True")
        self.vae = vae.copy_for_eval()
        image_size = vae.image_size
        num_image_tokens = vae.codebook_size
        self.text_max_seq_len = text_max_seq_len
        self.text_embedding = Embedding(text_num_tokens, dim, frac_gradient=embed_gradient_frac)
        self.text_abs_pos_emb = Embedding(text_max_seq_len, dim) if not text_rotary_pos_emb else None
        self.text_rotary_pos_emb = RotaryEmbedding(dim=min(32, text_enc_dim_head)) if text_rotary_pos_emb else None
        self.text_transformer = ReversibleTransformer(dim=dim, depth=text_enc_depth, heads=text_enc_heads,
                                                      dim_head=text_enc_dim_head, attn_dropout=attn_dropout,
                                                      ff_dropout=ff_dropout, rotary_pos_emb=text_rotary_pos_emb)
        self.video_bos = nn.Parameter(torch.randn(dim))
        self.image_embedding = Embedding(num_image_tokens, dim, frac_gradient=embed_gradient_frac)
        fmap_size = image_size // (2 ** vae.num_layers)
        self.video_fmap_size = fmap_size
        self.max_video_frames = max_video_frames
        video_shape = (max_video_frames, fmap_size, fmap_size)
        self.video_pos_emb = AxialPositionalEmbedding(dim, shape=video_shape)
        dilations = tuple(range(1, sparse_3dna_dilation + 1)) if not isinstance(sparse_3dna_dilation, (list, tuple)) \
            else tuple(sparse_3dna_dilation)
        klass = Transformer if not dec_reversible else ReversibleTransformer
        self.video_transformer = klass(dim=dim, depth=dec_depth, heads=dec_heads, dim_head=dec_dim_head, causal=True,
                                       cross_attend=True, attn_dropout=attn_dropout, ff_dropout=ff_dropout,
                                       ff_chunk_size=ff_chunk_size, shift_video_tokens=shift_video_tokens,
                                       sparse_3dna_video_shape=video_shape, sparse_3dna_attn=True,
                                       sparse_3dna_kernel_size=sparse_3dna_kernel_size, sparse_3dna_dilations=dilations,
                                       sparse_3dna_query_num_frames_chunk=sparse_3dna_query_num_frames_chunk,
                                       sparse_3dna_rel_pos_bias=sparse_3dna_rel_pos_bias)
        self.to_logits = nn.Linear(dim, num_image_tokens, bias=False)

    # ---- text encoder (:1821-1839) ----
    def embed_text(self, text, mask=None):
        batch, seq_len = text.shape
        assert seq_len <= self.text_max_seq_len, 'your input text has a greater length than what was designated on initialization'
        with torch.no_grad():
            axials, dims = (None, None, None), (1, 1, 1)
            if _exists(self.text_abs_pos_emb):
                axials, dims = (self.text_abs_pos_emb.embed.weight.detach().float().contiguous(), None, None), (seq_len, 1, 1)
            tokens = ops.embed_tokens(text.contiguous(), self.text_embedding.embed.weight.detach().float().contiguous(),
                                      nt=seq_len, axials=axials, dims=dims)
            rotary = None
            if _exists(self.text_rotary_pos_emb):
                rotary = (self.text_rotary_pos_emb.inv_freq.float().contiguous(), self.text_rotary_pos_emb.dim)
            km = mask.to(torch.uint8).contiguous() if _exists(mask) else None
            return engine.run_stack(self.text_transformer, tokens, key_mask=km, rotary=rotary)

    def _text_context(self, text, mask):
        with torch.no_grad():
            axials, dims = (None, None, None), (1, 1, 1)
            if _exists(self.text_abs_pos_emb):
                axials, dims = (self.text_abs_pos_emb.embed.weight.detach().float().contiguous(), None, None), (text.shape[1], 1, 1)
            tokens = ops.embed_tokens(text.contiguous(), self.text_embedding.embed.weight.detach().float().contiguous(),
                                      nt=text.shape[1], axials=axials, dims=dims)
            rotary = None
            if _exists(self.text_rotary_pos_emb):
                rotary = (self.text_rotary_pos_emb.inv_freq.float().contiguous(), self.text_rotary_pos_emb.dim)
            km = mask.to(torch.uint8).contiguous()
            _, e16 = engine.run_stack(self.text_transformer, tokens, key_mask=km, rotary=rotary, want_bf16=True)
        return engine.Context(e16, km)

    @torch.no_grad()
    @_eval_decorator
    def generate(self, *, text, filter_thres=0.9, temperature=1., decode_max_batchsize=10, cond_scale=2., num_frames=None,
                 _noise=None, _return_indices=False, _use_graph=True, _use_fused=True):
        batch = text.shape[0]
        context = self._text_context(text, text != 0)
        num_frames = num_frames if _exists(num_frames) else self.max_video_frames
        idx = self._generate_indices(context, batch, num_frames=num_frames, filter_thres=filter_thres,
                                     temperature=temperature, cond_scale=cond_scale, noise=_noise, use_graph=_use_graph,
                                     use_fused=_use_fused)
        if _return_indices:
            return idx
        return self._indices_to_video(idx, decode_max_batchsize)

    def forward(self, *, text, video=None, return_loss=False, cond_dropout_prob=0.2):
        batch, seq_len, frames, device = *text.shape, video.shape[1], text.device
        text_mask = text != 0
        if video.dtype == torch.long:
            frame_indices = video
        else:
            assert frames == self.max_video_frames, f'you must give the full video frames ({self.max_video_frames}) during training'
            assert _exists(self.vae), 'VAE must be passed in if you wish for video to be encoded to ids automatically'
            frame_indices = self.vae.get_video_indices(video)
        frame_indices = frame_indices.reshape(batch, -1)
        if self.training and cond_dropout_prob > 0:  # condition dropout (:1946-1950)
            uncond = torch.zeros((batch,), device=device).float().uniform_(0, 1) < cond_dropout_prob
            text_mask = text_mask & ~uncond[:, None]
        if return_loss and torch.is_grad_enabled() and any(p.requires_grad for p in self.to_logits.parameters()):
            # training step: CUDA forward with a tape + CUDA backward behind one autograd node (train.py)
            from . import train
            return train.nuwa_training_loss(self, text, frame_indices, text_mask.to(torch.uint8).contiguous())
        with torch.no_grad():
            # NB: the text encoder sees the un-dropped mask, the decoder the dropped one (reference order :1927-1950)
            context = self._text_context(text, text != 0).with_mask(text_mask.to(torch.uint8).contiguous())
            return self._decoder_logits(frame_indices, context, return_loss)


# ------------------------------------------------------------------------------------------------
# NUWASketch
# ------------------------------------------------------------------------------------------------
class NUWASketch(nn.Module, _VideoDecoderMixin):
    def __init__(self, *, vae, sketch_vae, dim, image_size, max_video_frames=5, sketch_max_video_frames=2,
                 sketch_enc_depth=6, sketch_enc_dim_head=64, sketch_enc_heads=8, sketch_enc_use_sparse_3dna=False,
                 enc_reversible=False, dec_depth=6, dec_dim_head=64, dec_heads=8, dec_reversible=False, attn_dropout=0.,
                 ff_dropout=0., ff_chunk_size=None, embed_gradient_frac=0.2, shift_video_tokens=True,
                 cross_2dna_kernel_size=3, cross_2dna_dilation=1, sparse_3dna_kernel_size=3, sparse_3dna_dilation=1,
                 sparse_3dna_query_num_frames_chunk=None):
        super().__init__()
        self.image_size = image_size
        self.sketch_vae = sketch_vae
        sketch_fmap_size = image_size // (2 ** sketch_vae.num_layers)
        sketch_shape = (sketch_max_video_frames, sketch_fmap_size, sketch_fmap_size)
        self.sketch_max_video_frames = sketch_max_video_frames
        self.sketch_embedding = Embedding(sketch_vae.codebook_size, dim, frac_gradient=embed_gradient_frac)
        self.sketch_pos_emb = AxialPositionalEmbedding(dim, shape=sketch_shape)
        dilations = tuple(range(1, sparse_3dna_dilation + 1)) if not isinstance(sparse_3dna_dilation, (list, tuple)) \
            else tuple(sparse_3dna_dilation)
        enc_klass = Transformer if not enc_reversible else ReversibleTransformer
        self.sketch_transformer = enc_klass(dim=dim, depth=sketch_enc_depth, heads=sketch_enc_heads,
                                            dim_head=sketch_enc_dim_head, attn_dropout=attn_dropout, ff_dropout=ff_dropout,
                                            shift_video_tokens=shift_video_tokens, sparse_3dna_video_shape=sketch_shape,
                                            sparse_3dna_kernel_size=sparse_3dna_kernel_size, sparse_3dna_dilations=dilations,
                                            sparse_3dna_query_num_frames_chunk=sparse_3dna_query_num_frames_chunk,
                                            sparse_3dna_attn=sketch_enc_use_sparse_3dna)
        self.vae = vae.copy_for_eval()
        self.video_bos = nn.Parameter(torch.randn(dim))
        self.image_embedding = Embedding(vae.codebook_size, dim, frac_gradient=embed_gradient_frac)
        fmap_size = image_size // (2 ** vae.num_layers)
        assert fmap_size == sketch_fmap_size, 'feature map size of video must be equal to the feature map size of sketches (VAEs must have same number of layers)'
        self.video_fmap_size = fmap_size
        self.max_video_frames = max_video_frames
        video_shape = (max_video_frames, fmap_size, fmap_size)
        self.video_pos_emb = AxialPositionalEmbedding(dim, shape=video_shape)
        cdil = tuple(range(1, cross_2dna_dilation + 1)) if not isinstance(cross_2dna_dilation, (list, tuple)) \
            else tuple(cross_2dna_dilation)
        dec_klass = Transformer if not dec_reversible else ReversibleTransformer
        self.video_transformer = dec_klass(dim=dim, depth=dec_depth, heads=dec_heads, dim_head=dec_dim_head, causal=True,
                                           cross_attend=True, cross_2dna_attn=True, cross_2dna_image_size=fmap_size,
                                           cross_2dna_kernel_size=cross_2dna_kernel_size, cross_2dna_dilations=cdil,
                                           attn_dropout=attn_dropout, ff_dropout=ff_dropout, ff_chunk_size=ff_chunk_size,
                                           shift_video_tokens=shift_video_tokens, sparse_3dna_video_shape=video_shape,
                                           sparse_3dna_kernel_size=sparse_3dna_kernel_size, sparse_3dna_dilations=dilations,
                                           sparse_3dna_query_num_frames_chunk=sparse_3dna_query_num_frames_chunk,
                                           sparse_3dna_attn=True)
        self.to_logits = nn.Linear(dim, vae.codebook_size, bias=False)

    def _embed_sketch_indices(self, sketch_indices, mask, want_bf16=False):
        """embed_sketch after tokenisation (:2420-2436): sketch_indices (b, f, h, w) int64, mask (b, f) bool | None."""
        b, f = sketch_indices.shape[:2]
        idx = sketch_indices.reshape(b, -1).contiguous()
        n = idx.shape[1]
        tokens = ops.embed_tokens(idx, self.sketch_embedding.embed.weight.detach().float().contiguous(), nt=n,
                                  axials=self.sketch_pos_emb.tables(), dims=self.sketch_pos_emb.full_shape)
        if _exists(mask):
            tok_mask = mask[:, :, None].expand(b, f, n // f).reshape(b, n)
        else:
            tok_mask = torch.ones((b, n), dtype=torch.bool, device=idx.device)
        km = tok_mask.to(torch.uint8).contiguous()
        out = engine.run_stack(self.sketch_transformer, tokens, key_mask=km, want_bf16=want_bf16)
        return out, tok_mask

    def embed_sketch(self, sketch, mask=None):
        batch, frames = sketch.shape[:2]
        if _exists(mask):
            assert mask.shape[:2] == (batch, frames), 'sketch mask must be in shape of (batch x frame)'
        with torch.no_grad():
            sketch_indices = self.sketch_vae.get_video_indices(sketch)
            return self._embed_sketch_indices(sketch_indices, mask)

    def _sketch_context(self, sketch, mask):
        batch, frames = sketch.shape[:2]
        if _exists(mask):
            assert mask.shape[:2] == (batch, frames), 'sketch mask must be in shape of (batch x frame)'
        sketch_indices = self.sketch_vae.get_video_indices(sketch)
        (_, e16), tok_mask = self._embed_sketch_indices(sketch_indices, mask, want_bf16=True)
        return engine.Context(e16, tok_mask.to(torch.uint8).contiguous())

    @torch.no_grad()
    @_eval_decorator
    def generate(self, *, sketch, sketch_mask=None, filter_thres=0.9, temperature=1., decode_max_batchsize=10,
                 cond_scale=2., num_frames=None, _noise=None, _return_indices=False, _use_graph=True):
        batch = sketch.shape[0]
        context = self._sketch_context(sketch, sketch_mask)
        num_frames = num_frames if _exists(num_frames) else self.max_video_frames
        idx = self._generate_indices(context, batch, num_frames=num_frames, filter_thres=filter_thres,
                                     temperature=temperature, cond_scale=cond_scale, noise=_noise, use_graph=_use_graph)
        if _return_indices:
            return idx
        return self._indices_to_video(idx, decode_max_batchsize)

    def forward(self, *, sketch, sketch_mask=None, video=None, return_loss=False, cond_dropout_prob=0.2):
        if sketch.ndim == 4:
            sketch = sketch[:, None]
        batch, sketch_frames, sketch_channels, sketch_image_size, _ = sketch.shape
        frames = video.shape[1]
        assert sketch_image_size == self.image_size, 'sketch image size must be equal'
        assert sketch_frames <= self.sketch_max_video_frames, 'sketch frames must be less than max sketch video frames'
        train_path = return_loss and torch.is_grad_enabled() and self.to_logits.weight.requires_grad
        with torch.no_grad():
            if not train_path:
                context = self._sketch_context(sketch, sketch_mask)
            else:  # training step (train.py): keep the sketch token ids, the sketch encoder runs with a tape
                if _exists(sketch_mask):
                    assert sketch_mask.shape[:2] == (batch, sketch_frames), 'sketch mask must be in shape of (batch x frame)'
                sketch_indices = self.sketch_vae.get_video_indices(sketch)
                per_frame = sketch_indices[0, 0].numel()
                tok_mask = (sketch_mask[:, :, None].expand(batch, sketch_frames, per_frame).reshape(batch, -1)
                            if _exists(sketch_mask) else
                            torch.ones((batch, sketch_frames * per_frame), dtype=torch.bool, device=sketch.device))
                tok_mask = tok_mask.to(torch.uint8).contiguous()
        assert frames == self.max_video_frames, f'you must give the full video frames ({self.max_video_frames}) during training'
        frame_indices = self.vae.get_video_indices(video).reshape(batch, -1)
        if self.training and cond_dropout_prob > 0:
            # reference: `sketch_mask *= ...` raises TypeError when sketch_mask is None and otherwise has no effect on
            # this call because the decoder mask was already derived
            if sketch_mask is None:
                raise TypeError("unsupported operand type(s) for *=: 'NoneType' and 'Tensor'")
            uncond = torch.zeros((batch,), device=sketch.device).float().uniform_(0, 1) < cond_dropout_prob
            sketch_mask *= ~uncond[:, None]
        if train_path:
            from . import train
            return train.sketch_training_loss(self, sketch_indices, tok_mask, frame_indices)
        with torch.no_grad():
            return self._decoder_logits(frame_indices, context, return_loss)
