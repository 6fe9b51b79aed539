"""Stack executor: walks a Transformer / ReversibleTransformer parameter container and runs it as a
sequence of C-ABI kernel launches (no torch arithmetic).

Data layout (B200-first):
  * residual streams are fp32 [B*n, D] in HBM; every GEMM operand is bf16,
  * one fused row kernel does  post-LayerNorm + residual add  of sub-block i  AND  pre-LayerNorm +
    ShiftVideoTokens scatter + bf16 cast of sub-block i+1  (sandwich_ln),
  * q|k|v live in ONE bf16 buffer [B, npos, 3*inner] written by a single fused projection GEMM; in decode
    mode that buffer is the KV cache and the shift-scatter operand buffers are persistent, so a generate()
    step touches only the new token's rows.
Two execution modes: full (all n positions, teacher forced) and decode (one position, with DecodeState).
"""
import ctypes
import os

import torch

from . import _lib, ops


def _f32(t):
    return t.detach().float().contiguous()


def _bf16(t):
    return t.detach().to(torch.bfloat16).contiguous()


class SubBlock:
    """Packed weights + static geometry of one SandwichNorm-wrapped sub-block.

    fp32 tensors (LayerNorm weights, biases, talking heads, null k/v) alias the live parameters.  The bf16 GEMM
    operands are COPIES; each is registered with the function that rebuilds it, so `refresh()` can re-derive every copy
    IN PLACE after an optimizer step: addresses stay fixed for the lifetime of the parameters, which is what lets a
    captured CUDA graph (graphs.GraphedTrainStep) keep replaying across weight updates."""

    def __init__(self, kind, sandwich, inner_mod, shift, read, write, **geom):
        self.kind, self.shift, self.read, self.write = kind, shift, read, write
        self.sandwich, self.mod = sandwich, inner_mod  # parameter owners (train.py maps gradients back to them)
        self.pre = (_f32(sandwich.prenorm.weight), _f32(sandwich.prenorm.bias))
        self.post = (_f32(sandwich.postnorm.weight), _f32(sandwich.postnorm.bias))
        self.__dict__.update(geom)
        self._builders = []
        self._bw = None
        m = inner_mod
        if kind in ('3dna', 'self'):
            self.H = m.heads
            self.inner = m.to_q.weight.shape[0]
            self.dh = self.inner // self.H
            self._packed('w_qkv', lambda: _bf16(torch.cat([m.to_q.weight, m.to_kv.weight], dim=0)))
            self._packed('w_out', lambda: _bf16(m.to_out.weight))
            self.b_out = _f32(m.to_out.bias) if m.to_out.bias is not None else None
            self.talk = _f32(m.talking_heads.weight.reshape(self.H, self.H))
        if kind in ('cross', 'x2dna'):
            self.H = m.heads
            self.inner = m.to_q.weight.shape[0]
            self.dh = self.inner // self.H
            self._packed('w_q', lambda: _bf16(m.to_q.weight))
            self._packed('w_kv', lambda: _bf16(m.to_kv.weight))
            self._packed('w_out', lambda: _bf16(m.to_out.weight))
            self.talk = _f32(m.talking_heads.weight.reshape(self.H, self.H))
        if kind in ('self', 'cross', 'x2dna'):
            self.null_k, self.null_v = _f32(m.null_k.reshape(-1)), _f32(m.null_v.reshape(-1))
        if kind == 'ff':
            w1, w2 = m.net[0].weight, m.net[3].weight
            self.ff_inner = w2.shape[1]
            self._packed('w1', lambda: _bf16(ops.pack_pairs(w1.detach().float())))   # (2*ip, D) pair packed
            ip = self.w1.shape[0] // 2

            def build_w2():
                w2p = torch.zeros(w2.shape[0], ip, dtype=torch.float32, device=w2.device)
                w2p[:, :self.ff_inner] = w2.detach().float()
                return _bf16(w2p)                                                     # (D, ip), zero padded K
            self._packed('w2', build_w2)

    def _packed(self, name, build):
        setattr(self, name, build())
        self._builders.append((name, build))

    def refresh(self):
        """Re-derive every bf16 copy from the fp32 masters, in place (capture safe: device copies only)."""
        for name, build in self._builders:
            getattr(self, name).copy_(build())
        if self._bw is not None:
            for name, build in self._bw_builders:
                self._bw[name].copy_(build())


class StackPack:
    """Everything run_stack needs, built once per parameter allocation (see pack_stack)."""

    def __init__(self, subs, norm_w, norm_b, reversible, dim):
        self.subs, self.norm_w, self.norm_b, self.reversible, self.dim = subs, norm_w, norm_b, reversible, dim

    def refresh(self):
        for s in self.subs:
            s.refresh()


def _signature(module):
    return (_lib.WEIGHTS_EPOCH[0],) + tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))


def _addresses(sig):
    return tuple(a for a, _ in sig[1:])


def pack_stack(stack):
    """Build (or fetch the cached) StackPack of a Transformer / ReversibleTransformer container."""
    sig = _signature(stack)
    cached = getattr(stack, '_nuwa_pack', None)
    if cached is not None and cached[0] == sig:
        return cached[1]
    if cached is not None and _addresses(cached[0]) == _addresses(sig):
        # same parameter storage, new values (optimizer step): refresh the bf16 copies in place -- a captured CUDA
        # graph holding their addresses stays valid and sees the new weights
        with torch.no_grad():
            cached[1].refresh()
        stack._nuwa_pack = (sig, cached[1])
        return cached[1]
    dev = next(stack.parameters()).device
    if dev.type != 'cuda':
        raise ops._lib.NuwaB200Error('transformer stacks must live on a CUDA device (no CPU path)')
    from .nuwa import Attention, FeedForward, ShiftVideoTokens, Sparse3DNA, SparseCross2DNA
    subs = []

    def add(sandwich, read, write):
        fn = sandwich.fn
        shift = False
        if isinstance(fn, ShiftVideoTokens):
            shift = fn.shift_space
            assert not fn.shift_time, 'shift_time is never enabled by NUWA / NUWASketch'
            fmap = fn.image_size
            fn = fn.fn
        else:
            fmap = None
        if isinstance(fn, Sparse3DNA):
            assert fn.rel_pos_bias is None, 'sparse_3dna_rel_pos_bias is broken in the reference for batch > 1 (SURVEY D11)'
            subs.append(SubBlock('3dna', sandwich, fn, shift, read, write, fmap=fn.video_shape[1],
                                 max_frames=fn.video_shape[0], kernel=fn.kernel_size, dilation=fn.dilation,
                                 causal=fn.causal))
        elif isinstance(fn, SparseCross2DNA):
            subs.append(SubBlock('x2dna', sandwich, fn, shift, read, write, fmap=fn.image_size, ck=fn.kernel_size,
                                 cdil=fn.dilation))
        elif isinstance(fn, Attention):
            subs.append(SubBlock('self' if not getattr(sandwich, '_is_cross', False) else 'cross', sandwich, fn, shift,
                                 read, write, fmap=fmap, causal=fn.causal))
        elif isinstance(fn, FeedForward):
            subs.append(SubBlock('ff', sandwich, fn, shift, read, write, fmap=fmap))
        else:
            raise TypeError(f'unsupported sub-block {type(fn)}')

    reversible = hasattr(stack, 'net')
    if not reversible:
        for attn, cross, ff in stack.layers:
            add(attn, 0, 0)
            if cross is not None:
                add(cross, 0, 0)
            add(ff, 0, 0)
    else:
        for f, g in stack.layers:  # y1 = x1 + f(x2) ; y2 = x2 + g(y1)   (reversible.py:61-68)
            add(f, 1, 0)
            add(g, 0, 1)
    ln = stack.norm.norm
    pack = StackPack(subs, _f32(ln.weight), _f32(ln.bias), reversible, ln.weight.shape[0])
    stack._nuwa_pack = (sig, pack)
    return pack


class DecodeState:
    """Persistent per-stack buffers for incremental decoding of `npos` positions (bos + video tokens)."""

    def __init__(self, pack, B, npos, device, t_dev=None):
        self.B, self.npos = B, npos
        # current position as a device-side int32 scalar (shared by the two sweeps of a generate step)
        self.t_dev = t_dev if t_dev is not None else torch.zeros(1, dtype=torch.int32, device=device)
        self.a, self.qkv, self.ctx_kv = {}, {}, {}
        for i, s in enumerate(pack.subs):
            if s.shift:
                self.a[i] = torch.zeros(B, npos, pack.dim, dtype=torch.bfloat16, device=device)
            if s.kind == '3dna':
                self.qkv[i] = torch.zeros(B, npos, 3 * s.inner, dtype=torch.bfloat16, device=device)


class Context:
    """Cross-attention context: bf16 copy of the context tokens + uint8 key mask; per-layer K/V are cached."""

    def __init__(self, ctx16, mask_u8):
        self.ctx16, self.mask = ctx16, mask_u8  # (B, nk, D) bf16 ; (B, nk) uint8 or None
        self.kv = {}
        self.kv_packed = {}  # per-(sample, head) repack of kv for the persistent decode kernel (FusedDecode)

    def with_mask(self, mask_u8):
        c = Context(self.ctx16, mask_u8)
        c.kv, c.kv_packed = self.kv, self.kv_packed  # K/V do not depend on the mask
        return c


def prime_context(stack, context):
    """Project the context to K/V for every cross-attention sub-block now (they are constant over a generate() call),
    so that a decode step allocates nothing context dependent and can be captured in a CUDA graph."""
    pack = pack_stack(stack)
    B, nk = context.ctx16.shape[:2]
    for i, s in enumerate(pack.subs):
        if s.kind in ('cross', 'x2dna') and i not in context.kv:
            context.kv[i] = ops.gemm(context.ctx16.view(B * nk, -1), s.w_kv, out_dtype=torch.bfloat16)


def _run_sub(i, s, a, B, nt, t0, npos, D, state, context, key_mask, rotary):
    """a: bf16 operand rows as a 2-D (B*nt, D) view (row stride may exceed D in decode mode).  Returns y fp32 (B*nt, D)."""
    dev = a.device
    M = B * nt
    if s.kind == 'ff':
        h = ops.gemm(a, s.w1, act='geglu', out_dtype=torch.bfloat16)
        return ops.gemm(h, s.w2, out_dtype=torch.float32)
    inner, H, dh = s.inner, s.H, s.dh
    o = torch.empty(M, inner, dtype=torch.bfloat16, device=dev)
    if s.kind == '3dna':
        if state is None:
            qkv = ops.gemm(a, s.w_qkv, out_dtype=torch.bfloat16)  # (B*n, 3*inner)
            ops.attn_sparse3dna(qkv, o, B=B, nq=nt, t0=0, npos=nt, H=H, dh=dh, talk=s.talk, fmap=s.fmap,
                                max_frames=s.max_frames, nv=nt - 1, kernel=s.kernel, dilation=s.dilation, causal=s.causal)
        else:
            # decode step: every address is position independent (the position lives in state.t_dev on the device),
            # so the whole step can be captured once in a CUDA graph and replayed per token
            cache = state.qkv[i]
            row = ops.gemm(a, s.w_qkv, out_dtype=torch.bfloat16)     # (B, 3*inner): the new token's q|k|v
            ops.cache_append(row, cache, state.t_dev)                 # KV cache row t
            ops.attn_sparse3dna_decode(row, cache, o, state.t_dev, B=B, npos=npos, H=H, dh=dh, talk=s.talk, fmap=s.fmap,
                                       max_frames=s.max_frames, kernel=s.kernel, dilation=s.dilation, causal=s.causal)
        return ops.gemm(o, s.w_out, bias=s.b_out, out_dtype=torch.float32)
    if s.kind == 'self':
        assert state is None, 'dense self-attention stacks (text encoder) run in full mode only'
        assert not s.causal, 'dense causal self-attention is not used by NUWA / NUWASketch'
        if rotary is not None:
            inv_freq, rot = rotary
            qkv32 = ops.gemm(a, s.w_qkv, out_dtype=torch.float32)
            qkv = ops.rotary_to_bf16(qkv32, inv_freq, nt, H, dh, rot)
        else:
            qkv = ops.gemm(a, s.w_qkv, out_dtype=torch.bfloat16)
        base = qkv.data_ptr()
        ops.attn_dense(base, base + inner * 2, base + 2 * inner * 2, o, B=B, nq=nt, nk=nt, H=H, dh=dh,
                       q_bs=nt * 3 * inner, q_rs=3 * inner, k_bs=nt * 3 * inner, k_rs=3 * inner, v_bs=nt * 3 * inner,
                       v_rs=3 * inner, o_bs=nt * inner, o_rs=inner, talk=s.talk, null_k=s.null_k, null_v=s.null_v,
                       key_mask=key_mask)
        return ops.gemm(o, s.w_out, out_dtype=torch.float32)
    # ---- cross attention (dense text context or sparse 2-D nearby sketch context) ----
    nk = context.ctx16.shape[1]
    kv = context.kv.get(i)
    if kv is None:
        kv = ops.gemm(context.ctx16.view(B * nk, -1), s.w_kv, out_dtype=torch.bfloat16)  # (B*nk, 2*inner), once per call
        context.kv[i] = kv
    q = ops.gemm(a, s.w_q, out_dtype=torch.bfloat16)  # (B*nt, inner)
    kb = kv.data_ptr()
    common = dict(B=B, H=H, dh=dh, q_bs=nt * inner, q_rs=inner, k_bs=nk * 2 * inner, k_rs=2 * inner,
                  v_bs=nk * 2 * inner, v_rs=2 * inner, o_bs=nt * inner, o_rs=inner, null_k=s.null_k, null_v=s.null_v,
                  key_mask=context.mask)
    if s.kind == 'cross':
        ops.attn_dense(q.data_ptr(), kb, kb + inner * 2, o, nq=nt, nk=nk, talk=s.talk, **common)
    else:
        first = 0
        if t0 == 0:  # bos query: dense over [null] + every context token, no talking heads (nuwa_pytorch.py:828-844)
            ops.attn_dense(q.data_ptr(), kb, kb + inner * 2, o, nq=1, nk=nk, talk=None, **common)
            first = 1
        if nt - first > 0:
            ops.attn_cross2dna(q.data_ptr() + first * inner * 2, kb, kb + inner * 2, o.data_ptr() + first * inner * 2,
                               nq=nt - first, t0=t0 + first, talk=s.talk, fmap=s.fmap, frames=nk // (s.fmap * s.fmap),
                               ck=s.ck, cdil=s.cdil, **common)
    return ops.gemm(o, s.w_out, out_dtype=torch.float32)


def run_stack(stack, x, *, context=None, key_mask=None, rotary=None, state=None, t0=0, want_bf16=False):
    """Run a whole stack.  x: fp32 (B, nt, D) contiguous (nt = all positions, or 1 with a DecodeState).
    context: Context or None; key_mask: uint8 (B, n) for dense self-attention; rotary: (inv_freq, rot_dim).
    Returns the StableLayerNorm output fp32 (B, nt, D) [and its bf16 copy]."""
    pack = pack_stack(stack)
    B, nt, D = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    npos = state.npos if state is not None else nt
    assert state is None or nt == 1
    if pack.reversible:
        streams = [x.view(B * nt, D).clone(), x.view(B * nt, D).clone()]  # X = [x, x]  (reversible.py:133)
    else:
        streams = [x.view(B * nt, D).clone()]
    subs = pack.subs

    t_dev = state.t_dev if state is not None else None

    def operand_buffer(i):
        """bf16 operand buffer of sub-block i and its addressing for sandwich_ln stage B."""
        s = subs[i]
        buf = torch.empty(B, nt, D, dtype=torch.bfloat16, device=x.device)
        if state is not None:
            # decode: dense fixed-address operand row; the token shift is a gather from the persistent shift cache
            addr = dict(a_out=buf, a_bs=D, a_rs=D, a_t0=0, a_npos=npos, gather=True, t_dev=t_dev)
            if s.shift:
                addr.update(shift_cache=state.a[i], sc_bs=npos * D)
            return buf, addr, buf.view(B, D)
        return buf, dict(a_out=buf, a_bs=nt * D, a_rs=D, a_t0=t0, a_npos=t0 + nt), buf.view(B * nt, D)

    # first sub-block: pre-norm only
    _, addr, a_view = operand_buffer(0)
    ops.sandwich_ln(B, nt, D, res_in=streams[subs[0].read], pre=subs[0].pre, shift=subs[0].shift,
                    fmap=subs[0].fmap or 0, t0=t0, **addr)
    for i, s in enumerate(subs):
        y = _run_sub(i, s, a_view, B, nt, t0, npos, D, state, context, key_mask, rotary)
        tgt = streams[s.write]
        if i + 1 < len(subs):
            nxt = subs[i + 1]
            assert nxt.read == s.write
            _, addr, a_view = operand_buffer(i + 1)
            ops.sandwich_ln(B, nt, D, y=y, post=s.post, res_in=tgt, x_out=tgt, pre=nxt.pre, shift=nxt.shift,
                            fmap=nxt.fmap or 0, t0=t0, **addr)
        else:
            ops.sandwich_ln(B, nt, D, y=y, post=s.post, res_in=tgt, x_out=tgt, t0=t0, t_dev=t_dev)
    b2 = streams[1] if pack.reversible else None
    o32, o16 = ops.stable_ln(streams[0], pack.norm_w, pack.norm_b, b2=b2, want_f32=True, want_bf16=want_bf16)
    o32 = o32.view(B, nt, D)
    return (o32, o16.view(B, nt, D)) if want_bf16 else o32


# ------------------------------------------------------------------------------------------------
# persistent decode step: the whole stack of one token step in ONE cooperative kernel (csrc/decode_stack.cu)
# ------------------------------------------------------------------------------------------------
_DEC_KIND = {'3dna': 0, 'cross': 1, 'ff': 2}


class FusedDecode:
    """Descriptor table + scratch for `nuwa_decode_stack` over one DecodeState.  `supported()` tells whether the stack
    fits the kernel's envelope (Sparse3DNA / dense cross-attention / FeedForward sub-blocks with one geometry,
    B <= 16, D <= 1024); other stacks (NUWASketch's sparse 2-D cross attention) keep the per-kernel decode path."""

    @staticmethod
    def supported(pack, B, context):
        if os.environ.get('NUWA_DECODE_FUSED', '1') == '0' or B > 16 or pack.dim % 32 or pack.dim > 1024:
            return False
        geo, heads = set(), set()
        for s in pack.subs:
            if s.kind not in _DEC_KIND:
                return False
            if s.kind == '3dna':
                geo.add((s.fmap, s.max_frames, bool(s.causal)))  # kernel / dilation may differ per layer
                if s.b_out is None:
                    return False
            if s.kind == 'ff' and s.shift and s.fmap is None:
                return False
            if s.kind != 'ff':
                heads.add((s.H, s.dh))
            if s.kind == 'cross' and context is None:
                return False
        if len(geo) != 1 or len(heads) != 1:
            return False
        H, dh = next(iter(heads))
        return H <= 16 and dh % 8 == 0 and H * dh <= 1024

    def __init__(self, pack, state, context, w_logits=None):
        dev = state.t_dev.device
        B, npos, D = state.B, state.npos, pack.dim
        self.pack, self.state, self.context = pack, state, context
        subs = (_lib.DecodeSub * len(pack.subs))()
        kmax, H, dh = D, None, None
        shift_fmap, j3max = None, 1
        for i, s in enumerate(pack.subs):
            d = subs[i]
            d.kind, d.shift, d.read, d.write = _DEC_KIND[s.kind], int(bool(s.shift)), s.read, s.write
            d.pre_w, d.pre_b, d.post_w, d.post_b = (t.data_ptr() for t in (*s.pre, *s.post))
            if s.shift:
                d.shift_cache = state.a[i].data_ptr()
                shift_fmap = s.fmap
            if s.kind == '3dna':
                H, dh = s.H, s.dh
                d.w_a, d.w_b, d.b_out, d.talk = s.w_qkv.data_ptr(), s.w_out.data_ptr(), s.b_out.data_ptr(), s.talk.data_ptr()
                d.cache = state.qkv[i].data_ptr()
                d.kt, d.kh, d.kw = s.kernel
                d.dt, d.dh_, d.dw = s.dilation
                j3max = max(j3max, 1 + s.kernel[0] * s.kernel[1] * s.kernel[2])
                geom = s
            elif s.kind == 'cross':
                H, dh = s.H, s.dh
                d.w_a, d.w_b, d.talk = s.w_q.data_ptr(), s.w_out.data_ptr(), s.talk.data_ptr()
                d.null_k, d.null_v = s.null_k.data_ptr(), s.null_v.data_ptr()
                # the kernel stages one (sample, head) slice per CTA: repack K|V rows [B*nk, 2*inner] as [B][H][2][nk][dh]
                # once per generate() call so that each slice is a single contiguous bulk copy
                packed = context.kv_packed
                if i not in packed:
                    nk_ = context.ctx16.shape[1]
                    packed[i] = context.kv[i].view(B, nk_, 2, s.H, s.dh).permute(0, 3, 2, 1, 4).contiguous()
                d.cache = packed[i].data_ptr()
            else:
                d.w_a, d.w_b, d.ip = s.w1.data_ptr(), s.w2.data_ptr(), s.w1.shape[0] // 2
                kmax = max(kmax, d.ip)
        inner = H * dh
        kmax = max(kmax, inner)
        raw = bytes(subs)
        self.subs_dev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        nk = context.ctx16.shape[1] if context is not None else 0
        self.out_f32 = torch.empty(B, 1, D, dtype=torch.float32, device=dev)
        self.out_bf16 = torch.empty(B, 1, D, dtype=torch.bfloat16, device=dev)
        self.y = torch.empty(B, D, dtype=torch.float32, device=dev)
        self.act = torch.empty(B, kmax, dtype=torch.bfloat16, device=dev)
        self.actq = torch.empty(B, inner, dtype=torch.bfloat16, device=dev)
        self.scores = torch.empty(B, H, nk + 1, dtype=torch.float32, device=dev)
        self.barrier = torch.zeros(2, dtype=torch.int32, device=dev)
        self.w_logits = w_logits
        self.logits = torch.empty(B, w_logits.shape[0], dtype=torch.float32, device=dev) if w_logits is not None else None
        p = _lib.DecodeParams()
        p.subs, p.nsubs = self.subs_dev.data_ptr(), len(pack.subs)
        p.B, p.D, p.H, p.dh, p.npos, p.reversible = B, D, H, dh, npos, int(pack.reversible)
        p.fmap, p.max_frames, p.causal, p.j3max = geom.fmap, geom.max_frames, int(bool(geom.causal)), j3max
        assert shift_fmap is None or shift_fmap == geom.fmap
        p.nk = nk
        if context is not None and context.mask is not None:
            p.key_mask, p.mask_bs = context.mask.data_ptr(), context.mask.shape[1]
        p.t_ptr = state.t_dev.data_ptr()
        p.norm_w, p.norm_b = pack.norm_w.data_ptr(), pack.norm_b.data_ptr()
        p.out_f32, p.out_bf16 = self.out_f32.data_ptr(), self.out_bf16.data_ptr()
        if w_logits is not None:
            p.w_logits, p.V, p.logits = w_logits.data_ptr(), w_logits.shape[0], self.logits.data_ptr()
        p.y, p.act, p.actq, p.scores = self.y.data_ptr(), self.act.data_ptr(), self.actq.data_ptr(), self.scores.data_ptr()
        p.barrier = self.barrier.data_ptr()
        p.kmax = kmax
        p.split_small = int(os.environ.get('NUWA_DECODE_SPLIT_SMALL', '0'))
        p.split_ff = int(os.environ.get('NUWA_DECODE_SPLIT_FF', '0'))
        p.split_logits = int(os.environ.get('NUWA_DECODE_SPLIT_LOGITS', '0'))
        p.max_ctas = int(os.environ.get('NUWA_DECODE_MAX_CTAS', '0'))
        p.debug_flags = int(os.environ.get('NUWA_DECODE_DEBUG', '0'))
        self.params = p
        self.cooperative = int(os.environ.get('NUWA_DECODE_COOP', '1'))

    def run(self, x):
        """x: fp32 (B, 1, D) contiguous.  Returns (normalised output fp32 (B,1,D), logits (B,V) or None); both are
        persistent buffers of this plan, overwritten by the next call."""
        assert x.dtype == torch.float32 and x.is_contiguous() and x.numel() == self.state.B * self.pack.dim
        self.params.x_in = x.data_ptr()
        _lib.check(_lib.lib().nuwa_decode_stack(ctypes.byref(self.params), self.cooperative, _lib.stream()),
                   "nuwa_decode_stack")
        return self.out_f32, self.logits
