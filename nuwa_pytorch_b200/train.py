"""Training step of the NUWA decoder path: `loss = nuwa(text=, video=, return_loss=True); loss.backward()`
(nuwa_pytorch.py:1917-1964; SURVEY §8d cfg 3) as hand-written CUDA forward + backward behind ONE autograd node.

Layout (B200-first): activations needed by the backward are simply kept (a cfg-3 step saves ~7 GB of 180 GB), so
nothing is recomputed except cheap row statistics; the reversible stacks (reversible.py:54-123) are differentiated
through the same saved-activation tape -- algebraically the gradient the reference obtains by reconstructing
activations, without its extra forward pass.  Every weight gradient is accumulated in fp32 inside one flat buffer
(one memset, one optional NCCL all-reduce per bucket, see parallel.py); the views handed to autograd alias it.

Per sub-block  x' = x + LN_post(fn(shift(LN_pre(x_r)))):
    dy   = LN_post backward of the stream gradient                     (ln_bwd, bf16 out; bias grad = column sum)
    da   = fn backward (GEMM dgrad on tcgen05, weight grads as split-K tcgen05 GEMMs reading dY and X contraction-major,
           attention backward kernels)
    G_r += LN_pre backward of da through the inverse token shift        (ln_bwd, accumulate)
"""
import torch

from . import engine, ops, ops_bwd
from .ops_bwd import _round_up


class GradStore:
    """fp32 gradient accumulators of a fixed parameter list, carved out of one flat zero-initialised buffer."""

    def __init__(self, params):
        self.params = list(params)
        dev = self.params[0].device
        sizes = [_round_up(p.numel(), 4) for p in self.params]  # 16-byte aligned views (vectorised atomics / float4)
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + s)
        self.flat = torch.zeros(self.offsets[-1], dtype=torch.float32, device=dev)
        self.index = {id(p): i for i, p in enumerate(self.params)}

    def __call__(self, p):
        i = self.index.get(id(p))
        if i is None:
            return None
        return self.flat[self.offsets[i]:self.offsets[i] + p.numel()].view(p.shape)

    def view2d(self, p, rows):
        g = self(p)
        return None if g is None else g.view(rows, -1)

    def grads(self):
        return [self(p) for p in self.params]

    def span(self, params):
        """[lo, hi) of the flat buffer covering `params` (contiguous when they are consecutive in the parameter list)."""
        ids = [self.index[id(p)] for p in params if id(p) in self.index]
        if not ids:
            return 0, 0
        return self.offsets[min(ids)], self.offsets[max(ids) + 1]


def _grad_store_of(model, params):
    """(store, persistent): the model's persistent gradient accumulator (set by trainer.TrainStep; every weight-gradient
    kernel ADDS into the flat buffer, so micro-batches accumulate without a per-parameter add) or a fresh zeroed store."""
    store = getattr(model, '_grad_store', None)
    if store is not None:
        assert len(store.params) == len(params) and all(a is b for a, b in zip(store.params, params)), \
            'the persistent gradient store was built for a different parameter list'
        return store, True
    return GradStore(params), False


def _bw(s):
    """Transposed bf16 weight copies for the dgrad GEMMs (built once per parameter allocation, cached on the SubBlock and
    refreshed in place by SubBlock.refresh() whenever the weights change)."""
    if getattr(s, '_bw', None) is None:
        t = ops_bwd.transpose
        names = {'3dna': ('w_qkv', 'w_out'), 'self': ('w_qkv', 'w_out'), 'cross': ('w_q', 'w_kv', 'w_out'),
                 'x2dna': ('w_q', 'w_kv', 'w_out'), 'ff': ('w1', 'w2')}[s.kind]
        builders = [(n + '_t', (lambda n=n: t(getattr(s, n)).contiguous())) for n in names]
        bw = {name: build() for name, build in builders}
        if s.kind == 'ff':
            inner, ip = s.ff_inner, s.w1.shape[0] // 2
            # packed row r of w1 -> row of net[0].weight (value rows [0, inner), gate rows [inner, 2*inner)), -1 = padding
            r = torch.arange(2 * ip)
            blk, j = r // 32, r % 32
            col = blk * 16 + (j % 16)
            src = torch.where(j < 16, col, col + inner)
            src = torch.where(col < inner, src, torch.full_like(src, -1))
            bw['w1_map'] = src.to(torch.int32).to(s.w1.device)
        s._bw_builders = builders
        s._bw = bw
    return s._bw


# ------------------------------------------------------------------------------------------------
# forward with tape
# ------------------------------------------------------------------------------------------------
def _sub_forward(i, s, a, B, nt, rec, context, key_mask, rotary):
    """a: bf16 (B*nt, D).  Returns y fp32 (B*nt, D); stores what the backward needs in rec."""
    M = B * nt
    dev = a.device
    if s.kind == 'ff':
        rec['h'] = ops.gemm(a, s.w1, out_dtype=torch.bfloat16)          # pair-packed pre-activation (M, 2*ip)
        rec['g'] = ops_bwd.geglu_fwd(rec['h'])
        return ops.gemm(rec['g'], s.w2, out_dtype=torch.float32)
    inner, H, dh = s.inner, s.H, s.dh
    o = torch.empty(M, inner, dtype=torch.bfloat16, device=dev)
    rec['o'] = o
    if s.kind == '3dna':
        qkv = ops.gemm(a, s.w_qkv, out_dtype=torch.bfloat16)
        rec['qkv'] = qkv
        ops.attn_sparse3dna(qkv, o, B=B, nq=nt, t0=0, npos=nt, H=H, dh=dh, talk=s.talk, fmap=s.fmap,
                            max_frames=s.max_frames, nv=nt - 1, kernel=s.kernel, dilation=s.dilation, causal=s.causal)
        return ops.gemm(o, s.w_out, bias=s.b_out, out_dtype=torch.float32)
    if s.kind == 'self':
        assert not s.causal
        if rotary is not None:
            inv_freq, rot = rotary
            qkv = ops.rotary_to_bf16(ops.gemm(a, s.w_qkv, out_dtype=torch.float32), inv_freq, nt, H, dh, rot)
        else:
            qkv = ops.gemm(a, s.w_qkv, out_dtype=torch.bfloat16)
        rec['qkv'] = qkv
        base = qkv.data_ptr()
        ops.attn_dense(base, base + inner * 2, base + 2 * inner * 2, o, B=B, nq=nt, nk=nt, H=H, dh=dh,
                       q_bs=nt * 3 * inner, q_rs=3 * inner, k_bs=nt * 3 * inner, k_rs=3 * inner, v_bs=nt * 3 * inner,
                       v_rs=3 * inner, o_bs=nt * inner, o_rs=inner, talk=s.talk, null_k=s.null_k, null_v=s.null_v,
                       key_mask=key_mask)
        return ops.gemm(o, s.w_out, out_dtype=torch.float32)
    if s.kind == 'cross':
        nk = context.ctx16.shape[1]
        kv = ops.gemm(context.ctx16.view(B * nk, -1), s.w_kv, out_dtype=torch.bfloat16)
        q = ops.gemm(a, s.w_q, out_dtype=torch.bfloat16)
        rec['q'], rec['kv'] = q, kv
        kb = kv.data_ptr()
        ops.attn_dense(q.data_ptr(), kb, kb + inner * 2, o, B=B, nq=nt, nk=nk, H=H, dh=dh, q_bs=nt * inner, q_rs=inner,
                       k_bs=nk * 2 * inner, k_rs=2 * inner, v_bs=nk * 2 * inner, v_rs=2 * inner, o_bs=nt * inner,
                       o_rs=inner, talk=s.talk, null_k=s.null_k, null_v=s.null_v, key_mask=context.mask)
        return ops.gemm(o, s.w_out, out_dtype=torch.float32)
    if s.kind == 'x2dna':
        nk = context.ctx16.shape[1]
        kv = ops.gemm(context.ctx16.view(B * nk, -1), s.w_kv, out_dtype=torch.bfloat16)
        q = ops.gemm(a, s.w_q, out_dtype=torch.bfloat16)
        rec['q'], rec['kv'] = q, kv
        kb = kv.data_ptr()
        common = dict(B=B, H=H, dh=dh, q_bs=nt * inner, q_rs=inner, k_bs=nk * 2 * inner, k_rs=2 * inner,
                      v_bs=nk * 2 * inner, v_rs=2 * inner, o_bs=nt * inner, o_rs=inner, null_k=s.null_k, null_v=s.null_v,
                      key_mask=context.mask)
        # bos query: dense over [null] + every context token, no talking heads (nuwa_pytorch.py:828-844)
        ops.attn_dense(q.data_ptr(), kb, kb + inner * 2, o, nq=1, nk=nk, talk=None, **common)
        if nt > 1:
            ops.attn_cross2dna(q.data_ptr() + inner * 2, kb, kb + inner * 2, o.data_ptr() + inner * 2, nq=nt - 1, t0=1,
                               talk=s.talk, fmap=s.fmap, frames=nk // (s.fmap * s.fmap), ck=s.ck, cdil=s.cdil, **common)
        return ops.gemm(o, s.w_out, out_dtype=torch.float32)
    raise NotImplementedError(s.kind)


def stack_forward(stack, x, *, context=None, key_mask=None, rotary=None):
    """Teacher-forced pass of a whole stack that keeps the tape.  x: fp32 (B, nt, D).
    Returns (out fp32 (B,nt,D), out bf16, tape)."""
    pack = engine.pack_stack(stack)
    B, nt, D = x.shape
    M = B * nt
    x2d = x.view(M, D)
    streams = [x2d, x2d] if pack.reversible else [x2d]   # X = [x, x]  (reversible.py:133)
    subs = pack.subs
    tape = dict(B=B, nt=nt, D=D, recs=[], context=context, key_mask=key_mask, rotary=rotary)

    def operand():
        return torch.empty(B, nt, D, dtype=torch.bfloat16, device=x.device)

    a = operand()
    ops.sandwich_ln(B, nt, D, res_in=streams[subs[0].read], pre=subs[0].pre, shift=subs[0].shift, fmap=subs[0].fmap or 0,
                    a_out=a, a_bs=nt * D, a_rs=D, a_t0=0, a_npos=nt)
    for i, s in enumerate(subs):
        rec = dict(x_read=streams[s.read], a=a.view(M, D))
        y = _sub_forward(i, s, rec['a'], B, nt, rec, context, key_mask, rotary)
        rec['y'] = y
        new = torch.empty(M, D, dtype=torch.float32, device=x.device)
        if i + 1 < len(subs):
            nxt = subs[i + 1]
            assert nxt.read == s.write
            a = operand()
            ops.sandwich_ln(B, nt, D, y=y, post=s.post, res_in=streams[s.write], x_out=new, pre=nxt.pre, shift=nxt.shift,
                            fmap=nxt.fmap or 0, a_out=a, a_bs=nt * D, a_rs=D, a_t0=0, a_npos=nt)
        else:
            ops.sandwich_ln(B, nt, D, y=y, post=s.post, res_in=streams[s.write], x_out=new)
        streams[s.write] = new
        tape['recs'].append(rec)
    tape['final'] = (streams[0], streams[1] if pack.reversible else None)
    o32, o16 = ops.stable_ln(streams[0], pack.norm_w, pack.norm_b, b2=tape['final'][1], want_f32=True, want_bf16=True)
    return o32.view(B, nt, D), o16.view(B, nt, D), tape


# ------------------------------------------------------------------------------------------------
# backward
# ------------------------------------------------------------------------------------------------
# Weight-gradient GEMMs feed nothing in the backward chain (only the optimiser / the all-reduce), so they run on a second
# stream: a persistent tcgen05 GEMM and the memory-bound kernels of the chain (LayerNorm backward, row kernels, GEGLU
# backward, batched attention products) co-reside on the SMs.  The fork / join is captured like any other edge when the
# step is a CUDA graph.  False = everything on one stream (A/B).
WGRAD_SIDE_STREAM = True
_side_streams = {}


class _SideWork:
    def __init__(self):
        self.keep, self.stream = [], None

    def run(self, fn, *keep):
        """fn() on the side stream after everything launched so far on the current stream; `keep` = tensors fn reads (held
        until join() so that the allocator cannot hand their memory out while the side stream still uses them)."""
        if not WGRAD_SIDE_STREAM:
            fn()
            return
        cur = torch.cuda.current_stream()
        if self.stream is None:
            dev = cur.device_index if hasattr(cur, 'device_index') else torch.cuda.current_device()
            if dev not in _side_streams:
                _side_streams[dev] = torch.cuda.Stream(device=dev)
            self.stream = _side_streams[dev]
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            fn()
        self.keep.append(keep)

    def join(self):
        if self.stream is not None and self.keep:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.keep.clear()


_side = _SideWork()


def _wgrad(dy16, a16, dst):
    """dst (N, K) fp32 += dy16 (M, N).T @ a16 (M, K)"""
    if dst is not None:
        _side.run(lambda: ops_bwd.gemm_splitk_tn(dy16, a16, dst), dy16, a16)


def _sub_backward(i, s, rec, dy16, B, nt, g, tape, dctx):
    """dy16: bf16 (M, D) gradient of the sub-block output y.  Returns da fp32 (M, D)."""
    bw = _bw(s)
    M = B * nt
    a16 = rec['a']
    if s.kind == 'ff':
        m = s.mod
        D = dy16.shape[1]
        ip = s.w1.shape[0] // 2
        dg = ops.gemm(dy16, bw['w2_t'], out_dtype=torch.bfloat16)                     # (M, ip)
        w2g = g(m.net[3].weight)
        if w2g is not None:
            def w2_grad(gact=rec['g']):
                tmp = torch.zeros(D, ip, dtype=torch.float32, device=dy16.device)
                ops_bwd.gemm_splitk_tn(dy16, gact, tmp)
                ops_bwd.add_rows(w2g, tmp, cols=s.ff_inner)
            _side.run(w2_grad, dy16, rec['g'])
        dh = ops_bwd.geglu_bwd(dg, rec['h'])                                         # (M, 2*ip) pair packed
        w1g = g(m.net[0].weight)
        if w1g is not None:
            def w1_grad():
                tmp = torch.zeros(2 * ip, D, dtype=torch.float32, device=dy16.device)
                ops_bwd.gemm_splitk_tn(dh, a16, tmp)
                ops_bwd.add_rows(w1g, tmp, row_map=bw['w1_map'])
            _side.run(w1_grad, dh, a16)
        return ops.gemm(dh, bw['w1_t'], out_dtype=torch.float32)
    m = s.mod
    inner, H, dh_ = s.inner, s.H, s.dh
    do = ops.gemm(dy16, bw['w_out_t'], out_dtype=torch.bfloat16)                      # (M, inner)
    _wgrad(dy16, rec['o'], g(m.to_out.weight))
    dtalk = g(m.talking_heads.weight)
    dtalk = dtalk.view(H, H) if dtalk is not None else torch.zeros(H, H, device=dy16.device)
    if s.kind == '3dna':
        dqkv = ops_bwd.attn_sparse3dna_bwd(rec['qkv'], do, B=B, n=nt, H=H, dh=dh_, talk=s.talk, dtalk=dtalk, fmap=s.fmap,
                                           max_frames=s.max_frames, kernel=s.kernel, dilation=s.dilation, causal=s.causal)
        dqkv = dqkv.view(M, 3 * inner)
    elif s.kind == 'self':
        qkv = rec['qkv']
        base = qkv.data_ptr()
        rotary = tape['rotary']
        dnk, dnv = g(m.null_k), g(m.null_v)
        dnk = dnk.view(-1) if dnk is not None else torch.zeros(inner, device=do.device)
        dnv = dnv.view(-1) if dnv is not None else torch.zeros(inner, device=do.device)
        out_dtype = torch.float32 if rotary is not None else torch.bfloat16
        esz = 4 if rotary is not None else 2
        dq_all = torch.empty(M, 3 * inner, dtype=out_dtype, device=do.device)
        ops_bwd.attn_dense_bwd(base, base + inner * 2, base + 2 * inner * 2, do.view(B, nt, inner), B=B, nq=nt, nk=nt, H=H,
                               dh=dh_, q_bs=nt * 3 * inner, q_rs=3 * inner, kv_bs=nt * 3 * inner, kv_rs=3 * inner,
                               talk=s.talk, dtalk=dtalk, null_k=s.null_k, null_v=s.null_v, dnull_k=dnk, dnull_v=dnv,
                               key_mask=tape['key_mask'], dq_out=(dq_all, dq_all.data_ptr()), dq_bs=nt * 3 * inner,
                               dq_rs=3 * inner, dk_ptr=dq_all.data_ptr() + inner * esz,
                               dv_ptr=dq_all.data_ptr() + 2 * inner * esz, dkv_bs=nt * 3 * inner, dkv_rs=3 * inner,
                               out_f32=rotary is not None)
        dqkv = ops_bwd.rotary_bwd_to_bf16(dq_all, rotary[0], nt, H, dh_, rotary[1]) if rotary is not None else dq_all
    elif s.kind in ('cross', 'x2dna'):
        ctx = tape['context']
        nk = ctx.ctx16.shape[1]
        q, kv = rec['q'], rec['kv']
        dnk, dnv = g(m.null_k), g(m.null_v)
        dnk = dnk.view(-1) if dnk is not None else torch.zeros(inner, device=do.device)
        dnv = dnv.view(-1) if dnv is not None else torch.zeros(inner, device=do.device)
        if s.kind == 'cross':
            dq = torch.empty(M, inner, dtype=torch.bfloat16, device=do.device)
            dkv = torch.empty(B * nk, 2 * inner, dtype=torch.bfloat16, device=do.device)
            kb = kv.data_ptr()
            ops_bwd.attn_dense_bwd(q.data_ptr(), kb, kb + inner * 2, do.view(B, nt, inner), B=B, nq=nt, nk=nk, H=H, dh=dh_,
                                   q_bs=nt * inner, q_rs=inner, kv_bs=nk * 2 * inner, kv_rs=2 * inner, talk=s.talk,
                                   dtalk=dtalk, null_k=s.null_k, null_v=s.null_v, dnull_k=dnk, dnull_v=dnv,
                                   key_mask=ctx.mask, dq_out=dq, dq_bs=nt * inner, dq_rs=inner, dk_ptr=dkv.data_ptr(),
                                   dv_ptr=dkv.data_ptr() + inner * 2, dkv_bs=nk * 2 * inner, dkv_rs=2 * inner,
                                   out_f32=False, side=_side, keep=(q, kv, dkv))
        else:
            dq, dkv = ops_bwd.attn_cross2dna_bwd(q.view(B, nt, inner), kv.view(B, nk, 2 * inner), do.view(B, nt, inner), B=B,
                                                 n=nt, nk=nk, H=H, dh=dh_, talk=s.talk, dtalk=dtalk, null_k=s.null_k,
                                                 null_v=s.null_v, dnull_k=dnk, dnull_v=dnv, key_mask=ctx.mask, fmap=s.fmap,
                                                 ck=s.ck, cdil=s.cdil, side=_side)
            dq, dkv = dq.view(M, inner), dkv.view(B * nk, 2 * inner)
        _wgrad(dq, a16, g(m.to_q.weight))
        if dctx is not None:
            _wgrad(dkv, ctx.ctx16.view(B * nk, -1), g(m.to_kv.weight))
            # dctx += dkv @ W_kv: the context gradient is consumed after this stack's backward (text / sketch encoder), and
            # dkv of the dense kind is produced on the second stream -> same stream, submission order
            _side.run(lambda: ops.gemm(dkv, bw['w_kv_t'], residual=dctx, out=dctx), dkv)
        return ops.gemm(dq, bw['w_q_t'], out_dtype=torch.float32)
    else:
        raise NotImplementedError(s.kind)
    _wgrad(dqkv[:, :inner], a16, g(m.to_q.weight))
    _wgrad(dqkv[:, inner:], a16, g(m.to_kv.weight))
    return ops.gemm(dqkv, bw['w_qkv_t'], out_dtype=torch.float32)


def stack_backward(stack, tape, dout, g, dctx=None, on_done=None):
    """dout: fp32 (B*nt, D) gradient of the stack output (after its StableLayerNorm).  g: GradStore.
    dctx: fp32 (B*nk, D) accumulator for the gradient w.r.t. the cross-attention context (or None).
    Returns the gradient w.r.t. the stack input, fp32 (B*nt, D)."""
    pack = engine.pack_stack(stack)
    B, nt, D = tape['B'], tape['nt'], tape['D']
    M = B * nt
    dev = dout.device
    ln = stack.norm.norm
    G = [torch.empty(M, D, dtype=torch.float32, device=dev) for _ in range(2 if pack.reversible else 1)]
    f0, f1 = tape['final']
    ops_bwd.ln_bwd(dout, f0, pack.norm_w, nt=nt, dw=g(ln.weight), db=g(ln.bias), x2=f1, stable=True, dx_f32=G[0],
                   dx2_f32=G[1] if pack.reversible else None)
    for i in range(len(pack.subs) - 1, -1, -1):
        s, rec = pack.subs[i], tape['recs'][i]
        sw = s.sandwich
        bias = getattr(s.mod, 'to_out', None)
        bias = bias.bias if (bias is not None and getattr(bias, 'bias', None) is not None) else None
        dy16 = ops_bwd.ln_bwd(G[s.write], rec['y'], s.post[0], nt=nt, dw=g(sw.postnorm.weight), db=g(sw.postnorm.bias),
                              dcol=g(bias) if bias is not None else None, dx_bf16=True)
        da = _sub_backward(i, s, rec, dy16, B, nt, g, tape, dctx)
        ops_bwd.ln_bwd(da, rec['x_read'], s.pre[0], nt=nt, dw=g(sw.prenorm.weight), db=g(sw.prenorm.bias), unshift=s.shift,
                       fmap=s.fmap or 0, dx_f32=G[s.read], accumulate=True)
        tape['recs'][i] = None  # free the saved activations of this sub-block
        if on_done is not None:
            _side.join()  # the weight gradients of this sub-block are complete
            on_done(s)  # every gradient of this sub-block's parameters is final (data-parallel all-reduce can start)
    _side.join()
    if pack.reversible:
        ops_bwd.add_rows(G[0], G[1])  # both streams start as x
    return G[0]


# ------------------------------------------------------------------------------------------------
# NUWA training step as one autograd node
# ------------------------------------------------------------------------------------------------
class NuwaStep:
    """Forward (with tape) and backward of NUWA.forward(return_loss=True) for one batch."""

    def __init__(self, model, text, frame_indices, text_mask_dec):
        self.model, self.text, self.idx, self.mask_dec = model, text.contiguous(), frame_indices.contiguous(), text_mask_dec
        self.params = [p for n, p in model.named_parameters() if p.requires_grad and not n.startswith('vae.')]

    def forward(self):
        m = self.model
        text, idx = self.text, self.idx
        B, N = idx.shape
        axials, dims = (None, None, None), (1, 1, 1)
        if m.text_abs_pos_emb is not None:
            axials = (m.text_abs_pos_emb.embed.weight.detach().float().contiguous(), None, None)
            dims = (text.shape[1], 1, 1)
        tokens = ops.embed_tokens(text, m.text_embedding.embed.weight.detach().float().contiguous(), nt=text.shape[1],
                                  axials=axials, dims=dims)
        rotary = None
        if m.text_rotary_pos_emb is not None:
            rotary = (m.text_rotary_pos_emb.inv_freq.float().contiguous(), m.text_rotary_pos_emb.dim)
        km = (text != 0).to(torch.uint8).contiguous()
        _, e16, self.tape_text = stack_forward(m.text_transformer, tokens, key_mask=km, rotary=rotary)
        self.context = engine.Context(e16, self.mask_dec)
        x = m._embed_video(idx, N)
        _, y16, self.tape_dec = stack_forward(m.video_transformer, x, context=self.context)
        self.y16 = y16.view(B * N, -1)
        self.logits = ops.gemm(self.y16, m._logits_weight(), out_dtype=torch.float32)
        self.targets = idx.reshape(-1).contiguous()
        return ops.cross_entropy_mean(self.logits, self.targets)

    def backward(self, gout, reducer=None):
        m = self.model
        g, persistent = _grad_store_of(m, self.params)
        on_done = None
        if reducer is not None:
            reducer.begin(g.flat)
            on_done = lambda s: reducer.ready(*g.span(list(s.sandwich.parameters())))  # noqa: E731
        B, N = self.idx.shape
        D = self.y16.shape[1]
        gscale = gout.detach().to(torch.float32).reshape(1).contiguous()
        dlogits = ops_bwd.ce_bwd(self.logits, self.targets, gscale)
        self.logits = None
        w_t = ops_bwd.transpose(m._logits_weight()).contiguous()                    # (D, V)
        dy = ops_bwd.linear_bwd(dlogits, self.y16, w_t, g(m.to_logits.weight))
        del dlogits
        nk = self.context.ctx16.shape[1]
        dctx = torch.zeros(B * nk, D, dtype=torch.float32, device=dy.device)
        dx = stack_backward(m.video_transformer, self.tape_dec, dy, g, dctx, on_done)
        frac = m.image_embedding.frac_gradient if m.training else 1.0
        pe = m.video_pos_emb
        dax, ax = [], 1
        for length in pe.full_shape:
            if length > 1:
                dax.append(g(getattr(pe, f'axial{ax}')))
                ax += 1
            else:
                dax.append(None)
        ops_bwd.embed_bwd(dx, self.idx, g(m.image_embedding.embed.weight), nt=N, frac=frac, dbos=g(m.video_bos),
                          daxials=tuple(dax), dims=pe.full_shape)
        dtok = stack_backward(m.text_transformer, self.tape_text, dctx, g)
        frac_t = m.text_embedding.frac_gradient if m.training else 1.0
        dpos = (None, None, None)
        if m.text_abs_pos_emb is not None:
            dpos = (g(m.text_abs_pos_emb.embed.weight), None, None)
        ops_bwd.embed_bwd(dtok, self.text, g(m.text_embedding.embed.weight), nt=self.text.shape[1], frac=frac_t,
                          daxials=dpos, dims=(self.text.shape[1], 1, 1))
        self.tape_dec = self.tape_text = None
        if reducer is not None:
            reducer.finish()
        # a persistent store already IS every parameter's .grad (trainer.TrainStep): the kernels accumulated into it in
        # place, handing the same views back to autograd would add them onto themselves
        return [None] * len(self.params) if persistent else g.grads()


class _StepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, step, *params):
        ctx.step = step
        with torch.no_grad():
            return step.forward()

    @staticmethod
    def backward(ctx, gout):
        with torch.no_grad():
            grads = ctx.step.backward(gout, getattr(ctx.step.model, '_grad_reducer', None))
        return (None, *grads)


def nuwa_training_loss(model, text, frame_indices, text_mask_dec):
    step = NuwaStep(model, text, frame_indices, text_mask_dec)
    return _StepFn.apply(step, *step.params)


# ------------------------------------------------------------------------------------------------
# NUWASketch training step (nuwa_pytorch.py:2514-2571): sketch tokens -> sketch encoder -> decoder with SparseCross2DNA
# ------------------------------------------------------------------------------------------------
class SketchStep(NuwaStep):
    def __init__(self, model, sketch_indices, tok_mask_u8, frame_indices):
        self.model, self.sidx, self.tok_mask, self.idx = model, sketch_indices, tok_mask_u8, frame_indices.contiguous()
        self.params = [p for n, p in model.named_parameters()
                       if p.requires_grad and not n.startswith('vae.') and not n.startswith('sketch_vae.')]

    def forward(self):
        m = self.model
        idx = self.idx
        B, N = idx.shape
        sflat = self.sidx.reshape(B, -1).contiguous()
        self.sflat = sflat
        ns = sflat.shape[1]
        tokens = ops.embed_tokens(sflat, m.sketch_embedding.embed.weight.detach().float().contiguous(), nt=ns,
                                  axials=m.sketch_pos_emb.tables(), dims=m.sketch_pos_emb.full_shape)
        _, e16, self.tape_ctx = stack_forward(m.sketch_transformer, tokens, key_mask=self.tok_mask)
        self.context = engine.Context(e16, self.tok_mask)
        x = m._embed_video(idx, N)
        _, y16, self.tape_dec = stack_forward(m.video_transformer, x, context=self.context)
        self.y16 = y16.view(B * N, -1)
        self.logits = ops.gemm(self.y16, m._logits_weight(), out_dtype=torch.float32)
        self.targets = idx.reshape(-1).contiguous()
        return ops.cross_entropy_mean(self.logits, self.targets)

    def backward(self, gout, reducer=None):
        m = self.model
        g, persistent = _grad_store_of(m, self.params)
        on_done = None
        if reducer is not None:
            reducer.begin(g.flat)
            on_done = lambda s: reducer.ready(*g.span(list(s.sandwich.parameters())))  # noqa: E731
        B, N = self.idx.shape
        D = self.y16.shape[1]
        gscale = gout.detach().to(torch.float32).reshape(1).contiguous()
        dlogits = ops_bwd.ce_bwd(self.logits, self.targets, gscale)
        self.logits = None
        w_t = ops_bwd.transpose(m._logits_weight()).contiguous()
        dy = ops_bwd.linear_bwd(dlogits, self.y16, w_t, g(m.to_logits.weight))
        del dlogits
        nk = self.context.ctx16.shape[1]
        dctx = torch.zeros(B * nk, D, dtype=torch.float32, device=dy.device)
        dx = stack_backward(m.video_transformer, self.tape_dec, dy, g, dctx, on_done)

        def axial_grads(pe):
            out, ax = [], 1
            for length in pe.full_shape:
                if length > 1:
                    out.append(g(getattr(pe, f'axial{ax}')))
                    ax += 1
                else:
                    out.append(None)
            return tuple(out)

        frac = m.image_embedding.frac_gradient if m.training else 1.0
        ops_bwd.embed_bwd(dx, self.idx, g(m.image_embedding.embed.weight), nt=N, frac=frac, dbos=g(m.video_bos),
                          daxials=axial_grads(m.video_pos_emb), dims=m.video_pos_emb.full_shape)
        dtok = stack_backward(m.sketch_transformer, self.tape_ctx, dctx, g)
        frac_s = m.sketch_embedding.frac_gradient if m.training else 1.0
        ops_bwd.embed_bwd(dtok, self.sflat, g(m.sketch_embedding.embed.weight), nt=self.sflat.shape[1], frac=frac_s,
                          daxials=axial_grads(m.sketch_pos_emb), dims=m.sketch_pos_emb.full_shape)
        self.tape_dec = self.tape_ctx = None
        if reducer is not None:
            reducer.finish()
        # a persistent store already IS every parameter's .grad (trainer.TrainStep): the kernels accumulated into it in
        # place, handing the same views back to autograd would add them onto themselves
        return [None] * len(self.params) if persistent else g.grads()


def sketch_training_loss(model, sketch_indices, tok_mask_u8, frame_indices):
    step = SketchStep(model, sketch_indices, tok_mask_u8, frame_indices)
    return _StepFn.apply(step, *step.params)
