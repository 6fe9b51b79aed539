"""NUWA / NUWASketch / Sparse3DNA through the CUDA path vs goldens of the unmodified reference.

Tolerance: the path computes with bf16 tensor-core operands and fp32 accumulation / residual streams /
norms / softmax; activations are compared by relative L2 against the fp32 reference.  BF16_TOL documents
the bound we hold end to end (the north_star's 1e-3 is met at op level on identical bf16 operands, see
tests/test_kernels_gpu.py and tests/test_gemm_gpu.py)."""
import pytest
import torch

from oracle import nuwa_oracle as O
from tests.helpers import gen, golden, nuwa_spec_from_kwargs, rel, synth

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2


def _load(model, fx, dev):
    sd = synth(fx)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected
    assert all(('mask' in k) or ('inv_freq' in k) or ('net.blocks' in k) for k in missing), missing
    return model.to(dev).eval(), sd


def _nuwa(name, dev):
    from nuwa_pytorch_b200 import NUWA, VQGanVAE
    fx = golden(name)
    vae = VQGanVAE(**fx['vae_kwargs'])
    model, sd = _load(NUWA(vae=vae, **fx['kwargs']), fx, dev)
    return fx, model, sd


def test_nuwa_small_forward(cuda_device):
    fx, model, sd = _nuwa("nuwa_small.pt", cuda_device)
    text, vidx = fx['text'].to(cuda_device), fx['video_indices'].to(cuda_device)
    emb = model.embed_text(text, mask=text != 0)
    r_text = rel(emb, fx['text_embeds'])
    loss = model(text=text, video=vidx, return_loss=True)
    # logits with return_loss=False need one position less than the full video (reference D6): use the loss path's
    # logits by re-running the decoder on the same inputs
    context = model._text_context(text, text != 0)
    x = model._embed_video(vidx.reshape(2, -1), 48)
    from nuwa_pytorch_b200 import engine, ops
    _, y16 = engine.run_stack(model.video_transformer, x, context=context, want_bf16=True)
    logits = ops.gemm(y16.view(96, -1), model._logits_weight(), out_dtype=torch.float32).view(2, 48, -1)
    r_log = rel(logits, fx['logits'])
    print(f"nuwa_small: text-emb rel {r_text:.3e} logits rel {r_log:.3e} loss {loss.item():.5f} vs {fx['loss'].item():.5f}")
    assert r_text < BF16_TOL and r_log < BF16_TOL
    assert abs(loss.item() - fx['loss'].item()) < 2e-2


def test_nuwa_rev_forward_and_incremental_generate_logits(cuda_device):
    fx, model, sd = _nuwa("nuwa_rev_small.pt", cuda_device)
    text, vidx = fx['text'].to(cuda_device), fx['video_indices'].to(cuda_device)
    loss = model(text=text, video=vidx, return_loss=True)
    assert abs(loss.item() - fx['loss'].item()) < 2e-2
    # teacher-forced incremental decode: force the sampled tokens to the golden sequence by feeding noise that makes
    # the forced token win is fragile; instead drive _generate_indices' building blocks directly
    from nuwa_pytorch_b200 import engine, ops
    seq = vidx.reshape(2, -1)
    context = model._text_context(text, text != 0)
    unc = context.with_mask(torch.zeros_like(context.mask))
    pack = engine.pack_stack(model.video_transformer)
    total = seq.shape[1]
    st_c, st_u = engine.DecodeState(pack, 2, total, cuda_device), engine.DecodeState(pack, 2, total, cuda_device)
    w = model._logits_weight()
    worst = 0.0
    for t in range(41):
        st_c.t_dev.fill_(t)  # decode kernels read the position from device memory
        st_u.t_dev.fill_(t)
        x = model._embed_video(seq, 1, t0=t)
        y32, y16 = engine.run_stack(model.video_transformer, x, context=context, state=st_c, t0=t, want_bf16=True)
        lc = ops.gemm(y16.view(2, -1), w, out_dtype=torch.float32)
        _, u16 = engine.run_stack(model.video_transformer, y32, context=unc, state=st_u, t0=t, want_bf16=True)
        lu = ops.gemm(u16.view(2, -1), w, out_dtype=torch.float32)
        guided = lu + (lc - lu) * 2.0
        if t in fx['step_logits']:
            r = rel(guided, fx['step_logits'][t])
            worst = max(worst, r)
            print(f"  step {t}: guided-logit rel {r:.3e}")
    assert worst < 3e-2  # guidance (x2) amplifies the bf16 difference of two sweeps


def test_generate_runs_and_matches_oracle_sampling(cuda_device):
    fx, model, sd = _nuwa("nuwa_rev_small.pt", cuda_device)
    text = fx['text'].to(cuda_device)
    g = torch.Generator().manual_seed(5)
    noise = torch.rand(32, 2, 64, generator=g).to(cuda_device)
    idx = model.generate(text=text, num_frames=2, _noise=noise, _return_indices=True)
    assert idx.shape == (2, 32) and idx.dtype == torch.int64 and int(idx.max()) < 64
    # the CUDA-graph replayed decode loop and the eager loop are the same computation
    idx_eager = model.generate(text=text, num_frames=2, _noise=noise, _return_indices=True, _use_graph=False)
    assert torch.equal(idx, idx_eager)
    video = model.generate(text=text, num_frames=2, _noise=noise)
    assert video.shape == (2, 2, 3, 64, 64) and torch.isfinite(video).all()
    # replay on the oracle with the SAME sampled prefix: each step's sampled token must be the oracle's choice
    # whenever the oracle's decision margin exceeds the bf16 noise
    spec = nuwa_spec_from_kwargs(fx['kwargs'], fx['vae_kwargs'])
    temb, tmask = O.nuwa_embed_text(fx['text'], sd, spec)
    agree = n_checked = 0
    for t in (0, 1, 7, 16, 25):
        lg = O.nuwa_generate_step_logits(temb, tmask, idx[:, :t].cpu(), sd, spec, 2.)
        filt = O.top_k_filter(lg, 0.9)
        u = noise[t].cpu()
        score = filt + (-torch.log((-torch.log(u.clamp(min=1e-20))).clamp(min=1e-20)))
        top2 = score.topk(2, dim=-1).values
        for b in range(2):
            if (top2[b, 0] - top2[b, 1]) > 0.3:
                n_checked += 1
                agree += int(score[b].argmax().item() == idx[b, t].item())
    print(f"generate: {agree}/{n_checked} confidently-decided samples agree with the oracle")
    assert n_checked == 0 or agree / n_checked >= 0.8


def test_sparse3dna_module_matches_reference(cuda_device):
    from nuwa_pytorch_b200 import Sparse3DNA
    from oracle.synth import synth_state_dict
    ops_fx = golden("sparse3dna_ops.pt")
    for name, c in ops_fx.items():
        mod = Sparse3DNA(dim=64, video_shape=(3, 4, 4), kernel_size=c['kernel'], dilation=c['dilation'], heads=2,
                         dim_head=32, causal=c['causal'])
        mod.load_state_dict(synth_state_dict(c['manifest'], c['seed']), strict=False)
        mod = mod.to(cuda_device)
        x = torch.randn(2, c['n'], 64, generator=gen(c['x_seed']))
        y = mod(x.to(cuda_device))
        r = rel(y, c['y'])
        print(f"  sparse3dna {name}: rel {r:.3e}")
        assert r < BF16_TOL, name


def test_sketch_small(cuda_device):
    from nuwa_pytorch_b200 import NUWASketch, VQGanVAE
    fx = golden("sketch_small.pt")
    vae, svae = VQGanVAE(**fx['vae_kwargs']), VQGanVAE(**fx['sketch_vae_kwargs'])
    from oracle.synth import manifest_of, synth_state_dict
    vae.load_state_dict(synth_state_dict(manifest_of(vae.state_dict()), fx['vae_seed']), strict=False)
    svae.load_state_dict(synth_state_dict(manifest_of(svae.state_dict()), fx['sketch_vae_seed']), strict=False)
    model = NUWASketch(vae=vae, sketch_vae=svae, **fx['kwargs'])
    model, sd = _load(model, fx, cuda_device)
    from nuwa_pytorch_b200 import engine
    fi = fx['video_indices'].reshape(2, -1).to(cuda_device)
    for nf, case in fx['cases'].items():
        (emb, e16), tok_mask = model._embed_sketch_indices(case['sketch_indices'].to(cuda_device),
                                                           case['sketch_mask'].to(cuda_device), want_bf16=True)
        r_emb = rel(emb, case['sketch_embeds'])
        ctx = engine.Context(e16, tok_mask.to(torch.uint8).contiguous())
        loss = model._decoder_logits(fi, ctx, True)
        print(f"  sketch nf={nf}: embed rel {r_emb:.3e} loss {loss.item():.5f} vs {case['loss'].item():.5f}")
        assert r_emb < BF16_TOL and abs(loss.item() - case['loss'].item()) < 2e-2
    # end to end through both VAEs (float sketch + float video)
    sketch = torch.randn(2, 3, 5, 64, 64, generator=gen(fx['e2e_sketch_seed'])).to(cuda_device)
    video = torch.randn(2, 3, 3, 64, 64, generator=gen(fx['e2e_video_seed'])).to(cuda_device)
    loss = model(sketch=sketch, sketch_mask=torch.ones(2, 3, dtype=torch.bool, device=cuda_device), video=video,
                 return_loss=True)
    print(f"  sketch e2e loss {loss.item():.5f} vs {fx['e2e_loss'].item():.5f}")
    assert abs(loss.item() - fx['e2e_loss'].item()) < 0.1  # token flips in the VAEs move single targets


def test_sketch_generate_incremental_matches_teacher_forced(cuda_device):
    """NUWASketch.generate (KV-cached decode incl. the sparse 2-D cross attention) against a teacher-forced full pass
    of the same model on the sampled sequence: the guided logits of every step must agree."""
    from nuwa_pytorch_b200 import NUWASketch, VQGanVAE, engine, ops
    from oracle.synth import manifest_of, synth_state_dict
    fx = golden("sketch_small.pt")
    vae, svae = VQGanVAE(**fx['vae_kwargs']), VQGanVAE(**fx['sketch_vae_kwargs'])
    vae.load_state_dict(synth_state_dict(manifest_of(vae.state_dict()), fx['vae_seed']), strict=False)
    svae.load_state_dict(synth_state_dict(manifest_of(svae.state_dict()), fx['sketch_vae_seed']), strict=False)
    model, sd = _load(NUWASketch(vae=vae, sketch_vae=svae, **fx['kwargs']), fx, cuda_device)
    sketch = torch.randn(2, 3, 5, 64, 64, generator=gen(fx['e2e_sketch_seed'])).to(cuda_device)
    smask = torch.ones(2, 3, dtype=torch.bool, device=cuda_device)
    noise = torch.rand(32, 2, 64, generator=gen(77)).to(cuda_device)
    video = model.generate(sketch=sketch, sketch_mask=smask, num_frames=2, _noise=noise)
    assert video.shape == (2, 2, 3, 64, 64) and torch.isfinite(video).all()
    ctx = model._sketch_context(sketch, smask)
    idx, step_logits = model._generate_indices(ctx, 2, num_frames=2, filter_thres=0.9, temperature=1., cond_scale=2.,
                                               noise=noise, return_step_logits=True)
    # teacher-forced: full-sequence passes (conditional, then the D8 second sweep on its output)
    x = model._embed_video(idx, 32)
    y32, y16 = engine.run_stack(model.video_transformer, x, context=ctx, want_bf16=True)
    w = model._logits_weight()
    lc = ops.gemm(y16.view(64, -1), w, out_dtype=torch.float32).view(2, 32, -1)
    unc = ctx.with_mask(torch.zeros_like(ctx.mask))
    _, u16 = engine.run_stack(model.video_transformer, y32, context=unc, want_bf16=True)
    lu = ops.gemm(u16.view(64, -1), w, out_dtype=torch.float32).view(2, 32, -1)
    guided = lu + (lc - lu) * 2.0
    worst = max(rel(step_logits[t], guided[:, t]) for t in range(32))
    print(f"  sketch generate: worst step-vs-teacher-forced guided-logit rel {worst:.3e}")
    assert worst < 2e-2


def test_cuda_graph_replay_matches_eager(cuda_device):
    from nuwa_pytorch_b200.graphs import GraphedCall
    fx, model, sd = _nuwa("nuwa_small.pt", cuda_device)
    text, vidx = fx['text'].to(cuda_device), fx['video_indices'].to(cuda_device)
    with torch.no_grad():  # the inference path (with autograd enabled, forward() takes the training path of train.py)
        eager = model(text=text, video=vidx, return_loss=True).item()
    g = GraphedCall(lambda t, v: model(text=t, video=v, return_loss=True), text, vidx)
    assert abs(g(text, vidx).item() - eager) < 1e-6
    vidx2 = (vidx + 7) % 64
    with torch.no_grad():
        eager2 = model(text=text, video=vidx2, return_loss=True).item()
    assert abs(g(text, vidx2).item() - eager2) < 1e-6 and abs(eager2 - eager) > 1e-4
