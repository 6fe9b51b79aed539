"""Training step (forward + backward) of NUWA through the CUDA path vs the gradients of the UNMODIFIED reference
(tests/golden/*_grads.pt, written by oracle/make_golden_grads.py from `loss.backward()` of the reference).

Tolerance: bf16 tensor-core operands with fp32 accumulation on the CUDA side against fp32 end to end in the
reference; gradients are compared per parameter tensor by relative L2 (GRAD_TOL) and over the concatenation of all of
them (GRAD_TOL_ALL)."""
import pytest
import torch

from tests.helpers import golden, rel, synth

pytestmark = pytest.mark.gpu
GRAD_TOL = 6e-2
GRAD_TOL_ALL = 2.5e-2
GRAD_TOL_TINY = 0.15


def _train_model(name, dev):
    from nuwa_pytorch_b200 import NUWA, VQGanVAE
    fx = golden(name)
    vae = VQGanVAE(**fx['vae_kwargs'])
    model = NUWA(vae=vae, **fx['kwargs'])
    missing, unexpected = model.load_state_dict(synth(fx), strict=False)
    assert not unexpected
    return fx, model.to(dev).train()


@pytest.mark.parametrize("name", ["nuwa_small", "nuwa_rev_small"])
def test_nuwa_loss_backward_matches_reference_gradients(cuda_device, name):
    fx, model = _train_model(name + ".pt", cuda_device)
    gold = golden(name + "_grads.pt")
    text, vidx = fx['text'].to(cuda_device), fx['video_indices'].to(cuda_device)
    loss = model(text=text, video=vidx, return_loss=True, cond_dropout_prob=0.)
    assert loss.requires_grad and abs(loss.item() - gold['loss'].item()) < 2e-2
    loss.backward()
    params = dict(model.named_parameters())
    worst, num, den, bad = (0., None), 0., 0., []
    for k, gref in gold['grads'].items():
        p = params[k]
        assert p.grad is not None, k
        r = rel(p.grad, gref)
        if r > worst[0]:
            worst = (r, k)
        num += (p.grad.double().cpu() - gref.double()).pow(2).sum().item()
        den += gref.double().pow(2).sum().item()
        # the (heads x heads) talking-heads gradients are sums with heavy cancellation over every (query, key) pair:
        # the fp32 oracle itself only reproduces the reference to 1e-4 there, bf16 operands cost ~3 decimal digits more
        bad = bad + [(k, r)] if r > (GRAD_TOL_TINY if gref.numel() <= 64 else GRAD_TOL) else bad
    total = (num / den) ** 0.5
    print(f"  {name}: {len(gold['grads'])} gradient tensors, worst rel {worst[0]:.3e} ({worst[1]}), all-params rel {total:.3e}")
    assert not bad, bad
    assert total < GRAD_TOL_ALL
    assert all(p.grad is None for k, p in params.items() if k.startswith('vae.'))  # the frozen VAE copy gets no gradient


def test_backward_scales_with_grad_output_and_accumulates(cuda_device):
    fx, model = _train_model("nuwa_small.pt", cuda_device)
    text, vidx = fx['text'].to(cuda_device), fx['video_indices'].to(cuda_device)
    model(text=text, video=vidx, return_loss=True, cond_dropout_prob=0.).backward()
    g1 = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    # gradient accumulation as NUWATrainer does it (train_nuwa.py:247-251): (loss / 2).backward() twice more
    for _ in range(2):
        (model(text=text, video=vidx, return_loss=True, cond_dropout_prob=0.) / 2).backward()
    for k, p in model.named_parameters():
        if p.grad is not None and g1[k].abs().max() > 0:
            assert rel(p.grad, 2 * g1[k]) < 2e-3, k  # split-K / atomic summation order is the only difference


def test_whole_step_cuda_graph_matches_eager(cuda_device):
    """bench.py replays the training step (forward + backward) from one CUDA graph: same loss and gradients as eager."""
    from nuwa_pytorch_b200.graphs import GraphedTrainStep
    fx, model = _train_model("nuwa_small.pt", cuda_device)
    text, vidx = fx['text'].to(cuda_device), fx['video_indices'].to(cuda_device)
    params = [p for n, p in model.named_parameters() if not n.startswith('vae.')]
    fn = lambda t, v: model(text=t, video=v, return_loss=True, cond_dropout_prob=0.)  # noqa: E731
    eager_loss = fn(text, vidx)
    eager_loss.backward()
    eager = [p.grad.clone() if p.grad is not None else None for p in params]
    eager_loss = eager_loss.detach()  # drop the eager autograd graph (and its stream-bound AccumulateGrad nodes)
    step = GraphedTrainStep(fn, params, text, vidx)
    for _ in range(2):
        loss = step(text, vidx)
    torch.cuda.synchronize()
    assert abs(loss.item() - eager_loss.item()) < 1e-5
    for p, e in zip(params, eager):
        if e is not None and e.abs().max() > 0:
            assert rel(p.grad, e) < 2e-3  # atomic summation order only
    # new inputs through the static buffers
    vidx2 = (vidx + 1) % 64
    l2 = step(text, vidx2).item()
    assert abs(l2 - fn(text, vidx2).item()) < 1e-4


def test_nuwa_sketch_loss_backward_matches_reference_gradients(cuda_device):
    """NUWASketch training step (BASELINE configs[4] shape of path): float sketch + video through both VAEs, 3DNA sketch
    encoder, decoder with SparseCross2DNA; gradients vs the unmodified reference."""
    from nuwa_pytorch_b200 import NUWASketch, VQGanVAE
    from oracle.synth import manifest_of, synth_state_dict
    from tests.helpers import gen
    fx = golden("sketch_small.pt")
    gold = golden("sketch_small_grads.pt")
    vae, svae = VQGanVAE(**fx['vae_kwargs']), VQGanVAE(**fx['sketch_vae_kwargs'])
    vae.load_state_dict(synth_state_dict(manifest_of(vae.state_dict()), fx['vae_seed']), strict=False)
    svae.load_state_dict(synth_state_dict(manifest_of(svae.state_dict()), fx['sketch_vae_seed']), strict=False)
    model = NUWASketch(vae=vae, sketch_vae=svae, **fx['kwargs'])
    model.load_state_dict(synth(fx), strict=False)
    model = model.to(cuda_device).train()
    sketch = torch.randn(2, 3, 5, 64, 64, generator=gen(fx['e2e_sketch_seed'])).to(cuda_device)
    video = torch.randn(2, 3, 3, 64, 64, generator=gen(fx['e2e_video_seed'])).to(cuda_device)
    smask = torch.ones(2, 3, dtype=torch.bool, device=cuda_device)
    # token ids exactly as the reference saw them (CPU oracle VAEs): a single flipped VQ token of the bf16 GPU VAE would
    # change the training inputs and hide / fake gradient differences, so the ids are pinned here and the public
    # forward() (which tokenises on the GPU) is checked separately below
    from oracle import nuwa_oracle as O
    from tests.helpers import vae_spec_from_kwargs
    full_sd = synth(fx)  # NB: loading the NUWASketch state dict also overwrites both VAEs (keys vae.* / sketch_vae.*)
    vsd = {k[len('vae.'):]: v for k, v in full_sd.items() if k.startswith('vae.')}
    ssd = {k[len('sketch_vae.'):]: v for k, v in full_sd.items() if k.startswith('sketch_vae.')}
    sidx = O.vae_get_video_indices(sketch.cpu(), ssd, vae_spec_from_kwargs(fx['sketch_vae_kwargs']))
    fi = O.vae_get_video_indices(video.cpu(), vsd, vae_spec_from_kwargs(fx['vae_kwargs'])).reshape(2, -1)
    with torch.no_grad():
        same_s = (model.sketch_vae.get_video_indices(sketch).cpu() == sidx).float().mean().item()
        same_v = (model.vae.get_video_indices(video).cpu().reshape(2, -1) == fi).float().mean().item()
    print(f"  GPU VAE token ids equal to the CPU oracle: sketch {same_s:.3f}, video {same_v:.3f}")
    from nuwa_pytorch_b200 import train
    tok_mask = torch.ones(2, 48, dtype=torch.uint8, device=cuda_device)
    loss = train.sketch_training_loss(model, sidx.to(cuda_device), tok_mask, fi.to(cuda_device))
    assert loss.requires_grad and abs(loss.item() - gold['loss'].item()) < 2e-2
    loss.backward()
    pub = model(sketch=sketch, sketch_mask=smask, video=video, return_loss=True, cond_dropout_prob=0.)  # public surface
    assert pub.requires_grad and abs(pub.item() - gold['loss'].item()) < 5e-2
    params = dict(model.named_parameters())
    worst, num, den, bad = (0., None), 0., 0., []
    for k, gref in gold['grads'].items():
        p = params[k]
        assert p.grad is not None, k
        r = rel(p.grad, gref)
        worst = max(worst, (r, k))
        num += (p.grad.double().cpu() - gref.double()).pow(2).sum().item()
        den += gref.double().pow(2).sum().item()
        if r > (GRAD_TOL_TINY if gref.numel() <= 64 else GRAD_TOL):
            bad.append((k, r))
    total = (num / den) ** 0.5
    print(f"  sketch_small: {len(gold['grads'])} gradient tensors, worst rel {worst[0]:.3e} ({worst[1]}), all-params rel {total:.3e}")
    assert not bad, bad
    assert total < GRAD_TOL_ALL


def _compare_grads(model, gold_grads, label):
    params = dict(model.named_parameters())
    worst, num, den, bad = (0., None), 0., 0., []
    for k, gref in gold_grads.items():
        p = params[k]
        assert p.grad is not None, k
        r = rel(p.grad, gref)
        worst = max(worst, (r, k))
        num += (p.grad.double().cpu() - gref.double()).pow(2).sum().item()
        den += gref.double().pow(2).sum().item()
        if r > (GRAD_TOL_TINY if gref.numel() <= 64 else GRAD_TOL):
            bad.append((k, r))
    total = (num / den) ** 0.5
    print(f"  {label}: {len(gold_grads)} gradient tensors, worst rel {worst[0]:.3e} ({worst[1]}), all-params rel {total:.3e}")
    assert not bad, bad
    assert total < GRAD_TOL_ALL


def test_training_variants_match_reference_gradients(cuda_device):
    """Code paths the main fixtures do not reach: learned absolute text positions + 4 heads x 16 (NUWA), and a dense
    (non-3DNA) sketch encoder + reversible decoder with SparseCross2DNA + a masked sketch frame (NUWASketch)."""
    from nuwa_pytorch_b200 import NUWA, NUWASketch, VQGanVAE, train
    from oracle.synth import synth_state_dict
    var = golden("train_variants.pt")
    a = var['nuwa_abs_pos']
    model = NUWA(vae=VQGanVAE(**a['vae_kwargs']), **a['kwargs'])
    _, unexpected = model.load_state_dict(synth_state_dict(a['manifest'], a['seed']), strict=False)
    assert not unexpected
    model = model.to(cuda_device).train()
    loss = model(text=a['text'].to(cuda_device), video=a['video_indices'].to(cuda_device), return_loss=True, cond_dropout_prob=0.)
    assert abs(loss.item() - a['loss'].item()) < 2e-2
    loss.backward()
    _compare_grads(model, a['grads'], 'nuwa_abs_pos')
    b = var['sketch_dense_rev']
    sk = NUWASketch(vae=VQGanVAE(**b['vae_kwargs']), sketch_vae=VQGanVAE(**b['sketch_vae_kwargs']), **b['kwargs'])
    _, unexpected = sk.load_state_dict(synth_state_dict(b['manifest'], b['seed']), strict=False)
    assert not unexpected
    sk = sk.to(cuda_device).train()
    tok_mask = b['sketch_mask'][:, :, None].expand(2, 2, 16).reshape(2, -1).to(torch.uint8).contiguous().to(cuda_device)
    loss = train.sketch_training_loss(sk, b['sketch_indices'].to(cuda_device), tok_mask, b['frame_indices'].to(cuda_device))
    assert abs(loss.item() - b['loss'].item()) < 2e-2
    loss.backward()
    _compare_grads(sk, b['grads'], 'sketch_dense_rev')
