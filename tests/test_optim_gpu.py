"""Fused trainer step (csrc/optim.cu behind nuwa_pytorch_b200.optim.FusedAdamW) against the CPU oracle and the golden
trajectory of the reference optimizer.  fp32 arithmetic in a different association order: 1e-5 relative."""
import pytest
import torch

from oracle import optim_oracle as OO
from tests.helpers import golden, rel

pytestmark = pytest.mark.gpu


def test_fused_adamw_matches_reference_trajectory(cuda_device):
    from nuwa_pytorch_b200.optim import FusedAdamW
    fx = golden("optim_small.pt")
    params = [torch.nn.Parameter(p.clone().to(cuda_device)) for p in fx['p0']]
    opt = FusedAdamW(params, lr=fx['lr'], wd=fx['wd'], max_grad_norm=fx['max_norm'])
    for t, grads in enumerate(fx['grads']):
        for p, g in zip(params, grads):
            p.grad = g.clone().to(cuda_device)      # not the flat layout: exercises the gather path
        norm = opt.step()
        assert abs(norm.item() - fx['norms'][t].item()) <= 1e-5 * fx['norms'][t].item()
        for p, want in zip(params, fx['traj'][t]):
            assert torch.allclose(p.detach().cpu(), want, rtol=1e-5, atol=1e-7), (t, (p.detach().cpu() - want).abs().max())
    assert int(opt.step_dev.item()) == len(fx['grads']) + 1


def test_fused_adamw_flat_gradient_buffer_and_zero_grad(cuda_device):
    """Gradients that already live in the flat GradStore layout are consumed in place and zeroed; a large ragged
    parameter set (tails that are not multiples of 4, chunk boundaries) against the oracle."""
    from nuwa_pytorch_b200.optim import FusedAdamW
    from nuwa_pytorch_b200.train import GradStore
    g = torch.Generator().manual_seed(3)
    shapes = [(1000, 37), (37,), (8193,), (3, 5, 7), (1,), (16384,), (513, 129)]
    p0 = [torch.randn(s, generator=g) for s in shapes]
    params = [torch.nn.Parameter(p.clone().to(cuda_device)) for p in p0]
    opt = FusedAdamW(params, lr=1e-2, wd=0.1, max_grad_norm=0.5)
    q = [p.clone() for p in p0]
    m, v = [torch.zeros_like(p) for p in q], [torch.zeros_like(p) for p in q]
    for step in range(1, 4):
        grads = [torch.randn(s, generator=g) * 10.0 ** (-step) for s in shapes]
        store = GradStore(params)
        for p, gr in zip(params, grads):
            store(p).copy_(gr.to(cuda_device))
            p.grad = store(p)
        flat, aliased = opt._flat_grads()
        assert aliased and flat.data_ptr() == store.flat.data_ptr()      # consumed in place, no gather
        norm = opt.step(grad_scale=0.5)
        want = OO.adamw_step(q, [gr * 0.5 for gr in grads], m, v, step, lr=1e-2, wd=0.1, max_grad_norm=0.5)
        assert abs(norm.item() - want.item()) <= 1e-5 * want.item()
        assert float(store.flat.abs().max()) == 0.0                    # zero_grad fused
        for p, w in zip(params, q):
            assert rel(p, w) < 1e-5
    for a, b in zip(opt.layout.offsets, opt.layout.offsets[1:]):
        assert a % 4 == 0 and b >= a


def test_fused_adamw_updates_repack_model_weights(cuda_device):
    """After a fused step the packed bf16 weights of the stack executor are rebuilt (weights epoch), so the next forward
    sees the new parameters: loss after a few steps on one batch must drop."""
    from nuwa_pytorch_b200 import NUWA, VQGanVAE
    from nuwa_pytorch_b200.optim import FusedAdamW, trainable_parameters
    from tests.helpers import synth
    fx = golden("nuwa_small.pt")
    model = NUWA(vae=VQGanVAE(**fx['vae_kwargs']), **fx['kwargs'])
    model.load_state_dict(synth(fx), strict=False)
    model = model.to(cuda_device).train()
    keys_before = list(model.state_dict().keys())
    opt = FusedAdamW(trainable_parameters(model), lr=3e-3, wd=0.01, max_grad_norm=0.5)
    assert list(model.state_dict().keys()) == keys_before
    text, vidx = fx['text'].to(cuda_device), fx['video_indices'].to(cuda_device)
    losses = []
    for _ in range(6):
        loss = model(text=text, video=vidx, return_loss=True, cond_dropout_prob=0.)
        loss.backward()
        opt.step()
        opt.zero_grad()
        losses.append(loss.item())
    print("  losses:", [round(x, 4) for x in losses])
    assert losses[-1] < losses[0] - 0.05


def _small_model(dev):
    from nuwa_pytorch_b200 import NUWA, VQGanVAE
    from tests.helpers import synth
    fx = golden("nuwa_small.pt")
    model = NUWA(vae=VQGanVAE(**fx['vae_kwargs']), **fx['kwargs'])
    model.load_state_dict(synth(fx), strict=False)
    return fx, model.to(dev).train()


def test_fused_step_consumes_autograd_gradients_in_place_and_zeroes_them(cuda_device):
    """gradients produced by a REAL backward() reach FusedAdamW as base-less views of the flat
    GradStore buffer (AccumulateGrad detaches).  They must still be consumed in place, and zero_grad=True must clear
    what autograd accumulates into -- two steps without opt.zero_grad() must not accumulate."""
    from nuwa_pytorch_b200.optim import FusedAdamW, trainable_parameters
    fx, model = _small_model(cuda_device)
    params = trainable_parameters(model)
    opt = FusedAdamW(params, lr=1e-3, wd=0.01, max_grad_norm=0.5)
    text, vidx = fx['text'].to(cuda_device), fx['video_indices'].to(cuda_device)
    norms = []
    for _ in range(2):
        model(text=text, video=vidx, return_loss=True, cond_dropout_prob=0.).backward()
        flat, aliased = opt._flat_grads()
        assert aliased, 'autograd-produced gradients must be recognised as the flat buffer (address check)'
        assert flat.data_ptr() == params[0].grad.data_ptr()
        norms.append(float(opt.step(zero_grad=True)))          # deliberately no opt.zero_grad()
        assert all(float(p.grad.abs().max()) == 0.0 for p in params if p.grad is not None)
    # had the first step's gradients survived, the second norm would be ~2x the first (lr is tiny)
    assert norms[1] < 1.5 * norms[0], norms
    # gather path (gradients not in the flat layout): zero_grad must clear the real .grad tensors too
    for p in params:
        p.grad = torch.ones_like(p)
    _, aliased = opt._flat_grads()
    assert not aliased
    opt.step(zero_grad=True)
    assert all(float(p.grad.abs().max()) == 0.0 for p in params)


def test_graphed_train_step_with_optimizer_tracks_eager_training(cuda_device):
    """a captured training step must keep computing with the CURRENT weights.  GraphedTrainStep with
    model= and optimizer= captures [refresh packed bf16 weights -> forward -> backward -> clip + AdamW] and is replayed
    for several steps; its losses must follow an eager loop doing the same thing on a twin model, and drop."""
    from nuwa_pytorch_b200.graphs import GraphedTrainStep
    from nuwa_pytorch_b200.optim import FusedAdamW, trainable_parameters
    fx, model_g = _small_model(cuda_device)
    _, model_e = _small_model(cuda_device)
    text, vidx = fx['text'].to(cuda_device), fx['video_indices'].to(cuda_device)
    kw = dict(lr=3e-3, wd=0.01, max_grad_norm=0.5)
    opt_e = FusedAdamW(trainable_parameters(model_e), **kw)
    eager = []
    for _ in range(6):
        loss = model_e(text=text, video=vidx, return_loss=True, cond_dropout_prob=0.)
        loss.backward()
        opt_e.step()
        eager.append(loss.item())
    opt_g = FusedAdamW(trainable_parameters(model_g), **kw)
    step = GraphedTrainStep(lambda t, v: model_g(text=t, video=v, return_loss=True, cond_dropout_prob=0.),
                            trainable_parameters(model_g), text, vidx, model=model_g, optimizer=opt_g)
    graphed = [step(text, vidx).item() for _ in range(6)]
    print("  eager  :", [round(x, 4) for x in eager])
    print("  graphed:", [round(x, 4) for x in graphed])
    assert graphed[-1] < graphed[0] - 0.05                       # it trains (stale weights would give a flat line)
    for a, b in zip(eager, graphed):
        assert abs(a - b) < 2e-2, (eager, graphed)                # atomics-order noise only
    # an eager forward between replays must not invalidate the graph (packs are refreshed in place, never freed)
    model_g.eval()
    with torch.no_grad():
        mid = model_g(text=text, video=vidx, return_loss=True, cond_dropout_prob=0.).item()
    model_g.train()
    nxt = step(text, vidx).item()
    assert abs(mid - nxt) < 2e-2 and nxt < graphed[0]
    # frozen-weight capture (no model=) refuses to replay once the weights moved
    _, model_f = _small_model(cuda_device)
    fstep = GraphedTrainStep(lambda t, v: model_f(text=t, video=v, return_loss=True, cond_dropout_prob=0.),
                             trainable_parameters(model_f), text, vidx)
    fstep(text, vidx)
    with torch.no_grad():
        next(iter(trainable_parameters(model_f))).add_(1e-3)
    with pytest.raises(Exception, match='captured without model'):
        fstep(text, vidx)
