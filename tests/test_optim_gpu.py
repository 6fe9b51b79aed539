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
        assert opt._flat_grads().data_ptr() == store.flat.data_ptr()   # consumed in place, no gather
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
