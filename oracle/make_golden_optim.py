"""TEST INFRASTRUCTURE ONLY.  Golden vectors of the trainer-step tail (clip_grad_norm_ + AdamW of get_optimizer +
zero_grad, reference train_nuwa.py:256-258 / optimizer.py:11-31) produced by the UNMODIFIED reference optimizer module
loaded from /root/reference (build container only; optimizer.py imports nothing but torch).

    python -m oracle.make_golden_optim      # writes tests/golden/optim_small.pt
"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import optim_oracle as OO  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "optim_small.pt")
SHAPES = [(37, 24), (24,), (5, 3, 2, 2), (1,), (130,), (64, 8), (3,)]
LR, WD, MAX_NORM, STEPS, SEED = 3e-4, 0.01, 0.5, 4, 7


def main():
    spec = importlib.util.spec_from_file_location("ref_optimizer", "/root/reference/nuwa_pytorch/optimizer.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = torch.Generator().manual_seed(SEED)
    p0 = [torch.randn(s, generator=g) for s in SHAPES]
    # step 0 has a large gradient (clipping active), later ones are small (clip coefficient 1)
    grads = [[torch.randn(s, generator=g) * (1.0 if t == 0 else 0.01) for s in SHAPES] for t in range(STEPS)]
    params = [torch.nn.Parameter(p.clone()) for p in p0]
    opt = ref.get_optimizer(params, lr=LR, wd=WD)
    norms, traj = [], []
    for t in range(STEPS):
        for p, gr in zip(params, grads[t]):
            p.grad = gr.clone()
        norms.append(torch.nn.utils.clip_grad_norm_(params, MAX_NORM).clone())
        opt.step()
        opt.zero_grad()
        traj.append([p.detach().clone() for p in params])
    # the oracle restatement reproduces the reference trajectory
    q = [p.clone() for p in p0]
    m, v = [torch.zeros_like(p) for p in p0], [torch.zeros_like(p) for p in p0]
    for t in range(STEPS):
        n = OO.adamw_step(q, grads[t], m, v, t + 1, lr=LR, wd=WD, max_grad_norm=MAX_NORM)
        assert abs(n.item() - norms[t].item()) <= 1e-5 * norms[t].item()
        for a, b in zip(q, traj[t]):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-7), (t, (a - b).abs().max())
    torch.save(dict(shapes=SHAPES, lr=LR, wd=WD, max_norm=MAX_NORM, seed=SEED, p0=p0, grads=grads, norms=norms, traj=traj), OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
