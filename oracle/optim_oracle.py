"""TEST INFRASTRUCTURE ONLY (imported by tests/ and the golden generator, never by the product).

CPU restatement of the tail of NUWATrainer.train_step (reference train_nuwa.py:256-258):
    torch.nn.utils.clip_grad_norm_(nuwa.parameters(), max_grad_norm)
    optim.step()          # get_optimizer(): AdamW, weight decay only on parameters with ndim >= 2 (optimizer.py:6-31)
    optim.zero_grad()
written as explicit tensor arithmetic (no torch.optim) so that it can check the fused CUDA step.  Pinned against the
unmodified reference optimizer (oracle/make_golden_optim.py -> tests/golden/optim_small.pt)."""
import torch


def clip_coef(grads, max_norm):
    """clip_grad_norm_: total 2-norm over all gradients; coefficient max_norm / (norm + 1e-6), clamped to 1."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    return total, torch.clamp(max_norm / (total + 1e-6), max=1.0)


def adamw_step(params, grads, exp_avg, exp_avg_sq, step, *, lr=3e-4, wd=0.01, betas=(0.9, 0.999), eps=1e-8,
               max_grad_norm=0.5):
    """One trainer step, in place on the fp32 tensors of `params`, `exp_avg`, `exp_avg_sq`; `step` counts from 1.
    Returns the pre-clip gradient norm."""
    total, coef = clip_coef(grads, max_grad_norm) if max_grad_norm else (None, 1.0)
    b1, b2 = betas
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        g = g * coef
        if wd != 0 and p.ndim >= 2:                      # optimizer.py:6-9: 1-D tensors (norms, biases) are not decayed
            p.mul_(1 - lr * wd)
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)
    return total
