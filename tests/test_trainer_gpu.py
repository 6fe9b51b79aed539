"""The end-to-end trainer step (reference train_nuwa.py:237-258 + optimizer.py:11-31) through
nuwa_pytorch_b200.trainer.TrainStep: gradient accumulation into the flat buffer, the NCCL mean all-reduce, clip 0.5,
fused AdamW -- against the trajectory of the UNMODIFIED reference (tests/golden/trainer_small.pt, written by
oracle/make_golden_trainer.py), against the unfused plumbing of the same kernels, as a captured CUDA graph, and on two
ranks under real NCCL (skipped when the box has one GPU).

Tolerances: forward / backward run with bf16 tensor-core operands (loss within 2e-2, gradient norm within 3 %); Adam
normalises every gradient element to a step of ~lr, so the parameter trajectory is compared through the cosine between
the reference's total update and ours and through the relative L2 of the parameters themselves."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import golden, rel, synth

pytestmark = pytest.mark.gpu


def _model(dev):
    from nuwa_pytorch_b200 import NUWA, VQGanVAE
    fx = golden("nuwa_small.pt")
    model = NUWA(vae=VQGanVAE(**fx['vae_kwargs']), **fx['kwargs'])
    sd = synth(fx)
    model.load_state_dict(sd, strict=False)
    return fx, model.to(dev).train(), sd


def _batches(tr, s, dev):
    return [dict(text=tr['texts'][s, a].to(dev), video=tr['videos'][s, a].to(dev)) for a in range(tr['accum'])]


def test_trainer_step_follows_the_reference_trajectory(cuda_device):
    from nuwa_pytorch_b200.trainer import TrainStep
    tr = golden("trainer_small.pt")
    fx, model, sd = _model(cuda_device)
    step = TrainStep(model, lr=tr['lr'], wd=tr['wd'], grad_accum_every=tr['accum'], max_grad_norm=tr['max_norm'],
                     forward_kwargs=dict(cond_dropout_prob=0.))
    for s in range(tr['steps']):
        loss, norm = step.step(_batches(tr, s, cuda_device))
        print(f"  step {s}: loss {loss.item():.5f} (reference {tr['losses'][s]:.5f})  grad norm {norm.item():.5f} "
              f"(reference {tr['norms'][s].item():.5f})")
        assert abs(loss.item() - tr['losses'][s]) < 5e-3                               # measured <= 1.7e-3
        assert abs(norm.item() - tr['norms'][s].item()) < 1.5e-2 * tr['norms'][s].item()   # measured <= 0.5 %
        assert float(step.store.flat.abs().max()) == 0.0          # zero_grad fused into the optimizer pass
    params = dict(model.named_parameters())
    num = den = dot = n1 = n2 = 0.
    for k, want in tr['final'].items():
        p0 = sd[k].double()
        got = params[k].detach().double().cpu()
        du, dw = got - p0, want.double() - p0
        dot += float((du * dw).sum()); n1 += float(du.pow(2).sum()); n2 += float(dw.pow(2).sum())
        num += float((got - want.double()).pow(2).sum()); den += float(want.double().pow(2).sum())
    cos = dot / (n1 ** 0.5 * n2 ** 0.5)
    print(f"  after {tr['steps']} steps: parameters rel {(num / den) ** 0.5:.3e}, cosine(update, reference update) {cos:.4f}, "
          f"|update| ratio {(n1 / n2) ** 0.5:.4f}")
    assert (num / den) ** 0.5 < 3e-3 and cos > 0.98 and abs((n1 / n2) ** 0.5 - 1) < 0.02   # measured 1.95e-3, 0.9967, 1.0000


def test_trainer_step_equals_unfused_plumbing_and_graph_replay(cuda_device):
    """(1) TrainStep (persistent flat gradient buffer, in-place accumulation, grads_flat hand-over) == the same kernels
    driven the long way (autograd accumulation over two backward() calls, FusedAdamW gathering p.grad).  (2) The captured
    whole-step graph reproduces the eager steps, including the weight refresh between replays."""
    from nuwa_pytorch_b200.optim import FusedAdamW, trainable_parameters
    from nuwa_pytorch_b200.trainer import TrainStep
    tr = golden("trainer_small.pt")
    kw = dict(lr=tr['lr'], wd=tr['wd'], max_grad_norm=tr['max_norm'])
    _, m_a, _ = _model(cuda_device)
    _, m_b, _ = _model(cuda_device)
    _, m_c, _ = _model(cuda_device)
    step_a = TrainStep(m_a, grad_accum_every=tr['accum'], forward_kwargs=dict(cond_dropout_prob=0.), **kw)
    opt_b = FusedAdamW(trainable_parameters(m_b), **kw)
    step_c = TrainStep(m_c, grad_accum_every=tr['accum'], forward_kwargs=dict(cond_dropout_prob=0.), **kw)
    step_c.capture(_batches(tr, 0, cuda_device))
    for s in range(tr['steps']):
        la, na = step_a.step(_batches(tr, s, cuda_device))
        lb = 0.
        for b in _batches(tr, s, cuda_device):
            loss = m_b(**b, return_loss=True, cond_dropout_prob=0.)
            (loss / tr['accum']).backward()
            lb += loss.item() / tr['accum']
        nb = opt_b.step()
        opt_b.zero_grad()
        lc, nc = step_c.step(_batches(tr, s, cuda_device))
        print(f"  step {s}: loss fused {la.item():.5f} unfused {lb:.5f} graph {lc.item():.5f}; norm {na.item():.5f} "
              f"{nb.item():.5f} {nc.item():.5f}")
        # the three runs execute the same kernels; what differs run to run is the order of the fp32 reduce-adds (split-K
        # weight gradients, LayerNorm / talking-heads parameter gradients), which three Adam steps at lr 3e-3 amplify:
        # observed over repeated runs: loss up to 4e-4, norm up to 3e-4 relative, parameters up to 5e-4
        assert abs(la.item() - lb) < 3e-3 and abs(la.item() - lc.item()) < 3e-3
        assert abs(na.item() - nb.item()) < 3e-3 * nb.item() and abs(na.item() - nc.item()) < 3e-3 * nb.item()
    pa, pb, pc = (dict(m.named_parameters()) for m in (m_a, m_b, m_c))
    worst_b = max(rel(pa[k], pb[k]) for k in pa if not k.startswith('vae.'))
    worst_c = max(rel(pa[k], pc[k]) for k in pa if not k.startswith('vae.'))
    print(f"  parameters after {tr['steps']} steps: fused vs unfused rel {worst_b:.2e}, fused vs graph rel {worst_c:.2e}")
    assert worst_b < 5e-3 and worst_c < 5e-3


# ---------------------------------------------------------------------------------------------------------
# two ranks under NCCL
# ---------------------------------------------------------------------------------------------------------
def _nccl_worker(rank, world, port, q, done):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from nuwa_pytorch_b200.parallel import GradAllReduce, rank_slice
    from nuwa_pytorch_b200.trainer import TrainStep
    tr = golden("trainer_small.pt")
    fx, model, sd = _model(dev)
    B = tr['texts'].shape[2]
    lo, hi = rank_slice(B, rank, world)
    # (a) one backward with the overlapped all-reduce: the averaged gradient of the rank shards
    model._grad_reducer = GradAllReduce(dist)
    loss = model(text=tr['texts'][0, 0, lo:hi].to(dev), video=tr['videos'][0, 0, lo:hi].to(dev), return_loss=True,
                 cond_dropout_prob=0.)
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}
    model._grad_reducer = None
    for p in model.parameters():
        p.grad = None
    # (b) full trainer steps on the sharded micro-batches: eager on one model, and on a twin model step 0 eager, then the
    #     whole step captured in a CUDA graph WITH the collective inside, replayed for steps 1 and 2
    shard = lambda s: [dict(text=tr['texts'][s, a, lo:hi].to(dev), video=tr['videos'][s, a, lo:hi].to(dev))  # noqa: E731
                       for a in range(tr['accum'])]
    kw = dict(lr=tr['lr'], wd=tr['wd'], grad_accum_every=tr['accum'], max_grad_norm=tr['max_norm'], dist=dist,
              forward_kwargs=dict(cond_dropout_prob=0.))
    step = TrainStep(model, **kw)
    out = []
    for s in range(3):
        l, n = step.step(shard(s))
        out.append((l.item(), n.item()))
    torch.cuda.synchronize()
    final = {k: p.detach().cpu() for k, p in model.named_parameters() if not k.startswith('vae.')}
    _, twin, _ = _model(dev)
    gstep = TrainStep(twin, **kw)
    gout = [tuple(x.item() for x in gstep.step(shard(0)))]
    graph_ok, graph_err = True, ''
    try:
        gstep.capture(shard(1))
    except Exception as e:  # capture of the collective is an optimisation of the launch path; report, fall back to eager
        graph_ok, graph_err = False, f'{type(e).__name__}: {e}'
        torch.cuda.synchronize()
        gstep.graph = None
    for s in (1, 2):
        gout.append(tuple(x.item() for x in gstep.step(shard(s))))
    torch.cuda.synchronize()
    gfinal = {k: p.detach().cpu() for k, p in twin.named_parameters() if not k.startswith('vae.')}
    q.put((rank, grads, out, final, graph_ok, graph_err, gout, gfinal))
    done.wait(300)                  # the parent has rebuilt the tensors (they travel as shared-memory handles)
    torch.cuda.synchronize()
    # leave without tearing NCCL down: destroying a communicator whose kernels live inside a captured graph can block,
    # and nothing here needs an orderly shutdown
    os._exit(0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL)")
def test_two_rank_nccl_gradients_and_trainer_step(cuda_device):
    """Batch sharded over 2 ranks: (a) the NCCL-averaged gradient equals the single-process gradient of the whole batch;
    (b) three trainer steps leave both ranks with the same parameters, which follow the single-process trajectory of the
    unmodified reference (golden); the second and third step run from the captured graph when NCCL capture works."""
    tr = golden("trainer_small.pt")
    fx, model, sd = _model(cuda_device)
    loss = model(text=tr['texts'][0, 0].to(cuda_device), video=tr['videos'][0, 0].to(cuda_device), return_loss=True,
                 cond_dropout_prob=0.)
    loss.backward()
    want = {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    done = ctx.Event()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q, done)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda t: t[0])
    done.set()
    for p in procs:
        p.join(timeout=60)
        if p.exitcode is None:
            p.kill()
    g0, g1 = res[0][1], res[1][1]
    worst = max(rel(g0[k], want[k]) for k in want if want[k].numel() > 64)
    same = max(float((g0[k] - g1[k]).abs().max()) for k in g0)
    print(f"  2-rank NCCL: averaged gradient vs single-process whole-batch gradient worst rel {worst:.3e}; "
          f"rank 0 vs rank 1 max abs difference {same:.1e}")
    assert same == 0.0            # every rank holds the same reduced buffer
    assert worst < 2e-2           # two bf16 backward passes over half batches vs one over the whole batch
    for s in range(3):
        (l0, n0), (l1, n1) = res[0][2][s], res[1][2][s]
        mean_loss = 0.5 * (l0 + l1)
        print(f"  step {s}: mean loss over ranks {mean_loss:.5f} (reference whole batch {tr['losses'][s]:.5f}), grad norm "
              f"{n0:.5f} / {n1:.5f} (reference {tr['norms'][s].item():.5f})")
        assert abs(mean_loss - tr['losses'][s]) < 2e-2 and n0 == n1
        assert abs(n0 - tr['norms'][s].item()) < 3e-2 * tr['norms'][s].item()
    f0, f1 = res[0][3], res[1][3]
    assert all(torch.equal(f0[k], f1[k]) for k in f0)               # replicas stay bit-identical
    r = (sum(float((f0[k].double() - tr['final'][k].double()).pow(2).sum()) for k in f0) /
         sum(float(tr['final'][k].double().pow(2).sum()) for k in f0)) ** 0.5
    print(f"  parameters after 3 sharded steps vs the reference trajectory: rel {r:.3e}")
    assert r < 3e-3
    # the captured whole-step graph (collective inside) against the eager steps of the same rank
    print(f"  whole trainer step captured with the NCCL all-reduce inside the CUDA graph: {res[0][4]} {res[0][5]}")
    for rk in (0, 1):
        print(f"    rank {rk}: eager (loss, norm) {res[rk][2]}  graph {res[rk][6]}")
    if res[0][4] and res[1][4]:
        for rk in (0, 1):
            for (le, ne), (lg, ng) in zip(res[rk][2], res[rk][6]):
                assert abs(le - lg) < 4e-3 and abs(ne - ng) < 4e-3 * ne   # reduce-add order noise, see above
            g0 = res[rk][7]
            assert max(rel(g0[k], res[rk][3][k]) for k in g0) < 5e-3
