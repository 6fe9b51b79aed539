"""TEST INFRASTRUCTURE ONLY.  Golden trajectory of the trainer step of the decoder path: the body of
NUWATrainer.train_step (reference train_nuwa.py:237-258) -- `grad_accum_every` micro-batches of
`loss = nuwa(text=, video=, return_loss=True); (loss / grad_accum_every).backward()`, then
`clip_grad_norm_(nuwa.parameters(), max_grad_norm)`, `optim.step()`, `optim.zero_grad()` with the optimizer of
`get_optimizer(nuwa.parameters(), lr, wd)` (optimizer.py:11-31) -- run with the UNMODIFIED reference model and optimizer
imported from /root/reference (build container only).  The NUWATrainer class itself wraps this body in a DataLoader, an
interactive prompt and checkpoint / sampling I/O, none of which is arithmetic; the micro-batches here are fixed tensors
and cond_dropout_prob is 0 (the default 0.2 draws an unseeded mask).

    python -m oracle.make_golden_trainer      # writes tests/golden/trainer_small.pt
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402
from oracle.synth import synth_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
LR, WD, MAX_NORM, ACCUM, STEPS, SEED = 3e-3, 0.01, 0.5, 2, 3, 41


def main():
    NP, VQ = import_reference()
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_optimizer", "/root/reference/nuwa_pytorch/optimizer.py")
    ref_opt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_opt)
    fx = torch.load(os.path.join(OUT, "nuwa_small.pt"))
    vae = VQ.VQGanVAE(**fx['vae_kwargs']).eval()
    nuwa = NP.NUWA(vae=vae, **fx['kwargs'])
    nuwa.load_state_dict(synth_state_dict(fx['manifest'], fx['seed']), strict=False)
    g = torch.Generator().manual_seed(SEED)
    B, T = fx['text'].shape
    vshape = fx['video_indices'].shape
    ntok = int(fx['kwargs']['text_num_tokens'])
    ncode = int(fx['vae_kwargs']['vq_codebook_size'])
    texts = torch.randint(1, ntok, (STEPS, ACCUM, B, T), generator=g)
    videos = torch.randint(0, ncode, (STEPS, ACCUM) + tuple(vshape), generator=g)
    optim = ref_opt.get_optimizer(nuwa.parameters(), lr=LR, wd=WD)
    losses, norms = [], []
    nuwa.train()
    for s in range(STEPS):
        acc = 0.
        for a in range(ACCUM):                                      # train_nuwa.py:243-254
            loss = nuwa(text=texts[s, a], video=videos[s, a], return_loss=True, cond_dropout_prob=0.)
            acc += loss.item() / ACCUM
            (loss / ACCUM).backward()
        norms.append(torch.nn.utils.clip_grad_norm_(nuwa.parameters(), MAX_NORM).clone())   # :256
        optim.step()                                                # :257
        optim.zero_grad()                                           # :258
        losses.append(acc)
        print(f"step {s}: loss {acc:.5f} grad norm {norms[-1].item():.5f}")
    final = {k: p.detach().clone() for k, p in nuwa.named_parameters() if not k.startswith('vae.')}
    path = os.path.join(OUT, "trainer_small.pt")
    torch.save(dict(lr=LR, wd=WD, max_norm=MAX_NORM, accum=ACCUM, steps=STEPS, texts=texts, videos=videos, losses=losses,
                    norms=norms, final=final), path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
