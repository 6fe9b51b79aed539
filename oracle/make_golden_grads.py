"""TEST INFRASTRUCTURE ONLY.  Golden GRADIENTS of the training step `loss = nuwa(text=, video=, return_loss=True);
loss.backward()` (SURVEY §8d cfg3 timed region), produced by the UNMODIFIED reference imported from /root/reference
(build container only) on the `nuwa_small` / `nuwa_rev_small` fixtures' weights and inputs, and used to pin the autograd
of the oracle restatement (oracle/nuwa_oracle.py) -- which in turn is the checker for the CUDA backward path.

    python -m oracle.make_golden_grads      # writes tests/golden/nuwa_small_grads.pt, nuwa_rev_small_grads.pt
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nuwa_oracle as O  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402
from oracle.synth import manifest_of, synth_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def oracle_grads(fx, spec, sd):
    """Autograd through the functional oracle; returns {state-dict key: grad} for every float leaf that got one."""
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and not k.startswith('vae.')}
    full = dict(sd)
    full.update(leaves)
    _, loss = O.nuwa_logits(fx['text'], fx['video_indices'].reshape(fx['text'].shape[0], -1), full, spec)
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in leaves.items() if v.grad is not None}


def main():
    NP, VQ = import_reference()
    for name, spec_kw in (("nuwa_small", dict(dec_depth=3, kernel=(5, 3, 3), dilation=(1, 2, 4))),
                          ("nuwa_rev_small", dict(dec_depth=2, dec_reversible=True, kernel=3, dilation=2))):
        fx = torch.load(os.path.join(OUT, name + ".pt"))
        vae = VQ.VQGanVAE(**fx['vae_kwargs']).eval()
        nuwa = NP.NUWA(vae=vae, **fx['kwargs'])
        sd = synth_state_dict(fx['manifest'], fx['seed'])
        nuwa.load_state_dict(sd, strict=False)
        nuwa.train()
        loss = nuwa(text=fx['text'], video=fx['video_indices'], return_loss=True, cond_dropout_prob=0.)
        loss.backward()
        assert abs(loss.item() - fx['loss'].item()) < 1e-6
        grads = {k: p.grad.clone() for k, p in nuwa.named_parameters() if p.grad is not None}
        assert not any(k.startswith('vae.') for k in grads)
        spec = O.NUWASpec(64, 4, 3, 64, text_enc_dim_head=32, text_enc_depth=2, text_enc_heads=2, dec_heads=2, **spec_kw)
        oloss, og = oracle_grads(fx, spec, sd)
        assert abs(oloss.item() - loss.item()) < 1e-6
        worst = 0.
        for k, g in grads.items():
            assert k in og, k
            worst = max(worst, rel(og[k], g))
            if rel(og[k], g) > 2e-4:
                print('  MISMATCH', k, rel(og[k], g))
        print(f"{name}: {len(grads)} gradient tensors, oracle-vs-reference worst rel {worst:.2e}")
        assert worst < 2e-4
        path = os.path.join(OUT, name + "_grads.pt")
        torch.save(dict(loss=loss.detach(), grads=grads), path)
        print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")

    # ---------------- NUWASketch (cfg-5 shape of path: 3DNA sketch encoder, SparseCross2DNA decoder), end to end ----------
    fx = torch.load(os.path.join(OUT, "sketch_small.pt"))
    vae = VQ.VQGanVAE(**fx['vae_kwargs']).eval()
    vae.load_state_dict(synth_state_dict(manifest_of(vae.state_dict()), fx['vae_seed']), strict=False)
    svae = VQ.VQGanVAE(**fx['sketch_vae_kwargs']).eval()
    svae.load_state_dict(synth_state_dict(manifest_of(svae.state_dict()), fx['sketch_vae_seed']), strict=False)
    sk = NP.NUWASketch(vae=vae, sketch_vae=svae, **fx['kwargs'])
    sd = synth_state_dict(fx['manifest'], fx['seed'])
    sk.load_state_dict(sd, strict=False)
    sk.train()
    g = lambda seed: torch.Generator().manual_seed(seed)  # noqa: E731
    sketch = torch.randn(2, 3, 5, 64, 64, generator=g(fx['e2e_sketch_seed']))
    video = torch.randn(2, 3, 3, 64, 64, generator=g(fx['e2e_video_seed']))
    smask = torch.ones(2, 3, dtype=torch.bool)
    loss = sk(sketch=sketch, sketch_mask=smask.clone(), video=video, return_loss=True, cond_dropout_prob=0.)
    loss.backward()
    assert abs(loss.item() - fx['e2e_loss'].item()) < 1e-6
    grads = {k: p.grad.clone() for k, p in sk.named_parameters() if p.grad is not None}
    assert not any(k.startswith('vae.') or k.startswith('sketch_vae.') for k in grads)
    with torch.no_grad():
        sidx = sk.sketch_vae.get_video_indices(sketch)
        fi = sk.vae.get_video_indices(video).reshape(2, -1)
    spec = O.SketchSpec(64, 4, 3, 3, 64, sketch_enc_depth=2, sketch_enc_heads=2, sketch_enc_use_sparse_3dna=True,
                        dec_depth=3, dec_heads=2, kernel=(5, 3, 3), dilation=(1, 2), cross_dilation=2)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and not k.startswith('vae.') and not k.startswith('sketch_vae.')}
    full = dict(sd)
    full.update(leaves)
    _, oloss = O.sketch_logits(sidx, smask, fi, full, spec)
    oloss.backward()
    worst = max(rel(leaves[k].grad, gr) for k, gr in grads.items())
    print(f"sketch_small: {len(grads)} gradient tensors, oracle-vs-reference worst rel {worst:.2e}")
    assert worst < 2e-4 and abs(oloss.item() - loss.item()) < 1e-6
    path = os.path.join(OUT, "sketch_small_grads.pt")
    torch.save(dict(loss=loss.detach(), grads=grads), path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")

    # ---------------- self-contained variants: code paths the three fixtures above do not reach ----------
    variants = {}
    gen_ = lambda seed: torch.Generator().manual_seed(seed)  # noqa: E731
    vae_kw = torch.load(os.path.join(OUT, "nuwa_small.pt"))['vae_kwargs']
    # (A) NUWA with learned absolute text positions (text_rotary_pos_emb=False), 4 heads x 16, dilation 1
    kwa = dict(dim=64, text_num_tokens=50, text_max_seq_len=10, text_enc_depth=1, text_enc_heads=4, text_enc_dim_head=16,
               text_rotary_pos_emb=False, enc_reversible=True, max_video_frames=2, dec_depth=2, dec_heads=4, dec_dim_head=16,
               sparse_3dna_kernel_size=3, sparse_3dna_dilation=1)
    ma = NP.NUWA(vae=VQ.VQGanVAE(**vae_kw).eval(), **kwa)
    mana = manifest_of(ma.state_dict())
    sda = synth_state_dict(mana, 41)
    ma.load_state_dict(sda, strict=False)
    ma.train()
    text = torch.randint(1, 50, (3, 10), generator=gen_(501))
    text[0, 6:] = 0
    vidx = torch.randint(0, 64, (3, 2, 4, 4), generator=gen_(502))
    loss = ma(text=text, video=vidx, return_loss=True, cond_dropout_prob=0.)
    loss.backward()
    ga = {k: p.grad.clone() for k, p in ma.named_parameters() if p.grad is not None}
    speca = O.NUWASpec(64, 4, 2, 64, text_enc_dim_head=16, text_enc_depth=1, text_enc_heads=4, dec_depth=2, dec_heads=4,
                       kernel=3, dilation=1)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sda.items() if v.is_floating_point() and not k.startswith('vae.')}
    full = dict(sda)
    full.update(leaves)
    _, ol = O.nuwa_logits(text, vidx.reshape(3, -1), full, speca)
    ol.backward()
    worst = max(rel(leaves[k].grad, gr) for k, gr in ga.items())
    print(f"variant abs_pos: {len(ga)} gradient tensors, oracle-vs-reference worst rel {worst:.2e}")
    assert worst < 3e-4 and abs(ol.item() - loss.item()) < 1e-6
    variants['nuwa_abs_pos'] = dict(kwargs=kwa, vae_kwargs=vae_kw, manifest=mana, seed=41, text=text, video_indices=vidx,
                                    loss=loss.detach(), grads=ga)
    # (B) NUWASketch: dense sketch encoder, reversible decoder (SparseCross2DNA inside a reversible block), masked frame
    fxs = torch.load(os.path.join(OUT, "sketch_small.pt"))
    kwb = dict(dim=64, image_size=64, sketch_max_video_frames=2, sketch_enc_depth=1, sketch_enc_heads=2, sketch_enc_dim_head=32,
               sketch_enc_use_sparse_3dna=False, max_video_frames=2, dec_depth=2, dec_heads=2, dec_dim_head=32,
               dec_reversible=True, sparse_3dna_kernel_size=3, sparse_3dna_dilation=2, cross_2dna_dilation=1)
    mb = NP.NUWASketch(vae=VQ.VQGanVAE(**fxs['vae_kwargs']).eval(), sketch_vae=VQ.VQGanVAE(**fxs['sketch_vae_kwargs']).eval(), **kwb)
    manb = manifest_of(mb.state_dict())
    sdb = synth_state_dict(manb, 42)
    mb.load_state_dict(sdb, strict=False)
    mb.train()
    sketch = torch.randn(2, 2, 5, 64, 64, generator=gen_(503))
    video = torch.randn(2, 2, 3, 64, 64, generator=gen_(504))
    smask = torch.tensor([[True, True], [True, False]])
    loss = mb(sketch=sketch, sketch_mask=smask.clone(), video=video, return_loss=True, cond_dropout_prob=0.)
    loss.backward()
    gb = {k: p.grad.clone() for k, p in mb.named_parameters() if p.grad is not None}
    with torch.no_grad():
        sidx = mb.sketch_vae.get_video_indices(sketch)
        fi = mb.vae.get_video_indices(video).reshape(2, -1)
    specb = O.SketchSpec(64, 4, 2, 2, 64, sketch_enc_depth=1, sketch_enc_heads=2, sketch_enc_use_sparse_3dna=False,
                         dec_depth=2, dec_heads=2, dec_reversible=True, kernel=3, dilation=2, cross_dilation=1)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sdb.items()
              if v.is_floating_point() and not k.startswith('vae.') and not k.startswith('sketch_vae.')}
    full = dict(sdb)
    full.update(leaves)
    _, ol = O.sketch_logits(sidx, smask, fi, full, specb)
    ol.backward()
    worst = max(rel(leaves[k].grad, gr) for k, gr in gb.items())
    print(f"variant sketch_dense_rev: {len(gb)} gradient tensors, oracle-vs-reference worst rel {worst:.2e}")
    assert worst < 3e-4 and abs(ol.item() - loss.item()) < 1e-6
    variants['sketch_dense_rev'] = dict(kwargs=kwb, vae_kwargs=fxs['vae_kwargs'], sketch_vae_kwargs=fxs['sketch_vae_kwargs'],
                                        manifest=manb, seed=42, sketch_indices=sidx, sketch_mask=smask, frame_indices=fi,
                                        loss=loss.detach(), grads=gb)
    path = os.path.join(OUT, "train_variants.pt")
    torch.save(variants, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
