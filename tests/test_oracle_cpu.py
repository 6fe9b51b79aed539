"""CPU suite: the oracle (oracle/nuwa_oracle.py) against the golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py).  This is the oracle's pin; it runs without a GPU and without
/root/reference."""
import torch
import torch.nn.functional as F

from oracle import nuwa_oracle as O
from tests.helpers import (gen, golden, nuwa_spec_from_kwargs, rel, sketch_spec_from_kwargs, synth,
                           vae_spec_from_kwargs)

TOL = 1e-5  # fp32 vs fp32 on the same ATen CPU ops


def test_unfold_standin_matches_torch_2d():
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "standins"))
    from unfoldNd import unfoldNd
    x = torch.randn(2, 3, 9, 7, generator=gen(1))
    for k, d, p in ((3, 1, 1), (3, 2, 2), (5, 1, 0)):
        assert torch.equal(unfoldNd(x, kernel_size=k, dilation=d, padding=p), F.unfold(x, k, dilation=d, padding=p))


def test_vae_cfg1_golden():
    fx = golden("vae_cfg1.pt")
    sd, spec = synth(fx), vae_spec_from_kwargs(fx['kwargs'])
    img = torch.randn(4, 3, 64, 64, generator=gen(fx['input_seed']))
    quant, ind, loss = O.vae_encode(img, sd, spec)
    assert torch.equal(ind, fx['indices'])  # bit-exact token ids
    assert rel(quant[:1], fx['quant_sample']) < TOL
    assert rel(O.vae_encode_fmap(img, sd, spec)[:1], fx['pre_vq_fmap']) < TOL
    assert rel(O.vae_forward(img, sd, spec), fx['recon']) < TOL
    assert fx['fmap_size_attr'] == 64 // 3 ** 2  # D5: the reference's (wrong) attribute value is 7


def test_vae_small_golden():
    fx = golden("vae_small_l4.pt")
    sd, spec = synth(fx), vae_spec_from_kwargs(fx['kwargs'])
    idx = torch.randint(0, 64, (2, 32), generator=gen(fx['idx_seed']))
    video = torch.randn(2, 3, 3, 64, 64, generator=gen(fx['video_seed']))
    assert rel(O.vae_codebook_indices_to_video(idx, sd, spec, fx['fmap_size_attr']), fx['video']) < TOL
    assert torch.equal(O.vae_get_video_indices(video, sd, spec), fx['video_indices'])


def test_vae_euclid_golden():
    fx = golden("vae_euclid.pt")
    sd, spec = synth(fx), vae_spec_from_kwargs(fx['kwargs'])
    img = torch.randn(3, 3, 32, 32, generator=gen(fx['input_seed']))
    assert torch.equal(O.vae_encode(img, sd, spec)[1], fx['indices'])
    assert rel(O.vae_forward(img, sd, spec), fx['recon']) < TOL


def test_nuwa_small_golden():
    fx = golden("nuwa_small.pt")
    sd, spec = synth(fx), nuwa_spec_from_kwargs(fx['kwargs'], fx['vae_kwargs'])
    text, vidx = fx['text'], fx['video_indices']
    emb, mask = O.nuwa_embed_text(text, sd, spec)
    assert rel(emb, fx['text_embeds']) < TOL
    logits, loss = O.nuwa_logits(text, vidx.reshape(2, -1), sd, spec)
    assert rel(logits, fx['logits']) < TOL and abs(loss.item() - fx['loss'].item()) < 1e-5


def test_nuwa_rev_golden_and_generate_steps():
    fx = golden("nuwa_rev_small.pt")
    sd, spec = synth(fx), nuwa_spec_from_kwargs(fx['kwargs'], fx['vae_kwargs'])
    text, vidx = fx['text'], fx['video_indices']
    _, loss = O.nuwa_logits(text, vidx.reshape(2, -1), sd, spec)
    assert abs(loss.item() - fx['loss'].item()) < 1e-5
    emb, mask = O.nuwa_embed_text(text, sd, spec)
    for plen, lg in fx['step_logits'].items():
        og = O.nuwa_generate_step_logits(emb, mask, vidx.reshape(2, -1)[:, :plen], sd, spec, 2.)
        assert rel(og, lg) < 5e-5, plen


def test_sketch_small_golden():
    fx = golden("sketch_small.pt")
    sd, spec = synth(fx), sketch_spec_from_kwargs(fx['kwargs'], fx['vae_kwargs'])
    fi = fx['video_indices'].reshape(2, -1)
    for nf, case in fx['cases'].items():
        emb, m = O.sketch_embed(case['sketch_indices'], case['sketch_mask'], sd, spec)
        assert rel(emb, case['sketch_embeds']) < 5e-5, nf
        lg, ls = O.sketch_logits(case['sketch_indices'], case['sketch_mask'], fi, sd, spec)
        assert rel(lg, case['logits']) < 5e-5 and abs(ls.item() - case['loss'].item()) < 1e-5


def test_sparse3dna_ops_golden_and_mask():
    ops = golden("sparse3dna_ops.pt")
    from oracle.synth import synth_state_dict
    for name, c in ops.items():
        sd = synth_state_dict(c['manifest'], c['seed'])
        x = torch.randn(2, c['n'], 64, generator=gen(c['x_seed']))
        y = O.sparse3dna(x, sd, 2, (3, 4, 4), c['kernel'], (c['dilation'],) * 3, c['causal'])
        assert rel(y, c['y']) < TOL, name
        # the persistent `mask` buffer (nuwa_pytorch.py:444-457) equals the oracle's out-of-grid predicate
        _, _, masked = O.sparse3dna_neighbours(48, (3, 4, 4), c['kernel'], (c['dilation'],) * 3, c['causal'], 3)
        assert torch.equal(F.pad(masked, (1, 0), value=False), c['mask']), name


def test_sparse3dna_is_prefix_consistent():
    """Property the incremental (KV-cached) decode relies on."""
    ops = golden("sparse3dna_ops.pt")
    from oracle.synth import synth_state_dict
    c = ops["causal_k533_d2_part"]
    sd = synth_state_dict(c['manifest'], c['seed'])
    x = torch.randn(1, 40, 64, generator=gen(7))
    full = O.sparse3dna(x, sd, 2, (3, 4, 4), c['kernel'], (2, 2, 2), True)
    for n in (1, 2, 17, 33):
        part = O.sparse3dna(x[:, :n], sd, 2, (3, 4, 4), c['kernel'], (2, 2, 2), True)
        assert rel(part, full[:, :n]) < 1e-6


def test_vq_tie_break_lowest_index():
    embed = F.normalize(torch.randn(8, 4, generator=gen(3)), dim=-1)
    embed[5] = embed[2]
    x = embed[2:3] * 3.0
    assert O.vq_lookup(x, embed).item() == 2


def test_topk_gumbel():
    lg = torch.randn(3, 100, generator=gen(5))
    f = O.top_k_filter(lg, 0.9)
    k = max(int((1 - 0.9) * 100), 1)  # = 9 (float rounding), nuwa_pytorch.py:1715
    assert (f > float("-inf")).sum(-1).tolist() == [k, k, k]
    u = torch.rand(3, 100, generator=gen(6))
    s = O.gumbel_argmax(f, u)
    assert all(f[i, s[i]] > float('-inf') for i in range(3))


def test_oracle_autograd_matches_reference_gradients():
    """The oracle is also the checker of the CUDA backward: its autograd must reproduce the gradients the UNMODIFIED
    reference produced for `loss.backward()` (tests/golden/*_grads.pt, oracle/make_golden_grads.py)."""
    import torch
    from tests.helpers import golden, nuwa_spec_from_kwargs, rel, synth
    for name in ("nuwa_small", "nuwa_rev_small"):
        fx, gold = golden(name + ".pt"), golden(name + "_grads.pt")
        sd = synth(fx)
        spec = nuwa_spec_from_kwargs(fx['kwargs'], fx['vae_kwargs'])
        leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and not k.startswith('vae.')}
        full = dict(sd)
        full.update(leaves)
        _, loss = O.nuwa_logits(fx['text'], fx['video_indices'].reshape(2, -1), full, spec)
        loss.backward()
        assert abs(loss.item() - gold['loss'].item()) < 1e-5
        worst = max(rel(leaves[k].grad, g) for k, g in gold['grads'].items())
        assert worst < 3e-4, (name, worst)
