"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference, build container only) on deterministic synthetic weights and inputs.

Each fixture stores: the constructor kwargs, the state-dict key list (name, shape, dtype) of the reference
module (the drop-in contract for checkpoints), the synth seed + manifest (oracle/synth.py) and the
reference's outputs.  Run:  PYTHONPATH=. python oracle/make_golden.py
The script also asserts that oracle/nuwa_oracle.py reproduces every stored output (the oracle's pin).
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


def keylist(sd):
    return [(k, tuple(v.shape), str(v.dtype)) for k, v in sd.items()]


def load_synth(module, seed):
    man = manifest_of(module.state_dict())
    sd = synth_state_dict(man, seed)
    module.load_state_dict(sd, strict=False)
    return man, sd


def gen(seed):
    return torch.Generator().manual_seed(seed)


def save(name, obj):
    path = os.path.join(OUT, name)
    torch.save(obj, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def main():
    NP, VQ = import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)

    # ---------------- config 1: VQGanVAE dim=64 image 64 L=3, batch 4 (BASELINE.json configs[0]) ----------
    kw = dict(dim=64, image_size=64, num_layers=3, use_vgg_and_gan=False, vq_kmeans_init=False)
    vae = VQ.VQGanVAE(**kw).eval()
    man, sd = load_synth(vae, 11)
    img = torch.randn(4, 3, 64, 64, generator=gen(101))
    with torch.no_grad():
        quant, ind, loss = vae.encode(img)
        recon = vae(img)
        fmap = img
        for e in vae.encoders:
            fmap = e(fmap)
    spec = O.VAESpec(64, 64, num_layers=3)
    oq, oi, _ = O.vae_encode(img, sd, spec)
    assert torch.equal(oi, ind) and rel(oq, quant) < 1e-6 and rel(O.vae_forward(img, sd, spec), recon) < 1e-6
    save("vae_cfg1.pt", dict(kwargs=kw, keys=keylist(vae.state_dict()), manifest=man, seed=11, input_seed=101,
                             fmap_size_attr=vae.fmap_size, indices=ind, recon=recon, pre_vq_fmap=fmap[:1].clone(),
                             quant_sample=quant[:1].clone()))

    # ---------------- small L=4 VAE, codebook_dim = dim*8 (D4) : indices -> video, video -> indices ----------
    kw2 = dict(dim=16, image_size=64, num_layers=4, use_vgg_and_gan=False, vq_kmeans_init=False,
               vq_codebook_size=64, vq_codebook_dim=128, attn_heads=2, attn_dim_head=16, num_resnet_blocks=2)
    vae2 = VQ.VQGanVAE(**kw2).eval()
    man2, sd2 = load_synth(vae2, 12)
    spec2 = O.VAESpec(16, 64, num_layers=4, codebook_dim=128, codebook_size=64, attn_heads=2, attn_dim_head=16,
                      num_resnet_blocks=2)
    idx_in = torch.randint(0, 64, (2, 2 * 16), generator=gen(102))
    video_in = torch.randn(2, 3, 3, 64, 64, generator=gen(103))
    with torch.no_grad():
        vid = vae2.codebook_indices_to_video(idx_in)
        vind = vae2.get_video_indices(video_in)
    assert rel(O.vae_codebook_indices_to_video(idx_in, sd2, spec2, vae2.fmap_size), vid) < 1e-6
    assert torch.equal(O.vae_get_video_indices(video_in, sd2, spec2), vind)
    save("vae_small_l4.pt", dict(kwargs=kw2, keys=keylist(vae2.state_dict()), manifest=man2, seed=12,
                                 idx_seed=102, video_seed=103, fmap_size_attr=vae2.fmap_size, video=vid,
                                 video_indices=vind))

    # ---------------- euclidean codebook VAE (vq_use_cosine_sim=False) ----------
    kw3 = dict(dim=16, image_size=32, num_layers=2, use_vgg_and_gan=False, vq_kmeans_init=False,
               vq_codebook_size=32, vq_codebook_dim=32, vq_use_cosine_sim=False, attn_heads=2, attn_dim_head=16)
    vae3 = VQ.VQGanVAE(**kw3).eval()
    man3, sd3 = load_synth(vae3, 13)
    spec3 = O.VAESpec(16, 32, num_layers=2, codebook_dim=32, codebook_size=32, use_cosine_sim=False, attn_heads=2,
                      attn_dim_head=16)
    img3 = torch.randn(3, 3, 32, 32, generator=gen(104))
    with torch.no_grad():
        _, ind3, _ = vae3.encode(img3)
        rec3 = vae3(img3)
    assert torch.equal(O.vae_encode(img3, sd3, spec3)[1], ind3) and rel(O.vae_forward(img3, sd3, spec3), rec3) < 1e-6
    save("vae_euclid.pt", dict(kwargs=kw3, keys=keylist(vae3.state_dict()), manifest=man3, seed=13, input_seed=104,
                               indices=ind3, recon=rec3))

    # ---------------- NUWA small (plain decoder, kernel (5,3,3), dilation cycle (1,2,4)) ----------
    nkw = dict(dim=64, text_num_tokens=100, text_max_seq_len=12, text_enc_depth=2, text_enc_heads=2,
               text_enc_dim_head=32, enc_reversible=True, max_video_frames=3, dec_depth=3, dec_heads=2,
               dec_dim_head=32, sparse_3dna_kernel_size=(5, 3, 3), sparse_3dna_dilation=(1, 2, 4))
    nuwa = NP.NUWA(vae=vae2, **nkw).eval()
    mann, sdn = load_synth(nuwa, 21)
    ns = O.NUWASpec(64, 4, 3, 64, text_enc_dim_head=32, text_enc_depth=2, text_enc_heads=2, dec_depth=3, dec_heads=2,
                    kernel=(5, 3, 3), dilation=(1, 2, 4))
    text = torch.randint(1, 100, (2, 12), generator=gen(201))
    text[1, 8:] = 0
    vidx = torch.randint(0, 64, (2, 3, 4, 4), generator=gen(202))
    with torch.no_grad():
        loss = nuwa(text=text, video=vidx, return_loss=True)
        temb = nuwa.embed_text(text, mask=text != 0)
        # logits (return_loss=True path recomputed to expose them)
        fi = vidx.reshape(2, -1)
        x = nuwa.image_embedding(fi[:, :-1]) + nuwa.video_pos_emb()[:-1]
        x = torch.cat((nuwa.video_bos[None, None].expand(2, 1, -1), x), 1)
        logits = nuwa.to_logits(nuwa.video_transformer(x, context=temb, context_mask=text != 0))
    ologits, oloss = O.nuwa_logits(text, fi, sdn, ns)
    assert rel(O.nuwa_embed_text(text, sdn, ns)[0], temb) < 1e-6 and rel(ologits, logits) < 1e-6
    assert abs(oloss.item() - loss.item()) < 1e-6
    save("nuwa_small.pt", dict(kwargs=nkw, vae_kwargs=kw2, keys=keylist(nuwa.state_dict()), manifest=mann, seed=21,
                               text=text, video_indices=vidx, loss=loss, text_embeds=temb, logits=logits))

    # ---------------- NUWA small, reversible decoder + generate-step logits (D8) ----------
    rkw = dict(dim=64, text_num_tokens=100, text_max_seq_len=12, text_enc_depth=2, text_enc_heads=2,
               text_enc_dim_head=32, enc_reversible=True, max_video_frames=3, dec_depth=2, dec_heads=2,
               dec_dim_head=32, dec_reversible=True, sparse_3dna_kernel_size=3, sparse_3dna_dilation=2)
    nuwar = NP.NUWA(vae=vae2, **rkw).eval()
    manr, sdr = load_synth(nuwar, 22)
    nsr = O.NUWASpec(64, 4, 3, 64, text_enc_dim_head=32, text_enc_depth=2, text_enc_heads=2, dec_depth=2,
                     dec_heads=2, dec_reversible=True, kernel=3, dilation=2)
    with torch.no_grad():
        lossr = nuwar(text=text, video=vidx, return_loss=True)
        tembr = nuwar.embed_text(text, mask=text != 0)
    steps = {}
    for plen in (0, 1, 5, 16, 21, 40):
        part = vidx.reshape(2, -1)[:, :plen]
        with torch.no_grad():
            pos = nuwar.video_pos_emb()
            fe = nuwar.image_embedding(part)
            fe = pos[:fe.shape[1]] + fe
            fe = torch.cat((nuwar.video_bos[None, None].expand(2, 1, -1), fe), 1)
            fe = nuwar.video_transformer(fe, context=tembr, context_mask=text != 0)
            lg = nuwar.to_logits(fe)
            ufe = nuwar.video_transformer(fe, context=tembr, context_mask=torch.zeros_like(text).bool())
            ul = nuwar.to_logits(ufe)
            lg = (ul + (lg - ul) * 2.)[:, -1]
        og = O.nuwa_generate_step_logits(O.nuwa_embed_text(text, sdr, nsr)[0], text != 0, part, sdr, nsr, 2.)
        assert rel(og, lg) < 1e-5, (plen, rel(og, lg))
        steps[plen] = lg
    assert abs(O.nuwa_logits(text, vidx.reshape(2, -1), sdr, nsr)[1].item() - lossr.item()) < 1e-6
    save("nuwa_rev_small.pt", dict(kwargs=rkw, vae_kwargs=kw2, keys=keylist(nuwar.state_dict()), manifest=manr,
                                   seed=22, text=text, video_indices=vidx, loss=lossr, step_logits=steps))

    # ---------------- NUWASketch small (3dna sketch encoder, cross-2dna decoder) ----------
    skw = dict(dim=16, image_size=64, num_layers=4, channels=5, use_vgg_and_gan=False, vq_kmeans_init=False,
               vq_codebook_size=32, vq_codebook_dim=128, attn_heads=2, attn_dim_head=16)
    svae = VQ.VQGanVAE(**skw).eval()
    load_synth(svae, 14)
    kkw = dict(dim=64, image_size=64, sketch_max_video_frames=3, sketch_enc_depth=2, sketch_enc_heads=2,
               sketch_enc_dim_head=32, sketch_enc_use_sparse_3dna=True, max_video_frames=3, dec_depth=3, dec_heads=2,
               dec_dim_head=32, sparse_3dna_kernel_size=(5, 3, 3), sparse_3dna_dilation=(1, 2), cross_2dna_dilation=2)
    sk = NP.NUWASketch(vae=vae2, sketch_vae=svae, **kkw).eval()
    mank, sdk = load_synth(sk, 23)
    ss = O.SketchSpec(64, 4, 3, 3, 64, sketch_enc_depth=2, sketch_enc_heads=2, sketch_enc_use_sparse_3dna=True,
                      dec_depth=3, dec_heads=2, kernel=(5, 3, 3), dilation=(1, 2), cross_dilation=2)
    out = {}
    for nf in (3, 2):  # 2 < sketch_max_video_frames exercises D16 (visible zero keys in the non-causal encoder)
        sketch_idx = torch.randint(0, 32, (2, nf, 4, 4), generator=gen(300 + nf))
        smask = torch.ones(2, nf, dtype=torch.bool)
        smask[1, nf - 1] = False
        with torch.no_grad():
            # bypass the sketch VAE (its tokenisation is covered by the VAE fixtures): embed_sketch body
            st = sk.sketch_embedding(sketch_idx.reshape(2, -1))
            st = st + sk.sketch_pos_emb()[:st.shape[1]]
            m = smask[:, :, None].expand(2, nf, 16).reshape(2, -1)
            semb = sk.sketch_transformer(st, mask=m)
            fi = vidx.reshape(2, -1)
            x = sk.image_embedding(fi[:, :-1]) + sk.video_pos_emb()[:-1]
            x = torch.cat((sk.video_bos[None, None].expand(2, 1, -1), x), 1)
            lg = sk.to_logits(sk.video_transformer(x, context=semb, context_mask=m))
            ls = torch.nn.functional.cross_entropy(lg.transpose(1, 2), fi)
        oemb, om = O.sketch_embed(sketch_idx, smask, sdk, ss)
        olg, ols = O.sketch_logits(sketch_idx, smask, fi, sdk, ss)
        assert rel(oemb, semb) < 1e-5 and rel(olg, lg) < 1e-5, (nf, rel(oemb, semb), rel(olg, lg))
        out[nf] = dict(sketch_indices=sketch_idx, sketch_mask=smask, sketch_embeds=semb, logits=lg, loss=ls)
    # end-to-end through the sketch VAE once (float sketch input)
    sketch = torch.randn(2, 3, 5, 64, 64, generator=gen(310))
    video = torch.randn(2, 3, 3, 64, 64, generator=gen(311))
    with torch.no_grad():
        e2e_loss = sk(sketch=sketch, sketch_mask=torch.ones(2, 3, dtype=torch.bool), video=video, return_loss=True)
    save("sketch_small.pt", dict(kwargs=kkw, vae_kwargs=kw2, sketch_vae_kwargs=skw, vae_seed=12, sketch_vae_seed=14,
                                 keys=keylist(sk.state_dict()), manifest=mank, seed=23, video_indices=vidx,
                                 cases=out, e2e_sketch_seed=310, e2e_video_seed=311, e2e_loss=e2e_loss))

    # ---------------- op-level: Sparse3DNA (exported class), partial sequences, non-causal ----------
    ops = {}
    for name, causal, ksz, dil, n in (("causal_k533_d1_full", True, (5, 3, 3), 1, 49), ("causal_k533_d2_part", True, (5, 3, 3), 2, 30),
                                      ("causal_k3_d4_part", True, 3, 4, 21), ("noncausal_k3_d1_part", False, 3, 1, 23),
                                      ("noncausal_k533_d2_full", False, (5, 3, 3), 2, 48), ("bos_only", True, 3, 1, 1)):
        mod = NP.Sparse3DNA(dim=64, video_shape=(3, 4, 4), kernel_size=ksz, dilation=dil, heads=2, dim_head=32,
                            causal=causal).eval()
        manm, sdm = load_synth(mod, 31)
        x = torch.randn(2, n, 64, generator=gen(400 + n))
        with torch.no_grad():
            y = mod(x)
        k3 = ksz if isinstance(ksz, tuple) else (ksz,) * 3
        oy = O.sparse3dna(x, sdm, 2, (3, 4, 4), k3, (dil,) * 3, causal)
        assert rel(oy, y) < 1e-5, (name, rel(oy, y))
        ops[name] = dict(causal=causal, kernel=k3, dilation=dil, n=n, x_seed=400 + n, y=y, mask=mod.mask.clone(),
                         manifest=manm, seed=31)
    save("sparse3dna_ops.pt", ops)
    print("all reference outputs reproduced by the oracle")


if __name__ == "__main__":
    main()
