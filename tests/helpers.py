"""Shared helpers for the parity tests (oracle = checker only)."""
import os

import torch

from oracle import nuwa_oracle as O
from oracle.synth import synth_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def gen(seed):
    return torch.Generator().manual_seed(seed)


def vae_spec_from_kwargs(kw):
    return O.VAESpec(kw['dim'], kw['image_size'], channels=kw.get('channels', 3), num_layers=kw.get('num_layers', 4),
                     num_resnet_blocks=kw.get('num_resnet_blocks', 1), codebook_dim=kw.get('vq_codebook_dim', 256),
                     codebook_size=kw.get('vq_codebook_size', 512), use_cosine_sim=kw.get('vq_use_cosine_sim', True),
                     attn_heads=kw.get('attn_heads', 8), attn_dim_head=kw.get('attn_dim_head', 64))


def nuwa_spec_from_kwargs(kw, vae_kw):
    fmap = vae_kw['image_size'] // 2 ** vae_kw.get('num_layers', 4)
    return O.NUWASpec(kw['dim'], fmap, kw.get('max_video_frames', 5), vae_kw.get('vq_codebook_size', 512),
                      text_enc_depth=kw.get('text_enc_depth', 6), text_enc_heads=kw.get('text_enc_heads', 8),
                      text_enc_dim_head=kw.get('text_enc_dim_head', 64), dec_depth=kw.get('dec_depth', 6),
                      dec_heads=kw.get('dec_heads', 8), dec_reversible=kw.get('dec_reversible', False),
                      enc_reversible=kw.get('enc_reversible', False), kernel=kw.get('sparse_3dna_kernel_size', 3),
                      dilation=kw.get('sparse_3dna_dilation', 1), shift_video_tokens=kw.get('shift_video_tokens', True))


def sketch_spec_from_kwargs(kw, vae_kw):
    fmap = kw['image_size'] // 2 ** vae_kw.get('num_layers', 4)
    return O.SketchSpec(kw['dim'], fmap, kw.get('max_video_frames', 5), kw.get('sketch_max_video_frames', 2),
                        vae_kw.get('vq_codebook_size', 512), sketch_enc_depth=kw.get('sketch_enc_depth', 6),
                        sketch_enc_heads=kw.get('sketch_enc_heads', 8),
                        sketch_enc_use_sparse_3dna=kw.get('sketch_enc_use_sparse_3dna', False),
                        enc_reversible=kw.get('enc_reversible', False), dec_depth=kw.get('dec_depth', 6),
                        dec_heads=kw.get('dec_heads', 8), dec_reversible=kw.get('dec_reversible', False),
                        kernel=kw.get('sparse_3dna_kernel_size', 3), dilation=kw.get('sparse_3dna_dilation', 1),
                        cross_kernel=kw.get('cross_2dna_kernel_size', 3), cross_dilation=kw.get('cross_2dna_dilation', 1),
                        shift_video_tokens=kw.get('shift_video_tokens', True))


def synth(fix):
    return synth_state_dict(fix['manifest'], fix['seed'])


def assert_ids_equal_up_to_fp32_ties(got, want, x, code, cosine, tie=2e-6):
    """VQ token ids must be EQUAL to the fp32 reference arg-max.  The only admissible difference is an fp32 tie: two
    codes whose exact (fp64) scores differ by less than `tie` relative -- below what fp32 summation order can resolve, on
    the CPU reference as much as on the GPU.  Every flip is printed with its fp64 margin; returns the flip count."""
    import torch.nn.functional as F
    got, want = got.reshape(-1).cpu(), want.reshape(-1).cpu()
    flips = (got != want).nonzero().reshape(-1).tolist()
    for m in flips:
        xr, cg, cw = x[m].double(), code[got[m]].double(), code[want[m]].double()
        if cosine:
            xr = F.normalize(xr, dim=-1)
            sg, sw = xr @ F.normalize(cg, dim=-1), xr @ F.normalize(cw, dim=-1)
            scale = 1.0
        else:
            sg, sw = -(xr - cg).pow(2).sum(), -(xr - cw).pow(2).sum()
            scale = float(xr.pow(2).sum() + cw.pow(2).sum())
        margin = abs(float(sg - sw)) / scale
        print(f"    id flip at token {m}: got {int(got[m])} want {int(want[m])}, fp64 score margin {margin:.2e}")
        assert margin < tie, f"token {m}: ids differ ({int(got[m])} vs {int(want[m])}) with a decisive fp64 margin {margin:.3e}"
    return len(flips)
