"""CPU suite: the C-ABI library loads and exports every symbol include/nuwa_b200.h declares; struct mirrors
match; host-side logic (weight packing, masks, state-dict keys) agrees with the reference goldens.  No compute
calls (there is no GPU here)."""
import os
import re

import pytest
import torch

from tests.helpers import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from nuwa_pytorch_b200 import _lib
    header = open(os.path.join(ROOT, "include", "nuwa_b200.h")).read()
    declared = set(re.findall(r"\b(nuwa_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    handle = _lib.lib()  # also checks struct sizes against the header's definitions
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/nuwa_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert handle.nuwa_abi_version() == 1
    assert handle.nuwa_strerror(-1).decode().startswith("invalid")


def test_product_fails_loudly_without_cuda():
    from nuwa_pytorch_b200 import VQGanVAE, _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    vae = VQGanVAE(dim=16, image_size=32, num_layers=2, use_vgg_and_gan=False, vq_kmeans_init=False).eval()
    with pytest.raises(_lib.NuwaB200Error):
        vae(torch.randn(1, 3, 32, 32))


def test_out_of_scope_paths_raise():
    from nuwa_pytorch_b200 import NUWA, VQGanVAE
    with pytest.raises(NotImplementedError):
        VQGanVAE(dim=16, image_size=32)  # default use_vgg_and_gan=True needs network VGG weights (D3)
    vae = VQGanVAE(dim=16, image_size=32, num_layers=2, use_vgg_and_gan=False)
    with pytest.raises(TypeError):
        NUWA(vae=vae, dim=32)  # enc_reversible=False cannot run in the reference either (D1)
    with pytest.raises(AttributeError):
        NUWA(image_size=32, dim=32)  # D2


def _keys(module):
    return [(k, tuple(v.shape), str(v.dtype)) for k, v in module.state_dict().items()]


def test_state_dict_keys_match_reference():
    from nuwa_pytorch_b200 import NUWA, NUWASketch, VQGanVAE
    for name in ("vae_cfg1.pt", "vae_small_l4.pt", "vae_euclid.pt"):
        fx = golden(name)
        assert _keys(VQGanVAE(**fx['kwargs'])) == [(k, tuple(s), d) for k, s, d in fx['keys']], name
    for name in ("nuwa_small.pt", "nuwa_rev_small.pt"):
        fx = golden(name)
        m = NUWA(vae=VQGanVAE(**fx['vae_kwargs']), **fx['kwargs'])
        assert _keys(m) == [(k, tuple(s), d) for k, s, d in fx['keys']], name
    fx = golden("sketch_small.pt")
    m = NUWASketch(vae=VQGanVAE(**fx['vae_kwargs']), sketch_vae=VQGanVAE(**fx['sketch_vae_kwargs']), **fx['kwargs'])
    assert _keys(m) == [(k, tuple(s), d) for k, s, d in fx['keys']]


def test_sparse3dna_mask_buffer_matches_reference():
    from nuwa_pytorch_b200 import Sparse3DNA
    for name, c in golden("sparse3dna_ops.pt").items():
        mod = Sparse3DNA(dim=64, video_shape=(3, 4, 4), kernel_size=c['kernel'], dilation=c['dilation'], heads=2,
                         dim_head=32, causal=c['causal'])
        assert torch.equal(mod.mask, c['mask']), name


def test_vae_fmap_size_attribute_quirk():
    from nuwa_pytorch_b200 import VQGanVAE
    assert VQGanVAE(dim=16, image_size=64, num_layers=3, use_vgg_and_gan=False).fmap_size == 7  # D5


def test_weight_packing_layouts():
    from nuwa_pytorch_b200 import ops
    w = torch.arange(2 * 20 * 3, dtype=torch.float32).reshape(40, 3)
    p = ops.pack_pairs(w)
    assert p.shape == (64, 3)
    assert torch.equal(p[:16], w[:16]) and torch.equal(p[16:32], w[20:36])      # value block, then its gates
    assert torch.equal(p[32:36], w[16:20]) and torch.equal(p[48:52], w[36:40])  # tail block, zero padded
    assert p[36:48].abs().sum() == 0 and p[52:].abs().sum() == 0
    cw = torch.randn(8, 5, 3, 3)
    pc = ops.pack_conv_weight(cw)
    assert pc.shape == (8, 9 * 64) and pc.dtype == torch.bfloat16
    assert torch.equal(pc.view(8, 9, 64)[:, 4, :5].float(), cw[:, :, 1, 1].bfloat16().float())
    assert pc.view(8, 9, 64)[:, :, 5:].abs().sum() == 0


def test_default_kernel_selection_switches():
    """The A/B switches the GPU tests flip (and restore) must ship in their product positions: tensor-core / fused paths on,
    kernel choice automatic."""
    from nuwa_pytorch_b200 import ops_bwd, train
    assert ops_bwd.WGRAD_TN is True            # weight gradients read dY / X contraction-major (no transposes)
    assert ops_bwd.DENSE_BWD_FUSED is True     # dense-attention backward: probability stage fused
    assert ops_bwd.Q1_KERNELS is True          # single-query dense attention kernels
    assert ops_bwd.SCORES_VARIANT == 'auto'    # Sparse3DNA backward scores / dq: tcgen05 kernel inside its envelope
    assert train.WGRAD_SIDE_STREAM is True     # weight-gradient GEMMs + key / value branches on the second stream
