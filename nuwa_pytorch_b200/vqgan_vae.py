"""VQGanVAE on sm_100a kernels -- drop-in for nuwa_pytorch/vqgan_vae.py:288-548 (inference hot path).

The nn.Module tree below is a *parameter container* that reproduces the reference's state-dict keys
(encoders.N..., decoders.N..., vq.project_in/out, vq._codebook.{initted,cluster_size,embed}); none of the
contained torch modules is ever called.  All arithmetic runs in libnuwa_b200.so:
  activations are NHWC bf16 in HBM, every convolution is an implicit-GEMM tcgen05 kernel with fused
  bias / LeakyReLU(0.1) / GLU / residual epilogues, GroupNorm / attention / VQ arg-max are dedicated
  kernels, and the VQ similarity is fp32 so token ids are bit-exact for identical inputs.

Out of scope (raises NotImplementedError, never a silent fallback): the GAN / perceptual training path
(vqgan_vae.py:481-548) and VectorQuantize's training-mode EMA / k-means updates.
"""
import copy
import math
from functools import wraps

import torch
import torch.nn.functional as F
from torch import nn

from . import ops


def _eval_decorator(fn):
    @wraps(fn)
    def inner(model, *args, **kwargs):
        was_training = model.training
        model.eval()
        out = fn(model, *args, **kwargs)
        model.train(was_training)
        return out
    return inner


def _split_prefixed(prefix, d):
    with_p = {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}
    without = {k: v for k, v in d.items() if not k.startswith(prefix)}
    return with_p, without


# ------------------------------------------------------------------------------------------------
# parameter containers (names == reference attribute names)
# ------------------------------------------------------------------------------------------------
class LayerNormChan(nn.Module):
    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.g = nn.Parameter(torch.ones(1, dim, 1, 1))
        self.b = nn.Parameter(torch.zeros(1, dim, 1, 1))


class ContinuousPositionBias(nn.Module):
    def __init__(self, *, dim, heads, layers=2):
        super().__init__()
        self.net = nn.ModuleList([nn.Sequential(nn.Linear(2, dim), nn.LeakyReLU(0.1))])
        for _ in range(layers - 1):
            self.net.append(nn.Sequential(nn.Linear(dim, dim), nn.LeakyReLU(0.1)))
        self.net.append(nn.Linear(dim, heads))

    @torch.no_grad()
    def bias_table(self, fmap, device):
        """(heads, n, n) fp32.  Input independent (vqgan_vae.py:192-210) => constant folded once per weight
        set at pack time instead of being recomputed by every attention call."""
        pos = torch.arange(fmap, device=device)
        gy, gx = torch.meshgrid(pos, pos, indexing='ij')
        grid = torch.stack([gy, gx]).reshape(2, -1).t()
        rel = grid[:, None, :] - grid[None, :, :]
        h = (torch.sign(rel) * torch.log(rel.abs() + 1)).float()
        for layer in self.net[:-1]:
            h = F.leaky_relu(F.linear(h, layer[0].weight.float(), layer[0].bias.float()), 0.1)
        h = F.linear(h, self.net[-1].weight.float(), self.net[-1].bias.float())
        return h.permute(2, 0, 1).contiguous()


class VQGanAttention(nn.Module):
    def __init__(self, *, dim, dim_head=64, heads=8, dropout=0.):
        super().__init__()
        self.heads, self.dim_head = heads, dim_head
        self.scale = nn.Parameter(torch.ones(1, heads, 1, 1) * math.log(0.01))
        inner = heads * dim_head
        self.dropout = nn.Dropout(dropout)
        self.post_norm = LayerNormChan(dim)
        self.cpb = ContinuousPositionBias(dim=dim // 4, heads=heads)
        self.to_qkv = nn.Conv2d(dim, inner * 3, 1, bias=False)
        self.to_out = nn.Conv2d(inner, dim, 1)


class GLUResBlock(nn.Module):
    def __init__(self, chan, groups=16):
        super().__init__()
        self.groups = groups
        self.net = nn.Sequential(nn.Conv2d(chan, chan * 2, 3, padding=1), nn.GLU(dim=1), nn.GroupNorm(groups, chan),
                                 nn.Conv2d(chan, chan * 2, 3, padding=1), nn.GLU(dim=1), nn.GroupNorm(groups, chan),
                                 nn.Conv2d(chan, chan, 1))


class ResBlock(nn.Module):
    def __init__(self, chan, groups=16):
        super().__init__()
        self.groups = groups
        self.net = nn.Sequential(nn.Conv2d(chan, chan, 3, padding=1), nn.GroupNorm(groups, chan), nn.LeakyReLU(0.1),
                                 nn.Conv2d(chan, chan, 3, padding=1), nn.GroupNorm(groups, chan), nn.LeakyReLU(0.1),
                                 nn.Conv2d(chan, chan, 1))


class _Codebook(nn.Module):
    """Buffers of the third-party codebook (cosine: no embed_avg) -- SURVEY.md §8(c)."""

    def __init__(self, dim, codebook_size, kmeans_init, cosine):
        super().__init__()
        if kmeans_init:
            embed = torch.zeros(codebook_size, dim)
        else:
            embed = F.normalize(torch.randn(codebook_size, dim), dim=-1) if cosine else torch.randn(codebook_size, dim)
        self.register_buffer('initted', torch.Tensor([not kmeans_init]))
        self.register_buffer('cluster_size', torch.zeros(codebook_size))
        self.register_buffer('embed', embed)
        if not cosine:
            self.register_buffer('embed_avg', embed.clone())


class VectorQuantize(nn.Module):
    """Parameter container for vector_quantize_pytorch.VectorQuantize (vqgan_vae.py:368-378)."""

    def __init__(self, dim, codebook_size, codebook_dim, use_cosine_sim, kmeans_init, **unused):
        super().__init__()
        proj = codebook_dim != dim
        self.project_in = nn.Linear(dim, codebook_dim) if proj else nn.Identity()
        self.project_out = nn.Linear(codebook_dim, dim) if proj else nn.Identity()
        self.use_cosine_sim = use_cosine_sim
        self.codebook_size = codebook_size
        self._codebook = _Codebook(codebook_dim, codebook_size, kmeans_init, use_cosine_sim)

    @property
    def codebook(self):
        return self._codebook.embed


# ------------------------------------------------------------------------------------------------
# VQGanVAE
# ------------------------------------------------------------------------------------------------
class VQGanVAE(nn.Module):
    def __init__(self, *, dim, image_size, channels=3, num_layers=4, layer_mults=None, l2_recon_loss=False,
                 use_hinge_loss=True, num_resnet_blocks=1, vgg=None, vq_codebook_dim=256, vq_codebook_size=512,
                 vq_decay=0.8, vq_commitment_weight=1., vq_kmeans_init=True, vq_use_cosine_sim=True, use_attn=True,
                 attn_dim_head=64, attn_heads=8, resnet_groups=16, attn_dropout=0., first_conv_kernel_size=5,
                 use_vgg_and_gan=True, **kwargs):
        super().__init__()
        assert dim % resnet_groups == 0, f'dimension {dim} must be divisible by {resnet_groups} (groups for the groupnorm)'
        vq_kwargs, kwargs = _split_prefixed('vq_', kwargs)  # unknown non-vq_ kwargs are swallowed (vqgan_vae.py:314)

        self.image_size = image_size
        self.channels = channels
        self.num_layers = num_layers
        self.fmap_size = image_size // (num_layers ** 2)  # sic (reference formula)
        self.codebook_size = vq_codebook_size
        self.first_conv_kernel_size = first_conv_kernel_size
        self.resnet_groups = resnet_groups

        self.encoders = nn.ModuleList([])
        self.decoders = nn.ModuleList([])
        layer_mults = layer_mults if layer_mults is not None else [2 ** t for t in range(num_layers)]
        assert len(layer_mults) == num_layers, 'layer multipliers must be equal to designated number of layers'
        layer_dims = [dim * m for m in layer_mults]
        dims = (dim, *layer_dims)
        if not isinstance(num_resnet_blocks, tuple):
            num_resnet_blocks = (*((0,) * (num_layers - 1)), num_resnet_blocks)
        if not isinstance(use_attn, tuple):
            use_attn = (*((False,) * (num_layers - 1)), use_attn)
        assert len(num_resnet_blocks) == num_layers, 'number of resnet blocks config must be equal to number of layers'
        assert len(use_attn) == num_layers

        for (d_in, d_out), n_res, attn in zip(zip(dims[:-1], dims[1:]), num_resnet_blocks, use_attn):
            self.encoders.append(nn.Sequential(nn.Conv2d(d_in, d_out, 4, stride=2, padding=1), nn.LeakyReLU(0.1)))
            self.decoders.insert(0, nn.Sequential(nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False),
                                                  nn.Conv2d(d_out, d_in, 3, padding=1), nn.LeakyReLU(0.1)))
            if attn:
                self.decoders.insert(0, VQGanAttention(dim=d_out, heads=attn_heads, dim_head=attn_dim_head, dropout=attn_dropout))
            for _ in range(n_res):
                self.encoders.append(ResBlock(d_out, groups=resnet_groups))
                self.decoders.insert(0, GLUResBlock(d_out, groups=resnet_groups))
            if attn:
                self.encoders.append(VQGanAttention(dim=d_out, heads=attn_heads, dim_head=attn_dim_head, dropout=attn_dropout))
        self.encoders.insert(0, nn.Conv2d(channels, dim, first_conv_kernel_size, padding=first_conv_kernel_size // 2))
        self.decoders.append(nn.Conv2d(dim, channels, 1))

        self.vq = VectorQuantize(dim=layer_dims[-1], codebook_dim=vq_codebook_dim, codebook_size=vq_codebook_size,
                                 use_cosine_sim=vq_use_cosine_sim, kmeans_init=vq_kmeans_init, **vq_kwargs)
        self.l2_recon_loss = l2_recon_loss
        self.vgg = None
        self.discr = None
        self.use_vgg_and_gan = use_vgg_and_gan
        if use_vgg_and_gan:
            raise NotImplementedError(
                'use_vgg_and_gan=True selects the GAN / perceptual training path (vqgan_vae.py:393-406, needs VGG16 '
                'weights from the network); it is outside the B200 hot path -- construct with use_vgg_and_gan=False')
        self._packed = None
        self._packed_sig = None

    # ---------------------------------------------------------------------------- reference surface
    def copy_for_eval(self):
        device = next(self.parameters()).device
        vae_copy = copy.deepcopy(self.cpu())  # NB: moves the caller's module to CPU, like the reference (D14)
        vae_copy.eval()
        return vae_copy.to(device)

    @property
    def codebook(self):
        return self.vq.codebook

    # ---------------------------------------------------------------------------- weight packing
    def _signature(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    @torch.no_grad()
    def _pack(self):
        sig = self._signature()
        if self._packed is not None and self._packed_sig == sig:
            return self._packed
        dev = next(self.parameters()).device
        if dev.type != 'cuda':
            raise ops._lib.NuwaB200Error('VQGanVAE must live on a CUDA device (no CPU path)')
        pk = {}

        def conv_pack(m, pairs=False):
            b = m.bias.detach().float()
            if pairs:
                b = ops.pack_pairs(b)
            return dict(w=ops.pack_conv_weight(m.weight.detach().float(), pairs=pairs), b=b.contiguous(),
                        cin=m.weight.shape[1])

        def lin_pack(w, b=None):
            w2 = w.detach().float().reshape(w.shape[0], -1)
            return dict(w=w2.to(torch.bfloat16).contiguous(), b=None if b is None else b.detach().float().contiguous())

        def attn_pack(m, fmap):
            return dict(qkv=lin_pack(m.to_qkv.weight), out=lin_pack(m.to_out.weight, m.to_out.bias),
                        bias=m.cpb.bias_table(fmap, dev), hscale=m.scale.detach().float().exp().reshape(-1).contiguous(),
                        g=m.post_norm.g.detach().float().reshape(-1).contiguous(),
                        b=m.post_norm.b.detach().float().reshape(-1).contiguous())

        fmap = self.image_size // (2 ** self.num_layers)
        enc = []
        for i, m in enumerate(self.encoders):
            if i == 0:
                enc.append(dict(kind='conv_in', w=ops.pack_im2col_weight(m.weight.detach().float()),
                                b=m.bias.detach().float().contiguous()))
            elif isinstance(m, nn.Sequential):
                enc.append(dict(kind='down', **conv_pack(m[0])))
            elif isinstance(m, ResBlock):
                n = m.net
                enc.append(dict(kind='res', c1=conv_pack(n[0]), g1=(n[1].weight.detach().float(), n[1].bias.detach().float()),
                                c2=conv_pack(n[3]), g2=(n[4].weight.detach().float(), n[4].bias.detach().float()),
                                c3=lin_pack(n[6].weight, n[6].bias), groups=m.groups))
            else:
                enc.append(dict(kind='attn', heads=m.heads, dh=m.dim_head, **attn_pack(m, fmap)))
        dec = []
        for i, m in enumerate(self.decoders):
            if isinstance(m, GLUResBlock):
                n = m.net
                dec.append(dict(kind='glures', c1=conv_pack(n[0], True), g1=(n[2].weight.detach().float(), n[2].bias.detach().float()),
                                c2=conv_pack(n[3], True), g2=(n[5].weight.detach().float(), n[5].bias.detach().float()),
                                c3=lin_pack(n[6].weight, n[6].bias), groups=m.groups))
            elif isinstance(m, VQGanAttention):
                dec.append(dict(kind='attn', heads=m.heads, dh=m.dim_head, **attn_pack(m, fmap)))
            elif isinstance(m, nn.Sequential):
                dec.append(dict(kind='up', **conv_pack(m[1])))
            else:
                dec.append(dict(kind='conv_out', w=m.weight.detach().float().reshape(m.weight.shape[0], -1).contiguous(),
                                b=m.bias.detach().float().contiguous()))
        vq = dict(cosine=self.vq.use_cosine_sim, embed=self.vq._codebook.embed.detach().float().contiguous())
        if self.vq.use_cosine_sim:
            vq['code'] = F.normalize(vq['embed'], dim=-1).contiguous()
            vq['code_sq'] = None
        else:
            vq['code'] = vq['embed']
            vq['code_sq'] = vq['embed'].pow(2).sum(-1).contiguous()
        # tensor-core arg-max (ops.vq_argmax 'auto'): bf16 copy of the compared codebook + its largest row norm (pack time)
        vq['code16'] = vq['code'].to(torch.bfloat16).contiguous()
        vq['emax'] = vq['code'].norm(dim=-1).max().reshape(1).float().contiguous()  # device scalar: no host sync
        if isinstance(self.vq.project_in, nn.Linear):
            # project_in feeds the arg-max: fp32-faithful (three exact bf16 terms per fp32 weight, ops.linear_f32x3),
            # so the token ids are those of the fp32 reference for the same pre-VQ map
            pw = self.vq.project_in.weight.detach().float().contiguous()
            vq['pin'] = dict(w3=ops.split3(pw), b=self.vq.project_in.bias.detach().float().contiguous())
            po = self.vq.project_out.weight.detach().float().contiguous()
            vq['pout'] = dict(w3=ops.split3(po), b=self.vq.project_out.bias.detach().float().contiguous())
        pk['enc'], pk['dec'], pk['vq'] = enc, dec, vq
        self._packed, self._packed_sig = pk, sig
        return pk

    # ---------------------------------------------------------------------------- kernels (NHWC)
    @staticmethod
    def _attn(x16, x32, a):
        B, H, W, C = x16.shape
        n, heads, dh = H * W, a['heads'], a['dh']
        inner = heads * dh
        qkv32 = ops.gemm(x16.view(B * n, C), a['qkv']['w'], out_dtype=torch.float32)
        qkv = ops.vae_attn_prep(qkv32, B, n, inner)
        o = torch.empty(B * n, inner, dtype=torch.bfloat16, device=x16.device)
        base = qkv.data_ptr()
        ops.attn_dense(base, base + inner * 2, base + 2 * inner * 2, o, B=B, nq=n, nk=n, H=heads, dh=dh,
                       q_bs=n * 3 * inner, q_rs=3 * inner, k_bs=n * 3 * inner, k_rs=3 * inner, v_bs=n * 3 * inner,
                       v_rs=3 * inner, o_bs=n * inner, o_rs=inner, head_scale=a['hscale'], bias=a['bias'], qscale=1.0)
        y = ops.gemm(o, a['out']['w'], bias=a['out']['b'], out_dtype=torch.float32)
        out32 = torch.empty(B, H, W, C, dtype=torch.float32, device=x16.device)
        out16 = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=x16.device)
        # LayerNormChan + residual (vqgan_vae.py:286) == row LayerNorm over the channel axis of NHWC rows
        ops.sandwich_ln(B * n, 1, C, y=y, post=(a['g'], a['b']), res_in=x32, x_out=out32, x_out_bf16=out16)
        return out16, out32

    @staticmethod
    def _resblock(x16, x32, r, glu):
        B, H, W, C = x16.shape
        act = 'glu' if glu else None
        h = ops.conv2d_nhwc(x16, r['c1']['w'], Cin=C, ksize=3, bias=r['c1']['b'], act=act, out_dtype=torch.float32)
        h = ops.groupnorm_nhwc(h, r['g1'][0], r['g1'][1], r['groups'], leaky=not glu)
        h = ops.conv2d_nhwc(h, r['c2']['w'], Cin=C, ksize=3, bias=r['c2']['b'], act=act, out_dtype=torch.float32)
        h = ops.groupnorm_nhwc(h, r['g2'][0], r['g2'][1], r['groups'], leaky=not glu)
        out32, out16 = ops.gemm(h.view(B * H * W, C), r['c3']['w'], bias=r['c3']['b'], residual=x32.view(B * H * W, C),
                                out_dtype=torch.float32, also_bf16=True)
        return out16.view(B, H, W, C), out32.view(B, H, W, C)

    def _encode_fmap_nhwc(self, img):
        """encoders (vqgan_vae.py:431-433): NCHW fp32 image -> (bf16, fp32) NHWC feature map."""
        pk = self._pack()
        B, C, H, W = img.shape
        x16 = x32 = None
        kinds = [e['kind'] for e in pk['enc']] + ['vq']
        for i, e in enumerate(pk['enc']):
            k = e['kind']
            if k == 'conv_in':
                a = ops.im2col(img.contiguous().float(), self.first_conv_kernel_size, e['w'].shape[1])
                x16 = ops.gemm(a, e['w'], bias=e['b'], out_dtype=torch.bfloat16).view(B, H, W, -1)
            elif k == 'down':
                if kinds[i + 1] != 'down':  # feeds residual blocks / VQ: keep an fp32 copy as the residual stream
                    x32, x16 = ops.conv2d_nhwc(x16, e['w'], Cin=e['cin'], ksize=4, stride=2, bias=e['b'], act='leaky',
                                               out_dtype=torch.float32, also_bf16=True)
                else:
                    x16 = ops.conv2d_nhwc(x16, e['w'], Cin=e['cin'], ksize=4, stride=2, bias=e['b'], act='leaky')
            elif k == 'res':
                x16, x32 = self._resblock(x16, x32, e, glu=False)
            else:
                x16, x32 = self._attn(x16, x32, e)
        return x16, x32

    def _quantize_nhwc(self, x16, x32):
        """self.vq(fmap) in eval mode: returns (quantised NHWC (bf16, fp32), indices (B,h,w) int64)."""
        pk = self._pack()['vq']
        B, H, W, C = x16.shape
        M = B * H * W
        flat = ops.linear_f32x3(x32.view(M, C), pk['pin']['w3'], pk['pin']['b']) if 'pin' in pk else x32.view(M, C)
        ind = ops.vq_argmax(flat, pk['code'], pk['code_sq'], cosine=pk['cosine'], code16=pk['code16'], emax=pk['emax'])
        if 'pout' in pk:
            e32 = ops.gather_rows(pk['embed'], ind, want_bf16=False, want_f32=True)
            q32, q16 = ops.linear_f32x3(e32, pk['pout']['w3'], pk['pout']['b'], also_bf16=True)
        else:
            q16, q32 = ops.gather_rows(pk['embed'], ind, want_bf16=True, want_f32=True)
        return q16.view(B, H, W, -1), q32.view(B, H, W, -1), ind.view(B, H, W)

    def _decode_nhwc(self, x16, x32):
        """decoders (vqgan_vae.py:437-441): NHWC -> NCHW fp32 image."""
        pk = self._pack()
        for d in pk['dec']:
            k = d['kind']
            if k == 'glures':
                x16, x32 = self._resblock(x16, x32, d, glu=True)
            elif k == 'attn':
                x16, x32 = self._attn(x16, x32, d)
            elif k == 'up':
                x16 = ops.upsample2x(x16)
                x16 = ops.conv2d_nhwc(x16, d['w'], Cin=d['cin'], ksize=3, bias=d['b'], act='leaky')
            else:
                return ops.conv1x1_to_nchw(x16, d['w'], d['b'])
        raise AssertionError('decoder has no output convolution')

    def _check_mode(self):
        if self.training and torch.is_grad_enabled():
            raise NotImplementedError('VQGanVAE training (EMA codebook update, straight-through gradients) is outside the '
                                      'B200 hot path; call .eval() / torch.no_grad()')

    # ---------------------------------------------------------------------------- public API
    def encode(self, fmap):
        """(quantized NCHW fp32, indices (b,h,w) int64, commit loss (1,)) -- vqgan_vae.py:431-435."""
        self._check_mode()
        with torch.no_grad():
            x16, x32 = self._encode_fmap_nhwc(fmap)
            q16, q32, ind = self._quantize_nhwc(x16, x32)
            quant = ops.nhwc_to_nchw_f32(q32)
        return quant, ind, torch.zeros(1, device=fmap.device)

    def decode(self, fmap):
        """NCHW fp32 feature map -> NCHW fp32 image -- vqgan_vae.py:437-441."""
        self._check_mode()
        with torch.no_grad():
            x16 = ops.nchw_to_nhwc_bf16(fmap)
            x32 = fmap.permute(0, 2, 3, 1).contiguous().float()  # un-rounded residual stream (layout change only)
            return self._decode_nhwc(x16, x32)

    @torch.no_grad()
    @_eval_decorator
    def codebook_indices_to_video(self, indices):
        """vqgan_vae.py:443-450 (raw codebook vectors go straight into decode, D4)."""
        b = indices.shape[0]
        fm = self.fmap_size
        flat = indices.reshape(-1).contiguous()
        c16, c32 = ops.gather_rows(self._pack()['vq']['embed'], flat, want_bf16=True, want_f32=True)
        d = c16.shape[1]
        video = self._decode_nhwc(c16.view(-1, fm, fm, d), c32.view(-1, fm, fm, d))
        return video.view(b, -1, *video.shape[1:])

    @torch.no_grad()
    @_eval_decorator
    def get_video_indices(self, video):
        """vqgan_vae.py:452-458: (b,f,c,h,w) float video -> (b,f,h',w') int64 token ids."""
        b, f = video.shape[:2]
        images = video.reshape(b * f, *video.shape[2:])
        x16, x32 = self._encode_fmap_nhwc(images)
        pk = self._pack()['vq']
        B, H, W, C = x16.shape
        flat = ops.linear_f32x3(x32.view(-1, C), pk['pin']['w3'], pk['pin']['b']) if 'pin' in pk else x32.view(-1, C)
        ind = ops.vq_argmax(flat, pk['code'], pk['code_sq'], cosine=pk['cosine'], code16=pk['code16'], emax=pk['emax'])  # project_out is not needed here
        return ind.view(b, f, H, W)

    def forward(self, img, return_loss=False, return_discr_loss=False, return_recons=False, apply_grad_penalty=False):
        batch, channels, height, width = img.shape
        assert height == self.image_size and width == self.image_size, 'height and width of input image must be equal to {self.image_size}'
        assert channels == self.channels, 'number of channels on image or sketch is not equal to the channels set on this VQGanVAE'
        self._check_mode()
        with torch.no_grad():
            x16, x32 = self._encode_fmap_nhwc(img)
            q16, q32, _ = self._quantize_nhwc(x16, x32)
            fmap = self._decode_nhwc(q16, q32)
        if not return_loss and not return_discr_loss:
            return fmap
        assert return_loss ^ return_discr_loss, 'you should either return autoencoder loss or discriminator loss, but not both'
        if return_discr_loss:
            assert self.discr is not None, 'discriminator must exist to train it'
        # reconstruction loss; without VGG / GAN the reference returns it as is (vqgan_vae.py:502-512) -- forward value
        # only here (the VAE backward / GAN path is outside the B200 hot path, _check_mode() rejects grad mode)
        recon_loss = ops.recon_loss(fmap, img.float(), l2=self.l2_recon_loss)
        if return_recons:
            return recon_loss, fmap
        return recon_loss
