"""A few eager decode steps of generate() with cfg-4 layer shapes (dim 512, 8 heads, batch 8, 256-token text context,
reversible decoder) but only 4 decoder layers and a 2x2 token grid, so that an ncu launch list of the per-kernel
durations of ONE decode step is cheap to obtain."""
import sys

import torch

sys.path.insert(0, '.')
from nuwa_pytorch_b200 import NUWA, VQGanVAE  # noqa: E402

dev = torch.device('cuda')
torch.manual_seed(0)
with torch.device(dev):
    vae = VQGanVAE(dim=64, image_size=32, num_layers=4, vq_codebook_size=8192, vq_codebook_dim=512, use_vgg_and_gan=False,
                   vq_kmeans_init=False)
    nuwa = NUWA(vae=vae, dim=512, dec_depth=4, dec_heads=8, dec_reversible=True, enc_reversible=True, max_video_frames=10,
                sparse_3dna_kernel_size=(5, 3, 3), sparse_3dna_dilation=(1, 2, 4)).eval()
text = torch.randint(1, 49408, (8, 256), device=dev)
idx = nuwa.generate(text=text, num_frames=2, _return_indices=True, _use_graph=False)
torch.cuda.synchronize()
print(idx.shape)
