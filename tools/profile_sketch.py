"""NUWASketch training step of BASELINE configs[4] (batch 4) for an ncu launch list: eager launches, 1 warm-up + 1 step."""
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
from nuwa_pytorch_b200 import NUWASketch, VQGanVAE  # noqa: E402

dev = torch.device('cuda')
torch.manual_seed(0)
with torch.device(dev):
    svae = VQGanVAE(**{**bench.DEC_VAE_KW, "channels": 5})
    vvae = VQGanVAE(**bench.DEC_VAE_KW)
    sk = NUWASketch(vae=vvae, sketch_vae=svae, dim=512, image_size=256, sketch_enc_depth=12, sketch_max_video_frames=3,
                    sketch_enc_use_sparse_3dna=True, max_video_frames=10, dec_depth=24, sparse_3dna_kernel_size=(5, 3, 3),
                    sparse_3dna_dilation=(1, 2, 4)).train()
SB = 4
g = torch.Generator(device=dev).manual_seed(300)
sketch = torch.randn(SB, 3, 5, 256, 256, device=dev, generator=g)
video = torch.randn(SB, 10, 3, 256, 256, device=dev, generator=g)
mask = torch.ones(SB, 3, dtype=torch.bool, device=dev)
for it in range(2):
    for p in sk.parameters():
        p.grad = None
    torch.cuda.synchronize()
    if it == 1:
        torch.cuda.profiler.start()
    loss = sk(sketch=sketch, sketch_mask=mask.clone(), video=video, return_loss=True)
    loss.backward()
    torch.cuda.synchronize()
    if it == 1:
        torch.cuda.profiler.stop()
print(float(loss))
