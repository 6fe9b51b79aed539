"""Decoder-only workload (BASELINE configs[2]: NUWA dim 512, depth 12, batch 8, forward loss) for ncu launch lists
and `--set full` captures of the attention kernels.  Eager launches (no CUDA graph) so every kernel is visible."""
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402

dev = torch.device('cuda')
nuwa = bench.build_decoder(dev)
g = torch.Generator(device=dev).manual_seed(100)
text = torch.randint(1, 49408, (bench.DEC_BATCH, 256), device=dev, generator=g)
video = torch.randint(0, 8192, (bench.DEC_BATCH, 10, 16, 16), device=dev, generator=g)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
with torch.no_grad():
    for _ in range(reps):
        loss = nuwa(text=text, video=video, return_loss=True)
torch.cuda.synchronize()
print(float(loss))
