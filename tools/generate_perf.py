"""generate() decode-step timing at the BASELINE configs[3] geometry: persistent decode kernel (csrc/decode_stack.cu)
vs the per-kernel decode path, both replayed from one CUDA graph per token, plus the kernel's tuning knobs.
Usage: python tools/generate_perf.py [depth=64] [frames=1] [out.json]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, '.')
from nuwa_pytorch_b200 import NUWA, VQGanVAE  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 64
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1
out = sys.argv[3] if len(sys.argv) > 3 else None
dev = torch.device('cuda')
torch.manual_seed(0)
with torch.device(dev):
    vae = VQGanVAE(dim=64, image_size=256, num_layers=4, vq_codebook_size=8192, vq_codebook_dim=512, use_vgg_and_gan=False,
                   vq_kmeans_init=False)
    nuwa = NUWA(vae=vae, dim=512, dec_depth=depth, dec_heads=8, dec_reversible=True, enc_reversible=True,
                max_video_frames=10, sparse_3dna_kernel_size=(5, 3, 3), sparse_3dna_dilation=(1, 2, 4)).eval()
B = 8
text = torch.randint(1, 49408, (B, 256), device=dev)
with torch.no_grad():
    nuwa.generate(text=text, num_frames=1, _return_indices=True)  # packs the weights
from nuwa_pytorch_b200 import engine  # noqa: E402
wbytes = nuwa._logits_weight().numel() * 2
for s_ in engine.pack_stack(nuwa.video_transformer).subs:
    for name in ('w_qkv', 'w_q', 'w_out', 'w1', 'w2'):
        if hasattr(s_, name):
            wbytes += getattr(s_, name).numel() * 2
res = dict(depth=depth, frames=frames, batch=B, weight_bytes_per_sweep=wbytes)


def run(label, **kw):
    nuwa.generate(text=text, num_frames=1, _return_indices=True, **kw)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    idx = nuwa.generate(text=text, num_frames=frames, _return_indices=True, **kw)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / (frames * 256)
    res[label] = dict(ms_per_token_step=round(ms, 4), tokens_per_s=round(B * 1e3 / ms, 1),
                      wall_s=round(time.perf_counter() - t0, 2))
    print(label, res[label], flush=True)
    return idx


run('fused_graph')
run('per_kernel_graph', _use_fused=False)
for ss, sf in ((1, 1), (1, 7), (7, 1)):
    os.environ['NUWA_DECODE_SPLIT_SMALL'], os.environ['NUWA_DECODE_SPLIT_FF'] = str(ss), str(sf)
    run(f'fused_split_{ss}_{sf}')
os.environ['NUWA_DECODE_SPLIT_SMALL'] = os.environ['NUWA_DECODE_SPLIT_FF'] = '0'
for ctas in (96, 128):
    os.environ['NUWA_DECODE_MAX_CTAS'] = str(ctas)
    run(f'fused_ctas_{ctas}')
os.environ['NUWA_DECODE_MAX_CTAS'] = '0'
os.environ['NUWA_DECODE_COOP'] = '0'
run('fused_plain_launch')
if out:
    with open(out, 'w') as f:
        json.dump(res, f, indent=1)
