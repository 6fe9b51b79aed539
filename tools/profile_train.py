"""Training step of BASELINE configs[2] (NUWA dim 512, depth 12, batch 8: forward loss + backward): CUDA-event timing
of the forward and backward halves, peak memory, and (under ncu) the per-kernel launch list.
    python tools/profile_train.py [reps]"""
import json
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
from nuwa_pytorch_b200 import _lib  # noqa: E402

dev = torch.device('cuda')
nuwa = bench.build_decoder(dev).train()
g = torch.Generator(device=dev).manual_seed(100)
text = torch.randint(1, 49408, (bench.DEC_BATCH, 256), device=dev, generator=g)
video = torch.randint(0, 8192, (bench.DEC_BATCH, 10, 16, 16), device=dev, generator=g)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rows = []
for it in range(reps + 2):
    for p in nuwa.parameters():
        p.grad = None
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    loss = nuwa(text=text, video=video, return_loss=True)
    e[1].record()
    loss.backward()
    e[2].record()
    torch.cuda.synchronize()
    if it >= 2:
        rows.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), _lib.launch_count() - l0))
fwd = sorted(r[0] for r in rows)[len(rows) // 2]
bwd = sorted(r[1] for r in rows)[len(rows) // 2]
tokens = bench.DEC_BATCH * 2560
out = dict(loss=float(loss), fwd_ms=round(fwd, 3), bwd_ms=round(bwd, 3), step_ms=round(fwd + bwd, 3),
           tokens_per_s=round(tokens / ((fwd + bwd) * 1e-3), 1), launches_per_step=rows[-1][2],
           peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 2),
           grad_norm=float(torch.sqrt(sum((p.grad.float() ** 2).sum() for p in nuwa.parameters() if p.grad is not None))))
print(json.dumps(out))
json.dump(out, open('gpurun_out/train_profile.json', 'w'))
