"""Where does a decode step of the persistent kernel (csrc/decode_stack.cu) spend its time?
(1) clock64 stamps of one CTA at every phase boundary, averaged per (sub-block kind, interval);
(2) whole-kernel CUDA-event times with parts of the work switched off (debug_flags), which isolates the cost of the
    device-wide barriers from the work between them.
Usage: python tools/decode_phase_profile.py [depth=16] [out.json]"""
import collections
import json
import os
import sys

import torch

sys.path.insert(0, '.')
from nuwa_pytorch_b200 import NUWA, VQGanVAE, engine  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 16
out = sys.argv[2] if len(sys.argv) > 2 else None
dev = torch.device('cuda')
torch.manual_seed(0)
with torch.device(dev):
    vae = VQGanVAE(dim=64, image_size=256, num_layers=4, vq_codebook_size=8192, vq_codebook_dim=512, use_vgg_and_gan=False,
                   vq_kmeans_init=False)
    nuwa = NUWA(vae=vae, dim=512, dec_depth=depth, dec_heads=8, dec_reversible=True, enc_reversible=True,
                max_video_frames=10, sparse_3dna_kernel_size=(5, 3, 3), sparse_3dna_dilation=(1, 2, 4)).eval()
B = 8
text = torch.randint(1, 49408, (B, 256), device=dev)
res = dict(depth=depth, batch=B)
with torch.no_grad():
    context = nuwa._text_context(text, text != 0)
    pack = engine.pack_stack(nuwa.video_transformer)
    engine.prime_context(nuwa.video_transformer, context)
    t_dev = torch.full((1,), 700, dtype=torch.int32, device=dev)
    state = engine.DecodeState(pack, B, 1280, dev, t_dev)
    for i in state.qkv:
        state.qkv[i].normal_(0, 0.5)
    x = torch.randn(B, 1, 512, device=dev)

    def timed(plan, n=20):
        for _ in range(3):
            plan.run(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            plan.run(x)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    nb = sum({'3dna': 3, 'cross': 4, 'ff': 2}[s.kind] for s in pack.subs)
    res['barriers_per_sweep'] = nb
    for flags, label in ((0, 'full'), (1, 'no_norms'), (2, 'no_products'), (4, 'no_attention'), (6, 'norms_only'),
                         (7, 'barriers_staging_prefetch_only')):
        os.environ['NUWA_DECODE_DEBUG'] = str(flags)
        plan = engine.FusedDecode(pack, state, context, nuwa._logits_weight())
        ms = timed(plan)
        res[label] = dict(ms_per_sweep=round(ms, 4), us_per_barrier=round(1e3 * ms / nb, 3))
        print(label, res[label], flush=True)
    os.environ['NUWA_DECODE_DEBUG'] = '0'
    for ctas in (148, 96, 64):
        os.environ['NUWA_DECODE_MAX_CTAS'] = str(ctas)
        os.environ['NUWA_DECODE_DEBUG'] = '7'
        plan = engine.FusedDecode(pack, state, context, nuwa._logits_weight())
        ms = timed(plan)
        res[f'barriers_only_ctas_{ctas}'] = dict(ms_per_sweep=round(ms, 4), us_per_barrier=round(1e3 * ms / nb, 3))
        print(f'barriers only, {ctas} CTAs', res[f'barriers_only_ctas_{ctas}'], flush=True)
    os.environ['NUWA_DECODE_MAX_CTAS'] = '0'
    os.environ['NUWA_DECODE_DEBUG'] = '0'

    # ---- per-phase stamps of CTA 0 (does attention work) and of the last CTA (only products) ----
    for cta in (0, 147):
        plan = engine.FusedDecode(pack, state, context, nuwa._logits_weight())
        prof = torch.zeros(2 * (24 * len(pack.subs) + 8), dtype=torch.int64, device=dev)
        plan.params.prof, plan.params.prof_cta = prof.data_ptr(), cta
        ms = timed(plan, 5)
        v = prof.cpu().tolist()
        stamps = []
        for i in range(0, len(v), 2):
            if v[i] < 0 or (i > 0 and v[i] == 0 and v[i + 1] == 0):
                break
            stamps.append((v[i] // 32, v[i] % 32, v[i + 1]))
        cyc_total = stamps[-1][2] - stamps[0][2]
        ghz = cyc_total / (ms * 1e6)
        agg = collections.defaultdict(list)
        for (s0, i0, c0), (s1, i1, c1) in zip(stamps, stamps[1:]):
            kind = pack.subs[s1].kind if s1 < len(pack.subs) else 'tail'
            agg[f'{kind}:{i0}->{i1}'].append(c1 - c0)
        table = {k: dict(n=len(c), mean_us=round(sum(c) / len(c) / ghz / 1e3, 3), max_us=round(max(c) / ghz / 1e3, 3))
                 for k, c in sorted(agg.items())}
        res[f'stamps_cta_{cta}'] = dict(kernel_ms=round(ms, 4), sm_ghz_estimate=round(ghz, 3), intervals=table)
        print(f'--- CTA {cta}: kernel {ms:.3f} ms, ~{ghz:.2f} GHz')
        for k, r in table.items():
            print(f'  {k:16s} n={r["n"]:4d} mean {r["mean_us"]:7.3f} us  max {r["max_us"]:7.3f} us')
if out:
    with open(out, 'w') as f:
        json.dump(res, f, indent=1)
