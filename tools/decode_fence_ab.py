"""Decode-stack sweep time (depth 16, batch 8) with all work (flags 0) and with barriers + staging only (flags 7): the A/B
harness used for the grid-barrier and proxy-fence variants (results: comments in csrc/decode_stack.cu)."""
import os, sys, torch
sys.path.insert(0, '.')
from nuwa_pytorch_b200 import NUWA, VQGanVAE, engine
dev = torch.device('cuda'); torch.manual_seed(0)
with torch.device(dev):
    vae = VQGanVAE(dim=64, image_size=256, num_layers=4, vq_codebook_size=8192, vq_codebook_dim=512, use_vgg_and_gan=False, vq_kmeans_init=False)
    nuwa = NUWA(vae=vae, dim=512, dec_depth=16, dec_heads=8, dec_reversible=True, enc_reversible=True, max_video_frames=10,
                sparse_3dna_kernel_size=(5, 3, 3), sparse_3dna_dilation=(1, 2, 4)).eval()
B = 8
text = torch.randint(1, 49408, (B, 256), device=dev)
with torch.no_grad():
    context = nuwa._text_context(text, text != 0)
    pack = engine.pack_stack(nuwa.video_transformer)
    engine.prime_context(nuwa.video_transformer, context)
    t_dev = torch.full((1,), 700, dtype=torch.int32, device=dev)
    state = engine.DecodeState(pack, B, 1280, dev, t_dev)
    for i in state.qkv: state.qkv[i].normal_(0, 0.5)
    x = torch.randn(B, 1, 512, device=dev)
    for flags in (0, 7, 0, 7):
        os.environ['NUWA_DECODE_DEBUG'] = str(flags)
        plan = engine.FusedDecode(pack, state, context, nuwa._logits_weight())
        for _ in range(5): plan.run(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50): plan.run(x)
        b.record(); torch.cuda.synchronize()
        print('flags', flags, 'ms per sweep', a.elapsed_time(b) / 50, flush=True)
