#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native NUWA hot paths.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run, one rank/GPU)
  python bench.py --impl reference ...                    (the reference algorithm's CPU path, oracle port)

Metric (BASELINE.json): "3DNA decoder video-tokens/sec + VQGanVAE frames/sec @256^2".  The JSON line's
`value` is VQGanVAE frames/sec on BASELINE configs[1] (dim=512, 256^2, L=4, 2 res blocks, codebook 8192,
batch 64 per GPU: encode + VQ + decode of every frame); the 3DNA-decoder video-tokens/sec of configs[2]
(NUWA dim 512, depth 12, 10 frames, kernel (5,3,3), dilation (1,2,4), forward loss, batch 8) rides along under
`"decoder"`.  One "step" = one pass of the hot path over one synthetic batch.  Weights are random-init of the
named architecture, inputs synthetic (there is no network for datasets / checkpoints).

Multi-GPU: the path shards by batch with no data-path collective (inference replicas); every rank
processes its own batch (weak scaling); timing = max over ranks between barriers.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VAE_KW = dict(dim=512, image_size=256, num_layers=4, num_resnet_blocks=2, vq_codebook_size=8192,
              use_vgg_and_gan=False, vq_kmeans_init=False)
VAE_BATCH = 64
VAE_GFLOP_PER_FRAME = 2095.0  # encode 679.5 + VQ 2.15 + decode 1413.4 (2*MAC)
DEC_VAE_KW = dict(dim=64, image_size=256, num_layers=4, vq_codebook_size=8192, vq_codebook_dim=512,
                  use_vgg_and_gan=False, vq_kmeans_init=False)
DEC_KW = dict(dim=512, dec_depth=12, dec_heads=8, max_video_frames=10, sparse_3dna_kernel_size=(5, 3, 3),
              sparse_3dna_dilation=(1, 2, 4), enc_reversible=True)
DEC_BATCH = 8
METRIC = "VQGanVAE frames/sec @256^2 (3DNA decoder video-tokens/sec under 'decoder')"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d["bf16_tflops_sustained"],
                    hbm_gbs=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["unavailable"])
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=float(self.samples[0][1]), reasons=reasons,
                    power_w_max=max(float(s[2]) for s in self.samples), samples=len(self.samples))


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_vae_frames_per_s(sd, frames, reps=1):
    """CPU fp32 reference-algorithm VAE recon (oracle port) on `frames` frames of the cfg-2 architecture."""
    from oracle import nuwa_oracle as O
    spec = O.VAESpec(512, 256, num_layers=4, num_resnet_blocks=2, codebook_size=8192)
    img = torch.randn(frames, 3, 256, 256, generator=torch.Generator().manual_seed(0))
    best = None
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            O.vae_forward(img, sd, spec)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return frames / best, best


def pick_cpu_threads(sd):
    """Give the CPU arm its best shot: PyTorch-eager conv stacks do not always scale to every core of a large host,
    so try all / half / a quarter of the cores on one frame (untimed calibration) and keep the fastest."""
    cores = os.cpu_count() or 1
    cands = sorted({cores, max(1, cores // 2), max(1, cores // 4)}, reverse=True)
    best_t, best_n = None, cores
    for n in cands:
        torch.set_num_threads(n)
        _, dt = cpu_vae_frames_per_s(sd, 1)
        if best_t is None or dt < best_t:
            best_t, best_n = dt, n
    torch.set_num_threads(best_n)
    return best_n


def cpu_state_dict_vae(seed=0):
    """Random-init weights of the cfg-2 VAE, generated on the CPU shape-by-shape (oracle/synth.py policy)."""
    from nuwa_pytorch_b200.vqgan_vae import VQGanVAE
    from oracle.synth import synth_state_dict
    with torch.device("meta"):
        vae = VQGanVAE(**VAE_KW)
    man = [(k, tuple(v.shape)) for k, v in vae.state_dict().items() if v.dtype == torch.float32]
    return synth_state_dict(man, seed)


def cpu_state_dict_nuwa(seed=0, **over):
    """Random-init weights of the cfg-3 / cfg-4 NUWA (decoder side only matters), CPU, oracle/synth.py policy."""
    from nuwa_pytorch_b200 import NUWA, VQGanVAE
    from oracle.synth import manifest_of, synth_state_dict
    m = NUWA(vae=VQGanVAE(**DEC_VAE_KW), **{**DEC_KW, **over})  # CPU container (the constructor deep-copies its VAE)
    man = [(k, shape) for k, shape in manifest_of(m.state_dict()) if not k.startswith("vae.")]
    del m
    return synth_state_dict(man, seed)


def nuwa_oracle_spec(**over):
    from oracle import nuwa_oracle as O
    kw = {**DEC_KW, **over}
    return O.NUWASpec(kw["dim"], 16, kw["max_video_frames"], 8192, dec_depth=kw["dec_depth"], dec_heads=kw["dec_heads"],
                      dec_reversible=kw.get("dec_reversible", False), enc_reversible=True,
                      kernel=kw["sparse_3dna_kernel_size"], dilation=kw["sparse_3dna_dilation"])


def cpu_decoder_train_tokens_per_s():
    """Reference algorithm (oracle port, PyTorch CPU fp32 autograd) on BASELINE configs[2] at B=1 (same per-sample shape:
    256 text tokens, 2560 video tokens, 12 layers): loss + loss.backward(), one pass.  SURVEY 8(d): the full B=8 step
    needs ~80 GB of RSS and minutes, so the per-unit rate at B=1 is the reported baseline."""
    from oracle import nuwa_oracle as O
    sd = cpu_state_dict_nuwa()
    for v in sd.values():
        v.requires_grad_(True)
    spec = nuwa_oracle_spec()
    g = torch.Generator().manual_seed(100)
    text = torch.randint(1, 49408, (1, 256), generator=g)
    video = torch.randint(0, 8192, (1, 2560), generator=g)
    t0 = time.perf_counter()
    _, loss = O.nuwa_logits(text, video, sd, spec)
    t_f = time.perf_counter() - t0
    loss.backward()
    dt = time.perf_counter() - t0
    return 2560 / dt, dt, 2560 / t_f, t_f


def cpu_generate_tokens_per_s(batch=1, prefixes=(1, 129, 257)):
    """Reference algorithm of generate() on BASELINE configs[3] (depth-64 reversible decoder): the reference recomputes
    the whole prefix every step and runs two guidance sweeps (nuwa_pytorch.py:1870-1908).  SURVEY 8(d) method: time single
    loop iterations at a few prefix lengths, fit step(t) = a + b t (every layer is window / 257-key attention + per-token
    GEMMs, so the cost is linear in the prefix length) and integrate over t = 1..1280."""
    from oracle import nuwa_oracle as O
    sd = cpu_state_dict_nuwa(dec_depth=64, dec_reversible=True)
    spec = nuwa_oracle_spec(dec_depth=64, dec_reversible=True)
    g = torch.Generator().manual_seed(200)
    text = torch.randint(1, 49408, (batch, 256), generator=g)
    seq = torch.randint(0, 8192, (batch, max(prefixes)), generator=g)
    pts = []
    with torch.no_grad():
        temb, tmask = O.nuwa_embed_text(text, sd, spec)
        for t in prefixes:
            t0 = time.perf_counter()
            O.nuwa_generate_step_logits(temb, tmask, seq[:, :t - 1], sd, spec, 2.)
            pts.append((t, time.perf_counter() - t0))
    n = len(pts)
    mx, my = sum(p[0] for p in pts) / n, sum(p[1] for p in pts) / n
    b = sum((p[0] - mx) * (p[1] - my) for p in pts) / sum((p[0] - mx) ** 2 for p in pts)
    a = my - b * mx
    total = sum(a + b * t for t in range(1, 1281))
    return batch * 1280 / total, pts, total


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    host_cores = os.cpu_count() or 1
    sd = cpu_state_dict_vae()
    frames = 2  # cfg 2 at B=2 on the CPU (B=64 would take ~10 min per pass)
    cores = pick_cpu_threads(sd)  # untimed calibration pass doubles as the warm-up
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_vae_frames_per_s(sd, frames)
    dt = time.perf_counter() - t0
    fps = args.steps * frames / dt
    del sd
    decoder = generate = None
    if not args.skip_decoder:
        tps, tdt, tps_f, tdt_f = cpu_decoder_train_tokens_per_s()
        decoder = dict(metric="3DNA decoder video-tokens/sec, forward loss + backward", value=round(tps, 2), unit="tokens/s",
                       forward_only=dict(value=round(tps_f, 2), unit="tokens/s", seconds=round(tdt_f, 1)),
                       cpu_baseline=dict(value=round(tps, 2), unit="tokens/s", cores=cores, host_cores=host_cores, kind="port",
                                         sample=f"BASELINE configs[2] at B=1 (2560 video tokens, 256 text tokens, 12 layers), one "
                                                f"forward+backward pass ({tdt:.1f} s), PyTorch CPU fp32 autograd through the oracle"))
    if not args.skip_generate:
        gps, pts, total = cpu_generate_tokens_per_s()
        generate = dict(metric="generate(): sampled video-tokens/sec (AR loop)", tokens_per_s=round(gps, 4), unit="tokens/s",
                        cpu_baseline=dict(value=round(gps, 4), unit="tokens/s", cores=cores, host_cores=host_cores, kind="port",
                                          sample="BASELINE configs[3] (depth-64 reversible), B=1: single loop iterations (full "
                                                 "prefix recompute, two sweeps) timed at prefix lengths %s -> %s s; step(t) = a + b t "
                                                 "fitted and integrated over t = 1..1280 (%.0f s per 1280-token sample)"
                                                 % ([p[0] for p in pts], [round(p[1], 2) for p in pts], total)))
    line = dict(metric=METRIC, value=fps, unit="frames/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload="VQGanVAE dim=512 image_size=256 num_layers=4 num_resnet_blocks=2 "
                            "vq_codebook_size=8192 encode+VQ+decode (BASELINE configs[1])", batch_per_step=frames,
                            note="reference algorithm (oracle port, PyTorch CPU fp32); the reference itself cannot be "
                                 "imported on the GPU box (two absent third-party deps, no /root/reference)"),
                cpu_baseline=dict(value=fps, unit="frames/s", cores=cores, host_cores=host_cores, kind="port",
                                  sample=f"{frames} frame(s) of the 64-frame batch per step, {cores} of {host_cores} host "
                                         f"cores (fastest of all / half / quarter, calibrated untimed)"),
                e2e=dict(value=fps, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0,
                decoder=decoder, generate=generate)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def build_vae(dev):
    from nuwa_pytorch_b200 import VQGanVAE
    torch.manual_seed(0)
    with torch.device(dev):
        vae = VQGanVAE(**VAE_KW)
    return vae.eval()


def build_decoder(dev):
    from nuwa_pytorch_b200 import NUWA, VQGanVAE
    torch.manual_seed(0)
    with torch.device(dev):
        vae = VQGanVAE(**DEC_VAE_KW)
        nuwa = NUWA(vae=vae, **DEC_KW)
    return nuwa.eval()


def timed(fn, steps, warmup, dist, flush):
    """W warm-up + K timed steps between barrier+sync; CUDA events; returns max-over-ranks seconds."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        if flush is not None:
            flush.zero_()
        fn()
    b.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sec = a.elapsed_time(b) / 1e3
    if dist is not None:
        t = torch.tensor([sec], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    return sec


def attention_kernel_rooflines(dev, pk):
    """The two attention cores of the decoder at the cfg-3 shapes, timed alone (CUDA events, L2 flushed between launches):
    Sparse3DNA against its HBM roofline (4096 algorithmic bytes per token per layer), the text cross
    attention against the tensor roofline.  A few ms of GPU time."""
    from nuwa_pytorch_b200 import ops
    B, H, dh, nv = DEC_BATCH, 8, 64, 2559
    inner, n = H * dh, nv + 1
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(7)
    qkv = torch.randn(B, n, 3 * inner, device=dev, generator=g).bfloat16()
    talk = torch.randn(H, H, device=dev, generator=g) / 2
    o = torch.empty(B, n, inner, dtype=torch.bfloat16, device=dev)

    def med_us(fn, iters=7):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2] * 1e3

    out = dict(sparse3dna=[], cross=None)
    for dil in (1, 2, 4):
        for variant, kname in (("halo", "attn_3dna_halo_kernel (mma.sync; 'auto' choice for causal passes)"),
                               ("umma", "attn_3dna_umma_kernel (tcgen05 / TMEM)")):
            us = med_us(lambda: ops.attn_sparse3dna(qkv, o, B=B, nq=n, t0=0, npos=n, H=H, dh=dh, talk=talk, fmap=16,
                                                    max_frames=10, nv=nv, kernel=(5, 3, 3), dilation=(dil,) * 3, causal=True,
                                                    variant=variant))
            gbs = B * nv * 4096 / (us * 1e-6) / 1e9
            out["sparse3dna"].append(dict(dilation=dil, us_per_launch=round(us, 1), bound="hbm", achieved=round(gbs, 1),
                                          peak=pk["hbm_gbs"], unit="GB/s", frac=round(gbs / pk["hbm_gbs"], 4),
                                          algorithmic_bytes_per_token=4096, kernel=kname))
    nq, nk = 2560, 256
    q = torch.randn(B, nq, inner, device=dev, generator=g).bfloat16()
    kv = torch.randn(B, nk, 2 * inner, device=dev, generator=g).bfloat16()
    nk_, nv_ = torch.randn(inner, device=dev, generator=g), torch.randn(inner, device=dev, generator=g)
    mask = torch.ones(B, nk, dtype=torch.uint8, device=dev)
    oc = torch.empty(B, nq, inner, dtype=torch.bfloat16, device=dev)
    us = med_us(lambda: ops.attn_dense(q.data_ptr(), kv.data_ptr(), kv.data_ptr() + inner * 2, oc, B=B, nq=nq, nk=nk, H=H, dh=dh,
                                       q_bs=nq * inner, q_rs=inner, k_bs=nk * 2 * inner, k_rs=2 * inner, v_bs=nk * 2 * inner,
                                       v_rs=2 * inner, o_bs=nq * inner, o_rs=inner, talk=talk, null_k=nk_, null_v=nv_, key_mask=mask))
    fl = 2 * 2 * B * H * nq * (nk + 1) * dh
    out["cross"] = dict(us_per_launch=round(us, 1), bound="tensor", achieved=round(fl / us / 1e6, 1), peak=pk["bf16_tflops_sustained"],
                        unit="TFLOP/s", frac=round(fl / us / 1e6 / pk["bf16_tflops_sustained"], 4), kernel="attn_dense_pres_kernel",
                        note="mma.sync + CUDA-core softmax at 8 warps / SM; the fraction is against the tcgen05 GEMM peak")
    return out


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: nuwa_pytorch_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    _saved_stdout_fd = None
    if world > 1:
        import torch.distributed as dist_mod
        # NCCL prints its version banner straight to stdout when NCCL_DEBUG is set in the environment (it is on the GPU
        # boxes) -- rank 0 must print ONE JSON line, so file descriptor 1 points at stderr until that line is written
        sys.stdout.flush()
        _saved_stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist_mod.init_process_group("nccl", device_id=dev)
        dist = dist_mod
    from nuwa_pytorch_b200 import _lib
    pk = peaks()

    # -------- VQGanVAE frames/s (configs[1]) --------
    vae = build_vae(dev)
    B = args.vae_batch
    g = torch.Generator(device=dev).manual_seed(rank)
    img = torch.randn(B, 3, 256, 256, device=dev, generator=g)
    host_in = torch.randn(B, 3, 256, 256).pin_memory()
    host_out = torch.empty(B, 3, 256, 256).pin_memory()
    dev_in = torch.empty_like(img)

    def step_dev():
        return vae(img)

    def step_e2e():
        dev_in.copy_(host_in, non_blocking=True)
        out = vae(dev_in)
        host_out.copy_(out, non_blocking=True)

    with torch.no_grad():
        step_dev()  # builds the packed bf16 weights (one-off) outside any timed region
        torch.cuda.synchronize()
        launches0 = _lib.launch_count()
        prof = _lib.GemmProfiler()
        with prof, ClockSampler(local) as clocks:
            sec = timed(step_dev, args.steps, args.warmup, dist, None)
        n_gemm, gemm_flops, gemm_ms = prof.collect()
        gemm_alg_bytes = prof.bytes()
        prof.close()
        launches = _lib.launch_count() - launches0
        sec_e2e = timed(step_e2e, args.steps, max(1, args.warmup // 2), dist, None)
    fps = world * B * args.steps / sec
    fps_e2e = world * B * args.steps / sec_e2e
    # roofline of the dominant kernel (tcgen05 implicit-GEMM conv / GEMM): algorithmic FLOPs / summed launch time.
    # gemm_prof counts warm-up + timed launches alike (same shapes), so the ratio is a per-launch average.
    ach = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    peak = pk["bf16_tflops_sustained"]
    # DRAM traffic of the same launches: dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu pass
    # (profiles/r01_vae_gemm_traffic.json, same command at --steps 1), next to the algorithmic bytes of those launches.
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", "r01_vae_gemm_traffic.json")
    if os.path.exists(tp) and B == VAE_BATCH:
        tj = json.load(open(tp))
        traffic = round(tj["dram_bytes_per_launch"] / 1e9, 3)
        traffic_note = ("GB per launch, ncu dram__bytes_read.sum + dram__bytes_write.sum averaged over the %d launches of "
                        "one step (profiles/r01_vae_gemm_traffic.json)" % tj["launches_per_step"])
    alg_gb = gemm_alg_bytes / max(1, n_gemm) / 1e9
    roof = dict(bound="tensor", kernel="gemm_tcgen05_kernel (implicit-GEMM conv + GEMM)", achieved=round(ach, 1),
                peak=peak, unit="TFLOP/s", frac=round(ach / peak, 4), traffic=traffic, traffic_unit=traffic_note,
                algorithmic_gb_per_launch=round(alg_gb, 3),
                flop_per_byte=round(gemm_flops / max(1.0, gemm_alg_bytes), 1),
                peak_source=pk["source"] + ", sustained figure (kernel timed inside a long step)",
                launches_per_step=n_gemm // (args.steps + args.warmup),
                share_of_step=round((gemm_ms / (args.steps + args.warmup)) / (1e3 * sec / args.steps), 4),
                step_level=dict(gflop_per_frame=VAE_GFLOP_PER_FRAME,
                                achieved_tflops=round(fps / world * VAE_GFLOP_PER_FRAME / 1e3, 1),
                                frac=round(fps / world * VAE_GFLOP_PER_FRAME / 1e3 / peak, 4)))
    del vae, img, dev_in
    torch.cuda.empty_cache()

    # -------- 3DNA decoder video-tokens/s (configs[2], forward loss) --------
    decoder = None
    if not args.skip_decoder:
        nuwa = build_decoder(dev)
        gt = torch.Generator(device=dev).manual_seed(100 + rank)
        text = torch.randint(1, 49408, (DEC_BATCH, 256), device=dev, generator=gt)
        video = torch.randint(0, 8192, (DEC_BATCH, 10, 16, 16), device=dev, generator=gt)
        h_text, h_video = text.cpu().pin_memory(), video.cpu().pin_memory()

        from nuwa_pytorch_b200.graphs import GraphedCall

        def fwd(t, v):
            return nuwa(text=t, video=v, return_loss=True)

        with torch.no_grad():
            fwd(text, video)
            torch.cuda.synchronize()
            l0 = _lib.launch_count()
            fwd(text, video)
            dl = _lib.launch_count() - l0  # kernels per forward pass (the graph replays exactly these)
            graphed = GraphedCall(fwd, text, video)

        def dstep():
            return graphed(text, video)

        def dstep_e2e():
            return float(graphed(h_text, h_video).item())  # H2D of the ids, graph replay, D2H of the loss

        with torch.no_grad():
            dsec = timed(dstep, args.steps, args.warmup, dist, None)
            dsec_e2e = timed(dstep_e2e, args.steps, 1, dist, None)
        ntok = DEC_BATCH * 2560
        decoder = dict(metric="3DNA decoder video-tokens/sec", value=round(world * ntok * args.steps / dsec, 1),
                       unit="tokens/s", ms_per_step=round(1e3 * dsec / args.steps, 3),
                       e2e=dict(value=round(world * ntok * args.steps / dsec_e2e, 1), unit="tokens/s",
                                h2d_bytes_per_step=int(h_text.numel() * 8 + h_video.numel() * 8), d2h_bytes_per_step=4),
                       gpu_launches=int(dl),
                       config=dict(workload="NUWA dim=512 dec_depth=12 heads=8 max_video_frames=10 kernel (5,3,3) "
                                   "dilation (1,2,4), forward loss incl. 6-layer text encoder + logits + CE "
                                   "(BASELINE configs[2])", batch_per_gpu=DEC_BATCH, tokens_per_sample=2560,
                                   launch="one CUDA graph replay per step (nuwa_pytorch_b200.graphs.GraphedCall)",
                                   backward="not included (forward loss only; autograd kernels are next-round work)"),
                       flops=dict(mflop_per_token_fwd=104.4,
                                  achieved_tflops=round(world * ntok * args.steps / dsec * 104.4e6 / 1e12 / world, 1)))
        peak_t = pk["bf16_tflops_sustained"]
        fwd_tf = ntok * args.steps / dsec * 104.4e6 / 1e12
        decoder["roofline"] = dict(bound="tensor", achieved=round(fwd_tf, 1), peak=peak_t, unit="TFLOP/s",
                                   frac=round(fwd_tf / peak_t, 4),
                                   algorithmic="104.4 MFLOP per token forward (SURVEY 8d: 12 x 8.0 + logits 8.39) x 20480 tokens",
                                   peak_source=pk["source"] + ", sustained figure", traffic=None,
                                   note="whole forward step (219 launches); per-kernel rooflines of the attention cores under "
                                        "kernel_rooflines")
        del graphed
        # ---- training step of the same config: forward loss + backward (SURVEY §8d cfg 3 timed region) ----
        from nuwa_pytorch_b200.graphs import GraphedTrainStep
        from nuwa_pytorch_b200.parallel import GradAllReduce
        nuwa.train()
        tparams = [p for n, p in nuwa.named_parameters() if not n.startswith('vae.')]

        def loss_fn(t, v):
            return nuwa(text=t, video=v, return_loss=True)  # default cond_dropout_prob = 0.2, as NUWATrainer calls it

        def eager_step(t, v):
            for p in tparams:
                p.grad = None
            loss = loss_fn(t, v)
            loss.backward()
            return loss

        eager_step(text, video)
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        eager_step(text, video)
        tl = _lib.launch_count() - l0
        if world > 1:
            nuwa._grad_reducer = GradAllReduce(dist)  # the ONE collective: fp32 gradient all-reduce (mean) over NCCL
        try:  # whole step in ONE CUDA graph; at N > 1 the per-sub-block NCCL all-reduces are captured inside it
            tstepper = GraphedTrainStep(loss_fn, tparams, text, video)
            launch_mode = "one CUDA graph replay per step (forward + backward" + (
                ", overlapped NCCL all-reduce of the flat gradient buffer inside the graph" if world > 1 else "") + ")"
        except Exception as e:
            print(f"[bench] training step not captured ({type(e).__name__}: {e}); eager launches", file=sys.stderr)
            torch.cuda.synchronize()
            tstepper = eager_step
            launch_mode = "eager launches; per-sub-block NCCL all-reduce of the flat gradient buffer overlapping the backward"

        def tstep():
            return tstepper(text, video)

        def tstep_e2e():
            return float(tstepper(h_text.to(dev, non_blocking=True), h_video.to(dev, non_blocking=True)).item())

        tsec = timed(tstep, args.steps, args.warmup, dist, None)
        tsec_e2e = timed(tstep_e2e, args.steps, 1, dist, None)
        gnorm = float(torch.sqrt(sum((p.grad.float() ** 2).sum() for p in tparams if p.grad is not None)))
        decoder["train"] = dict(
            metric="3DNA decoder video-tokens/sec, forward loss + backward", value=round(world * ntok * args.steps / tsec, 1),
            unit="tokens/s", ms_per_step=round(1e3 * tsec / args.steps, 3),
            e2e=dict(value=round(world * ntok * args.steps / tsec_e2e, 1), unit="tokens/s",
                     h2d_bytes_per_step=int(h_text.numel() * 8 + h_video.numel() * 8), d2h_bytes_per_step=4),
            gpu_launches=int(tl), launch=launch_mode, grad_norm=round(gnorm, 5),
            collective=None if world == 1 else "all_reduce(AVG) of %.1f M fp32 gradients per step" % (
                sum(p.numel() for p in tparams) / 1e6),
            flops=dict(mflop_per_token_fwd_bwd=3 * 104.4,
                       achieved_tflops=round(ntok * args.steps / tsec * 3 * 104.4e6 / 1e12, 1)))
        tr_tf = ntok * args.steps / tsec * 3 * 104.4e6 / 1e12
        decoder["train"]["roofline"] = dict(bound="tensor", achieved=round(tr_tf, 1), peak=peak_t, unit="TFLOP/s",
                                            frac=round(tr_tf / peak_t, 4), traffic=None,
                                            algorithmic="3 x 104.4 MFLOP per token (forward + dgrad + wgrad) x 20480 tokens",
                                            peak_source=pk["source"] + ", sustained figure")
        decoder["config"]["backward"] = "see 'train' (same model and batch, loss.backward() included)"
        # ---- the whole trainer step (reference NUWATrainer.train_step, train_nuwa.py:237-258: grad_accum_every = 8
        # micro-batches -> [N>1: NCCL mean all-reduce overlapped with the last backward] -> clip 0.5 -> AdamW -> zero_grad)
        # through nuwa_pytorch_b200.trainer.TrainStep, captured into ONE CUDA graph when the capture succeeds ----
        try:
            from nuwa_pytorch_b200.trainer import TrainStep
            del tstepper
            for p_ in tparams:
                p_.grad = None
            ACC = 8
            trainer = TrainStep(nuwa, lr=3e-4, wd=0.01, grad_accum_every=ACC, max_grad_norm=0.5, dist=dist if world > 1 else None)
            mb = [dict(text=torch.randint(1, 49408, (DEC_BATCH, 256), device=dev, generator=gt),
                       video=torch.randint(0, 8192, (DEC_BATCH, 10, 16, 16), device=dev, generator=gt)) for _ in range(ACC)]
            trainer.step(mb)
            torch.cuda.synchronize()
            tmode = "eager launches"
            try:
                trainer.capture(mb)
                tmode = "one CUDA graph replay per trainer step (8 x (forward + backward), all-reduce, clip + AdamW inside)"
            except Exception as e:
                print(f"[bench] trainer step not captured ({type(e).__name__}: {e}); eager launches", file=sys.stderr)
                torch.cuda.synchronize()
                trainer.graph = None
            l_first = float(trainer.step(mb)[0])
            trsec = timed(lambda: trainer.step(mb), args.steps, args.warmup, dist, None)
            l_last = float(trainer.step(mb)[0])
            ttok = ACC * ntok
            decoder["trainer_step"] = dict(
                metric="3DNA decoder video-tokens/sec through the whole trainer step (8 micro-batches, clip, AdamW)",
                value=round(world * ttok * args.steps / trsec, 1), unit="tokens/s", ms_per_step=round(1e3 * trsec / args.steps, 3),
                grad_accum_every=ACC, launch=tmode, loss_first=round(l_first, 4), loss_last=round(l_last, 4),
                collective=None if world == 1 else "all_reduce(AVG) of the flat fp32 gradient buffer once per trainer step, "
                                                   "issued per finished sub-block during the last micro-batch's backward",
                reference="train_nuwa.py:237-258 + optimizer.py:11-31")
            del trainer
        except Exception as e:
            print(f"[bench] trainer step failed ({type(e).__name__}: {e})", file=sys.stderr)
            decoder["trainer_step"] = dict(error=f"{type(e).__name__}: {e}")
            tstepper = None
        with torch.no_grad():
            decoder["kernel_rooflines"] = attention_kernel_rooflines(dev, pk)
        del nuwa
        torch.cuda.empty_cache()

        # ---- NUWASketch training step (BASELINE configs[4]): 12-layer 3DNA sketch encoder over 3 sketch frames, ----
        # ---- 24-layer decoder with SparseCross2DNA, batch 4 per GPU, float sketch + video through both VAEs     ----
        if not args.skip_sketch:
            from nuwa_pytorch_b200 import NUWASketch, VQGanVAE
            torch.manual_seed(0)
            with torch.device(dev):
                svae = VQGanVAE(**{**DEC_VAE_KW, "channels": 5})
                vvae = VQGanVAE(**DEC_VAE_KW)
                sk = NUWASketch(vae=vvae, sketch_vae=svae, dim=512, image_size=256, sketch_enc_depth=12,
                                sketch_max_video_frames=3, sketch_enc_use_sparse_3dna=True, max_video_frames=10, dec_depth=24,
                                sparse_3dna_kernel_size=(5, 3, 3), sparse_3dna_dilation=(1, 2, 4)).train()
            SB = 4
            gs = torch.Generator(device=dev).manual_seed(300 + rank)
            sketch = torch.randn(SB, 3, 5, 256, 256, device=dev, generator=gs)
            svideo = torch.randn(SB, 10, 3, 256, 256, device=dev, generator=gs)
            smask = torch.ones(SB, 3, dtype=torch.bool, device=dev)
            h_sketch, h_svideo = sketch.cpu().pin_memory(), svideo.cpu().pin_memory()
            sparams = [p for n, p in sk.named_parameters() if not n.startswith('vae.') and not n.startswith('sketch_vae.')]
            if world > 1:
                sk._grad_reducer = GradAllReduce(dist)

            def sk_step(s_, v_):
                for p in sparams:
                    p.grad = None
                loss = sk(sketch=s_, sketch_mask=smask.clone(), video=v_, return_loss=True)
                loss.backward()
                return loss

            sk_step(sketch, svideo)
            torch.cuda.synchronize()
            l0 = _lib.launch_count()
            sk_step(sketch, svideo)
            sl = _lib.launch_count() - l0
            sk_mode = "eager launches" + ("" if world == 1 else " + overlapped NCCL gradient all-reduce")
            sk_runner = sk_step
            if True:  # whole step (both VAE encodes, forward, backward [, NCCL all-reduce]) as ONE CUDA graph
                try:
                    def sk_loss(s_, v_):
                        return sk(sketch=s_, sketch_mask=smask.clone(), video=v_, return_loss=True)
                    sk_runner = GraphedTrainStep(sk_loss, sparams, sketch, svideo)
                    sk_mode = "one CUDA graph replay per step (VAE encodes + forward + backward" + (
                        ", NCCL all-reduce inside the graph" if world > 1 else "") + ", GraphedTrainStep)"
                except Exception as e:  # capture is an optimisation of the launch path only
                    print(f"[bench] NUWASketch step not captured ({type(e).__name__}: {e}); eager launches", file=sys.stderr)
                    torch.cuda.synchronize()
                    sk_runner = sk_step
            ssec = timed(lambda: sk_runner(sketch, svideo), args.steps, args.warmup, dist, None)
            ssec_e2e = timed(lambda: float(sk_runner(h_sketch.to(dev, non_blocking=True), h_svideo.to(dev, non_blocking=True)).item()),
                             args.steps, 1, dist, None)
            stok = SB * 2560
            decoder["sketch_train"] = dict(
                metric="NUWASketch video-tokens/sec, forward loss + backward (incl. both VAE encodes)",
                value=round(world * stok * args.steps / ssec, 1), unit="tokens/s", ms_per_step=round(1e3 * ssec / args.steps, 3),
                e2e=dict(value=round(world * stok * args.steps / ssec_e2e, 1), unit="tokens/s",
                         h2d_bytes_per_step=int(h_sketch.numel() * 4 + h_svideo.numel() * 4), d2h_bytes_per_step=4),
                gpu_launches=int(sl), launch=sk_mode,
                config=dict(workload="NUWASketch dim=512 sketch_enc_depth=12 (Sparse3DNA) sketch_max_video_frames=3 dec_depth=24 "
                            "(Sparse3DNA + SparseCross2DNA) max_video_frames=10, loss.backward() (BASELINE configs[4])",
                            batch_per_gpu=SB, tokens_per_sample=2560, context_tokens=768))
            del sk, svae, vvae, sk_runner
            torch.cuda.empty_cache()

    # -------- generate() (configs[3]): depth-64 reversible decoder, 5 frames = 1280 AR steps, KV-cached, --------
    # -------- one CUDA-graph replay per token (both guidance sweeps + sampling inside the graph)          --------
    generate = None
    if not args.skip_generate:
        from nuwa_pytorch_b200 import NUWA, VQGanVAE
        torch.manual_seed(0)
        with torch.device(dev):
            gvae = VQGanVAE(**DEC_VAE_KW)
            gnuwa = NUWA(vae=gvae, **{**DEC_KW, "dec_depth": 64, "dec_reversible": True}).eval()
        gt = torch.Generator(device=dev).manual_seed(200 + rank)
        gtext = torch.randint(1, 49408, (DEC_BATCH, 256), device=dev, generator=gt)
        frames = args.gen_frames
        with torch.no_grad():
            gnuwa.generate(text=gtext, num_frames=1, _return_indices=True)  # warm-up (packs weights)
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c = torch.cuda.Event(enable_timing=True)
            a.record()
            idx = gnuwa.generate(text=gtext, num_frames=frames, _return_indices=True)
            b.record()
            vid = gnuwa._indices_to_video(idx, 10)
            c.record()
            torch.cuda.synchronize()
        t_ar, t_dec = a.elapsed_time(b) / 1e3, b.elapsed_time(c) / 1e3
        if dist is not None:
            tt = torch.tensor([t_ar, t_dec], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_ar, t_dec = float(tt[0]), float(tt[1])
        ntok = DEC_BATCH * frames * 256
        # algorithmic HBM bytes of ONE token step: both guidance sweeps stream every decoder weight once in bf16
        # (data dependent, cannot share a read; 126 MB of L2 cannot hold them), + the K/V windows of the 64 3DNA layers
        # (<= 46 keys x 2 x 512 x 2 B per layer per sample) and the cross-attention K/V (257 x 2 x 512 x 2 B) per sweep,
        # + the bf16 logits weight twice
        w_bytes = 2 * sum(p.numel() for n_, p in gnuwa.named_parameters()
                          if n_.startswith("video_transformer.layers.") or n_ == "to_logits.weight")
        n3 = sum(1 for n_, _ in gnuwa.named_parameters() if n_.endswith("fn.fn.to_kv.weight") and n_.startswith("video_transformer.layers."))
        nx = sum(1 for n_, _ in gnuwa.named_parameters() if n_.endswith(".fn.null_k") and n_.startswith("video_transformer.layers."))
        kv_bytes = DEC_BATCH * (n3 * 46 + nx * 257) * 2 * 512 * 2
        step_bytes = 2 * (w_bytes + kv_bytes)
        step_s = t_ar / (frames * 256)
        gen_roof = dict(bound="hbm", achieved=round(step_bytes / step_s / 1e9, 1), peak=pk["hbm_gbs"], unit="GB/s",
                        frac=round(step_bytes / step_s / 1e9 / pk["hbm_gbs"], 4), traffic=None,
                        algorithmic_mb_per_step=round(step_bytes / 1e6, 1),
                        algorithmic="2 sweeps x (%.1f MB bf16 decoder + logits weights + %.1f MB K/V windows) per token step at "
                                    "batch %d" % (w_bytes / 1e6, kv_bytes / 1e6, DEC_BATCH),
                        kernel="decode_stack_kernel (persistent, one launch per sweep)", peak_source=pk["source"])
        generate = dict(roofline=gen_roof, metric="generate(): sampled video-tokens/sec (AR loop) and decoded frames/sec",
                        tokens_per_s=round(world * ntok / t_ar, 1), ms_per_token_step=round(1e3 * t_ar / (frames * 256), 3),
                        vae_decode_frames_per_s=round(world * DEC_BATCH * frames / t_dec, 1),
                        config=dict(workload="NUWA dim=512 dec_depth=64 dec_reversible=True, generate(num_frames=%d), "
                                    "cond_scale=2 (two sweeps per token, SURVEY D8), filter_thres=0.9 (BASELINE configs[3])"
                                    % frames, batch_per_gpu=DEC_BATCH, steps=frames * 256, video_shape=list(vid.shape),
                                    launch="KV-cached incremental decode, one CUDA graph replay per token"))
        del gnuwa, gvae, idx, vid
        torch.cuda.empty_cache()

    # -------- CPU baseline (rank 0, N == 1 only): the reference algorithm's CPU path, bounded sample --------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        host_cores = os.cpu_count() or 1
        sd = cpu_state_dict_vae()
        cores = pick_cpu_threads(sd)
        v, dt = cpu_vae_frames_per_s(sd, 2)
        cpu = dict(value=round(v, 4), unit="frames/s", cores=cores, host_cores=host_cores, kind="port",
                   sample=f"2 frames of the 64-frame batch, one pass ({dt:.1f} s), PyTorch CPU fp32 oracle port of the reference; "
                          f"{cores} threads = the fastest of all / half / a quarter of the {host_cores} host cores")
        del sd
        if decoder is not None:
            tps, tdt, tps_f, tdt_f = cpu_decoder_train_tokens_per_s()
            decoder["cpu_baseline"] = dict(
                value=round(tps_f, 2), unit="tokens/s", cores=cores, host_cores=host_cores, kind="port",
                sample=f"BASELINE configs[2] at B=1 (2560 video + 256 text tokens, 12 layers): forward loss, one pass "
                       f"({tdt_f:.1f} s), PyTorch CPU fp32 oracle port")
            decoder["train"]["cpu_baseline"] = dict(
                value=round(tps, 2), unit="tokens/s", cores=cores, host_cores=host_cores, kind="port",
                sample=f"the same pass with loss.backward() through torch autograd ({tdt:.1f} s in total)")
        if generate is not None:
            gps, pts, total = cpu_generate_tokens_per_s()
            generate["cpu_baseline"] = dict(
                value=round(gps, 4), unit="tokens/s", cores=cores, host_cores=host_cores, kind="port",
                sample="BASELINE configs[3], B=1: single iterations of the reference loop (full prefix recompute, two sweeps) "
                       "timed at prefix lengths %s -> %s s; step(t) = a + b t fitted and integrated over t = 1..1280 "
                       "(%.0f s per 1280-token sample)" % ([p[0] for p in pts], [round(p[1], 2) for p in pts], total))

    if rank == 0:
        line = dict(metric=METRIC, value=round(fps, 2), unit="frames/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=round(1e3 * sec / args.steps, 3), higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                    config=dict(workload="VQGanVAE dim=512 image_size=256 num_layers=4 num_resnet_blocks=2 "
                                "vq_codebook_size=8192: encode + VQ + decode of every frame (BASELINE configs[1])",
                                batch_per_gpu=B, parallelism=f"replicas x{world} (batch sharded, no collective)",
                                l2="inputs larger than L2: 50 MB image batch, 4.4 GB bf16 weights, multi-GB activations"),
                    e2e=dict(value=round(fps_e2e, 2), unit="frames/s", h2d_bytes_per_step=int(host_in.numel() * 4),
                             d2h_bytes_per_step=int(host_out.numel() * 4)),
                    gpu_launches=int(launches), roofline=roof, cpu_baseline=cpu, clocks=clocks.summary(),
                    decoder=decoder, generate=generate)
        if _saved_stdout_fd is not None:  # give stdout back for the ONE line
            sys.stdout.flush()
            os.dup2(_saved_stdout_fd, 1)
        print(json.dumps(line), flush=True)
        if _saved_stdout_fd is not None:
            os.dup2(2, 1)  # anything the teardown prints goes to stderr again
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--vae-batch", type=int, default=VAE_BATCH)
    ap.add_argument("--skip-sketch", action="store_true", help="skip the NUWASketch training leg")
    ap.add_argument("--skip-decoder", action="store_true")
    ap.add_argument("--skip-generate", action="store_true")
    ap.add_argument("--gen-frames", type=int, default=5)
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
