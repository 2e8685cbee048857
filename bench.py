#!/usr/bin/env python
"""bench.py — LeftRefill DDIM/UNet hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path over one batch of synthetic input = DDIMSampler.sample of `--batch` stitched
512x1024 canvases (64x128 latents, 9-channel UNet input, cfg 2.5 -> UNet batch 2*batch) for `--ddim-steps` (50) DDIM
steps: BASELINE.json configs[1] ("batch=4 ref-inpainting, 50 DDIM steps, cfg=2.5, fp16, 1xB200").
Metric: stitched 512x1024 images/sec @ 50 DDIM steps (whole job, all ranks).

  value  : inputs resident in HBM when the timed region starts (device-timed with CUDA events, max over ranks)
  e2e    : the same through the public API with HOST (pinned) buffers: H2D of x_T / c_concat / contexts and D2H of the
           samples inside the timed region, every step
  roofline     : gemm_conv_kernel (all convs + linears, the dominant kernel by FLOPs), tensor bound: algorithmic FLOPs of
                 its launches / their summed CUDA-event durations in profiled UNet forwards run inside this process
  cpu_baseline : the oracle (a port: /root/reference does not exist on the GPU box) timed on the host cores on a bounded
                 sample (one CFG step of one canvas), extrapolated to the metric's unit

  gpu_reference: CONTEXT ONLY (never routed through repo kernels, never the product): the oracle (the reference's
                 PyTorch ops) run on the same GPU under torch.autocast, once with the reference's materialised-logits
                 attention (attention.py:168-196) and once with F.scaled_dot_product_attention swapped in - i.e. what
                 cuDNN + cuBLAS + a library flash kernel do on this box for the same UNet batch.

`--impl reference` times the reference's CPU implementation of the path (the oracle port) with all host threads.
`--config c2|c4|c5` selects the workload (BASELINE.json configs[1] (default) / [3] multiview 4-reference stitched /
[4] NVS 32x64 latent); the driver contract (no flag) is c2 at N = 1 and c3 = c2 per GPU at N = 8.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "stitched 512x1024 images/sec @ 50 DDIM steps"
UNIT = "images/s"


def _ncu_traffic():
    """DRAM bytes (read + write) per gemm_conv_kernel launch from the committed ncu launch list of one UNet forward
    (profiles/r1_ncu_forward_launches_summary.json: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
    dram__bytes_write.sum` over the same workload, cold cache per launch). None if the file is missing."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_forward_launches_summary.json")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r1_ncu_forward_launches_summary.json")
    try:
        d = json.load(open(p))
        rd = wr = n = 0
        for k, v in d.items():
            if "gemm_conv_kernel" in k:
                rd += v["dram_read"]
                wr += v["dram_write"]
                n += v["launches"]
        return ((rd + wr) / n if n else None), os.path.basename(p)
    except Exception:  # noqa: BLE001
        return None, None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = str(gpu_index)
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.idx, f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(threads=None, canvases=1, repeats=1, warmup=1, decode=True):
    """Oracle (CPU fp32 restatement of the reference UNet + DDIM step [+ first-stage decoder]) on the host cores, bounded
    sample: one CFG step (UNet batch 2) of `canvases` canvas at 64x128 latent [+ one decode of that canvas], extrapolated:
    images/s = canvases / (50 * t_step + t_decode). Returns (step times, decode seconds or 0, threads)."""
    import torch
    from helpers import O, synthetic_inputs
    from oracle import vae_oracle as V
    if threads:
        torch.set_num_threads(threads)
    cfg = O.DEFAULT_CFG
    sd = O.make_state_dict(cfg, seed=0)
    xT, c_cat, ctx, uc = synthetic_inputs(canvases)
    xc = torch.cat([torch.cat([xT, xT]), torch.cat([c_cat, c_cat])], dim=1)
    cc = torch.cat([uc, ctx])
    t = torch.full((2 * canvases,), 981, dtype=torch.long)
    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            t0 = time.perf_counter()
            e = O.unet_forward(sd, cfg, xc, t, cc)
            e_u, e_c = e.chunk(2)
            O.ddim_step(xT, e_u, e_c, torch.zeros_like(xT), 2.5, 0.5, 0.6, 0.0)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
        t_dec = 0.0
        if decode:
            del sd
            vsd = V.make_state_dict(V.DEFAULT_CFG, seed=0)
            t0 = time.perf_counter()
            V.decode(vsd, V.DEFAULT_CFG, xT * 0.7, scale_factor=V.SCALE_FACTOR)
            t_dec = time.perf_counter() - t0
    return times, t_dec, torch.get_num_threads()


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path (oracle port), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    ddim_steps = args.ddim_steps
    batch = args.batch if args.batch > 0 else CONFIGS["c2"]["batch"]
    times, t_dec, cores = cpu_baseline(threads=os.cpu_count(), canvases=1, repeats=args.steps,
                                       warmup=max(1, min(args.warmup, 1)), decode=not args.no_decode)
    t = sum(times) / len(times)
    value = 1.0 / (ddim_steps * t + t_dec)
    sample = (f"each step = 1 CFG DDIM step (UNet batch 2, 64x128 latent, fp32) of 1 canvas on {cores} threads = {t:.2f} s"
              f"{'' if args.no_decode else f'; + one first-stage decode of that canvas = {t_dec:.1f} s'}; "
              f"images/s = 1 / ({ddim_steps} x step time + decode time)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": (t * ddim_steps + t_dec) * 1e3 * batch,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"batch={batch} ref-inpainting canvases, {ddim_steps} DDIM steps, cfg=2.5, "
                                   "stitched 512x1024 (64x128 latent), SD2-inpainting UNet 865.9M params, random init",
                       "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


CONFIGS = {
    # BASELINE.json configs[1] (and configs[2] = the same per GPU on 8 GPUs)
    "c2": dict(H=64, W=128, batch=4, view=None, rows_per_sample=1,
               name="batch={B} ref-inpainting canvases per GPU, {S} DDIM steps, cfg=2.5, eta=1.0, stitched 512x1024 "
                    "(64x128 latent, 9-ch input, UNet batch 2*batch), SD2-inpainting UNet 865.9M params random init "
                    "(BASELINE.json configs[1])"),
    # configs[3]: multiview, 4 reference views stitched with the target: MultiViewUnetModel(view_num=5, concat_target=True),
    # every sample = 4 stitched [ref_i | target] canvases whose self-attention runs over 5*64*64 = 20480 tokens
    "c4": dict(H=64, W=128, batch=2, view=(5, True), rows_per_sample=4,
               name="batch={B} multiview samples per GPU x 4 stitched [ref_i | target] 512x1024 canvases (view_num=5, "
                    "concat_target=True: self-attention over 20480 tokens), {S} DDIM steps, cfg=2.5, UNet batch "
                    "2*4*batch (BASELINE.json configs[3])"),
    # configs[4]: NVS config (novel_view_synthesis.yaml: the plain UNetModel at a 32x64 latent), 4 canvases per GPU
    "c5": dict(H=32, W=64, batch=4, view=None, rows_per_sample=1,
               name="batch={B} NVS canvases per GPU (novel_view_synthesis.yaml, 256x512 stitched = 32x64 latent), {S} DDIM "
                    "steps, cfg=2.5, UNet batch 2*batch (BASELINE.json configs[4]: 16 canvases over 4 GPUs)"),
}


def device_unet(cls, cfg, dev, seed, **extra):
    """The reference architecture with synthetic weights drawn ON THE DEVICE (no checkpoint offline): the module is
    built on the meta device and filled in place, so a run does not spend a minute of host time initialising 866 M
    parameters. Same scaling rules as the oracle's make_state_dict (variance preserving; norm gains around 1; the
    reference's zero-initialised tensors drawn like the others)."""
    import torch
    with torch.device("meta"):
        m = cls(**cfg, **extra)
    m = m.to_empty(device=dev)
    g = torch.Generator(device=dev).manual_seed(seed)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if p.dim() == 1:
                p.normal_(0.0, 0.05, generator=g)
                if name.endswith("weight"):
                    p.mul_(2.0).add_(1.0)
            else:
                fan_in = p[0].numel()
                p.normal_(0.0, 1.0 / fan_in ** 0.5, generator=g)
    return m.eval()


def gpu_reference(unet, cfg, xc, tt, cc, iters=3):
    """CONTEXT ONLY - never the product, never through repo kernels: the oracle (the reference's PyTorch op sequence)
    on this GPU under torch.autocast with the same weights and UNet batch, (a) with the reference's own attention
    (materialised fp32 logits, attention.py:168-196), (b) with F.scaled_dot_product_attention swapped in. ms per UNet
    forward, CUDA events. This - cuDNN + cuBLAS (+ a library flash kernel) - is what the hand-written kernels compete
    with on the same box."""
    import torch
    from helpers import O
    sd = {k: p.detach() for k, p in unet.named_parameters()}
    out = {}
    for impl in ("vanilla", "sdpa"):
        O.ATTN_IMPL = impl
        try:
            with torch.no_grad(), torch.autocast("cuda"):
                O.unet_forward(sd, cfg, xc, tt, cc)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    O.unet_forward(sd, cfg, xc, tt, cc)
                e1.record()
                torch.cuda.synchronize()
            out[impl + "_ms_per_unet_forward"] = e0.elapsed_time(e1) / iters
        except Exception as e:  # noqa: BLE001 - context leg: report, never fail the bench
            out[impl + "_error"] = str(e)[:200]
        finally:
            O.ATTN_IMPL = "vanilla"
        torch.cuda.empty_cache()
    return out


def _ensure_built():
    """liblr_b200.so is a build artefact (git-ignored, shipped with the gpurun snapshot): build it if this checkout does
    not have it yet. The product path itself never falls back to anything when the library is missing."""
    from leftrefill_b200 import build as b
    if not b.is_stale():
        return
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        b.build(force=True, verbose=False)
    else:  # another local rank is building: wait for it
        t0 = time.time()
        while b.is_stale() and time.time() - t0 < 300:
            time.sleep(1.0)


def run_native(args):
    _ensure_built()
    import torch
    import torch.distributed as dist
    from helpers import FakeLDM, O, synthetic_inputs
    import leftrefill_b200 as lr
    from leftrefill_b200 import _native as N
    from leftrefill_b200 import parallel as P

    os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout: one JSON line only
    rank, world = P.init_distributed()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (sm_100a); there is no CPU fallback for the native arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    wl = CONFIGS[args.config]
    S = args.ddim_steps
    nsamp = args.batch if args.batch > 0 else wl["batch"]   # samples (images) per GPU and batch
    B = nsamp * wl["rows_per_sample"]                        # rows the sampler sees (stitched canvases)
    cfg = O.DEFAULT_CFG
    H, W = wl["H"], wl["W"]

    # ---- model: reference architecture, random init on the device (no checkpoint offline), replicated per rank ----
    if wl["view"] is None:
        unet = device_unet(lr.UNetModel, cfg, dev, seed=0)
    else:
        unet = device_unet(lr.MultiViewUnetModel, cfg, dev, seed=0, view_num=wl["view"][0], concat_target=wl["view"][1])
    ldm = FakeLDM(unet, dev)
    # first-stage decoder (SD2 VAE, configs/ref_inpainting.yaml:38-58), native: the reference's log_images ends with
    # decode_first_stage (inpainting_ldm/ref_inpainting_ldm.py:57), the metric counts decoded 512x1024 images
    from oracle import vae_oracle as V  # config constants only (ch, ch_mult, scale_factor); weights are drawn on device
    vae = None
    if not args.no_decode:
        vcfg = V.DEFAULT_CFG
        vae = device_unet(lr.AutoencoderKL, {}, dev, seed=1, ddconfig={k: v for k, v in vcfg.items() if k != "embed_dim"},
                          embed_dim=vcfg["embed_dim"])
    z_scale = 1.0 / V.SCALE_FACTOR

    # ---- synthetic inputs, rank-specific seed: weak scaling, B canvases per GPU ----
    xT_h, ccat_h, ctx_h, uc_h = [t.pin_memory() for t in synthetic_inputs(B, h=H, w=W, seed=1234 + rank)]
    xT, ccat, ctx, uc = [t.to(dev) for t in (xT_h, ccat_h, ctx_h, uc_h)]
    out_h = (torch.empty(B, 3, 8 * H, 8 * W) if vae is not None else torch.empty(B, 4, H, W)).pin_memory()

    def one_batch(x_T, c_cat, c_ctx, c_uc, decode=True):
        sampler = lr.DDIMSampler(ldm)
        cond = {"c_concat": [c_cat], "c_crossattn": [c_ctx]}
        ucond = {"c_concat": [c_cat], "c_crossattn": [c_uc]}
        samples, _ = sampler.sample(S, B, (4, H, W), cond, eta=1.0, x_T=x_T, verbose=False,
                                    unconditional_guidance_scale=2.5, unconditional_conditioning=ucond)
        if vae is not None and decode:
            return vae.decode(samples, z_scale=z_scale)          # [B, 3, 8H, 8W] images
        return samples

    def step_resident():
        s = one_batch(xT, ccat, ctx, uc)
        return P.gather_outputs(s, B * world, rank, world)      # the one collective of a batch

    def step_e2e():
        a, b, c, d = [t.to(dev, non_blocking=True) for t in (xT_h, ccat_h, ctx_h, uc_h)]
        s = one_batch(a, b, c, d)
        s = P.gather_outputs(s, B * world, rank, world)
        out_h.copy_(s[rank * B:(rank + 1) * B], non_blocking=True)
        torch.cuda.current_stream().synchronize()               # the caller holds the result on the host
        return s

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return P.max_over_ranks(e0.elapsed_time(e1) / 1e3, dev)

    for _ in range(args.warmup):
        step_resident()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    # the end-to-end arm (the headline) is timed first, then the HBM-resident arm; both under the clock sampler
    step_e2e()
    t_e2e = timed(step_e2e, args.steps)
    N.lib().lr_launch_count_reset()
    t_res = timed(step_resident, args.steps)
    launches = N.lib().lr_launch_count()
    # the same batch without the decode (latents gathered instead of images): the UNet / sampler share of the metric
    t_nodec = None
    if vae is not None:
        t_nodec = timed(lambda: P.gather_outputs(one_batch(xT, ccat, ctx, uc, decode=False), B * world, rank, world),
                        max(1, args.steps // 2))
        t_nodec /= max(1, args.steps // 2)
    clk = clocks.stop() if rank == 0 else None
    dec_ms = None
    if vae is not None:
        zz = torch.randn(B, 4, H, W, device=dev) * 0.7
        vae.decode(zz, z_scale=z_scale)
        torch.cuda.synchronize()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(5):
            vae.decode(zz, z_scale=z_scale)
        d1.record()
        torch.cuda.synchronize()
        dec_ms = d0.elapsed_time(d1) / 5

    # ---- per-kernel-class profile of the UNet forward (CUDA events between plan steps, same process) ----
    import ctypes
    h = unet.engine()
    xc = torch.cat([torch.cat([xT, xT]), torch.cat([ccat, ccat])], dim=1).contiguous()
    tt = torch.full((2 * B,), 981, dtype=torch.long, device=dev)
    unet.set_context(torch.cat([uc, ctx]).contiguous())
    N.lib().lr_unet_set_profiling(h, 1)
    ms_c = (ctypes.c_double * 5)()
    fl_c = (ctypes.c_double * 5)()
    n_c = (ctypes.c_int * 5)()
    acc_ms, acc_fl, acc_n = [0.0] * 5, [0.0] * 5, [0] * 5
    prof_iters = 5
    self_ms, self_fl = 0.0, 0.0  # self-attention only (tq == tk): the cross-attention launches are HBM-, not tensor-bound
    for i in range(prof_iters + 1):
        unet.forward_native(xc, tt, None)
        N.check(N.lib().lr_unet_read_profile(h, ms_c, fl_c, n_c), "read_profile")
        if i == 0:
            continue  # warm-up of the event pool
        for k in range(N.lib().lr_unet_num_steps(h)):
            ms_k, fl_k, cls_k = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
            buf = ctypes.create_string_buffer(256)
            N.check(N.lib().lr_unet_step_info(h, k, ctypes.byref(ms_k), ctypes.byref(fl_k), ctypes.byref(cls_k), buf,
                                              256), "step_info")
            if cls_k.value == 1:
                d = buf.value.decode()
                tq, tk = [int(d.split(key + "=")[1].split()[0]) for key in ("tq", "tk")]
                if tq == tk and ms_k.value > 0:
                    self_ms += ms_k.value
                    self_fl += fl_k.value
        for c in range(5):
            acc_ms[c] += ms_c[c]
            acc_fl[c] += fl_c[c]
            acc_n[c] += n_c[c]
    N.lib().lr_unet_set_profiling(h, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        unet.forward_native(xc, tt, None)
    e1.record()
    torch.cuda.synchronize()
    unet_ms = e0.elapsed_time(e1) / 10
    unet_flops = unet.last_flops()
    gref = None
    if rank == 0 and world == 1 and wl["view"] is None and not args.no_gpu_reference:
        gref = gpu_reference(unet, cfg, xc, tt, torch.cat([uc, ctx]).contiguous())
        if "vanilla_ms_per_unet_forward" in gref:
            gref["native_ms_per_unet_forward"] = unet_ms
            gref["speedup_vs_vanilla"] = gref["vanilla_ms_per_unet_forward"] / unet_ms
        if "sdpa_ms_per_unet_forward" in gref:
            gref["speedup_vs_sdpa"] = gref["sdpa_ms_per_unet_forward"] / unet_ms
        gref["note"] = ("context only: the oracle's PyTorch ops (cuDNN/cuBLAS, autocast fp16) on this GPU, same weights "
                        f"and UNet batch {xc.shape[0]} at {H}x{W}; not the product path")

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    sustained, burst, hbm, peak_src = _peaks()
    traffic, traffic_src = _ncu_traffic()
    names = ["gemm_conv_kernel", "attention_kernel", "groupnorm", "layernorm", "other"]
    classes = {names[c]: {"ms_per_forward": acc_ms[c] / prof_iters, "tflops": (acc_fl[c] / acc_ms[c] / 1e9) if acc_ms[c] > 0 and acc_fl[c] > 0 else None,
                          "steps_per_forward": acc_n[c] // prof_iters} for c in range(5)}
    gemm_launches = acc_n[0]
    achieved = acc_fl[0] / acc_ms[0] / 1e9
    sm_mhz = (clk or {}).get("sm_mhz") or 0
    mufu_bound = 148 * 16 * sm_mhz * 1e6 * 256 / 1e12 if sm_mhz else None
    roofline = {"kernel": "gemm_conv_kernel (tcgen05 implicit-GEMM conv3x3 + linear, all launches of a UNet forward)",
                "bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                "frac": achieved / sustained, "traffic": traffic,
                "traffic_note": "DRAM read+write bytes per launch, mean over the gemm_conv_kernel launches of one forward "
                                f"(ncu launch list of this build, cold cache per launch; profiles/{traffic_src})",
                "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)",
                "flops_per_launch": acc_fl[0] / gemm_launches, "ms_per_launch": acc_ms[0] / gemm_launches,
                "launches_timed": gemm_launches,
                "attention_kernel": {"achieved": classes["attention_kernel"]["tflops"], "peak": sustained,
                                     "frac": (classes["attention_kernel"]["tflops"] or 0) / sustained,
                                     "self_attention_tflops": (self_fl / self_ms / 1e9) if self_ms > 0 else None,
                                     "mufu_bound_tflops": mufu_bound,
                                     "self_attention_frac_of_mufu_bound":
                                         (self_fl / self_ms / 1e9 / mufu_bound) if self_ms > 0 and mufu_bound else None,
                                     "note": "d_head = 64: every score costs 256 tensor FLOP and one ex2; the MUFU unit "
                                             "issues 16 ex2/clk/SM, which caps the kernel at 148 SMs x 16 x clock x 256 "
                                             "FLOP (clock = median SM clock sampled during the run)"},
                "by_class": classes}

    cb = None
    if world == 1 and args.config == "c2":  # the CPU baseline is reported at N = 1 only (and for the headline config)
        cb_times, cb_dec, cores = cpu_baseline(threads=os.cpu_count(), canvases=1, repeats=1, warmup=1,
                                               decode=vae is not None)
        cb_t = sum(cb_times) / len(cb_times)
        cb = {"value": 1.0 / (S * cb_t + cb_dec), "unit": UNIT, "cores": cores, "kind": "port",
              "sample": f"1 CFG DDIM step (UNet batch 2, 64x128 latent, fp32 oracle) of 1 canvas = {cb_t:.2f} s on "
                        f"{cores} threads, first-stage decode of 1 canvas = {cb_dec:.1f} s; images/s = 1/({S} x step + "
                        "decode)"}

    images = nsamp * world * args.steps
    h2d = sum(t.numel() * t.element_size() for t in (xT_h, ccat_h, ctx_h, uc_h))
    line = {"metric": METRIC, "value": images / t_res, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_res / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": wl["name"].format(B=nsamp, S=S), "name": args.config,
                       "l2": "inputs larger than L2: 1.73 GB fp16 weights + >2 GB activations per UNet forward vs 126 MB",
                       "parallelism": f"dp{world}: canvases sharded, weights replicated, 1 all-gather of the decoded "
                                      "images per batch",
                       "precision": "fp16 operands, fp32 accumulate / norm statistics / softmax / DDIM state"},
            "decode": None if vae is None else {
                "included_in_value_and_e2e": True, "ms_per_batch": dec_ms, "tflop_per_batch": vae.last_flops() / 1e12,
                "tflops": vae.last_flops() / dec_ms / 1e9,
                "value_without_decode": nsamp * world / t_nodec, "unit": UNIT,
                "note": "native first-stage decoder (lr_vae_decode) after the 50 DDIM steps; value_without_decode "
                        "gathers the latents instead (the round-1 definition of the metric)"},
            "unet_ms_per_ddim_step": unet_ms, "unet_tflops": unet_flops / unet_ms / 1e9,
            "unet_tflop_per_forward": unet_flops / 1e12,
            "e2e": {"value": images / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": out_h.numel() * 4,
                    "result": "decoded fp32 images" if vae is not None else "fp32 latents"},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cb,
            "gpu_reference": gref}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU (0 = the config's: 4 canvases for c2 / c5, "
                                                         "2 multiview samples for c4); the UNet batch is 2x rows with CFG")
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the PyTorch-ops-on-GPU context leg")
    ap.add_argument("--no-decode", action="store_true", help="stop at the latents (no first-stage decode)")
    ap.add_argument("--ddim-steps", type=int, default=50)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
