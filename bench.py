#!/usr/bin/env python
"""bench.py — denoised MRI slices/sec of the cDDPM reconstruction hot path on B200.

Workload (BASELINE.json configs[1]): conditioned UNet (43.9 M params) + ResNet-50 condition encoder, batch of 32
synthetic 96x96 slices per GPU, full reverse loop from T0 = test_timesteps = 500 (GaussianDiffusion.sample ->
p_sample_loop, cond_DDPM.py:446-464) with a fresh simplex-noise field per step, random-init weights.
One "step" = encoder + 500 UNet forwards + 500 posterior steps for one batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value` = device-timed whole-job throughput with inputs resident in HBM; `e2e` = the same
through the public drop-in API with host buffers (H2D of the slices and D2H of the reconstructions inside the timed
region); `roofline` = tensor-core rate of the dominant kernel (the tcgen05 implicit-GEMM convolution) from CUDA events
bracketing its launches inside the timed region; `cpu_baseline` = the oracle port (plain PyTorch fp32) on the host cores
for a bounded sample.  `--impl reference` times that CPU implementation alone.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "conditioned-diffusion-models-uad_b200")]

METRIC = "denoised MRI slices/sec (cDDPM: ResNet-50 encoder + conditioned UNet, reverse loop from T0)"
UNET_GFLOP = 149.138  # algorithmic GFLOP per UNet forward per 96x96 slice (SURVEY.md §8d / BASELINE.md §2)
ENC_GFLOP = 1.473


class Cfg(dict):
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


def model_cfg():
    return Cfg(imageDim=[192, 192, 100], rescaleFactor=2, unet_dim=128, dim_mults=[1, 2, 2], condition=True,
               backbone="Spark_Encoder_2D", version="resnet50", cond_dim=128, noisetype="simplex", noise_ensemble=True,
               test_timesteps=500, lr=1e-4, resizedEvaluation=True, erodeBrainmask=True, medianFiltering=True,
               saveOutputImages=False, evalSeg=True, threshold="auto", spatial_transformer=False,
               pretrained_encoder=False, objective="pred_x0")


def synthetic_volume(seed: int, depth: int = 50, size: int = 96):
    """One synthetic BraTS21-shaped case, [1,1,H,W,D] float32 host tensors: a smooth 0.2-0.8
    field inside an ellipsoid brain mask, 1-3 spherical lesions (+0.3) as the segmentation."""
    import torch

    g = torch.Generator().manual_seed(7000 + seed)
    H = W = size
    coarse = torch.rand(1, 1, 12, 12, max(2, depth // 5), generator=g)
    field = 0.2 + 0.6 * torch.nn.functional.interpolate(coarse, size=(H, W, depth), mode="trilinear", align_corners=True)[0, 0]
    y, x, z = torch.meshgrid(torch.arange(H), torch.arange(W), torch.arange(depth), indexing="ij")
    c = ((H - 1) / 2, (W - 1) / 2, (depth - 1) / 2)
    mask = ((y - c[0]) / (0.42 * H)) ** 2 + ((x - c[1]) / (0.36 * W)) ** 2 + ((z - c[2]) / (0.48 * depth)) ** 2 <= 1
    seg = torch.zeros_like(mask)
    for _ in range(int(torch.randint(1, 4, (1,), generator=g))):
        r = float(torch.randint(3, 9, (1,), generator=g))
        p = [c[i] + float(torch.randn(1, generator=g)) * f * n for i, (f, n) in enumerate(((0.12, H), (0.1, W), (0.12, depth)))]
        seg |= (y - p[0]) ** 2 + (x - p[1]) ** 2 + (z - p[2]) ** 2 <= r * r
    seg &= mask
    vol = (field * mask + 0.3 * seg.float()).clamp(0, 1)
    u = lambda t: t.float()[None, None].contiguous()  # noqa: E731
    return {"vol": u(vol), "mask_orig": u(mask), "seg_orig": u(seg)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops": float(p.get("bf16_tflops_sustained", 1395.6)), "tflops_burst": float(p.get("bf16_tflops", 1619.1)),
                "hbm_gbs": float(p.get("hbm_gbs", 6557.8)), "source": "MEASURED_PEAKS.json (sustained, of measured)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "B200_PROFILING.md fallback (of fallback)"}


# ---------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for nme, val in zip(names, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nme)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_sample(batch: int, rev_steps: int, start_t: int, threads: int):
    """Oracle port (plain PyTorch fp32, the restatement of the reference's own CPU path) on the host cores: encoder +
    `rev_steps` reverse steps for `batch` slices, extrapolated to the full start_t-step loop.  Returns slices/s."""
    import numpy as np
    import torch

    from oracle import diffusion_port, resnet_port, unet_port
    from oracle.simplex_port import gen_noise_port
    from oracle.weights import make_state_dict, synthetic_slices

    torch.set_num_threads(threads)
    spec = unet_port.UNetSpec()
    sd = make_state_dict(unet_port.param_shapes(spec), seed=1)
    esd = make_state_dict(resnet_port.param_shapes(128), seed=3)
    sched = diffusion_port.schedule_buffers()
    x = synthetic_slices(batch, 96, seed=0)
    np.random.seed(0)
    model = lambda xt, t, c: unet_port.unet_forward(sd, spec, xt, t, c)  # noqa: E731
    with torch.no_grad():
        t0 = time.perf_counter()
        cond = resnet_port.resnet_forward(esd, x)
        t_enc = time.perf_counter() - t0
        t0 = time.perf_counter()
        # reverse_loop runs `rev_steps` steps when started at index rev_steps; the schedule index only changes scalars
        diffusion_port.reverse_loop(model, sched, x * 2 - 1, cond, rev_steps, lambda: gen_noise_port((batch, 1, 96, 96)))
        t_rev = time.perf_counter() - t0
    per_step = t_rev / rev_steps
    total = t_enc + per_step * start_t
    return batch / total, {"t_encoder_s": t_enc, "t_per_reverse_step_s": per_step}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def best_cpu_sample(start_t, threads, rev_steps=2):
    """oneDNN picks very different convolution paths for batch 1 and batch 2 on some hosts; report the faster one."""
    best = None
    for b in (1, 2):
        v, parts = cpu_reference_sample(b, rev_steps, start_t, threads)
        if best is None or v > best[0]:
            best = (v, parts, b)
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    S = 2
    for _ in range(args.warmup):
        cpu_reference_sample(1, 1, args.start_t, threads)
    vals = []
    t0 = time.perf_counter()
    B = 1
    for _ in range(args.steps):
        v, _, B = best_cpu_sample(args.start_t, threads, S)
        vals.append(v)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    v = statistics.median(vals)
    sample = (f"oracle port (fp32 PyTorch restatement of the reference CPU path), best of batch 1 / batch 2 (here {B}) x {S} "
              f"reverse steps + encoder per step, extrapolated x{args.start_t // S} to the {args.start_t}-step loop")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "slices/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 32),
            "cpu_baseline": {"value": v, "unit": "slices/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(args, batch):
    return {"workload": f"configs[1]: cDDPM (DDPM_cond_spark_2D) conditioned UNet + ResNet-50 encoder, batch {batch} "
                        f"synthetic 96x96 slices per GPU, full reverse loop from T0={args.start_t} on B200",
            "batch_per_gpu": batch, "start_t": args.start_t, "image": "1x96x96", "objective": "pred_x0",
            "noise": "simplex (GPU, one field per step)", "parallelism": f"slice-sharded x{args.gpus}, no data-path collective",
            "l2": "streaming working set ~1.9 GB per UNet forward >> 126 MB L2 (no flush needed)"}


# ---------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the cDDPM engine has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from cddpm.ddpm_2d import DDPM_2D

    B, T0 = args.batch, args.start_t
    torch.manual_seed(1234 + rank)
    np.random.seed(1234 + rank)
    cfg = model_cfg()
    if args.dtype == "bf16":
        cfg["engine_dtype"] = "bf16"  # same tcgen05 rate; fp16 is the default because it is 10x closer to the fp32 reference
    model = DDPM_2D(cfg, prefix="bench/")
    with torch.no_grad():  # random-init every tensor, including the reference's zero-initialised output convolutions
        for name, p in model.named_parameters():
            if p.dim() >= 2 and float(p.abs().sum()) == 0.0:
                fan_in = p[0].numel()
                p.normal_(0.0, 1.0 / fan_in ** 0.5)
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.fill_(1.0)
    model = model.to(dev).eval()
    diffusion = model.diffusion
    unet = diffusion.model

    g = torch.Generator().manual_seed(7 + rank)
    x_host = torch.rand(B, 1, 96, 96, generator=g).pin_memory()
    out_host = torch.empty(B, 1, 96, 96).pin_memory()
    x_dev = x_host.to(dev)

    def step(x):
        with torch.no_grad():
            cond = model(x)
            return diffusion.sample(cond=cond, x_start=x * 2 - 1, start_t=T0, noise=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(x_dev)
    barrier()
    eng = unet.engine()

    # ---------------- timed region 1: inputs resident in HBM
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    conv_ms, conv_launches = [], 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        eng.profile_arm()  # the first UNet forward of this step is bracketed per convolution launch
        res = step(x_dev)
        ms, n = eng.profile_read()
        conv_ms.append(ms)
        conv_launches = n
    e1.record()
    barrier()
    dt_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else {}

    # ---------------- timed region 2: end to end through the public API with host buffers
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        xd = x_host.to(dev, non_blocking=True)
        res = step(xd)
        out_host.copy_(res, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        _ = float(out_host[0, 0, 0, 0])  # the host consumes the result
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)

    t = torch.tensor([dt_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # the only collective of the sweep: gather per-slice scores (here the mean reconstruction) to every rank
        score = res.mean(dim=(1, 2, 3)).contiguous()
        gathered = [torch.empty_like(score) for _ in range(world)]
        dist.all_gather(gathered, score)
    dt_ms, e2e_ms = float(t[0]), float(t[1])

    # ---------------- sub-records (every rank takes part: the volume / sweep workloads carry collectives)
    extras = {}
    if not args.no_extras:
        import copy

        sub = copy.copy(args)
        sub.steps, sub.warmup, sub.no_cpu_baseline = 2, 3, True
        if world == 1:
            extras["forward"] = forward_latency(unet, dev)
            try:
                extras["gpu_eager_baseline"] = gpu_eager_sample(B, T0, dev)
            except Exception as e:  # a baseline leg must never take the bench line down
                extras["gpu_eager_baseline"] = {"unavailable": repr(e)[:200]}
        def guarded(name, fn):  # a sub-record must never take the headline line down (failures are symmetric over ranks)
            try:
                extras[name] = fn()
            except Exception as e:
                extras[name] = {"unavailable": repr(e)[:300]}
                print(f"bench.py: sub-record {name} failed: {e!r}", file=sys.stderr)

        sub.batch = 8
        guarded("volume", lambda: run_volume(sub, sub=True))
        sub.batch = 64
        guarded("train", lambda: run_train(sub, sub=True))
        if world > 1:
            guarded("sweep", lambda: sweep_record(args, dev, world, rank))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    slices = B * world * args.steps
    value = slices / (dt_ms / 1e3)
    e2e_value = slices / (e2e_ms / 1e3)
    pk = peaks()
    conv_flops_fwd = eng.conv_flops_per_sample * B
    conv_ms_fwd = statistics.median(conv_ms)
    achieved = conv_flops_fwd / (conv_ms_fwd / 1e3) / 1e12
    enc_launches = 1 + 1 + 16 * 4 + 4 + 3 + 2  # stem, pool, (3 GEMM + im2col) x 16 blocks, 4 downsample GEMMs, 3 strided gathers, pool + fc
    launches_per_step = enc_launches + 2 + T0 * (eng.launches_per_forward + 2)
    # traffic: dram__bytes_read.sum + dram__bytes_write.sum of the timed launches of one B=32 forward at HEAD (ncu --set
    # full, profiles/r02_v4_ncu_conv.csv: 55 of the 56 conv_igemm2 launches, 4.462 GB read + 0.575 GB written =
    # 91.6 MB per launch, like `achieved`; the FiLM GEMM, profiles/r02_v3_ncu_misc.csv, reads 24.3 MB).  Algorithmic
    # bytes of the 57 launches (every operand read once, every result written once, 16-bit): 5.21 GB = 91.4 MB each.
    traffic = 91.6e6 * B / 32
    roofline = {"bound": "tensor",
                "kernel": "conv_igemm2_kernel / conv_igemm_kernel (tcgen05 implicit GEMM: every convolution and GEMM launch "
                          "of a UNet forward)",
                "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
                "peak_source": pk["source"], "flops_per_launch": conv_flops_fwd / max(1, conv_launches),
                "avg_launch_ms": conv_ms_fwd / max(1, conv_launches), "launches_timed": conv_launches * len(conv_ms),
                "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read + write)",
                "unet_forward_tflops": UNET_GFLOP * B * T0 * args.steps / 1e3 / (dt_ms / 1e3),  # per GPU
                "unet_forward_frac_of_peak": UNET_GFLOP * B * T0 * args.steps / 1e3 / (dt_ms / 1e3) / pk["tflops"]}
    line = {"metric": METRIC, "value": value, "unit": "slices/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.dtype == "bf16" else "f16", "data": "synthetic",
            "config": workload_config(args, B),
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": "slices/s", "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
            "single_step_equiv_slices_per_s": value * T0}
    if extras:
        def brief(rec):  # the sub-records keep their own roofline / e2e / clocks; drop the long prose
            if not isinstance(rec, dict):
                return rec
            return {k: v for k, v in rec.items() if k not in ("higher_is_better", "vs_baseline", "data", "scaling")}
        line["extras"] = {k: brief(v) for k, v in extras.items()}
    if world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        v, parts, bb = best_cpu_sample(T0, threads, 2)
        line["cpu_baseline"] = {"value": v, "unit": "slices/s", "cores": threads, "kind": "port",
                                "sample": f"oracle port (fp32 PyTorch) on the host: best of batch 1 / batch 2 (here {bb}) x 2 "
                                          f"reverse steps + encoder, extrapolated to the {T0}-step loop", **parts}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------- sub-records
def gpu_eager_sample(B, T0, dev, rev_steps=10):
    """SURVEY.md §8(d) "the real bar": the reference's own op sequence as eager PyTorch on THIS GPU - the oracle's
    restatement of UNetModel.forward / the reverse loop (/root/reference does not exist on the GPU box) with the weights
    on the device, torch.autocast(float16) (the reference's `precision: 16`), cudnn.benchmark, channels-last left to
    cuDNN.  `rev_steps` reverse steps of batch B, device-timed, extrapolated to the T0-step loop.  A reported baseline
    (like cpu_baseline), never the product path."""
    import numpy as np
    import torch

    from oracle import diffusion_port, resnet_port, unet_port
    from oracle.weights import make_state_dict

    prev = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    try:
        spec = unet_port.UNetSpec()
        sd = {k: v.to(dev) for k, v in make_state_dict(unet_port.param_shapes(spec), seed=1).items()}
        esd = {k: v.to(dev) for k, v in make_state_dict(resnet_port.param_shapes(128), seed=3).items()}
        sched = {k: v.to(dev) for k, v in diffusion_port.schedule_buffers().items()}
        x = torch.rand(B, 1, 96, 96, device=dev)
        g = torch.Generator(device=dev).manual_seed(0)
        noise_fn = lambda: torch.randn(1, 1, 96, 96, device=dev, generator=g).expand(B, 1, 96, 96).half()  # noqa: E731
        model = lambda xt, t, c: unet_port.unet_forward(sd, spec, xt, t, c).float()  # noqa: E731

        def loop(n):
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                cond = resnet_port.resnet_forward(esd, x).float()
                return diffusion_port.reverse_loop(model, sched, x * 2 - 1, cond, n, noise_fn)

        loop(3)  # cudnn.benchmark picks its algorithms here
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loop(rev_steps)
        e1.record()
        torch.cuda.synchronize()
        per_step_ms = e0.elapsed_time(e1) / rev_steps
    finally:
        torch.backends.cudnn.benchmark = prev
    del np
    return {"value": B / (per_step_ms * T0 / 1e3), "unit": "slices/s", "ms_per_reverse_step": per_step_ms,
            "kind": "oracle port of the reference's op sequence, eager PyTorch on this GPU (fp16 autocast, cuDNN/cuBLAS, "
                    "cudnn.benchmark)",
            "sample": f"batch {B}: encoder + {rev_steps} reverse steps device-timed, extrapolated to the {T0}-step loop"}


def forward_latency(unet, dev, batches=(1, 8, 32)):
    """Device-timed UNet forward at several batch sizes (the small-batch regime of SURVEY.md §7)."""
    import torch

    out = {}
    for B in batches:
        x = torch.randn(B, 1, 96, 96, device=dev)
        t = torch.full((B,), 499, device=dev, dtype=torch.long)
        c = torch.randn(B, 128, device=dev)
        with torch.no_grad():
            for _ in range(4):
                unet(x, t, cond=c)
            torch.cuda.synchronize()
            n = 50 if B < 16 else 20
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                unet(x, t, cond=c)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out[f"B{B}"] = {"ms": ms, "slices_per_s": B / ms * 1e3, "tflops": UNET_GFLOP * B / ms}
    return out


def sweep_record(args, dev, world, rank, vols_per_gpu=16):
    """BASELINE configs[3] inside the driver's record: cddpm.sweep.test_sweep (validation stage, then test stage) over
    `vols_per_gpu` x world synthetic volumes dealt round-robin to the ranks, WITH its collectives inside the timed region
    (all-gather of the per-volume result lists per stage, all-reduced counts of the global Dice-threshold bisection),
    device-timed (CUDA events between barriers, max over ranks); then rank 0 repeats the sweep alone on all volumes and
    compares every per-volume entry.  The simplex noise of a volume is seeded from its index (the reference draws from
    numpy's global stream, which a sharded run cannot share)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from cddpm import eval_tail, sweep
    from cddpm.ddpm_2d import DDPM_2D

    keys = ("IDs", "DiceScorePerVol", "BestDicePerVol", "BestThresholdPerVol", "AUCPerVol", "AUPRCPerVol", "HausPerVol",
            "TPPerVol", "FPPerVol", "TNPerVol", "FNPerVol", "AnomalyScoreRecoPerVol", "AnomalyScoreRegPerVol",
            "l1recoErrorAll", "lesionSizePerVol")
    nvol = vols_per_gpu * world
    torch.manual_seed(1234)  # identical replicas
    cfg = model_cfg()
    cfg["force_num_eval_slices"] = False
    model = DDPM_2D(cfg, prefix="sweep/")
    with torch.no_grad():
        for _, p in model.named_parameters():
            if p.dim() >= 2 and float(p.abs().sum()) == 0.0:
                p.normal_(0.0, 1.0 / p[0].numel() ** 0.5)
    model = model.to(dev).eval()
    inner = model.test_step_reconstruct

    def seeded(batch):
        np.random.seed(5000 + int(batch["ID"][0][1:]))
        return inner(batch)

    model.test_step_reconstruct = seeded

    def loader(stage, first):
        out = []
        for i in range(first, first + nvol):
            if i % world == rank or rank == 0:  # rank 0 keeps everything for the one-process check
                v = synthetic_volume(i % 24, 50)  # 24 distinct cases, reused
                out.append({"Dataset": ["Brats21"], "vol": {"data": v["vol"]}, "vol_orig": {"data": v["vol"]},
                            "seg_orig": {"data": v["seg_orig"]}, "mask_orig": {"data": v["mask_orig"]},
                            "seg_available": True, "ID": [f"v{i}"], "stage": stage, "label": torch.tensor([1])})
            else:
                out.append(None)
        return out

    sets = {"Datamodules_eval.Brats21": (loader("val", 0), loader("test", 1000))}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sweep.test_sweep(model, sets, pickle_preds=False)  # plans, graph captures, communicator warm-up
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    preds, _ = sweep.test_sweep(model, sets, pickle_preds=False)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_sharded = float(t[0])
    # the collectives alone: one all-gather of result lists per stage + the bisection's all-reduces, timed on a dummy
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    g0.record()
    if world > 1:
        payload = {k: list(v) for k, v in preds["val"]["Datamodules_eval.Brats21"].items() if type(v) is list}
        gathered = [None] * world
        dist.all_gather_object(gathered, (vols_per_gpu, {k: v[:vols_per_gpu] for k, v in payload.items()}))
        cnt = torch.zeros(9, dtype=torch.int64, device=dev)
        for _ in range(10):
            dist.all_reduce(cnt)
    g1.record()
    barrier()
    gather_ms = g0.elapsed_time(g1)
    rec = None
    if rank == 0:
      try:
        saved = (sweep._world, eval_tail._dist_sum, eval_tail._dist_max)
        sweep._world = lambda: (0, 1)
        eval_tail._dist_sum = eval_tail._dist_max = lambda t: None
        try:
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record()
            preds1, _ = sweep.test_sweep(model, sets, pickle_preds=False)
            h1.record()
            torch.cuda.synchronize()
            ms_serial = h0.elapsed_time(h1)
        finally:
            sweep._world, eval_tail._dist_sum, eval_tail._dist_max = saved
        bad = []
        for stage in ("val", "test"):
            a, b = preds[stage]["Datamodules_eval.Brats21"], preds1[stage]["Datamodules_eval.Brats21"]
            for k in keys:
                xa, xb = list(a[k]), list(b[k])
                if not (len(xa) == len(xb) and all(x == y or (x != x and y != y) for x, y in zip(xa, xb))):
                    bad.append(f"{stage}/{k}")
            for k in ("DicePerVolMean", "AUPRCPerVolMean", "HausPerVolMean"):
                if not (a[k] == b[k] or (a[k] != a[k] and b[k] != b[k])):
                    bad.append(f"{stage}/{k}")
        v = preds["val"]["Datamodules_eval.Brats21"]
        rec = {"metric": "volumes/sec of the sharded test sweep (validation + test stage, full tail, NCCL collectives inside "
                         "the timed region)",
               "value": 2 * nvol / (ms_sharded / 1e3), "unit": "volumes/s", "n_gpus": world, "volumes": 2 * nvol,
               "volumes_per_gpu_per_stage": vols_per_gpu, "ms": ms_sharded, "collectives_ms": gather_ms,
               "one_process_volumes_per_s": 2 * nvol / (ms_serial / 1e3), "equals_one_process_result": not bad,
               "mismatches": bad[:8], "val_DicePerVolMean": float(v["DicePerVolMean"]),
               "val_AUPRCPerVolMean": float(v["AUPRCPerVolMean"])}
      except Exception as e:  # the other ranks are waiting at the barrier below: always get there
        rec = {"unavailable": repr(e)[:300], "value": 2 * nvol / (ms_sharded / 1e3), "unit": "volumes/s", "ms": ms_sharded}
    barrier()
    return rec


# ---------------------------------------------------------------------------------------------- training workload
TRAIN_METRIC = "training slices/sec (cDDPM DDPM_2D.training_step: encoder + conditioned UNet fwd+bwd, L1 pred_x0 loss, Adam)"


def cpu_train_sample(threads: int):
    """Oracle port, one UNet forward + backward (fp32 autograd) for one slice on the host cores -> slices/s."""
    import torch

    from oracle import unet_port
    from oracle.weights import make_state_dict, synthetic_slices

    torch.set_num_threads(threads)
    spec = unet_port.UNetSpec()
    sd = {k: v.requires_grad_(True) for k, v in make_state_dict(unet_port.param_shapes(spec), seed=1).items()}
    x = synthetic_slices(1, 96, seed=0)
    t0 = time.perf_counter()
    out = unet_port.unet_forward(sd, spec, x * 2 - 1, torch.tensor([500]), torch.zeros(1, 128))
    (out - (x * 2 - 1)).abs().mean().backward()
    return 1.0 / (time.perf_counter() - t0)


def run_train(args, sub=False):
    """BASELINE.json configs[4] (not the headline line): `python bench.py --workload train [--batch 64]`."""
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the cDDPM engine has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not sub:
        dist.init_process_group("nccl", device_id=dev)
    from cddpm.ddpm_2d import DDPM_2D
    from cddpm.dist_train import sync_gradients

    B = args.batch
    torch.manual_seed(1234)  # identical replicas
    np.random.seed(1234 + rank)
    cfg = model_cfg()
    cfg["engine_dtype"] = "bf16"
    model = DDPM_2D(cfg, prefix="bench/")
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2 and float(p.abs().sum()) == 0.0:
                p.normal_(0.0, 1.0 / p[0].numel() ** 0.5)
    model = model.to(dev).train()
    opt = model.configure_optimizers()
    g = torch.Generator().manual_seed(7 + rank)
    x_host = torch.rand(B, 1, 96, 96, 1, generator=g).pin_memory()
    x_dev = x_host.to(dev)

    def step(x):
        opt.zero_grad(set_to_none=True)
        loss = model.training_step({"vol": {"data": x}}, 0)["loss"]
        loss.backward()
        if world > 1:
            sync_gradients(model)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(x_dev)
    barrier()
    eng = model.diffusion.model.train_engine()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(x_dev)
    e1.record()
    barrier()
    dt_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else {}
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    f0.record()
    for _ in range(args.steps):
        loss = step(x_host.to(dev, non_blocking=True))
        _ = float(loss.detach())  # the host reads the loss every step
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    # the engine's forward + backward launch lists alone, for the roofline
    xt = torch.randn(B, 1, 96, 96, device=dev)
    tt = torch.randint(0, 1000, (B,), device=dev)
    cc = torch.randn(B, 128, device=dev)
    dd = torch.randn(B, 1, 96, 96, device=dev) / (B * 9216)
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.forward(xt, tt, cc)
    eng.backward(dd, want_dcond=True)
    torch.cuda.synchronize()
    h0.record()
    for _ in range(args.steps):
        eng.forward(xt, tt, cc)
        eng.backward(dd, want_dcond=True)
    h1.record()
    torch.cuda.synchronize()
    unet_ms = h0.elapsed_time(h1) / args.steps
    t = torch.tensor([dt_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1 and not sub:
            dist.destroy_process_group()
        return None
    pk = peaks()
    flops = (eng.conv_flops_per_sample + eng.bwd_flops_per_sample) * B
    achieved = flops / (unet_ms / 1e3) / 1e12
    slices = B * world * args.steps
    line = {"metric": TRAIN_METRIC, "value": slices / (dt_ms / 1e3), "unit": "slices/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"configs[4]: cDDPM training step, batch {B} synthetic 96x96 slices per GPU, random t, "
                                   "simplex noise, L1 pred_x0, Adam(1e-4), bf16 activations / gradients",
                       "batch_per_gpu": B, "parallelism": f"data-parallel x{world}: one all-reduce of the UNet's flat "
                                                          "gradient buffer + one for the encoder per step",
                       "l2": "streaming working set ~3.8 GB of activations per step >> 126 MB L2 (no flush needed)"},
            "roofline": {"bound": "tensor", "kernel": "UNet forward + backward launch lists (tcgen05 conv / dgrad / "
                         "wgrad + the HBM-bound GroupNorm kernels between them), timed as a whole with CUDA events",
                         "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
                         "peak_source": pk["source"], "flops_per_step": flops, "unet_fwd_bwd_ms": unet_ms, "traffic": None},
            "e2e": {"value": slices / (e2e_ms / 1e3), "unit": "slices/s", "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": (eng.launches_per_forward + int(lib_bwd_launches(eng))) * args.steps, "clocks": clocks}
    if world == 1 and not args.no_cpu_baseline and not sub:
        threads = host_threads()
        cpu_train_sample(threads)
        v = max(cpu_train_sample(threads) for _ in range(2))
        line["cpu_baseline"] = {"value": v, "unit": "slices/s", "cores": threads, "kind": "port",
                                "sample": "oracle port (fp32 PyTorch autograd) on the host: UNet forward + backward of one "
                                          "slice, best of 2 after a warm-up"}
    if sub:
        return line
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return line


# ---------------------------------------------------------------------------------------------- volume workload
VOLUME_METRIC = ("volumes/sec (cDDPM full-volume reconstruction + anomaly map: encoder, 3-step noise ensemble over all "
                 "slices, residual, eroded brain mask, 5x5x5 median, threshold search, component filter, Dice/AUROC/AUPRC/"
                 "Hausdorff)")
# launches of the anomaly-scoring tail per volume: residual_erode, median, max, 10 bisection count launches, ranking
# (gather, sort, 2 scans, 3 small), threshold mask, component filter, confusion counts, Hausdorff (6), row statistics
TAIL_LAUNCHES = 1 + 1 + 1 + 10 + 8 + 1 + 1 + 1 + 6 + 1


def cpu_volume_sample(depth: int, threads: int):
    """Oracle port on the host cores: encoder + the three ensemble reconstructions for a 2-slice sample (extrapolated to
    `depth` slices) + the whole numpy / scipy tail of one volume.  Returns volumes/s."""
    import numpy as np
    import torch

    from oracle import diffusion_port, resnet_port, tail_port, unet_port
    from oracle.simplex_port import gen_noise_port
    from oracle.weights import make_state_dict

    torch.set_num_threads(threads)
    spec = unet_port.UNetSpec()
    sd = make_state_dict(unet_port.param_shapes(spec), seed=1)
    esd = make_state_dict(resnet_port.param_shapes(128), seed=3)
    sched = diffusion_port.schedule_buffers()
    v = synthetic_volume(0, depth)
    vol = v["vol"][0, 0]
    x = vol[..., depth // 2:depth // 2 + 2].permute(2, 0, 1)[:, None].contiguous()
    np.random.seed(0)
    with torch.no_grad():
        t0 = time.perf_counter()
        cond = resnet_port.resnet_forward(esd, x)
        recos = []
        for t in (250, 500, 750):
            noise = gen_noise_port(tuple(x.shape))
            tt = torch.full((x.shape[0],), t - 1, dtype=torch.long)
            xt = diffusion_port.q_sample(sched, x * 2 - 1, tt, noise)
            recos.append((unet_port.unet_forward(sd, spec, xt, tt, cond).clamp(-1, 1) + 1) / 2)
        t_model = (time.perf_counter() - t0) * depth / x.shape[0]
    reco = (vol * v["mask_orig"][0, 0]).numpy()
    t0 = time.perf_counter()
    tail_port.volume_tail(reco, vol.numpy(), v["seg_orig"][0, 0].numpy(), v["mask_orig"][0, 0].numpy(), stage="val")
    t_tail = time.perf_counter() - t0
    return 1.0 / (t_model + t_tail), {"t_model_s": t_model, "t_tail_s": t_tail}


def run_volume(args, sub=False):
    """BASELINE.json configs[2] (not the headline line): `python bench.py --workload volume [--batch 8]`: per step,
    `batch` synthetic BraTS21-shaped [1,1,96,96,50] volumes per GPU through the validation stage of the test sweep
    (on_test_start, test_step per volume, on_test_end with the global threshold search)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the cDDPM engine has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not sub:
        dist.init_process_group("nccl", device_id=dev)
    from cddpm.ddpm_2d import DDPM_2D
    from cddpm.sweep import run_stage

    NV, D = args.batch, 50
    torch.manual_seed(1234)
    np.random.seed(1234 + rank)
    cfg = model_cfg()
    cfg["force_num_eval_slices"] = False  # all 50 slices, not the fork's hard-coded 4
    if args.dtype == "bf16":
        cfg["engine_dtype"] = "bf16"
    model = DDPM_2D(cfg, prefix="bench/")
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2 and float(p.abs().sum()) == 0.0:
                p.normal_(0.0, 1.0 / p[0].numel() ** 0.5)
    model = model.to(dev).eval()
    unet = model.diffusion.model

    def batch_of(v):
        return {"Dataset": ["Brats21"], "vol": {"data": v["vol"]}, "vol_orig": {"data": v["vol"]},
                "seg_orig": {"data": v["seg_orig"]}, "mask_orig": {"data": v["mask_orig"]}, "seg_available": True,
                "ID": ["synthetic"], "stage": "val", "label": torch.tensor([1])}

    # run_stage deals the list round-robin to ranks: NV volumes per GPU; only this rank's entries carry data
    host, resident = [], []
    for i in range(NV * world):
        if i % world == rank:
            v = {k: t.pin_memory() for k, t in synthetic_volume(i, D).items()}
            host.append(batch_of(v))
            resident.append(batch_of({k: t.to(dev) for k, t in v.items()}))
        else:
            host.append(None)
            resident.append(None)

    def step(batches):
        ed = run_stage(model, batches, dev)
        return float(model.threshold["total"]), ed

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(resident)
    barrier()
    eng = unet.engine()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    conv_ms, conv_launches = [], 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for it in range(args.steps):
        eng.profile_arm()  # the first UNet forward of EVERY step is bracketed per convolution launch (median reported)
        thr, ed = step(resident)
        ms, n = eng.profile_read()
        conv_ms.append(ms)
        conv_launches = n
    e1.record()
    barrier()
    dt_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else {}
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        thr, ed = step(host)  # pinned host tensors in, eval_dict scalars out
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    t = torch.tensor([dt_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1 and not sub:
            dist.destroy_process_group()
        return None
    vols = NV * world * args.steps
    pk = peaks()
    # the 3-member noise ensemble of a volume is ONE UNet forward over the 3 x D stacked slices unless the reference's
    # member loop is selected (cfg stack_ensemble: False)
    k_ens = len(cfg.get("step_ensemble", [250, 500, 750]))
    stacked = bool(cfg.get("stack_ensemble", True))
    fwd_batch = D * k_ens if stacked else D
    conv_flops_fwd = eng.conv_flops_per_sample * fwd_batch
    conv_ms_fwd = statistics.median(conv_ms)
    achieved = conv_flops_fwd / (conv_ms_fwd / 1e3) / 1e12
    enc_launches = 1 + 1 + 16 * 4 + 4 + 3 + 2
    per_volume = enc_launches + (1 if stacked else k_ens) * eng.launches_per_forward + k_ens * 3 + TAIL_LAUNCHES
    line = {"metric": VOLUME_METRIC, "value": vols / (dt_ms / 1e3), "unit": "volumes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.dtype == "bf16" else "f16", "data": "synthetic",
            "config": {"workload": f"configs[2]: full-volume reconstruction + anomaly map, {NV} synthetic BraTS21-shaped "
                                   f"[1,1,96,96,{D}] volumes per GPU per step, validation stage of the test sweep "
                                   "(noise ensemble 250/500/750, all slices), global threshold search at the end",
                       "volumes_per_gpu": NV, "slices_per_volume": D, "image": "1x96x96",
                       "parallelism": f"volume-sharded x{world}; all-gather of per-volume results + all-reduce of the "
                                      "threshold-search counts at the end of the stage",
                       "ensemble": ("one UNet forward over the 3 x D stacked slices" if stacked else "one UNet forward per member"),
                       "l2": f"streaming working set ~{0.06 * fwd_batch:.0f} GB per B={fwd_batch} UNet forward >> 126 MB L2 (no flush needed)"},
            "roofline": {"bound": "tensor", "kernel": "conv_igemm2_kernel / conv_igemm_kernel (every convolution and GEMM "
                                                       f"launch of the first B={fwd_batch} UNet forward of each step, median over steps)",
                         "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
                         "peak_source": pk["source"], "flops_per_launch": conv_flops_fwd / max(1, conv_launches),
                         "avg_launch_ms": conv_ms_fwd / max(1, conv_launches), "launches_timed": conv_launches * len(conv_ms),
                         "traffic": None},
            "e2e": {"value": vols / (e2e_ms / 1e3), "unit": "volumes/s", "h2d_bytes_per_step": NV * 4 * 96 * 96 * D * 4,
                    "d2h_bytes_per_step": NV * (128 * 4 + 4 + 7 * 8 + 2 * 8 + 10 * 5 * 8 + 3 * 8 + 4 * 8 + 96 * 5 * 8 + 4),
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": per_volume * NV * args.steps, "clocks": clocks,
            "slices_per_s": vols * D / (dt_ms / 1e3), "unet_slice_evaluations_per_s": vols * D * 3 / (dt_ms / 1e3),
            "val_threshold": thr}
    if not args.no_cpu_baseline and world == 1 and not sub:
        threads = host_threads()
        v, parts = cpu_volume_sample(D, threads)
        line["cpu_baseline"] = {"value": v, "unit": "volumes/s", "cores": threads, "kind": "port",
                                "sample": "oracle port (fp32 PyTorch + numpy/scipy) on the host: encoder + 3 ensemble "
                                          f"reconstructions of 2 slices extrapolated to {D}, plus the full tail of one volume",
                                **parts}
    if sub:
        return line
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return line


def lib_bwd_launches(eng):
    from cddpm._lib import lib

    return lib().cddpm_unet_bwd_launches(eng._h)


_JSON_OUT = None


def emit(line):
    """The one JSON line of this run, on the process's ORIGINAL stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries exactly one JSON line.  NCCL prints its banner ("NCCL version ...") to file descriptor 1 from native
    # code (NCCL_DEBUG_FILE does not move it), so keep a private copy of the real stdout for the JSON line and point
    # descriptor 1 at stderr for everything else.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--workload", default="reverse", choices=["reverse", "train", "volume"],
                    help="reverse = BASELINE configs[1] (the headline line, default); volume = configs[2] (--batch = "
                         "volumes per GPU per step); train = configs[4]")
    ap.add_argument("--start-t", dest="start_t", type=int, default=500)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the sub-records of the headline line (configs[2] volume, configs[4] train, the eager-GPU "
                         "baseline, the forward latencies; at N > 1 the configs[3] sharded sweep)")
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"],
                    help="16-bit storage / tensor-core operand type of the inference engines (reverse, volume); fp32 "
                         "accumulation either way.  The reference runs fp16 autocast (trainer/default.yaml:7); training is "
                         "always bf16")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print(f"note: --warmup {args.warmup} < 3 breaks the timing rules; use >= 3 for a reportable number", file=sys.stderr)
    if args.batch is None:
        args.batch = {"train": 64, "volume": 8}.get(args.workload, 32)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "train":
        run_train(args)
    elif args.workload == "volume":
        run_volume(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
