"""ORACLE (test infrastructure).  Generates tests/golden/*.npz by importing the UNMODIFIED reference live from
/root/reference (only possible in the build container; the GPU box never sees /root/reference).

    NUMBA_CACHE_DIR=/tmp/numba_cache PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Stub packages under oracle/stubs/ shadow the third-party imports that are not installed here.
Weights come from oracle/weights.make_state_dict (deterministic, per-key seeded) and are loaded into the reference
modules with strict=True — which also pins our statement of the state_dict key/shape layout.
"""
from __future__ import annotations

import json
import os
import sys

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CDDPM_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "oracle", "stubs"), REF, ROOT]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import diffusion_port, resnet_port, unet_port  # noqa: E402
from oracle.weights import make_state_dict, synthetic_slices, synthetic_volume  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class Cfg(dict):
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


def base_cfg(**over):
    c = Cfg(imageDim=[192, 192, 100], rescaleFactor=2, unet_dim=128, dim_mults=[1, 2, 2], condition=True,
            backbone="Spark_Encoder_2D", version="resnet50", cond_dim=128, noisetype="simplex", noise_ensemble=True,
            test_timesteps=500, lr=1e-4, resizedEvaluation=True, erodeBrainmask=True, medianFiltering=True,
            saveOutputImages=False, evalSeg=True, threshold="auto", spatial_transformer=False,
            pretrained_encoder=False)
    c.update(over)
    return c


def save(name, **arrays):
    path = os.path.join(GOLD, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrays.items()})
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


def ref_unet(spec: unet_port.UNetSpec, image=96):
    from src.models.modules.OpenAI_Unet import UNetModel

    m = UNetModel(image_size=(image, image), in_channels=spec.in_channels, model_channels=spec.model_channels,
                  out_channels=spec.out_channels, num_res_blocks=spec.num_res_blocks,
                  attention_resolutions=tuple(spec.attention_resolutions), dropout=0,
                  channel_mult=list(spec.channel_mult), conv_resample=True, dims=2, num_classes=spec.num_classes,
                  use_checkpoint=False, use_fp16=True, num_heads=1, num_head_channels=spec.num_head_channels,
                  num_heads_upsample=-1, use_scale_shift_norm=True, resblock_updown=True,
                  use_new_attention_order=True, use_spatial_transformer=False, transformer_depth=1)
    shapes = unet_port.param_shapes(spec)
    ref_shapes = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert ref_shapes == shapes, "oracle.unet_port.param_shapes disagrees with the reference state_dict"
    return m.eval(), shapes


def golden_schedule():
    from src.models.modules.cond_DDPM import GaussianDiffusion

    d = GaussianDiffusion(torch.nn.Identity(), image_size=(96, 96), timesteps=1000, sampling_timesteps=1000,
                          objective="pred_x0", channels=1, loss_type="l1", p2_loss_weight_gamma=0, cfg=base_cfg())
    bufs = {k: v for k, v in d.state_dict().items()}
    assert list(bufs.keys()) == diffusion_port.BUFFER_NAMES, list(bufs.keys())
    mine = diffusion_port.schedule_buffers()
    for k in bufs:
        assert torch.equal(bufs[k], mine[k]), k
    save("schedule.npz", **bufs)


def golden_simplex():
    from src.utils.generate_noise import gen_noise

    from oracle.simplex_port import gen_noise_port

    out = {}
    for seed in (0, 7):
        np.random.seed(seed)
        ref = gen_noise(base_cfg(), (2, 1, 96, 96))
        np.random.seed(seed)
        mine = gen_noise_port((2, 1, 96, 96))
        assert ref.dtype == torch.float16 and tuple(ref.shape) == (2, 1, 96, 96)
        assert torch.equal(ref, mine), f"simplex port differs from the reference for seed {seed}"
        out[f"field_seed{seed}"] = ref[0, 0]
    save("simplex.npz", **out)


def golden_unet():
    # 1. the production geometry, conditioned and unconditioned, B=2 at 96x96
    for tag, ncls in (("cond", 128), ("uncond", None)):
        spec = unet_port.UNetSpec(num_classes=ncls)
        m, shapes = ref_unet(spec)
        sd = make_state_dict(shapes, seed=1)
        m.load_state_dict(sd, strict=True)
        x = synthetic_slices(2, 96, seed=3) * 2 - 1 + 0.3 * torch.randn(2, 1, 96, 96, generator=torch.Generator().manual_seed(4))
        t = torch.tensor([499, 37])
        cond = torch.randn(2, 128, generator=torch.Generator().manual_seed(5)) if ncls else None
        with torch.no_grad():
            y = m(x, t, cond=cond)
            mine = unet_port.unet_forward(sd, spec, x, t, cond)
        err = (y - mine).abs().max().item()
        print(f"unet {tag}: |ref| max {y.abs().max().item():.3f}, port-vs-ref max abs {err:.3g}")
        assert err < 1e-4
        arrays = dict(x=x, t=t, y=y)
        if cond is not None:
            arrays["cond"] = cond
        save(f"unet_{tag}_96.npz", **arrays)
    with open(os.path.join(GOLD, "unet_cond_keys.json"), "w") as f:
        json.dump([[k, list(s)] for k, s in unet_port.param_shapes(unet_port.UNetSpec())], f)
    # 2. a small geometry for quick tests (64 base channels, 2 levels, 32x32, attention in the middle only)
    spec = unet_port.UNetSpec(model_channels=64, channel_mult=(1, 2), num_res_blocks=1, num_classes=128)
    m, shapes = ref_unet(spec, image=32)
    sd = make_state_dict(shapes, seed=2)
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(3, 1, 32, 32, generator=g)
    t = torch.tensor([0, 250, 999])
    cond = torch.randn(3, 128, generator=g)
    with torch.no_grad():
        y = m(x, t, cond=cond)
        mine = unet_port.unet_forward(sd, spec, x, t, cond)
    assert (y - mine).abs().max().item() < 1e-4
    save("unet_small_32.npz", x=x, t=t, cond=cond, y=y)


def golden_encoder():
    from src.models.modules.DDPM_encoder import get_encoder

    cfg = base_cfg()
    enc, out_features = get_encoder(cfg)
    assert out_features == 128
    shapes = resnet_port.param_shapes(128)
    ref_shapes = [(k, tuple(v.shape)) for k, v in enc.state_dict().items()]
    assert ref_shapes == [("encoder." + k, s) for k, s in shapes], "encoder state_dict layout mismatch"
    sd = make_state_dict(shapes, seed=3)
    enc.load_state_dict({"encoder." + k: v for k, v in sd.items()}, strict=True)
    enc.eval()
    x = synthetic_slices(2, 96, seed=8)
    with torch.no_grad():
        c = enc(x)
        mine = resnet_port.resnet_forward(sd, x)
    err = (c - mine).abs().max().item()
    print(f"encoder: |c| max {c.abs().max().item():.3f}, port-vs-ref {err:.3g}")
    assert err < 1e-4
    save("encoder_96.npz", x=x, c=c)


def golden_diffusion():
    from src.models.modules import cond_DDPM
    from src.models.modules.cond_DDPM import GaussianDiffusion

    spec = unet_port.UNetSpec(model_channels=64, channel_mult=(1, 2), num_res_blocks=1, num_classes=128)
    m, shapes = ref_unet(spec, image=32)
    sd = make_state_dict(shapes, seed=2)
    m.load_state_dict(sd, strict=True)
    sched = diffusion_port.schedule_buffers()
    g = torch.Generator().manual_seed(9)
    img = torch.rand(3, 1, 32, 32, generator=g)
    cond = torch.randn(3, 128, generator=g)
    noise = (0.6 * torch.randn(1, 1, 32, 32, generator=g)).repeat(3, 1, 1, 1).half()
    out = {"img": img, "cond": cond, "noise": noise}
    model = lambda x, t, c: unet_port.unet_forward(sd, spec, x, t, c)  # noqa: E731
    for objective in ("pred_x0", "pred_noise"):
        d = GaussianDiffusion(m, image_size=(32, 32), timesteps=1000, sampling_timesteps=1000, objective=objective,
                              channels=1, loss_type="l1", p2_loss_weight_gamma=0, cfg=base_cfg())
        with torch.no_grad():
            loss, reco = d(img, cond=cond, t=499, noise=noise)
            ploss, preco = diffusion_port.reconstruct(model, sched, img, 499, cond, noise, objective=objective)
        assert (reco - preco).abs().max().item() < 1e-4 and abs(loss.item() - ploss.item()) < 1e-5
        out[f"reco_{objective}"] = reco
        out[f"loss_{objective}"] = loss
    # reverse loop (cond_DDPM.py:517-530, :446-464); the shipped code needs use_spatial_transformer set by the caller
    d = GaussianDiffusion(m, image_size=(32, 32), timesteps=1000, sampling_timesteps=1000, objective="pred_x0",
                          channels=1, loss_type="l1", p2_loss_weight_gamma=0, cfg=base_cfg())
    d.use_spatial_transformer = False
    T0 = 4
    noises = [(0.6 * torch.randn(1, 1, 32, 32, generator=g)).repeat(3, 1, 1, 1).half() for _ in range(T0 + 1)]
    feed = list(noises)
    orig = cond_DDPM.gen_noise
    cond_DDPM.gen_noise = lambda cfg, shape: feed.pop(0)
    try:
        with torch.no_grad():
            rec = d.sample(cond=cond, x_start=img * 2 - 1, start_t=T0, noise=True)
    finally:
        cond_DDPM.gen_noise = orig
    feed2 = list(noises)
    with torch.no_grad():
        prec = diffusion_port.reverse_loop(model, sched, img * 2 - 1, cond, T0, lambda: feed2.pop(0))
    assert (rec - prec).abs().max().item() < 1e-4
    out["reverse_T0"] = T0
    out["reverse_noises"] = torch.stack(noises)
    out["reverse_out"] = rec
    save("diffusion_small_32.npz", **out)


def golden_tail():
    """utils_eval._test_step / _test_end on synthetic volumes with a GIVEN reconstruction (val then test stage)."""
    from src.utils import utils_eval

    from oracle import tail_port

    class Host:
        pass

    res = {}
    for depth, tag in ((50, "d50"), (4, "d4")):
        host = Host()
        host.cfg = base_cfg()
        host.eval_dict = utils_eval.get_eval_dictionary()
        host.threshold = {}
        host.dataset = ["Brats21"]
        host.new_size = [160, 190, 160]
        host.diffs_list, host.seg_list = [], []
        vols = [synthetic_volume(s, depth=depth) for s in (0, 1)]
        # ---- val stage
        host.stage = "val"
        ports = []
        for i, v in enumerate(vols):
            utils_eval._test_step(host, v["reco"].clone(), v["vol"].clone(), v["seg_orig"].clone(),
                                  v["mask_orig"].clone(), i, [f"v{i}"], torch.tensor([1]))
            ports.append(tail_port.volume_tail(v["reco"][0, 0].numpy(), v["vol"][0, 0].numpy(),
                                               v["seg_orig"][0, 0].numpy(), v["mask_orig"][0, 0].numpy(), stage="val"))
        val_dict = {k: list(host.eval_dict[k]) for k in
                    ("DiceScorePerVol", "BestDicePerVol", "BestThresholdPerVol", "AUCPerVol", "AUPRCPerVol", "TPPerVol",
                     "FPPerVol", "TNPerVol", "FNPerVol", "TPRPerVol", "FPRPerVol", "l1recoErrorAll", "l2recoErrorAll",
                     "l1recoErrorUnhealthy", "l1recoErrorHealthy", "l2recoErrorUnhealthy", "l2recoErrorHealthy",
                     "AccuracyPerVol", "PrecisionPerVol", "RecallPerVol", "SpecificityPerVol", "lesionSizePerVol",
                     "DiceScorePerSlice", "AnomalyScoreRecoPerSlice", "labelPerSlice", "AnomalyScoreRecoPerVol",
                     "AUCAnomalyRecoPerSlice", "AUPRCAnomalyRecoPerSlice", "HausPerVol")}
        utils_eval._test_end(host)
        total = host.threshold["total"]
        # port parity on the val stage
        for i, p in enumerate(ports):
            assert p["BestThreshold"] == val_dict["BestThresholdPerVol"][i], (p["BestThreshold"], val_dict["BestThresholdPerVol"][i])
            assert p["BestDice"] == val_dict["BestDicePerVol"][i]
            assert p["Dice"] == val_dict["DiceScorePerVol"][i]
            assert (p["TP"], p["FP"], p["TN"], p["FN"]) == tuple(int(val_dict[k][i]) for k in ("TPPerVol", "FPPerVol", "TNPerVol", "FNPerVol"))
            assert abs(p["AUC"] - val_dict["AUCPerVol"][i]) < 1e-12 and abs(p["AUPRC"] - val_dict["AUPRCPerVol"][i]) < 1e-12
            assert abs(p["l1recoErrorAll"] - val_dict["l1recoErrorAll"][i]) < 1e-7
            assert p["Haus"] == val_dict["HausPerVol"][i] or (np.isnan(p["Haus"]) and np.isnan(val_dict["HausPerVol"][i]))
            assert abs(float(p["AnomalyScoreRecoPerVol"]) - float(val_dict["AnomalyScoreRecoPerVol"][i])) < 1e-7
        flat = np.concatenate([p["diff_filtered"].flatten() for p in ports])
        gflat = np.concatenate([(v["seg_orig"][0, 0].numpy() > 0).flatten() for v in vols])
        _, ptotal = tail_port.find_best_val(flat, gflat, val_range=(0, np.max(flat)), max_steps=10)
        assert ptotal == total, (ptotal, total)
        # ---- test stage with the global threshold
        host.eval_dict = utils_eval.get_eval_dictionary()
        host.stage = "test"
        v = vols[0]
        utils_eval._test_step(host, v["reco"].clone(), v["vol"].clone(), v["seg_orig"].clone(), v["mask_orig"].clone(),
                              0, ["v0"], torch.tensor([1]))
        pt = tail_port.volume_tail(v["reco"][0, 0].numpy(), v["vol"][0, 0].numpy(), v["seg_orig"][0, 0].numpy(),
                                   v["mask_orig"][0, 0].numpy(), stage="test", threshold_total=total)
        assert pt["Dice"] == host.eval_dict["DiceScorePerVol"][0]
        res[tag] = dict(
            val={k: [float(x) for x in vals] for k, vals in val_dict.items()},
            threshold_total=float(total),
            test_dice=float(host.eval_dict["DiceScorePerVol"][0]),
            test_counts=[int(host.eval_dict[k][0]) for k in ("TPPerVol", "FPPerVol", "TNPerVol", "FNPerVol")],
            filtered_sum=[float(p["diff_filtered"].astype(np.float64).sum()) for p in ports],
            masked_sum=[float(p["diff_masked"].astype(np.float64).sum()) for p in ports],
            thresholded_count=[int(p["thresholded"].sum()) for p in ports],
        )
        print(f"tail {tag}: total threshold {total:.6f}, val dice {val_dict['DiceScorePerVol']}, test dice {res[tag]['test_dice']:.4f}")
    with open(os.path.join(GOLD, "tail.json"), "w") as f:
        json.dump(res, f, indent=1)
    # small dense fixtures for the stencils (erosion + median) incl. borders: W=32 -> 1 erosion, W=56 -> 2,
    # W=24 -> W//25 == 0 which scipy reads as "erode until stable" (everything vanishes)
    arrays = {}
    for hw, d in ((32, 10), (56, 6), (24, 7)):
        g = np.random.default_rng(hw)
        vol = g.random((hw, hw, d), dtype=np.float32)
        mask = (g.random((hw, hw, d)) > 0.04).astype(np.float32)
        t = torch.from_numpy(vol.copy())[None, None]
        masked = utils_eval.apply_brainmask_volume(t.clone(), torch.from_numpy(mask)[None, None]).squeeze().numpy()
        filt = utils_eval.apply_3d_median_filter(masked.copy(), kernelsize=5)
        assert np.array_equal(masked, tail_port.apply_brainmask_volume(vol, mask)), hw
        assert np.array_equal(filt, tail_port.median_filter_3d(masked, 5)), hw
        print(f"stencil {hw}: nonzero after erosion {np.count_nonzero(masked)}")
        arrays.update({f"vol{hw}": vol, f"mask{hw}": mask.astype(np.uint8), f"masked{hw}": masked, f"filtered{hw}": filt})
    save("stencil.npz", **arrays)


def golden_test_step():
    """The whole reference DDPM_2D (encoder -> UNet -> diffusion -> test_step -> _test_step) on one synthetic volume.
    Only inputs that cannot be regenerated (none) and the outputs are stored: reco volume + eval_dict scalars."""
    from src.models.DDPM_2D import DDPM_2D

    cfg = base_cfg(noise_ensemble=True)
    model = DDPM_2D(cfg, prefix="t/")
    enc_sd = make_state_dict(resnet_port.param_shapes(128), seed=3)
    unet_sd = make_state_dict(unet_port.param_shapes(unet_port.UNetSpec()), seed=1)
    sched = diffusion_port.schedule_buffers()
    full = {"encoder.encoder." + k: v for k, v in enc_sd.items()}
    full.update({"diffusion." + k: v for k, v in sched.items()})
    full.update({"diffusion.model." + k: v for k, v in unet_sd.items()})
    assert list(model.state_dict().keys()) == list(full.keys()), "DDPM_2D state_dict layout mismatch"
    model.load_state_dict(full, strict=True)
    model.eval()
    v = synthetic_volume(0, depth=8)
    batch = {"Dataset": ["Brats21"], "vol": {"data": v["vol"]}, "vol_orig": {"data": v["vol"].clone()},
             "seg_orig": {"data": v["seg_orig"]}, "mask_orig": {"data": v["mask_orig"]}, "ID": ["v0"],
             "age": torch.tensor([50]), "stage": "val", "label": torch.tensor([1]), "seg_available": True}
    np.random.seed(11)
    captured = {}
    from src.utils import utils_eval as ue
    import src.models.DDPM_2D as mod

    orig_ts = mod._test_step

    def spy(self, final_volume, *a, **k):
        captured["reco"] = final_volume.detach().clone()
        return orig_ts(self, final_volume, *a, **k)

    mod._test_step = spy
    try:
        with torch.no_grad():
            model.on_test_start()
            model.test_step(batch, 0)
    finally:
        mod._test_step = orig_ts
    ed = model.eval_dict
    keys = ("DiceScorePerVol", "BestDicePerVol", "BestThresholdPerVol", "AUCPerVol", "AUPRCPerVol", "l1recoErrorAll",
            "AnomalyScoreRecoPerVol", "AnomalyScoreRegPerVol")
    save("test_step_96.npz", reco=captured["reco"][0, 0], latent=ed["latentSpace"][0],
         **{k: np.asarray([float(x) for x in ed[k]]) for k in keys})
    print({k: [float(x) for x in ed[k]] for k in keys})


def golden_train_step():
    """The reference's training arithmetic end to end (DDPM_2D.py:114-138 with a fixed timestep instead of the random
    draw): encoder in train() mode (batch-statistics BatchNorm) -> gen_noise -> GaussianDiffusion.forward -> L1 loss ->
    loss.backward() on CPU in fp32.  Stored: the loss, the L2 norm of every parameter gradient (649 - 13 buffers - BN
    statistics) and a few small gradients in full."""
    from src.models.DDPM_2D import DDPM_2D
    from src.utils.generate_noise import gen_noise

    cfg = base_cfg()
    model = DDPM_2D(cfg, prefix="t/")
    enc_sd = make_state_dict(resnet_port.param_shapes(128), seed=3)
    unet_sd = make_state_dict(unet_port.param_shapes(unet_port.UNetSpec()), seed=1)
    full = {"encoder.encoder." + k: v for k, v in enc_sd.items()}
    full.update({"diffusion." + k: v for k, v in diffusion_port.schedule_buffers().items()})
    full.update({"diffusion.model." + k: v for k, v in unet_sd.items()})
    model.load_state_dict(full, strict=True)
    model.train()
    x = synthetic_slices(2, 96, seed=21)
    np.random.seed(13)
    features = model(x)
    noise = gen_noise(cfg, x.shape)
    loss, _ = model.diffusion(x, t=300, cond=features, noise=noise)
    loss.backward()
    names, norms = [], []
    for n, p in model.named_parameters():
        names.append(n)
        norms.append(float(p.grad.norm()) if p.grad is not None else -1.0)
    small = {"diffusion.model.out.2.weight", "diffusion.model.time_embed.0.bias", "diffusion.model.label_emb.2.bias",
             "diffusion.model.middle_block.1.qkv.bias", "diffusion.model.input_blocks.0.0.weight",
             "diffusion.model.output_blocks.11.0.in_layers.0.weight", "encoder.encoder.fc.bias"}
    grads = {n.replace(".", "__"): p.grad for n, p in model.named_parameters() if n in small}
    save("train_step_96.npz", loss=loss.detach(), features=features.detach(), grad_norms=np.asarray(norms),
         names=np.asarray(names), **grads)
    print("loss", float(loss), "grad norms", min(norms), max(norms))


def golden_train_step_b4():
    """configs[4] arithmetic at B = 4 (the B = 2 golden leaves 18 samples per channel for layer4's
    batch statistics).  Two runs of the unmodified reference on the same weights, slices and numpy seed:
      * fp32 (the golden proper): loss, features, every parameter-gradient norm, a few small gradients in full;
      * torch.autocast('cpu', bfloat16) - the reference's own arithmetic at the operand width the benchmarked training
        step uses (BASELINE configs[4]: bf16): loss, features, gradient norms.  Its deviation from the fp32 run is the
        envelope the bf16 engines are held to.
    timm's DropPath is stochastic: the encoder stand-in has none, i.e. drop_path_rate = 0 on both sides."""
    from src.models.DDPM_2D import DDPM_2D
    from src.utils.generate_noise import gen_noise

    cfg = base_cfg()
    enc_sd = make_state_dict(resnet_port.param_shapes(128), seed=3)
    unet_sd = make_state_dict(unet_port.param_shapes(unet_port.UNetSpec()), seed=1)
    full = {"encoder.encoder." + k: v for k, v in enc_sd.items()}
    full.update({"diffusion." + k: v for k, v in diffusion_port.schedule_buffers().items()})
    full.update({"diffusion.model." + k: v for k, v in unet_sd.items()})
    x = synthetic_slices(4, 96, seed=22)
    out = {}
    for tag, amp in (("", False), ("_amp", True)):
        model = DDPM_2D(cfg, prefix="t/")
        model.load_state_dict(full, strict=True)
        model.train()
        np.random.seed(14)
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=amp):
            features = model(x)
            noise = gen_noise(cfg, x.shape)
            loss, _ = model.diffusion(x, t=300, cond=features, noise=noise)
        loss.backward()
        names, norms = [], []
        for n, p in model.named_parameters():
            names.append(n)
            norms.append(float(p.grad.float().norm()) if p.grad is not None else -1.0)
        out["loss" + tag] = loss.detach().float()
        out["features" + tag] = features.detach().float()
        out["grad_norms" + tag] = np.asarray(norms)
        if not amp:
            out["names"] = np.asarray(names)
            small = {"diffusion.model.out.2.weight", "diffusion.model.time_embed.0.bias", "diffusion.model.label_emb.2.bias",
                     "diffusion.model.middle_block.1.qkv.bias", "diffusion.model.input_blocks.0.0.weight",
                     "diffusion.model.output_blocks.11.0.in_layers.0.weight", "encoder.encoder.fc.bias",
                     "encoder.encoder.conv1.weight", "encoder.encoder.layer4.2.bn3.weight"}
            out.update({n.replace(".", "__"): p.grad for n, p in model.named_parameters() if n in small})
        print(tag or "fp32", "loss", float(loss), "grad norms", min(norms), max(norms), flush=True)
    a, b = out["grad_norms"], out["grad_norms_amp"]
    rel = np.abs(a - b) / np.maximum(a, 1e-12)
    print("reference bf16-autocast vs fp32: loss rel", abs(float(out["loss_amp"]) - float(out["loss"])) / float(out["loss"]),
          "grad-norm rel max", rel.max(), "median", np.median(rel), "features max-abs",
          float((out["features"] - out["features_amp"]).abs().max()))
    save("train_step_96_b4.npz", **out)


def _full_ddpm2d(cfg):
    """The unmodified reference DDPM_2D with the deterministic synthetic weights (oracle.weights) loaded strictly."""
    from src.models.DDPM_2D import DDPM_2D

    model = DDPM_2D(cfg, prefix="t/")
    unet_sd = make_state_dict(unet_port.param_shapes(unet_port.UNetSpec(
        num_classes=128 if cfg.get("condition", True) else None)), seed=1)
    full = {}
    if cfg.get("condition", True):
        enc_sd = make_state_dict(resnet_port.param_shapes(128), seed=3)
        full.update({"encoder.encoder." + k: v for k, v in enc_sd.items()})
    full.update({"diffusion." + k: v for k, v in diffusion_port.schedule_buffers().items()})
    full.update({"diffusion.model." + k: v for k, v in unet_sd.items()})
    assert list(model.state_dict().keys()) == list(full.keys()), "DDPM_2D state_dict layout mismatch"
    model.load_state_dict(full, strict=True)
    return model.eval()


def fit_denoiser_readout(n_samples=32, ridge=1e-3):
    """Synthetic weights under which the reverse process is WELL-CONDITIONED.  With every tensor random (oracle.weights)
    the x0-prediction is a random function of x_t with a large Lipschitz constant and the reverse loop amplifies any
    per-step rounding difference (measured on the fp32 reference itself: a 1e-3 perturbation of each UNet output grows
    to 2.6e-2 over the last 50 steps) - unlike a trained model, whose x0-prediction is a denoiser.  So the LAST layer
    (out.2: 128 -> 1 channels, 3x3, 1153 numbers) is fitted by ridge least squares to predict x_0 from x_t over the
    random features of the layers below it (32 synthetic slices, random t < 500, simplex noise).  Everything below stays
    the seeded random draw; only these 1153 numbers are stored with the golden.  Returns (weight [1,128,3,3], bias [1])."""
    import torch.nn.functional as F

    from oracle.simplex_port import gen_noise_port

    spec = unet_port.UNetSpec()
    sd = make_state_dict(unet_port.param_shapes(spec), seed=1)
    enc = make_state_dict(resnet_port.param_shapes(128), seed=3)
    sched = diffusion_port.schedule_buffers()
    K = 128 * 9 + 1
    AtA = torch.zeros(K, K, dtype=torch.float64)
    Aty = torch.zeros(K, dtype=torch.float64)
    np.random.seed(101)
    tg = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for b in range(n_samples // 4):
            x = synthetic_slices(4, 96, seed=500 + b)
            cond = resnet_port.resnet_forward(enc, x)
            t = torch.randint(0, 500, (4,), generator=tg)
            xt = diffusion_port.q_sample(sched, x * 2 - 1, t, gen_noise_port((4, 1, 96, 96)).float())
            taps = {}
            unet_port.unet_forward(sd, spec, xt, t, cond, taps=taps)
            phi = F.silu(unet_port._gn(taps["output_blocks.11.0"], sd, "out.0", spec.groups))
            A = F.unfold(phi, 3, padding=1).permute(0, 2, 1).reshape(-1, 128 * 9).double()
            A = torch.cat([A, torch.ones(A.shape[0], 1, dtype=torch.float64)], 1)
            AtA += A.T @ A
            Aty += A.T @ (x * 2 - 1).reshape(-1).double()
    reg = ridge * torch.diag(AtA).mean() * torch.eye(K, dtype=torch.float64)
    reg[-1, -1] = 0
    w = torch.linalg.solve(AtA + reg, Aty)
    return w[:-1].reshape(1, 128, 3, 3).float().contiguous(), w[-1:].float().contiguous()


def golden_reverse_96():
    """BASELINE configs[1] at its OWN geometry (VERDICT r1 'next' #1): the live reference's encoder -> conditioned
    128-channel UNet at 96x96 -> GaussianDiffusion.sample(cond, x_start, start_t=T0, noise=True) (cond_DDPM.py:517-530,
    :446-464) for T0 in {50, 500}, batch 2, on CPU.  The simplex fields come from gen_noise under a fixed
    np.random.seed, so the GPU test regenerates them bit-exactly and nothing but inputs-by-recipe and outputs are stored.
    x_t is stored at every 50th step (p_sample return values) so a deviation can be located.

    Two weight sets: "rand" (every tensor the seeded random draw) and "fit" (same, with the last layer fitted so the
    model is a denoiser, fit_denoiser_readout).  Two precisions of the SAME unmodified reference code: fp32, and
    torch.autocast(fp16) - the precision the reference is configured to run at (configs/trainer/default.yaml:7
    `precision: 16` -> Lightning native AMP).  The autocast run gives the yardstick for an iterated map: how far the
    reference's own production precision moves from its fp32 path over T0 steps."""
    import time

    cfg = base_cfg()
    model = _full_ddpm2d(cfg)
    model.diffusion.use_spatial_transformer = False  # the reference forgets this attribute (cond_DDPM.py:401)
    x = synthetic_slices(2, 96, seed=31)
    out = {}
    W, b = fit_denoiser_readout()
    out["fit_out2_weight"], out["fit_out2_bias"] = W, b
    rand_w = model.diffusion.model.out[2].weight.detach().clone()
    rand_b = model.diffusion.model.out[2].bias.detach().clone()
    with torch.no_grad():
        cond = model(x)
        out["cond"] = cond
        for wtag in ("rand", "fit"):
            model.diffusion.model.out[2].weight.copy_(rand_w if wtag == "rand" else W)
            model.diffusion.model.out[2].bias.copy_(rand_b if wtag == "rand" else b)
            for T0, seed in ((50, 17), (500, 19)):
                for prec in ("fp32", "amp16"):
                    snaps = {}
                    orig = model.diffusion.p_sample

                    def spy(x_, t, *a, _orig=orig, _snaps=snaps, **k):
                        r = _orig(x_, t, *a, **k)
                        if t % 50 == 0 and t > 0:
                            _snaps[t] = r.detach().float().clone()
                        return r

                    model.diffusion.p_sample = spy
                    np.random.seed(seed)
                    t0 = time.time()
                    with torch.autocast("cpu", dtype=torch.float16, enabled=(prec == "amp16")):
                        rec = model.diffusion.sample(cond=cond, x_start=x * 2 - 1, start_t=T0, noise=True).float()
                    model.diffusion.p_sample = orig
                    key = f"{wtag}_{prec}_T{T0}"
                    print(f"reverse_96 {key}: {time.time() - t0:.1f} s, out range [{rec.min():.3f}, {rec.max():.3f}], "
                          f"std {rec.std():.4f}", flush=True)
                    out["out_" + key] = rec
                    ts = sorted(snaps)
                    if ts and prec == "fp32":  # x_t of sample 0, fp32 (the drift is compared at the 1e-4 level)
                        out["snap_t_" + key] = np.asarray(ts)
                        out["snap_x_" + key] = torch.stack([snaps[t][0] for t in ts])
                    if prec == "amp16":
                        ref = out[f"out_{wtag}_fp32_T{T0}"]
                        print(f"   reference amp16 vs its own fp32: max-abs {(rec - ref).abs().max():.4g}", flush=True)
                        if ts:  # the yardstick trajectory: per-snapshot deviation of amp16 from fp32 (sample 0)
                            rs = out[f"snap_x_{wtag}_fp32_T{T0}"]
                            out["amp_drift_" + key] = np.asarray([float((snaps[t][0] - rs[i]).abs().max())
                                                                  for i, t in enumerate(ts)])
                            print("   amp16 drift by t:", dict(zip(ts, out["amp_drift_" + key].round(5))), flush=True)
            out[f"seed_T50"], out[f"seed_T500"] = 17, 19
    save("reverse_96.npz", **out)


class _NoSliceCrop(Cfg):
    """A config object that refuses the fork's hard-coded `cfg['num_eval_slices'] = 4` (DDPM_2D.py:193), so the
    UNMODIFIED reference evaluates the whole volume - the behaviour of the upstream code and of BASELINE configs[2]."""

    def __setitem__(self, k, v):
        if k == "num_eval_slices":
            return
        super().__setitem__(k, v)


def golden_test_step_d50():
    """BASELINE configs[2]: the whole reference DDPM_2D.test_step on ONE full-depth 96x96x50 synthetic volume
    (50 slices x 3-member noise ensemble, then _test_step's residual / erosion / median / thresholds / metrics)."""
    import src.models.DDPM_2D as mod

    cfg = _NoSliceCrop(base_cfg(noise_ensemble=True))
    model = _full_ddpm2d(cfg)
    v = synthetic_volume(2, depth=50)
    batch = {"Dataset": ["Brats21"], "vol": {"data": v["vol"]}, "vol_orig": {"data": v["vol"].clone()},
             "seg_orig": {"data": v["seg_orig"]}, "mask_orig": {"data": v["mask_orig"]}, "ID": ["v2"],
             "age": torch.tensor([50]), "stage": "val", "label": torch.tensor([1]), "seg_available": True}
    np.random.seed(23)
    captured = {}
    orig_ts = mod._test_step

    def spy(self, final_volume, *a, **k):
        captured["reco"] = final_volume.detach().clone()
        return orig_ts(self, final_volume, *a, **k)

    mod._test_step = spy
    try:
        with torch.no_grad():
            model.on_test_start()
            model.test_step(batch, 0)
    finally:
        mod._test_step = orig_ts
    assert captured["reco"].shape[-1] == 50
    ed = model.eval_dict
    keys = ("DiceScorePerVol", "BestDicePerVol", "BestThresholdPerVol", "AUCPerVol", "AUPRCPerVol", "l1recoErrorAll",
            "l2recoErrorAll", "AnomalyScoreRecoPerVol", "AnomalyScoreRegPerVol", "TPPerVol", "FPPerVol", "TNPerVol",
            "FNPerVol", "HausPerVol", "lesionSizePerVol")
    save("test_step_96_d50.npz", reco=captured["reco"][0, 0].half(), latent=ed["latentSpace"][0],
         **{k: np.asarray([float(x) for x in ed[k]]) for k in keys})
    print({k: [float(x) for x in ed[k]] for k in keys})


def golden_uncond_step():
    """BASELINE configs[0]: the unconditioned DDPM_2D (cfg.condition=False -> no encoder, UNet num_classes=None),
    batch 1, single-step reconstruction from t = 499 with a simplex field (DDPM_2D.py:233-239 -> cond_DDPM.py:647-655)."""
    from src.utils.generate_noise import gen_noise

    cfg = base_cfg(condition=False, noise_ensemble=False)
    model = _full_ddpm2d(cfg)
    assert not hasattr(model, "encoder")
    x = synthetic_slices(1, 96, seed=41)
    np.random.seed(29)
    with torch.no_grad():
        feats = model(x)
        assert feats is None
        noise = gen_noise(cfg, x.shape)
        loss, reco = model.diffusion(x, cond=feats, t=cfg.test_timesteps - 1, noise=noise)
    print(f"uncond step: loss {float(loss):.6f}, reco range [{reco.min():.3f}, {reco.max():.3f}]")
    save("uncond_step_96.npz", reco=reco, loss=loss, seed=29, n_state=len(model.state_dict()))



def golden_ddim():
    """GaussianDiffusion.ddim_sample (cond_DDPM.py:466-515; reached through sample() when sampling_timesteps < timesteps)
    and p_sample with clip_denoised=False, on the small geometry.  Two DDIM cases: start_t = 300 with the simplex flag
    (fully determined by np.random.seed: the flag True itself is q_sample's noise, the per-step fields are gen_noise
    draws) and start_t = 0 from a stored Gaussian x_T."""
    from src.models.modules import cond_DDPM
    from src.models.modules.cond_DDPM import GaussianDiffusion

    from oracle.simplex_port import gen_noise_port

    spec = unet_port.UNetSpec(model_channels=64, channel_mult=(1, 2), num_res_blocks=1, num_classes=128)
    m, shapes = ref_unet(spec, image=32)
    sd = make_state_dict(shapes, seed=2)
    m.load_state_dict(sd, strict=True)
    sched = diffusion_port.schedule_buffers()
    model = lambda x, t, c: unet_port.unet_forward(sd, spec, x, t, c)  # noqa: E731
    g = torch.Generator().manual_seed(33)
    img = torch.rand(2, 1, 32, 32, generator=g)
    cond = torch.randn(2, 128, generator=g)
    out = {"img": img, "cond": cond}
    for objective in ("pred_x0", "pred_noise"):
        d = GaussianDiffusion(m, image_size=(32, 32), timesteps=1000, sampling_timesteps=6, objective=objective,
                              channels=1, loss_type="l1", p2_loss_weight_gamma=0, ddim_sampling_eta=0.7, cfg=base_cfg())
        d.use_spatial_transformer = False
        assert d.is_ddim_sampling
        with torch.no_grad():
            np.random.seed(5)
            rec = d.sample(cond=cond, x_start=img * 2 - 1, start_t=300, noise=True)
            np.random.seed(5)
            prec = diffusion_port.ddim_sample(model, sched, (2, 1, 32, 32), img * 2 - 1, cond, 300, 6, 0.7, True,
                                              lambda: gen_noise_port((2, 1, 32, 32)), None, objective=objective)
        assert (rec - prec).abs().max().item() < 1e-4, (rec - prec).abs().max().item()
        out[f"ddim_simplex_{objective}"] = rec
        # yardstick for an iterated map: the same reference code under its configured precision (fp16 autocast)
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.float16):
            np.random.seed(5)
            amp = d.sample(cond=cond, x_start=img * 2 - 1, start_t=300, noise=True).float()
        out[f"ddim_simplex_{objective}_amp16_dev"] = float((amp - rec).abs().max())
        print(f"ddim {objective}: reference amp16 vs its own fp32 max-abs {out[f'ddim_simplex_{objective}_amp16_dev']:.4g}")
    # start_t = 0: x_T is the SECOND torch.randn draw (the first is discarded); per-step noise is simplex (cfg.noisetype)
    d = GaussianDiffusion(m, image_size=(32, 32), timesteps=1000, sampling_timesteps=4, objective="pred_x0", channels=1,
                          loss_type="l1", p2_loss_weight_gamma=0, ddim_sampling_eta=1.0, cfg=base_cfg())
    d.use_spatial_transformer = False
    with torch.no_grad():
        torch.manual_seed(77)
        torch.randn(2, 1, 32, 32)
        x_T = torch.randn(2, 1, 32, 32)
        torch.manual_seed(77)
        np.random.seed(6)
        rec = d.sample(cond=cond, x_start=img * 2 - 1, start_t=0, noise=None)
        draws = [torch.zeros(2, 1, 32, 32), x_T]
        np.random.seed(6)
        prec = diffusion_port.ddim_sample(model, sched, (2, 1, 32, 32), img * 2 - 1, cond, 0, 4, 1.0, None,
                                          lambda: gen_noise_port((2, 1, 32, 32)), lambda: draws.pop(0))
    assert (rec - prec).abs().max().item() < 1e-4, (rec - prec).abs().max().item()
    out["ddim_xT"] = x_T
    out["ddim_gauss_start"] = rec
    # p_sample without clipping (clip_denoised=False, cond_DDPM.py:422-444), one step with Gaussian noise handed in
    d = GaussianDiffusion(m, image_size=(32, 32), timesteps=1000, sampling_timesteps=1000, objective="pred_x0",
                          channels=1, loss_type="l1", p2_loss_weight_gamma=0, cfg=base_cfg())
    d.use_spatial_transformer = False
    with torch.no_grad():
        x = 1.5 * torch.randn(2, 1, 32, 32, generator=g)
        torch.manual_seed(78)
        nz = torch.randn(2, 1, 32, 32)
        torch.manual_seed(78)
        out["noclip_x"] = x
        out["noclip_noise"] = nz
        out["noclip_out"] = d.p_sample(x.clone(), 400, clip_denoised=False, cond=cond, noise=None)
        torch.manual_seed(78)
        out["clip_out"] = d.p_sample(x.clone(), 400, clip_denoised=True, cond=cond, noise=None)
        assert (out["noclip_out"] - out["clip_out"]).abs().max().item() > 1e-3  # the clamp is active on this input
    save("ddim_small_32.npz", **out)



def golden_tail_fullres():
    """utils_eval._test_step with cfg.resizedEvaluation=False (:24-25): the [96,96,50] reconstruction is resized to
    new_size = [160,190,160] with F.interpolate(trilinear, align_corners=True) and scored against full-resolution
    originals.  Stored: probes and the float64 sum of the resized volume, the eval_dict scalars."""
    import torch.nn.functional as F

    from src.utils import utils_eval

    from oracle.weights import synthetic_fullres_case

    class Host:
        pass

    host = Host()
    host.cfg = base_cfg(resizedEvaluation=False)
    host.eval_dict = utils_eval.get_eval_dictionary()
    host.threshold = {}
    host.dataset = ["Brats21"]
    host.new_size = [160, 190, 160]
    host.diffs_list, host.seg_list = [], []
    host.stage = "val"
    c = synthetic_fullres_case(4)
    utils_eval._test_step(host, c["reco"].clone(), c["vol_orig"].clone(), c["seg_orig"].clone(), c["mask_orig"].clone(),
                          0, ["v4"], torch.tensor([1]))
    resized = F.interpolate(c["reco"], size=host.new_size, mode="trilinear", align_corners=True)[0, 0]
    probes = [(0, 0, 0), (159, 189, 159), (80, 95, 80), (17, 101, 33), (123, 7, 150), (64, 64, 64)]
    keys = ("DiceScorePerVol", "BestDicePerVol", "BestThresholdPerVol", "AUCPerVol", "AUPRCPerVol", "l1recoErrorAll",
            "l2recoErrorAll", "l1recoErrorUnhealthy", "l1recoErrorHealthy", "TPPerVol", "FPPerVol", "TNPerVol",
            "FNPerVol", "HausPerVol", "lesionSizePerVol", "AnomalyScoreRecoPerVol")
    save("tail_fullres.npz", probes=np.asarray(probes), probe_values=np.asarray([float(resized[p]) for p in probes]),
         resized_sum=float(resized.double().sum()), resized_slice=resized[:, :, 80],
         **{k: np.asarray([float(x) for x in host.eval_dict[k]]) for k in keys})
    print({k: [float(x) for x in host.eval_dict[k]] for k in keys})



def golden_vol2slice():
    """create_dataset.vol2slice (:143-193) of the live reference on fake subjects (objects with a `.data` tensor): the
    slice index it draws (torch.randint under a fixed seed) and a checksum of the slices it returns, for the index rules
    the configs use (random slice, fixed start slice, sequential window, onlyBrain, unique_slice)."""
    from src.datamodules.create_dataset import vol2slice

    class Img:
        def __init__(self, t):
            self.data = t

    def make_ds(n=6, depth=20):
        g = torch.Generator().manual_seed(77)
        ds = []
        for i in range(n):
            vol = torch.rand(1, 8, 8, depth, generator=g)
            mask = torch.zeros(1, 8, 8, depth)
            mask[..., 3 + i % 3:15 - i % 2] = 1.0
            ds.append({"vol": Img(vol), "mask": Img(mask)})
        return ds

    cases = {"random": dict(), "start": dict(slice=7), "window": dict(slice=5, seq_slices=6), "brain": dict(onlyBrain=True),
             "unique": dict(cfg=dict(unique_slice=True, batch_size=3))}
    out = {}
    for name, kw in cases.items():
        kw = dict(kw)
        cfg = Cfg(kw.pop("cfg", {}))
        torch.manual_seed(123)
        v2s = vol2slice(make_ds(), cfg, **kw)
        rec = []
        for i in range(len(v2s)):
            s = v2s[i]
            rec.append([int(s["ind"]) if not isinstance(s["ind"], int) else s["ind"], float(s["vol"].data.double().sum()),
                        float(s["mask"].data.double().sum()), list(s["vol"].data.shape)])
        out[name] = rec
    with open(os.path.join(GOLD, "vol2slice.json"), "w") as f:
        json.dump(out, f)
    print(out)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["schedule", "simplex", "unet", "encoder", "diffusion", "tail", "test_step", "train_step", "train_step_b4", "reverse_96",
                             "test_step_d50", "uncond_step", "ddim", "tail_fullres", "vol2slice"]
    torch.manual_seed(0)
    for w in which:
        print(f"== {w}")
        globals()["golden_" + w]()
