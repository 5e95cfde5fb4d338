"""BASELINE.json's configs at THEIR OWN geometry against the LIVE reference (VERDICT r1 "next" #1).

Golden outputs come from the unmodified reference run on CPU in fp32 (oracle/make_golden.py: reverse_96, test_step_d50,
uncond_step); the simplex fields are regenerated here from the same np.random.seed (the GPU generator is bit-exact,
tests/test_diffusion_gpu.py), so no noise is stored.

* configs[1]  encoder -> conditioned 128-channel UNet, 96x96 -> GaussianDiffusion.sample(start_t = T0, noise=True)
              for T0 = 50 and T0 = 500 (cond_DDPM.py:517-530, :446-464); x_t snapshots every 50 steps locate any drift;
              tolerance = the envelope of the reference's own fp16-autocast path (see the test's docstring).
* configs[2]  DDPM_2D.test_step on one FULL-depth 96x96x50 volume (DDPM_2D.py:171-286 without the fork's 4-slice crop).
* configs[0]  DDPM_2D(condition=False), batch 1, single-step reconstruction from t = 499.

Tolerance (north_star): reconstruction max-abs <= 1e-2; Dice / AUPRC / AUC within 1e-3; counts and thresholds as stated
per assertion."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

pytestmark = pytest.mark.gpu

TOL = 1e-2  # north_star: reconstructions within max-abs 1e-2 of the reference's fp32 path


class Cfg(dict):
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


def _cfg(**over):
    c = Cfg(imageDim=[192, 192, 100], rescaleFactor=2, unet_dim=128, dim_mults=[1, 2, 2], condition=True,
            backbone="Spark_Encoder_2D", version="resnet50", cond_dim=128, noisetype="simplex", noise_ensemble=True,
            test_timesteps=500, lr=1e-4, resizedEvaluation=True, erodeBrainmask=True, medianFiltering=True,
            saveOutputImages=False, evalSeg=True, threshold="auto", spatial_transformer=False, pretrained_encoder=False)
    c.update(over)
    return c


def _state_dict(condition=True):
    from oracle import diffusion_port, resnet_port, unet_port
    from oracle.weights import make_state_dict

    full = {}
    if condition:
        enc = make_state_dict(resnet_port.param_shapes(128), seed=3)
        full.update({"encoder.encoder." + k: v for k, v in enc.items()})
    unet = make_state_dict(unet_port.param_shapes(unet_port.UNetSpec(num_classes=128 if condition else None)), seed=1)
    full.update({"diffusion." + k: v for k, v in diffusion_port.schedule_buffers().items()})
    full.update({"diffusion.model." + k: v for k, v in unet.items()})
    return full


def _model(cfg, condition=True):
    from src.models.DDPM_2D import DDPM_2D

    full = _state_dict(condition)
    m = DDPM_2D(cfg, prefix="t/")
    assert list(m.state_dict().keys()) == list(full.keys())
    m.load_state_dict(full, strict=True)
    return m.cuda().eval()


@pytest.fixture(scope="module")
def cond_model():
    return _model(_cfg())


@pytest.mark.parametrize("weights", ["rand", "fit"])
@pytest.mark.parametrize("T0", [50, 500])
def test_reverse_loop_headline_geometry_vs_reference(cond_model, T0, weights):
    """configs[1] at its own geometry.  The reverse loop is an ITERATED map: T0 UNet evaluations feed each other, so a
    per-step rounding difference is amplified by the model's own sensitivity.  The golden therefore holds
    the unmodified reference twice - in fp32 and under torch.autocast(fp16), the precision it is configured to run at
    (configs/trainer/default.yaml:7 `precision: 16`) - and the criterion is: our deviation from the reference's fp32
    path stays within the envelope of the reference's OWN fp16-autocast deviation (x1.5 in the maximum and in the mean;
    never asked to be below the single-step tolerance 1e-2), at the output and (x2) at every stored x_t.  Two weight sets: "rand" (all tensors random: a
    chaotic map, the envelope itself reaches 0.45 at T0 = 500) and "fit" (last layer fitted so the UNet is a denoiser,
    oracle/make_golden.py fit_denoiser_readout: the well-conditioned case a trained model is)."""
    from oracle.weights import synthetic_slices

    g = np.load(os.path.join(GOLD, "reverse_96.npz"))
    x = synthetic_slices(2, 96, seed=31).cuda()
    d = cond_model.diffusion
    out2 = d.model.out[2]
    saved = (out2.weight.detach().clone(), out2.bias.detach().clone())
    if weights == "fit":
        with torch.no_grad():
            out2.weight.copy_(torch.from_numpy(g["fit_out2_weight"]).cuda())
            out2.bias.copy_(torch.from_numpy(g["fit_out2_bias"]).cuda())
    snaps = {}
    orig = d.p_sample

    def spy(x_, t, *a, **k):
        k.pop("_out", None)  # fresh tensors so the snapshots survive the loop's ping-pong buffers
        r = orig(x_, t, *a, **k)
        if t % 50 == 0 and t > 0:
            snaps[t] = r.detach().clone()
        return r

    key = f"{weights}_fp32_T{T0}"
    ref = torch.from_numpy(g["out_" + key])
    amp_d = (torch.from_numpy(g[f"out_{weights}_amp16_T{T0}"]) - ref).abs()
    amp, amp_mean = amp_d.max().item(), amp_d.mean().item()
    try:
        with torch.no_grad():
            cond = cond_model(x)
            cerr = (cond.cpu() - torch.from_numpy(g["cond"])).abs().max().item()
            np.random.seed(int(g[f"seed_T{T0}"]))
            d.p_sample = spy
            try:
                rec = d.sample(cond=cond, x_start=x * 2 - 1, start_t=T0, noise=True)
            finally:
                del d.p_sample
    finally:
        with torch.no_grad():
            out2.weight.copy_(saved[0])
            out2.bias.copy_(saved[1])
    err = (rec.cpu() - ref).abs().max().item()
    mean_err = (rec.cpu() - ref).abs().mean().item()
    line = (f"reverse loop 96x96 T0={T0} [{weights}]: max-abs {err:.4g} (mean {mean_err:.3g}) vs reference fp32; the "
            f"reference's own fp16-autocast path: max-abs {amp:.4g} (mean {amp_mean:.3g}); encoder cond err {cerr:.3g}")
    # envelope: within 1.5x of the reference's own mixed-precision deviation, in the maximum and in the mean.  Once the
    # envelope itself is O(1) (random weights at T0 = 500: single pixels flip between the clamp limits) the maximum of
    # one pixel carries no information any more and is allowed 2x; the mean and the x_t trajectory stay at their bounds.
    ok = err <= max(TOL, (2.0 if amp > 0.25 else 1.5) * amp) and mean_err <= max(1e-3, 1.5 * amp_mean)
    if T0 == 500:
        ts = [int(t) for t in g["snap_t_" + key]]
        sx = torch.from_numpy(g["snap_x_" + key])
        ad = {int(t): float(a) for t, a in zip(ts, g[f"amp_drift_{weights}_amp16_T{T0}"])}
        drift = {t: (snaps[t][0].cpu() - sx[i]).abs().max().item() for i, t in enumerate(ts)}
        line += "\n   x_t deviation (ours | reference fp16-autocast) by t: " + ", ".join(
            f"{t}: {drift[t]:.3g} | {ad[t]:.3g}" for t in sorted(drift, reverse=True))
        for t in ts:
            ok = ok and drift[t] <= max(5e-3, 2.0 * ad[t])
    print(line)
    assert torch.isfinite(rec).all()
    assert cerr <= 5e-3
    assert ok, line


def test_full_depth_test_step_vs_reference():
    from oracle.weights import synthetic_volume

    cfg = _cfg(noise_ensemble=True)
    cfg["force_num_eval_slices"] = False  # BASELINE configs[2]: the whole volume, not the fork's 4 centre slices
    model = _model(cfg)
    v = synthetic_volume(2, depth=50)
    batch = {"Dataset": ["Brats21"], "vol": {"data": v["vol"]}, "vol_orig": {"data": v["vol"].clone()},
             "seg_orig": {"data": v["seg_orig"]}, "mask_orig": {"data": v["mask_orig"]}, "ID": ["v2"],
             "age": torch.tensor([50]), "stage": "val", "label": torch.tensor([1]), "seg_available": True}
    g = np.load(os.path.join(GOLD, "test_step_96_d50.npz"))
    np.random.seed(23)
    model.on_test_start()
    final = model.test_step(batch, 0)
    reco = final[0, 0].float().cpu()
    ref = torch.from_numpy(g["reco"]).float()  # stored as fp16: +2.5e-4 of representation error on values <= 1
    assert tuple(reco.shape) == tuple(ref.shape) == (96, 96, 50)
    err = (reco - ref).abs().max().item()
    ed = model.eval_dict
    print(f"test_step D=50: reco max-abs {err:.4g}; "
          + "; ".join(f"{k[:-6]} {float(ed[k][0]):.6g} vs {float(g[k][0]):.6g}"
                      for k in ("DiceScorePerVol", "AUPRCPerVol", "AUCPerVol", "BestThresholdPerVol", "TPPerVol",
                                "FPPerVol", "FNPerVol", "HausPerVol")))
    assert err <= TOL + 1e-3  # fp16 storage of the golden can add its half-ulp (4.9e-4 at 1.0 <= |x| < 2)
    for k in ("DiceScorePerVol", "BestDicePerVol", "AUCPerVol", "AUPRCPerVol", "l1recoErrorAll", "l2recoErrorAll",
              "AnomalyScoreRecoPerVol", "AnomalyScoreRegPerVol"):
        assert abs(float(ed[k][0]) - float(g[k][0])) <= 1e-3, (k, ed[k][0], g[k][0])
    assert abs(float(ed["BestThresholdPerVol"][0]) - float(g["BestThresholdPerVol"][0])) <= 2e-3
    assert float(ed["lesionSizePerVol"][0]) == float(g["lesionSizePerVol"][0])  # ground truth only: exact
    # confusion counts follow the thresholded reconstruction: within 0.1 % of the 460 800 voxels
    for k in ("TPPerVol", "FPPerVol", "TNPerVol", "FNPerVol"):
        assert abs(float(ed[k][0]) - float(g[k][0])) <= 461, (k, ed[k][0], g[k][0])


def test_unconditioned_single_step_vs_reference():
    from cddpm.noise import gen_noise
    from oracle.weights import synthetic_slices

    cfg = _cfg(condition=False, noise_ensemble=False)
    model = _model(cfg, condition=False)
    assert not hasattr(model, "encoder")
    g = np.load(os.path.join(GOLD, "uncond_step_96.npz"))
    assert len(model.state_dict()) == int(g["n_state"])
    x = synthetic_slices(1, 96, seed=41).cuda()
    np.random.seed(int(g["seed"]))
    with torch.no_grad():
        feats = model(x)
        assert feats is None
        noise = gen_noise(cfg, x.shape)
        loss, reco = model.diffusion(x, cond=feats, t=cfg.test_timesteps - 1, noise=noise)
    err = (reco.cpu() - torch.from_numpy(g["reco"])).abs().max().item()
    lerr = abs(float(loss) - float(g["loss"]))
    print(f"unconditioned B=1 t=499: reco max-abs {err:.4g}, loss err {lerr:.3g}")
    assert err <= TOL and lerr <= 1e-3
