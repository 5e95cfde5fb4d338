"""GPU parity of the condition encoder and of the complete DDPM_2D.test_step (encoder -> simplex noise -> q_sample ->
UNet -> ensemble mean -> residual -> eroded mask -> median -> threshold search -> metrics) against golden outputs of the
live reference (tests/golden/encoder_96.npz, test_step_96.npz; oracle/make_golden.py).

Reference: src/models/DDPM_2D.py:171-286 (test_step), src/models/modules/spark/Spark_2D.py:285-290 (encoder).
Tolerances (north_star): reconstruction max-abs <= 1e-2; Dice / AUPRC / AUC within 1e-3."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

pytestmark = pytest.mark.gpu


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


def _full_state_dict():
    from oracle import diffusion_port, resnet_port, unet_port
    from oracle.weights import make_state_dict

    enc = make_state_dict(resnet_port.param_shapes(128), seed=3)
    unet = make_state_dict(unet_port.param_shapes(unet_port.UNetSpec()), seed=1)
    full = {"encoder.encoder." + k: v for k, v in enc.items()}
    full.update({"diffusion." + k: v for k, v in diffusion_port.schedule_buffers().items()})
    full.update({"diffusion.model." + k: v for k, v in unet.items()})
    return full, enc


def test_encoder_matches_reference_golden():
    from cddpm.encoder import get_encoder

    _, enc_sd = _full_state_dict()
    enc, feat = get_encoder(_cfg())
    assert feat == 128
    enc.load_state_dict({"encoder." + k: v for k, v in enc_sd.items()}, strict=True)
    enc = enc.cuda().eval()
    g = np.load(os.path.join(GOLD, "encoder_96.npz"))
    with torch.no_grad():
        c = enc(torch.from_numpy(g["x"]).cuda()).cpu()
    err = (c - torch.from_numpy(g["c"])).abs().max().item()
    print(f"encoder: max-abs {err:.4g} (ref max {np.abs(g['c']).max():.3g})")
    assert err <= 2e-2
    # batch of one goes through a different (re-planned) tile geometry
    with torch.no_grad():
        c1 = enc(torch.from_numpy(g["x"][:1]).cuda()).cpu()
    assert (c1 - c[:1]).abs().max().item() <= 1e-3


def test_full_test_step_matches_reference_golden():
    from oracle.weights import synthetic_volume
    from src.models.DDPM_2D import DDPM_2D  # the drop-in overlay path Hydra would resolve

    full, _ = _full_state_dict()
    model = DDPM_2D(_cfg(), prefix="t/")
    assert list(model.state_dict().keys()) == list(full.keys())
    model.load_state_dict(full, strict=True)
    model = model.cuda().eval()
    # state_dict round-trips bit-exactly (fp32 masters are untouched by the engine's 16-bit re-layout)
    for k, v in model.state_dict().items():
        assert torch.equal(v.cpu(), full[k]), k
    v = synthetic_volume(0, depth=8)
    batch = {"Dataset": ["Brats21"], "vol": {"data": v["vol"]}, "vol_orig": {"data": v["vol"].clone()},
             "seg_orig": {"data": v["seg_orig"]}, "mask_orig": {"data": v["mask_orig"]}, "ID": ["v0"],
             "age": torch.tensor([50]), "stage": "val", "label": torch.tensor([1]), "seg_available": True}
    g = np.load(os.path.join(GOLD, "test_step_96.npz"))
    np.random.seed(11)
    model.on_test_start()
    final = model.test_step(batch, 0)
    reco = final[0, 0].cpu()
    ref = torch.from_numpy(g["reco"])
    assert tuple(reco.shape) == tuple(ref.shape) == (96, 96, 4)  # the fork evaluates the 4 centre slices
    err = (reco - ref).abs().max().item()
    ed = model.eval_dict
    lat_err = (ed["latentSpace"][0] - torch.from_numpy(g["latent"])).abs().max().item()
    print(f"test_step: reco max-abs {err:.4g}; latent err {lat_err:.4g}; "
          f"Dice {ed['DiceScorePerVol'][0]:.5f} vs {g['DiceScorePerVol'][0]:.5f}; "
          f"AUPRC {ed['AUPRCPerVol'][0]:.5f} vs {g['AUPRCPerVol'][0]:.5f}; AUC {ed['AUCPerVol'][0]:.5f} vs {g['AUCPerVol'][0]:.5f}")
    assert err <= 1e-2
    assert lat_err <= 2e-2
    for k in ("DiceScorePerVol", "BestDicePerVol", "AUCPerVol", "AUPRCPerVol", "l1recoErrorAll",
              "AnomalyScoreRecoPerVol", "AnomalyScoreRegPerVol"):
        assert abs(float(ed[k][0]) - float(g[k][0])) <= 1e-3, (k, ed[k][0], g[k][0])
    assert abs(float(ed["BestThresholdPerVol"][0]) - float(g["BestThresholdPerVol"][0])) <= 2e-3
    model.on_test_end()
    assert "total" in model.threshold


def test_cpu_module_fails_loudly():
    from cddpm import CddpmError
    from src.models.DDPM_2D import DDPM_2D

    model = DDPM_2D(_cfg(), prefix="t/").eval()
    with pytest.raises(CddpmError):
        with torch.no_grad():
            model(torch.zeros(1, 1, 96, 96))


def test_sweep_driver_val_then_test():
    """train.py:182-237 through cddpm.sweep.test_sweep: the validation stage finds threshold['total'], the test stage
    consumes it; eval_dict keeps the reference's keys; preds_dict.pkl is written."""
    import pickle
    import tempfile

    from cddpm import sweep
    from oracle.weights import synthetic_volume
    from src.models.DDPM_2D import DDPM_2D

    full, _ = _full_state_dict()
    model = DDPM_2D(_cfg(), prefix="t/")
    model.load_state_dict(full, strict=True)
    model = model.cuda().eval()

    def loader(stage, seeds):
        out = []
        for sd in seeds:
            v = synthetic_volume(sd, depth=8)
            out.append({"Dataset": ["Brats21"], "vol": {"data": v["vol"]}, "vol_orig": {"data": v["vol"].clone()},
                        "seg_orig": {"data": v["seg_orig"]}, "mask_orig": {"data": v["mask_orig"]}, "ID": [f"{stage}{sd}"],
                        "age": torch.tensor([50]), "stage": stage, "label": torch.tensor([1]), "seg_available": True})
        return out

    np.random.seed(3)
    with tempfile.TemporaryDirectory() as d:
        preds, logs = sweep.test_sweep(model, {"Datamodules_eval.Brats21": (loader("val", [0, 1]), loader("test", [2, 3]))},
                                       fold=0, log_dir=d)
        with open(os.path.join(d, "1_preds_dict.pkl"), "rb") as f:
            assert set(pickle.load(f)) == {"val", "test"}
    val, test = preds["val"]["Datamodules_eval.Brats21"], preds["test"]["Datamodules_eval.Brats21"]
    assert len(val["DiceScorePerVol"]) == 2 and len(test["DiceScorePerVol"]) == 2
    assert np.isfinite(test["DicePerVolMean"]) and np.isfinite(val["AUPRCPerVolMean"])
    assert "1/Datamodules_eval.Brats21/test/DicePerVolMean" in logs
    assert not hasattr(model, "threshold")  # deleted after the test stage, as in the reference (utils_eval.py:258-259)


def test_pipelined_sweep_equals_plain_loop(monkeypatch):
    """run_stage enqueues volume i+1 before it scores volume i (side stream for the tail): same eval_dict, same global
    threshold as the one-volume-at-a-time loop, bit for bit."""
    from cddpm import sweep
    from oracle.weights import synthetic_volume
    from src.models.DDPM_2D import DDPM_2D

    full, _ = _full_state_dict()
    cfg = _cfg()
    cfg["force_num_eval_slices"] = False
    model = DDPM_2D(cfg, prefix="t/")
    model.load_state_dict(full, strict=True)
    model = model.cuda().eval()
    batches = []
    for sd in range(5):
        v = synthetic_volume(sd, depth=6)
        batches.append({"Dataset": ["Brats21"], "vol": {"data": v["vol"]}, "vol_orig": {"data": v["vol"].clone()},
                        "seg_orig": {"data": v["seg_orig"]}, "mask_orig": {"data": v["mask_orig"]}, "ID": [f"v{sd}"],
                        "stage": "val", "label": torch.tensor([1]), "seg_available": True})
    keys = ("DiceScorePerVol", "BestThresholdPerVol", "AUPRCPerVol", "HausPerVol", "TPPerVol", "FNPerVol",
            "AnomalyScoreRegPerVol", "AnomalyScoreRecoPerVol", "DiceScorePerSlice", "IDs")
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("CDDPM_SWEEP_PIPELINE", mode)
        np.random.seed(11)
        ed = sweep.run_stage(model, batches)
        res[mode] = ({k: [float(x) if not isinstance(x, str) else x for x in ed[k]] for k in keys},
                     float(model.threshold["total"]))
    assert res["0"][1] == res["1"][1]
    for k in keys:
        a, b = res["0"][0][k], res["1"][0][k]
        assert len(a) == len(b) and all(x == y or (x != x and y != y) for x, y in zip(a, b)), k


def test_stacked_ensemble_equals_member_loop():
    """reconstruct_slices with the noise ensemble as ONE UNet forward over the 3 x D stacked slices
    (GaussianDiffusion.ensemble_reconstruct) against the reference's loop of three forwards (DDPM_2D.py:214-232;
    cfg stack_ensemble=False): same noise draws, every slice an independent sample, so only the summation order of the
    GroupNorm partials may move - reconstruction within 2e-3, the last member's loss within 1e-4."""
    from src.models.DDPM_2D import DDPM_2D

    full, _ = _full_state_dict()
    model = DDPM_2D(_cfg(), prefix="t/")
    model.load_state_dict(full, strict=True)
    model = model.cuda().eval()
    g = torch.Generator().manual_seed(5)
    x = torch.rand(6, 1, 96, 96, generator=g).cuda()
    out = {}
    with torch.no_grad():
        for stacked in (True, False):
            model.cfg["stack_ensemble"] = stacked
            np.random.seed(21)
            reco, loss, feats = model.reconstruct_slices(x)
            out[stacked] = (reco.cpu(), float(loss), feats.cpu())
    assert torch.equal(out[True][2], out[False][2])
    err = (out[True][0] - out[False][0]).abs().max().item()
    print(f"stacked vs loop: reco max-abs {err:.3g}, loss {out[True][1]:.6f} vs {out[False][1]:.6f}")
    assert err <= 2e-3
    assert abs(out[True][1] - out[False][1]) <= 1e-4
