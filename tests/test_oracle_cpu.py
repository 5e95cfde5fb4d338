"""CPU tests (no GPU): the oracle restatements against the golden vectors generated from the live reference
(oracle/make_golden.py), and against scipy / sklearn where the reference calls those libraries directly."""
import json
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_schedule_port_bit_exact():
    from oracle import diffusion_port

    g = np.load(os.path.join(GOLD, "schedule.npz"))
    s = diffusion_port.schedule_buffers()
    assert list(s.keys()) == diffusion_port.BUFFER_NAMES == list(g.keys())
    for k in s:
        assert np.array_equal(s[k].numpy(), g[k]), k
    # survey probe values
    assert abs(float(s["sqrt_alphas_cumprod"][499]) - 0.70274) < 1e-5


def test_simplex_port_bit_exact():
    from oracle.simplex_port import gen_noise_port

    g = np.load(os.path.join(GOLD, "simplex.npz"))
    for seed in (0, 7):
        np.random.seed(seed)
        n = gen_noise_port((2, 1, 96, 96))
        assert n.dtype == torch.float16
        assert np.array_equal(n[0, 0].numpy(), g[f"field_seed{seed}"])
        assert torch.equal(n[0], n[1])  # one field shared by the whole batch


def test_unet_port_small_matches_reference():
    from oracle import unet_port
    from oracle.weights import make_state_dict

    spec = unet_port.UNetSpec(model_channels=64, channel_mult=(1, 2), num_res_blocks=1, num_classes=128)
    sd = make_state_dict(unet_port.param_shapes(spec), seed=2)
    g = np.load(os.path.join(GOLD, "unet_small_32.npz"))
    with torch.no_grad():
        y = unet_port.unet_forward(sd, spec, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]), torch.from_numpy(g["cond"]))
    assert (y - torch.from_numpy(g["y"])).abs().max().item() < 1e-4


def test_unet_key_layout_matches_reference():
    from oracle import unet_port

    keys = json.load(open(os.path.join(GOLD, "unet_cond_keys.json")))
    mine = unet_port.param_shapes(unet_port.UNetSpec())
    assert [[k, list(s)] for k, s in mine] == keys
    assert len(mine) == 316 and sum(int(np.prod(s)) for _, s in mine) == 43871873


def test_diffusion_port_matches_reference():
    from oracle import diffusion_port, unet_port
    from oracle.weights import make_state_dict

    spec = unet_port.UNetSpec(model_channels=64, channel_mult=(1, 2), num_res_blocks=1, num_classes=128)
    sd = make_state_dict(unet_port.param_shapes(spec), seed=2)
    g = np.load(os.path.join(GOLD, "diffusion_small_32.npz"))
    img, cond, noise = (torch.from_numpy(g[k]) for k in ("img", "cond", "noise"))
    model = lambda x, t, c: unet_port.unet_forward(sd, spec, x, t, c)  # noqa: E731
    sched = diffusion_port.schedule_buffers()
    with torch.no_grad():
        for obj in ("pred_x0", "pred_noise"):
            loss, reco = diffusion_port.reconstruct(model, sched, img, 499, cond, noise, objective=obj)
            assert (reco - torch.from_numpy(g[f"reco_{obj}"])).abs().max().item() < 1e-4
            assert abs(loss.item() - float(g[f"loss_{obj}"])) < 1e-5
        feed = list(torch.from_numpy(g["reverse_noises"]))
        rec = diffusion_port.reverse_loop(model, sched, img * 2 - 1, cond, int(g["reverse_T0"]), lambda: feed.pop(0))
    assert (rec - torch.from_numpy(g["reverse_out"])).abs().max().item() < 1e-4


def test_stencil_ports_vs_golden_and_scipy():
    from scipy import ndimage

    from oracle import tail_port

    g = np.load(os.path.join(GOLD, "stencil.npz"))
    for hw in (32, 56, 24):
        vol, mask = g[f"vol{hw}"], g[f"mask{hw}"].astype(np.float32)
        masked = tail_port.apply_brainmask_volume(vol, mask)
        assert np.array_equal(masked, g[f"masked{hw}"])
        assert np.array_equal(tail_port.median_filter_3d(masked, 5), g[f"filtered{hw}"])
    # the equivalences the restatement relies on, directly against scipy (the reference's own callee)
    rng = np.random.default_rng(0)
    m = rng.random((40, 40)) > 0.1
    assert np.array_equal(tail_port.erode_cross(m, 3),
                          ndimage.binary_erosion(m, structure=ndimage.generate_binary_structure(2, 1), iterations=3))
    v = rng.random((20, 18, 7)).astype(np.float32)
    assert np.array_equal(tail_port.median_filter_3d(v, 5), ndimage.median_filter(v, (5, 5, 5)))
    assert np.array_equal(tail_port.median_filter_3d(v, 3), ndimage.median_filter(v, (3, 3, 3)))


def test_ranking_ports_vs_sklearn():
    from sklearn.metrics import auc, average_precision_score, roc_curve

    from oracle import tail_port

    rng = np.random.default_rng(1)
    x = rng.random(5000).astype(np.float32)
    x[rng.random(5000) < 0.6] = 0.0  # heavy ties, like a masked residual volume
    y = rng.random(5000) < (0.05 + 0.4 * x)
    fpr, tpr, _ = roc_curve(y.astype(int), x, pos_label=1)
    assert abs(tail_port.roc_auc(x, y) - auc(fpr, tpr)) < 1e-12
    assert abs(tail_port.average_precision(x, y) - average_precision_score(y.astype(int), x)) < 1e-12


def test_find_best_val_quirks():
    from oracle import tail_port

    x = np.array([0.0, 0.1, 0.2, 0.9, 0.95], dtype=np.float32)
    y = np.array([0, 0, 0, 1, 1], dtype=bool)
    d, t = tail_port.find_best_val(x, y, val_range=(0, np.max(x)), max_steps=10)
    assert d == 1.0 and 0.2 <= t < 0.9
    # degenerate range (all-zero residual): (lo, 1) is searched instead (utils_eval.py:512-513)
    d0, t0 = tail_port.find_best_val(np.zeros(4, np.float32), np.array([1, 0, 0, 0], bool), val_range=(0, 0.0), max_steps=3)
    assert d0 == 0.0
    # empty segmentation: Dice is NaN everywhere and the (0, 0) initialisation is returned (Appendix B.16)
    with np.errstate(all="ignore"):
        dn, tn = tail_port.find_best_val(np.zeros(4, np.float32), np.zeros(4, bool), val_range=(0, 1.0), max_steps=3)
    assert (dn, tn) == (0, 0)


def test_volume_tail_port_matches_reference_golden():
    from oracle import tail_port
    from oracle.weights import synthetic_volume

    gold = json.load(open(os.path.join(GOLD, "tail.json")))["d4"]
    for i in (0, 1):
        v = synthetic_volume(i, depth=4)
        p = tail_port.volume_tail(v["reco"][0, 0].numpy(), v["vol"][0, 0].numpy(), v["seg_orig"][0, 0].numpy(),
                                  v["mask_orig"][0, 0].numpy(), stage="val")
        assert p["BestThreshold"] == np.float32(gold["val"]["BestThresholdPerVol"][i])
        assert p["Dice"] == gold["val"]["DiceScorePerVol"][i]
        assert (p["TP"], p["FP"], p["TN"], p["FN"]) == tuple(int(gold["val"][k][i]) for k in ("TPPerVol", "FPPerVol", "TNPerVol", "FNPerVol"))
        assert abs(p["AUPRC"] - gold["val"]["AUPRCPerVol"][i]) < 1e-12
        assert abs(float(p["diff_filtered"].astype(np.float64).sum()) - gold["filtered_sum"][i]) < 1e-9
        assert p["Haus"] == gold["val"]["HausPerVol"][i]
        assert int(p["thresholded"].sum()) == gold["thresholded_count"][i]


def test_component_filter_and_hausdorff_ports():
    """The two restatements without a live library behind them (scikit-image, monai absent): checked against
    independent brute-force statements of the published definitions."""
    from scipy import ndimage

    from oracle import tail_port

    rng = np.random.default_rng(3)
    v = rng.random((6, 14, 14)) < 0.12
    got = tail_port.filter_small_components(v)
    lab, n = ndimage.label(v, structure=np.ones((3, 3, 3)))
    sizes = ndimage.sum(v, lab, index=np.arange(1, n + 1))
    want = np.isin(lab, 1 + np.flatnonzero(sizes > 7))
    assert np.array_equal(got, want) and 0 < want.sum() < v.sum()
    # octahedron shell + 1: seven voxels around an empty centre are NOT a hole under skimage's full structuring element
    o = np.zeros((7, 7, 7), bool)
    for d in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)):
        o[3 + d[0], 3 + d[1], 3 + d[2]] = True
    o[5, 3, 3] = True
    assert not tail_port.filter_small_components(o).any()

    def brute(a, b):
        def surf(m):
            p = np.pad(m, 1)
            inner = p[1:-1, 1:-1, 1:-1].copy()
            for ax in range(3):
                for sh in (-1, 1):
                    inner &= np.roll(p, sh, axis=ax)[1:-1, 1:-1, 1:-1]
            return np.argwhere(m & ~inner)
        pa, pb = surf(a), surf(b)
        if len(pa) == 0 and len(pb) == 0:
            return float("nan")
        if len(pa) == 0 or len(pb) == 0:
            return float("inf")
        d2 = ((pa[:, None, :] - pb[None, :, :]) ** 2).sum(-1)
        return float(np.sqrt(np.float64(max(d2.min(1).max(), d2.min(0).max()))))

    for seed in range(4):
        r = np.random.default_rng(seed)
        a = ndimage.binary_dilation(r.random((8, 12, 10)) < 0.01, iterations=2)
        b = ndimage.binary_dilation(r.random((8, 12, 10)) < 0.01, iterations=1)
        w, g = brute(a, b), tail_port.hausdorff_distance(a, b)
        assert w == g or (np.isnan(w) and np.isnan(g))
    z = np.zeros((4, 4, 4), bool)
    assert np.isnan(tail_port.hausdorff_distance(z, z)) and tail_port.hausdorff_distance(~z, z) == float("inf")


def test_encoder_port_matches_golden():
    from oracle import resnet_port
    from oracle.weights import make_state_dict

    sd = make_state_dict(resnet_port.param_shapes(128), seed=3)
    g = np.load(os.path.join(GOLD, "encoder_96.npz"))
    with torch.no_grad():
        c = resnet_port.resnet_forward(sd, torch.from_numpy(g["x"]))
    assert (c - torch.from_numpy(g["c"])).abs().max().item() < 1e-4
    assert len(resnet_port.param_shapes(128)) == 320


def test_noise_ensemble_members_are_independent_batch_samples():
    """Why the product may run the noise ensemble (DDPM_2D.py:214-232) as ONE UNet forward over the stacked members:
    in the reference's own arithmetic (oracle port, fp32 CPU) every sample of a batch is independent - per-sample
    timestep embedding, GroupNorm statistics per sample, attention per sample - so k forwards at (t_i, noise_i) over D
    slices equal one forward over the k x D stacked slices with a per-sample t vector."""
    from oracle import diffusion_port, unet_port
    from oracle.weights import make_state_dict

    spec = unet_port.UNetSpec(model_channels=64, channel_mult=(1, 2), num_res_blocks=1, num_classes=128)
    sd = make_state_dict(unet_port.param_shapes(spec), seed=2)
    sched = diffusion_port.schedule_buffers()
    g = torch.Generator().manual_seed(0)
    D, steps = 3, (249, 499, 749)
    img = torch.rand(D, 1, 32, 32, generator=g)
    cond = torch.randn(D, 128, generator=g)
    noises = [0.6 * torch.randn(D, 1, 32, 32, generator=g) for _ in steps]
    model = lambda x, t, c: unet_port.unet_forward(sd, spec, x, t, c)  # noqa: E731
    with torch.no_grad():
        loop = [diffusion_port.reconstruct(model, sched, img, t, cond, n)[1] for t, n in zip(steps, noises)]
        x0 = img * 2 - 1
        tt = torch.cat([torch.full((D,), t, dtype=torch.long) for t in steps])
        xt = torch.cat([diffusion_port.q_sample(sched, x0, tt[i * D:(i + 1) * D], n) for i, n in enumerate(noises)])
        out = model(xt, tt, cond.repeat(len(steps), 1))
        stacked = ((out + 1) * 0.5).reshape(len(steps), D, 1, 32, 32)
    for i in range(len(steps)):
        assert (stacked[i] - loop[i]).abs().max().item() <= 2e-5
    mean_loop = sum(loop) / len(steps)
    assert (stacked.mean(0) - mean_loop).abs().max().item() <= 2e-5


def test_ports_reproduce_headline_geometry_goldens():
    """The oracle ports at BASELINE's own geometry (conditioned 128-channel UNet, 96x96) against goldens of the LIVE
    reference: the encoder's condition vector, the T0 = 50 reverse loop of configs[1] (50 UNet forwards of batch 2,
    ~30 s of host time) and configs[0]'s unconditioned single step - all bit-identical, simplex noise regenerated from
    the stored numpy seed."""
    from oracle import diffusion_port, resnet_port, unet_port
    from oracle.simplex_port import gen_noise_port
    from oracle.weights import make_state_dict, synthetic_slices

    sched = diffusion_port.schedule_buffers()
    with torch.no_grad():
        g = np.load(os.path.join(GOLD, "uncond_step_96.npz"))
        spec = unet_port.UNetSpec(num_classes=None)
        sd = make_state_dict(unet_port.param_shapes(spec), seed=1)
        x = synthetic_slices(1, 96, seed=41)
        np.random.seed(int(g["seed"]))
        loss, reco = diffusion_port.reconstruct(lambda a, t, c: unet_port.unet_forward(sd, spec, a, t, c), sched, x, 499,
                                                None, gen_noise_port((1, 1, 96, 96)))
        assert (reco - torch.from_numpy(g["reco"])).abs().max().item() <= 1e-6
        assert abs(float(loss) - float(g["loss"])) <= 1e-6

        g = np.load(os.path.join(GOLD, "reverse_96.npz"))
        spec = unet_port.UNetSpec()
        sd = make_state_dict(unet_port.param_shapes(spec), seed=1)
        enc = make_state_dict(resnet_port.param_shapes(128), seed=3)
        x = synthetic_slices(2, 96, seed=31)
        cond = resnet_port.resnet_forward(enc, x)
        assert (cond - torch.from_numpy(g["cond"])).abs().max().item() <= 1e-6
        np.random.seed(int(g["seed_T50"]))
        rec = diffusion_port.reverse_loop(lambda a, t, c: unet_port.unet_forward(sd, spec, a, t, c), sched, x * 2 - 1,
                                          cond, 50, lambda: gen_noise_port((2, 1, 96, 96)))
        assert (rec - torch.from_numpy(g["out_rand_fp32_T50"])).abs().max().item() <= 1e-5
        # the yardstick the GPU test uses is part of the golden: the reference under fp16 autocast against its fp32 self
        for wtag, T0, lo, hi in (("rand", 50, 0.02, 0.3), ("rand", 500, 0.1, 1.0), ("fit", 50, 0.002, 0.05),
                                 ("fit", 500, 0.02, 0.5)):
            dev = np.abs(g[f"out_{wtag}_amp16_T{T0}"] - g[f"out_{wtag}_fp32_T{T0}"]).max()
            assert lo <= dev <= hi, (wtag, T0, dev)


def test_ddim_port_matches_reference_golden():
    """oracle.diffusion_port.ddim_sample against the live reference's ddim_sample (tests/golden/ddim_small_32.npz)."""
    from oracle import diffusion_port, unet_port
    from oracle.simplex_port import gen_noise_port
    from oracle.weights import make_state_dict

    g = np.load(os.path.join(GOLD, "ddim_small_32.npz"))
    spec = unet_port.UNetSpec(model_channels=64, channel_mult=(1, 2), num_res_blocks=1, num_classes=128)
    sd = make_state_dict(unet_port.param_shapes(spec), seed=2)
    model = lambda x, t, c: unet_port.unet_forward(sd, spec, x, t, c)  # noqa: E731
    sched = diffusion_port.schedule_buffers()
    img, cond = torch.from_numpy(g["img"]), torch.from_numpy(g["cond"])
    with torch.no_grad():
        for objective in ("pred_x0", "pred_noise"):
            np.random.seed(5)
            rec = diffusion_port.ddim_sample(model, sched, (2, 1, 32, 32), img * 2 - 1, cond, 300, 6, 0.7, True,
                                             lambda: gen_noise_port((2, 1, 32, 32)), None, objective=objective)
            assert (rec - torch.from_numpy(g[f"ddim_simplex_{objective}"])).abs().max().item() < 1e-4
        draws = [torch.zeros(2, 1, 32, 32), torch.from_numpy(g["ddim_xT"])]
        np.random.seed(6)
        rec = diffusion_port.ddim_sample(model, sched, (2, 1, 32, 32), img * 2 - 1, cond, 0, 4, 1.0, None,
                                         lambda: gen_noise_port((2, 1, 32, 32)), lambda: draws.pop(0))
        assert (rec - torch.from_numpy(g["ddim_gauss_start"])).abs().max().item() < 1e-4
