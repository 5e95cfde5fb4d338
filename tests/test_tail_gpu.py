"""GPU parity of the anomaly-scoring tail (residual, eroded brain mask, 5x5x5 median, Dice bisection, AUC/AP, counts)
against (1) golden vectors from the live reference's utils_eval (tests/golden/stencil.npz, tail.json) and (2) the numpy
oracle port on full-size 96x96x50 volumes.  Integer / index / threshold results must be bit-exact.

Reference: src/utils/utils_eval.py:18-194 (_test_step), :196-297 (_test_end), :447-464, :508-557."""
import json
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

pytestmark = pytest.mark.gpu


class Cfg(dict):
    __getattr__ = dict.get


def _cfg(**over):
    c = Cfg(resizedEvaluation=True, erodeBrainmask=True, medianFiltering=True, saveOutputImages=False, evalSeg=True,
            threshold="auto")
    c.update(over)
    return c


class Host:
    pass


def _host(stage):
    from cddpm.eval_tail import get_eval_dictionary

    h = Host()
    h.cfg = _cfg()
    h.eval_dict = get_eval_dictionary()
    h.threshold = {}
    h.dataset = ["Brats21"]
    h.stage = stage
    h.diffs_list, h.seg_list = [], []
    return h


@pytest.mark.parametrize("hw", [32, 56, 24])
def test_stencils_bit_exact_vs_reference_golden(hw):
    from cddpm.eval_tail import apply_3d_median_filter, apply_brainmask_volume

    g = np.load(os.path.join(GOLD, "stencil.npz"))
    vol = torch.from_numpy(g[f"vol{hw}"]).cuda()
    mask = torch.from_numpy(g[f"mask{hw}"].astype(np.float32)).cuda()
    masked = apply_brainmask_volume(vol.clone()[None, None], mask[None, None]).squeeze()
    assert np.array_equal(masked.cpu().numpy(), g[f"masked{hw}"])
    filt = apply_3d_median_filter(masked)
    assert np.array_equal(filt.cpu().numpy(), g[f"filtered{hw}"])
    # numpy in / numpy out form used by the reference's call site (utils_eval.py:69)
    filt_np = apply_3d_median_filter(g[f"masked{hw}"], kernelsize=5)
    assert isinstance(filt_np, np.ndarray) and np.array_equal(filt_np, g[f"filtered{hw}"])


def test_median_properties_full_size():
    from cddpm.eval_tail import apply_3d_median_filter

    g = torch.Generator().manual_seed(0)
    x = torch.rand(96, 96, 50, generator=g).cuda()
    m = apply_3d_median_filter(x)
    assert torch.equal(apply_3d_median_filter(x * 2), m * 2)  # commutes with exact monotone maps
    assert torch.equal(apply_3d_median_filter(-x), -m)  # odd window: median(-x) = -median(x)
    c = torch.full((96, 96, 50), 0.37, device="cuda")
    assert torch.equal(apply_3d_median_filter(c), c)
    assert float(m.min()) >= float(x.min()) and float(m.max()) <= float(x.max())
    # kernel 3 against scipy directly (scipy is the reference's own implementation and ships in this image)
    from scipy import ndimage

    xs = x[:40, :40, :9].contiguous()
    ref3 = ndimage.median_filter(xs.cpu().numpy(), (3, 3, 3))
    assert np.array_equal(apply_3d_median_filter(xs, kernelsize=3).cpu().numpy(), ref3)


@pytest.mark.parametrize("depth", [50, 4])
def test_volume_tail_matches_oracle_port_and_reference(depth):
    from cddpm import eval_tail
    from oracle import tail_port
    from oracle.weights import synthetic_volume

    gold = json.load(open(os.path.join(GOLD, "tail.json")))[f"d{depth}"]
    host = _host("val")
    vols = [synthetic_volume(s, depth=depth) for s in (0, 1)]
    for i, v in enumerate(vols):
        out = eval_tail._test_step(host, v["reco"].cuda(), v["vol"].cuda(), v["seg_orig"].cuda(), v["mask_orig"].cuda(), i,
                                   [f"v{i}"], torch.tensor([1]))
        port = tail_port.volume_tail(v["reco"][0, 0].numpy(), v["vol"][0, 0].numpy(), v["seg_orig"][0, 0].numpy(),
                                     v["mask_orig"][0, 0].numpy(), stage="val")
        # bit-exact volumes
        assert np.array_equal(out.diff_hwd().cpu().numpy(), port["diff_filtered"])
        ed = host.eval_dict
        assert ed["BestThresholdPerVol"][i] == port["BestThreshold"] == np.float32(gold["val"]["BestThresholdPerVol"][i])
        assert ed["BestDicePerVol"][i] == port["BestDice"] == gold["val"]["BestDicePerVol"][i]
        assert ed["DiceScorePerVol"][i] == port["Dice"] == gold["val"]["DiceScorePerVol"][i]
        for k, pk in (("TPPerVol", "TP"), ("FPPerVol", "FP"), ("TNPerVol", "TN"), ("FNPerVol", "FN")):
            assert ed[k][i] == port[pk] == int(gold["val"][k][i])
        assert abs(ed["AUCPerVol"][i] - gold["val"]["AUCPerVol"][i]) < 1e-9
        assert abs(ed["AUPRCPerVol"][i] - gold["val"]["AUPRCPerVol"][i]) < 1e-9
        for k in ("l1recoErrorAll", "l2recoErrorAll", "l1recoErrorUnhealthy", "l1recoErrorHealthy"):
            assert abs(ed[k][i] - gold["val"][k][i]) < 1e-6 * max(1.0, abs(gold["val"][k][i]))
        assert abs(ed["AnomalyScoreRecoPerVol"][i] - gold["val"]["AnomalyScoreRecoPerVol"][i]) < 1e-6
        assert ed["lesionSizePerVol"][i] == int(gold["val"]["lesionSizePerVol"][i])
        assert ed["HausPerVol"][i] == port["Haus"] == gold["val"]["HausPerVol"][i]  # sqrt of an exact integer
        assert abs(ed["TPRPerVol"][i] - gold["val"]["TPRPerVol"][i]) < 1e-12
        assert abs(ed["SpecificityPerVol"][i] - gold["val"]["SpecificityPerVol"][i]) < 1e-12
    ed = host.eval_dict
    assert np.allclose(ed["DiceScorePerSlice"], gold["val"]["DiceScorePerSlice"], rtol=0, atol=1e-12)
    assert ed["labelPerSlice"] == [int(v) for v in gold["val"]["labelPerSlice"]]
    assert np.allclose(ed["AnomalyScoreRecoPerSlice"], gold["val"]["AnomalyScoreRecoPerSlice"], rtol=1e-5, atol=1e-7)
    assert np.allclose(ed["AUCAnomalyRecoPerSlice"], gold["val"]["AUCAnomalyRecoPerSlice"], atol=1e-9)
    assert np.allclose(ed["AUPRCAnomalyRecoPerSlice"], gold["val"]["AUPRCAnomalyRecoPerSlice"], atol=1e-9)
    # global threshold over all validation voxels (_test_end), then the test stage with it
    eval_tail._test_end(host)
    assert host.threshold["total"] == np.float32(gold["threshold_total"])
    total = host.threshold["total"]
    host2 = _host("test")
    host2.threshold = {"total": total}
    v = vols[0]
    eval_tail._test_step(host2, v["reco"].cuda(), v["vol"].cuda(), v["seg_orig"].cuda(), v["mask_orig"].cuda(), 0, ["v0"],
                         torch.tensor([1]))
    assert host2.eval_dict["DiceScorePerVol"][0] == gold["test_dice"]
    assert [host2.eval_dict[k][0] for k in ("TPPerVol", "FPPerVol", "TNPerVol", "FNPerVol")] == gold["test_counts"]
    eval_tail._test_end(host2)
    assert not hasattr(host2, "threshold")  # the reference deletes it after the test stage (utils_eval.py:259-260)


def test_reco_in_unet_layout_is_read_in_place():
    """final_volume arrives as reco[D,1,H,W].squeeze().permute(1,2,0) (DDPM_2D.py:256-257): a strided view."""
    from cddpm import eval_tail
    from oracle.weights import synthetic_volume

    v = synthetic_volume(2, depth=12)
    reco_dhw = v["reco"][0, 0].permute(2, 0, 1).contiguous().cuda()  # what the UNet produces: [D,H,W]
    view = reco_dhw.permute(1, 2, 0)[None, None]  # [1,1,H,W,D] non-contiguous
    a, _ = eval_tail.residual_and_filter(view, v["vol"].cuda(), v["seg_orig"].cuda(), v["mask_orig"].cuda())
    b, _ = eval_tail.residual_and_filter(v["reco"].cuda(), v["vol"].cuda(), v["seg_orig"].cuda(), v["mask_orig"].cuda())
    assert torch.equal(a.diff, b.diff)


def test_find_best_val_and_ranking_standalone():
    from cddpm import eval_tail
    from oracle import tail_port

    rng = np.random.default_rng(3)
    x = rng.random(20000).astype(np.float32)
    x[rng.random(20000) < 0.5] = 0.0  # many exact ties, like a masked residual
    y = (rng.random(20000) < 0.1 + 0.5 * x)
    bd, bt = eval_tail.find_best_val(x, y, val_range=(0, np.max(x)), max_steps=10)
    pd, pt = tail_port.find_best_val(x, y, val_range=(0, np.max(x)), max_steps=10)
    assert bd == pd and bt == pt
    auc, *_ = eval_tail.compute_roc(torch.from_numpy(x).cuda(), y)
    ap, *_ = eval_tail.compute_prc(torch.from_numpy(x).cuda(), y)
    assert abs(auc - tail_port.roc_auc(x, y)) < 1e-12 and abs(ap - tail_port.average_precision(x, y)) < 1e-12


def _octahedron_plus_one():
    """Six face neighbours of an empty centre plus one attached voxel: 7 voxels.  With skimage's full structuring
    element the centre is not a hole (filled_area 7 -> removed); a 6-connected fill would call it 8 and keep it."""
    v = np.zeros((7, 7, 7), bool)
    for d in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)):
        v[3 + d[0], 3 + d[1], 3 + d[2]] = True
    v[5, 3, 3] = True
    return v


@pytest.mark.parametrize("shape,density", [((50, 96, 96), 0.02), ((50, 96, 96), 0.08), ((4, 96, 96), 0.05),
                                            ((3, 5, 7), 0.3), ((1, 1, 9), 0.6), ((20, 32, 32), 0.15)])
def test_small_component_filter_bit_exact_vs_oracle(shape, density):
    from cddpm.eval_tail import filter_3d_connected_components
    from oracle import tail_port

    rng = np.random.default_rng(hash((shape, density)) % (1 << 31))
    v = rng.random(shape) < density
    want = tail_port.filter_small_components(v)
    got = filter_3d_connected_components(v.copy())
    assert isinstance(got, np.ndarray) and got.dtype == bool
    assert np.array_equal(got, want)
    assert 0 < want.sum() < v.sum() or v.size < 200
    # tensor in -> tensor out on the same device; idempotent; only removes voxels
    t = filter_3d_connected_components(torch.from_numpy(v).cuda())
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), want)
    assert np.array_equal(filter_3d_connected_components(want.copy()), want)
    assert not (want & ~v).any()


def test_small_component_filter_edge_cases():
    from cddpm.eval_tail import filter_3d_connected_components
    from oracle import tail_port

    o = _octahedron_plus_one()
    assert not tail_port.filter_small_components(o).any()
    assert not filter_3d_connected_components(o.copy()).any()
    # 8 voxels in a diagonal chain (26-connectivity only) stay, 7 go; components touching the border
    chain = np.zeros((10, 10, 10), bool)
    for i in range(8):
        chain[i, i, i] = True
    chain[9, 0, 0:7] = True
    want = tail_port.filter_small_components(chain)
    assert want.sum() == 8
    assert np.array_equal(filter_3d_connected_components(chain.copy()), want)
    # empty, full, 4-D input folded like the reference does (:491-493)
    z = np.zeros((4, 6, 6), bool)
    assert not filter_3d_connected_components(z).any()
    assert filter_3d_connected_components(~z).all()
    v4 = np.random.default_rng(5).random((2, 3, 8, 8)) < 0.2
    want4 = tail_port.filter_small_components(v4.reshape(6, 8, 8)).reshape(v4.shape)
    assert np.array_equal(filter_3d_connected_components(v4.copy()), want4)


def _blobs(shape, seed, n, rmax):
    rng = np.random.default_rng(seed)
    zz, yy, xx = np.meshgrid(*(np.arange(s) for s in shape), indexing="ij")
    v = np.zeros(shape, bool)
    for _ in range(n):
        c = [rng.integers(0, s) for s in shape]
        r = rng.uniform(1, rmax)
        v |= (zz - c[0]) ** 2 + (yy - c[1]) ** 2 + (xx - c[2]) ** 2 <= r * r
    return v


@pytest.mark.parametrize("shape", [(96, 96, 50), (96, 96, 4), (5, 7, 3), (1, 1, 1), (33, 17, 2)])
def test_hausdorff_bit_exact_vs_oracle(shape):
    from cddpm.eval_tail import compute_hausdorff_distance
    from oracle import tail_port

    for seed in range(3):
        a = _blobs(shape, seed, 3, 9.0)
        b = _blobs(shape, 100 + seed, 2, 7.0)
        want = tail_port.hausdorff_distance(a, b)
        got = compute_hausdorff_distance(torch.from_numpy(a).float().cuda(), torch.from_numpy(b).float().cuda())
        assert got == want or (np.isnan(got) and np.isnan(want)), (shape, seed, got, want)
        # symmetric, zero against itself
        assert compute_hausdorff_distance(b.astype(np.float32), a.astype(np.float32)) == got or np.isnan(got)
        if a.any():
            assert compute_hausdorff_distance(a.astype(np.float32), a.astype(np.float32)) == 0.0
    z = np.zeros(shape, np.float32)
    one = z.copy()
    one[0, 0, 0] = 1
    assert np.isnan(compute_hausdorff_distance(z, z))          # monai: both empty -> nan
    assert compute_hausdorff_distance(one, z) == float("inf")  # one empty -> inf
    assert compute_hausdorff_distance(z, one) == float("inf")
    far = z.copy()
    far[-1, -1, -1] = 1
    d2 = sum((s - 1) ** 2 for s in shape)
    assert compute_hausdorff_distance(one, far) == float(np.sqrt(np.float64(d2)))


def test_hausdorff_on_strided_seg_view_and_counts():
    """seg arrives in the dataloader layout [H,W,D] while the prediction buffer is [D,H,W]."""
    import ctypes

    from cddpm import eval_tail
    from cddpm._lib import check, current_stream, lib, ptr
    from oracle import tail_port

    shape = (40, 48, 12)
    pred = _blobs(shape, 7, 4, 6.0)
    seg = _blobs(shape, 8, 3, 8.0)
    segt = torch.from_numpy(seg.astype(np.float32) * 3.0).cuda()  # any value > 0 counts
    predt = torch.from_numpy(pred).cuda()
    vol = eval_tail._Volume(predt.permute(2, 0, 1).float().contiguous(), segt, segt, shape)
    p8 = predt.permute(2, 0, 1).to(torch.uint8).contiguous()
    assert eval_tail._hausdorff_device(p8, vol) == tail_port.hausdorff_distance(pred, seg)
    cc = torch.zeros(3, dtype=torch.int64, device="cuda")
    sv = vol.seg_view
    check(lib().cddpm_confusion_counts(ptr(p8), ctypes.byref(sv), shape[0], shape[1], shape[2], ptr(cc), current_stream()),
          "cddpm_confusion_counts")
    assert cc.tolist() == [int((pred & seg).sum()), int((pred & ~seg).sum()), int((~pred & seg).sum())]


def test_error_sums_are_run_to_run_identical_and_match_float64():
    """The l1 / l2 error sums of a volume (utils_eval.py:36-49) must not depend on block scheduling: a sharded sweep
    reports the same volume from whichever rank it lands on, and the 8-GPU sweep check compares to the last bit.  The
    kernel adds its per-block partials as fixed point in integer atomics; here: 25 launches on a full-size volume give
    identical bits, and agree with a float64 numpy sum to 1e-9."""
    from cddpm.eval_tail import residual_and_filter
    from oracle.weights import synthetic_volume

    v = synthetic_volume(3, depth=50)
    args = [v[k].cuda() for k in ("reco", "vol", "seg_orig", "mask_orig")]
    runs = [residual_and_filter(*args, median=False)[1] for _ in range(25)]
    for r in runs[1:]:
        assert np.array_equal(r.view(np.uint64), runs[0].view(np.uint64))
    d = v["vol"][0, 0].numpy().astype(np.float32) - v["reco"][0, 0].numpy().astype(np.float32)
    seg = v["seg_orig"][0, 0].numpy() > 0
    a = np.abs(d).astype(np.float64)
    q = d.astype(np.float64) ** 2
    want = [a.sum(), q.sum(), a[seg].sum(), q[seg].sum(), a[~seg].sum(), q[~seg].sum(), float(seg.sum())]
    assert runs[0][6] == want[6]
    for got, w in zip(runs[0][:6], want[:6]):
        assert abs(got - w) <= 1e-9 * max(1.0, abs(w))


def test_full_resolution_evaluation_vs_reference_golden():
    """cfg.resizedEvaluation=False (utils_eval.py:24-25): the [96,96,50] reconstruction is resized to new_size =
    [160,190,160] (trilinear, align_corners=True) on the GPU and scored against full-resolution originals.  Golden of the
    live reference (oracle/make_golden.py tail_fullres).  The resize is compared at 1e-6 (fp32 fused-multiply-add
    contraction on the host is not specified), thresholds / Dice / counts exactly as far as the resize allows."""
    from cddpm import eval_tail
    from oracle.weights import synthetic_fullres_case

    g = np.load(os.path.join(GOLD, "tail_fullres.npz"))
    c = synthetic_fullres_case(4)
    # the reconstruction as the UNet hands it over: a permuted [D,1,H,W] view, read in place
    reco_dchw = c["reco"][0, 0].permute(2, 0, 1).unsqueeze(1).contiguous().cuda()
    final = reco_dchw.squeeze(1).permute(1, 2, 0).unsqueeze(0).unsqueeze(0)
    resized = eval_tail.trilinear_resize(final, [160, 190, 160])
    assert tuple(resized.shape) == (1, 1, 160, 190, 160) and resized.is_contiguous()
    r = resized[0, 0].cpu()
    for p, v in zip(g["probes"], g["probe_values"]):
        assert abs(float(r[tuple(int(i) for i in p)]) - float(v)) <= 1e-6, (p, v)
    assert (r[:, :, 80] - torch.from_numpy(g["resized_slice"])).abs().max().item() <= 1e-6
    assert abs(float(r.double().sum()) - float(g["resized_sum"])) <= 1e-6 * abs(float(g["resized_sum"]))
    # end points are copied exactly (align_corners)
    assert float(r[0, 0, 0]) == float(c["reco"][0, 0, 0, 0, 0]) and float(r[-1, -1, -1]) == float(c["reco"][0, 0, -1, -1, -1])

    h = _host("val")
    h.cfg = _cfg(resizedEvaluation=False)
    h.new_size = [160, 190, 160]
    eval_tail._test_step(h, final, c["vol_orig"].cuda(), c["seg_orig"].cuda(), c["mask_orig"].cuda(), 0, ["v4"],
                         torch.tensor([1]))
    ed = h.eval_dict
    print({k: float(ed[k][0]) for k in ("DiceScorePerVol", "BestThresholdPerVol", "AUPRCPerVol", "TPPerVol", "HausPerVol")})
    for k in ("DiceScorePerVol", "BestDicePerVol", "AUCPerVol", "AUPRCPerVol", "l1recoErrorAll", "l2recoErrorAll",
              "l1recoErrorUnhealthy", "l1recoErrorHealthy", "AnomalyScoreRecoPerVol"):
        assert abs(float(ed[k][0]) - float(g[k][0])) <= 1e-5, (k, ed[k][0], g[k][0])
    assert abs(float(ed["BestThresholdPerVol"][0]) - float(g["BestThresholdPerVol"][0])) <= 1e-6
    assert float(ed["lesionSizePerVol"][0]) == float(g["lesionSizePerVol"][0])
    assert float(ed["HausPerVol"][0]) == float(g["HausPerVol"][0])
    for k in ("TPPerVol", "FPPerVol", "TNPerVol", "FNPerVol"):  # 4.86 M voxels; a 1-ulp resize difference may flip a few
        assert abs(float(ed[k][0]) - float(g[k][0])) <= 5, (k, ed[k][0], g[k][0])


@pytest.mark.parametrize("case", ["dense", "sparse_ties", "empty_label", "all_zero"])
def test_device_bisection_equals_host_loop(case, monkeypatch):
    """cddpm_dice_bisect (ten find_best_val decisions on the device over the ranking pass's sorted scores, one launch)
    against the host loop that counts over the volume once per step (the round-1 path, itself pinned to the live
    reference by tests/golden/tail.json): identical best Dice, identical float32 threshold, identical AUC / AUPRC."""
    from cddpm import eval_tail

    g = torch.Generator().manual_seed({"dense": 1, "sparse_ties": 2, "empty_label": 3, "all_zero": 4}[case])
    H, W, D = 40, 48, 21
    x = torch.rand(D, H, W, generator=g)
    seg = (torch.rand(H, W, D, generator=g) > 0.93).float()
    if case == "sparse_ties":  # mostly exact zeros and a few repeated values, like a masked residual
        x = torch.where(torch.rand(D, H, W, generator=g) > 0.2, torch.zeros(()), (x * 8).round() / 8)
    elif case == "empty_label":
        seg.zero_()
    elif case == "all_zero":
        x.zero_()
    vol = eval_tail._Volume(x.cuda().contiguous(), seg.cuda(), seg.cuda(), (H, W, D))
    monkeypatch.setenv("CDDPM_DEVICE_BISECT", "1")
    dev = eval_tail._ranking_and_bisect(vol, 10)
    monkeypatch.setenv("CDDPM_DEVICE_BISECT", "0")
    host = eval_tail._ranking_and_bisect(vol, 10)
    print(case, dev, host)
    for a, b in zip(dev, host):
        assert (a == b) or (a != a and b != b), (case, dev, host)
    assert type(dev[3]) is type(host[3])  # np.float32 (or the int 0 when no step updated the optimum)


def test_log_images_writes_the_reference_grid_files(tmp_path, monkeypatch):
    """saveOutputImages=True (utils_eval.py:72-74, log_images :586-628): `grid/{ID}_{j}_Grid.png` for every 10th axial
    slice, written off the critical path; each file decodes to the four-panel row of oracle.tail_port.compose_grid_port
    (rot90(., 3) panels, per-panel gray normalisation, inferno difference normalised to [0, max + 0.01]) within one
    grey level, and the metrics of the step are unchanged by the logging."""
    from PIL import Image

    from cddpm import eval_tail
    from oracle import tail_port
    from oracle.weights import synthetic_volume

    monkeypatch.chdir(tmp_path)
    v = synthetic_volume(0, depth=24)
    args = (v["reco"].cuda(), v["vol"].cuda(), v["seg_orig"].cuda(), v["mask_orig"].cuda(), 0, ["caseA"], torch.tensor([1]))
    plain = _host("val")
    eval_tail._test_step(plain, *args)
    host = _host("val")
    host.cfg["saveOutputImages"] = True
    out = eval_tail._test_step(host, *args)
    eval_tail._test_end(host)  # flushes the writer
    for k in ("DiceScorePerVol", "AUCPerVol", "BestThresholdPerVol", "HausPerVol"):
        assert host.eval_dict[k] == plain.eval_dict[k]
    files = sorted(os.listdir(tmp_path / "grid"))
    assert files == sorted(f"caseA_{j}_Grid.png" for j in range(0, 24, 10))
    diff = out.diff_hwd().cpu().numpy()
    orig, reco, seg = (v[k][0, 0].numpy() for k in ("vol", "reco", "seg_orig"))
    for j in range(0, 24, 10):
        panels = np.stack([orig[..., j], reco[..., j], diff[..., j], seg[..., j]]).astype(np.float32)
        ranges = np.stack([panels.min(axis=(1, 2)), panels.max(axis=(1, 2))], axis=1)
        ranges[2] = (0.0, np.float32(diff.max()) + np.float32(0.01))
        want = tail_port.compose_grid_port(panels, ranges)
        got = np.asarray(Image.open(tmp_path / "grid" / f"caseA_{j}_Grid.png"))
        assert got.shape == want.shape == (96, 4 * 96, 3)
        assert np.abs(got.astype(np.int16) - want.astype(np.int16)).max() <= 1
        assert got[:, 2 * 96:3 * 96].std() > 0  # the difference panel is not blank


@pytest.mark.parametrize("shape", [(13, 17, 9), (33, 7, 5), (40, 41, 12)])
def test_median5_pair_kernel_vs_scipy_on_odd_shapes(shape):
    """The 5x5x5 median (csrc/tail.cu:median5_pair_kernel: two x-neighbours per thread, forgetful selection) against
    scipy.ndimage.median_filter itself - the reference's implementation (utils_eval.py:462-464) - on volumes whose sides
    are odd, smaller than a tile, or not multiples of the pair width; values mix exact zeros, ties and negatives."""
    from scipy import ndimage

    from cddpm.eval_tail import apply_3d_median_filter

    g = torch.Generator().manual_seed(sum(shape))
    x = torch.rand(*shape, generator=g)
    x = torch.where(torch.rand(*shape, generator=g) < 0.4, torch.zeros(()), x)       # many exact zeros (masked residual)
    x = torch.where(torch.rand(*shape, generator=g) < 0.1, torch.full((), 0.25), x)  # ties
    x = torch.where(torch.rand(*shape, generator=g) < 0.05, -x, x)                   # a few negatives
    ref = ndimage.median_filter(x.numpy(), (5, 5, 5))
    got = apply_3d_median_filter(x.cuda(), kernelsize=5).cpu().numpy()
    assert np.array_equal(got, ref)


def test_compose_grid_kernel_matches_port_on_a_non_square_slice():
    import ctypes  # noqa: F401

    from cddpm._lib import check, current_stream, lib, ptr
    from oracle.tail_port import compose_grid_port

    H, W = 40, 56
    g = torch.Generator().manual_seed(5)
    panels = torch.rand(4, H, W, generator=g)
    panels[3] = (panels[3] > 0.7).float()
    ranges = torch.tensor([[0.1, 0.9], [0.0, 1.0], [0.0, 0.8], [0.0, 1.0]])
    rgb = torch.empty(W, 4 * H, 3, dtype=torch.uint8, device="cuda")
    panels_dev, ranges_dev = panels.cuda(), ranges.cuda()  # named: the launch is asynchronous
    check(lib().cddpm_compose_grid(ptr(panels_dev), ptr(ranges_dev), H, W, ptr(rgb), current_stream()),
          "cddpm_compose_grid")
    torch.cuda.synchronize()
    want = compose_grid_port(panels.numpy(), ranges.numpy())
    diff = np.abs(rgb.cpu().numpy().astype(np.int16) - want.astype(np.int16))
    assert diff.max() <= 1 and (diff > 0).mean() < 0.2  # one grey level where the fused multiply-add rounds the other way
