"""BASELINE configs[2] timings (development aid): DDPM_2D.test_step on synthetic BraTS21-shaped volumes, plus the
anomaly-scoring tail kernel by kernel (CUDA events, back-to-back launches after warm-up, algorithmic bytes -> GB/s).
   python tools/time_volume.py [depth=50] [n_vol=8]"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "conditioned-diffusion-models-uad_b200")]
import bench  # noqa: E402
from cddpm import eval_tail  # noqa: E402
from cddpm._lib import check, current_stream, lib, ptr  # noqa: E402
from cddpm.ddpm_2d import DDPM_2D  # noqa: E402


def ev_time(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3  # us


def main():
    D = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    NV = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    H = W = 96
    torch.manual_seed(0)
    np.random.seed(0)
    vols = [bench.synthetic_volume(s, D) for s in range(NV)]
    cfg = bench.model_cfg()
    cfg["force_num_eval_slices"] = False
    model = DDPM_2D(cfg, prefix="p/").cuda().eval()

    def batch(v, stage):
        return {"Dataset": ["Brats21"], "vol": {"data": v["vol"]}, "vol_orig": {"data": v["vol"]},
                "seg_orig": {"data": v["seg_orig"]}, "mask_orig": {"data": v["mask_orig"]}, "seg_available": True,
                "ID": ["v"], "stage": stage, "label": torch.tensor([1])}

    def sweep():
        model.on_test_start()
        for i, v in enumerate(vols):
            model.test_step(batch(v, "val"), i)

    for _ in range(2):
        sweep()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        sweep()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3 / NV
    print(f"test_step (ensemble {cfg['step_ensemble'] if cfg.get('step_ensemble') else '250/500/750'}, D={D}): "
          f"{dt * 1e3:.2f} ms per volume = {1 / dt:.1f} volumes/s = {D / dt:.0f} slices/s", flush=True)

    # model part only (encoder + 3 single-step reconstructions of D slices)
    x = vols[0]["vol"].cuda().squeeze(0).permute(3, 0, 1, 2).contiguous()
    with torch.no_grad():
        us = ev_time(lambda: model.reconstruct_slices(x), iters=10)
    print(f"reconstruct_slices (encoder + 3 x [q_sample, UNet, recon]) B={D}: {us / 1e3:.2f} ms", flush=True)
    with torch.no_grad():
        reco, _, _ = model.reconstruct_slices(x)
    final = reco.squeeze(1).permute(1, 2, 0).unsqueeze(0).unsqueeze(0)
    v = vols[0]
    orig, seg, mask = v["vol"].cuda(), v["seg_orig"].cuda(), v["mask_orig"].cuda()
    host = model
    host.stage = "val"
    host.dataset = ["Brats21"]
    model.eval_dict = eval_tail.get_eval_dictionary()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(10):
        eval_tail._test_step(host, final, orig, seg, mask, i, ["v"], torch.tensor([1]))
    torch.cuda.synchronize()
    print(f"_test_step tail (host wall clock, all syncs included): {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms per volume",
          flush=True)

    # ---- kernel by kernel
    n = H * W * D
    vol, _ = eval_tail.residual_and_filter(final, orig, seg, mask)
    s = current_stream()
    L = lib()
    views = [eval_tail._view3(t)[0] for t in (orig, final, seg, mask)]
    diffm = torch.empty(D, H, W, device="cuda")
    filt = torch.empty(D, H, W, device="cuda")
    sums = torch.zeros(7, dtype=torch.float64, device="cuda")
    rows = []

    def add(name, us, nbytes):
        rows.append((name, us, nbytes))
        print(f"  {name:34s} {us:9.1f} us   {nbytes / 1e6:8.2f} MB algorithmic   {nbytes / us / 1e3:8.1f} GB/s", flush=True)

    add("residual_erode (4 reads, 1 write)", ev_time(lambda: check(L.cddpm_residual_erode(
        ctypes.byref(views[0]), ctypes.byref(views[1]), ctypes.byref(views[2]), ctypes.byref(views[3]), H, W, D, W // 25, 1,
        ptr(diffm), ptr(sums), s), "residual")), 5 * 4 * n)
    add("median3d k=5", ev_time(lambda: check(L.cddpm_median3d(ptr(diffm), ptr(filt), H, W, D, 5, s), "median")), 2 * 4 * n)
    mx = torch.zeros(1, device="cuda")
    add("max", ev_time(lambda: check(L.cddpm_max(ptr(filt), n, ptr(mx), s), "max")), 4 * n)
    q = (ctypes.c_float * 2)(0.01, 0.02)
    cnt = torch.zeros(5, dtype=torch.int64, device="cuda")
    add("threshold_counts (2 thresholds)", ev_time(lambda: check(L.cddpm_threshold_counts(
        ptr(filt), ctypes.byref(views[2]), H, W, D, q, 2, ptr(cnt), s), "counts")), 8 * n)
    tm = torch.empty(D, H, W, dtype=torch.uint8, device="cuda")
    thr = float(filt.max().item()) * 0.3
    add("threshold_mask", ev_time(lambda: check(L.cddpm_threshold_mask(ptr(filt), n, thr, ptr(tm), s), "tmask")), 5 * n)
    tf = torch.empty_like(tm)
    add("filter_small_components", ev_time(lambda: check(L.cddpm_filter_small_components(ptr(tm), ptr(tf), H, W, D, 7, s),
                                                         "cc")), 2 * n)
    cc = torch.zeros(3, dtype=torch.int64, device="cuda")
    add("confusion_counts", ev_time(lambda: check(L.cddpm_confusion_counts(ptr(tf), ctypes.byref(views[2]), H, W, D, ptr(cc), s),
                                                  "conf")), 5 * n)
    nb = int(L.cddpm_hausdorff_workspace_bytes(H, W, D))
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    res = torch.empty(4, dtype=torch.int64, device="cuda")
    add("hausdorff (edges + 3-axis EDT x2)", ev_time(lambda: check(L.cddpm_hausdorff(
        ptr(tf), ctypes.byref(views[2]), H, W, D, ptr(ws), nb, ptr(res), s), "haus")), 5 * n + 2 * 4 * 4 * 2 * n)
    rws = torch.zeros(H, 4, dtype=torch.int64, device="cuda")
    rsum = torch.zeros(H, dtype=torch.float64, device="cuda")
    add("row_stats", ev_time(lambda: check(L.cddpm_row_stats(ptr(filt), ctypes.byref(views[2]), ctypes.byref(views[3]), H, W, D,
                                                             thr, ptr(rws), ptr(rsum), s), "rows")), 12 * n)
    nbr = int(L.cddpm_ranking_workspace_bytes(n))
    wsr = torch.empty(nbr, dtype=torch.uint8, device="cuda")
    rr = torch.empty(2, dtype=torch.float64, device="cuda")
    add("ranking_metrics (radix sort + scans)", ev_time(lambda: check(L.cddpm_ranking_metrics(
        ptr(filt), ctypes.byref(views[2]), H, W, D, ptr(wsr), nbr, ptr(rr), s), "rank")), 8 * n)
    print(f"  sum of tail kernels: {sum(r[1] for r in rows):.1f} us per volume", flush=True)


if __name__ == "__main__":
    main()
