"""BASELINE configs[3] check on real GPUs (run under torchrun on N GPUs): the volume-sharded test sweep — volumes dealt
round-robin to ranks, NCCL all-gather of the per-volume results, all-reduced counts for the global Dice threshold —
reports exactly what one process reports on all volumes.  The simplex noise of a volume is seeded from its index here
(the reference draws from numpy's global stream, which a sharded run cannot share), nothing else differs from
`cddpm.sweep.test_sweep`.
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sweep_check.py [n_volumes=8]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "conditioned-diffusion-models-uad_b200")]
import bench  # noqa: E402
from cddpm import eval_tail, sweep  # noqa: E402
from cddpm.ddpm_2d import DDPM_2D  # noqa: E402

KEYS = ("IDs", "DiceScorePerVol", "BestDicePerVol", "BestThresholdPerVol", "AUCPerVol", "AUPRCPerVol", "HausPerVol",
        "TPPerVol", "FPPerVol", "TNPerVol", "FNPerVol", "AnomalyScoreRecoPerVol", "AnomalyScoreRegPerVol",
        "l1recoErrorAll", "lesionSizePerVol")


def main():
    nvol = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234)  # identical replicas
    cfg = bench.model_cfg()
    cfg["force_num_eval_slices"] = False
    model = DDPM_2D(cfg, prefix="c/")
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
            v = bench.synthetic_volume(i, 50)
            out.append({"Dataset": ["Brats21"], "vol": {"data": v["vol"]}, "vol_orig": {"data": v["vol"]},
                        "seg_orig": {"data": v["seg_orig"]}, "mask_orig": {"data": v["mask_orig"]}, "seg_available": True,
                        "ID": [f"v{i}"], "stage": stage, "label": torch.tensor([1])})
        return out

    sets = {"Datamodules_eval.Brats21": (loader("val", 0), loader("test", 100))}
    # first pass: plans, graph captures, NCCL communicator warm-up; the second (identical) pass is the one timed and
    # compared (host wall clock, barrier on both sides)
    sweep.test_sweep(model, sets, pickle_preds=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    preds, logs = sweep.test_sweep(model, sets, pickle_preds=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_sharded = time.perf_counter() - t0
    if rank == 0:
        # the same sweep in this process alone: no sharding, no collectives
        saved = (sweep._world, eval_tail._dist_sum, eval_tail._dist_max)
        sweep._world = lambda: (0, 1)
        eval_tail._dist_sum = eval_tail._dist_max = lambda t: None
        t0 = time.perf_counter()
        preds1, logs1 = sweep.test_sweep(model, sets, pickle_preds=False)
        torch.cuda.synchronize()
        t_serial = time.perf_counter() - t0
        sweep._world, eval_tail._dist_sum, eval_tail._dist_max = saved
        bad = 0
        for stage in ("val", "test"):
            a, b = preds[stage]["Datamodules_eval.Brats21"], preds1[stage]["Datamodules_eval.Brats21"]
            for k in KEYS:
                xa, xb = list(a[k]), list(b[k])
                same = len(xa) == len(xb) and all(x == y or (x != x and y != y) for x, y in zip(xa, xb))
                if not same:
                    bad += 1
                    print(f"MISMATCH {stage} {k}: {xa} vs {xb}")
            for k in ("DicePerVolMean", "AUPRCPerVolMean", "HausPerVolMean"):
                if not (a[k] == b[k] or (a[k] != a[k] and b[k] != b[k])):
                    bad += 1
                    print(f"MISMATCH {stage} {k}: {a[k]} vs {b[k]}")
        v = preds["val"]["Datamodules_eval.Brats21"]
        print(f"world={world}: {2 * nvol} volumes (val + test) sharded in {t_sharded * 1e3:.1f} ms "
              f"({2 * nvol / t_sharded:.1f} volumes/s), one process {t_serial * 1e3:.1f} ms ({2 * nvol / t_serial:.1f} volumes/s)")
        print(f"val DicePerVolMean {v['DicePerVolMean']:.6f}, AUPRCPerVolMean {v['AUPRCPerVolMean']:.6f}, "
              f"HausPerVolMean {v['HausPerVolMean']:.4f}")
        print("sharded == serial: " + ("OK (all compared entries identical)" if bad == 0 else f"{bad} MISMATCHES"))
        if bad:
            raise SystemExit(1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
