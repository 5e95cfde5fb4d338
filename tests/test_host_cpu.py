"""CPU tests (no GPU) of the host side: the C-ABI library loads and exports every symbol include/cddpm_b200.h declares,
the drop-in modules keep the reference's state_dict layout, the product path fails loudly without CUDA, host-only logic
(permutation LCG, bisection control flow, eval dictionary), and the world-size-2 gloo path of the global threshold."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "conditioned-diffusion-models-uad_b200")
GOLD = os.path.join(ROOT, "tests", "golden")


class Cfg(dict):
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


def _cfg(**over):
    c = Cfg(imageDim=[192, 192, 100], rescaleFactor=2, unet_dim=128, dim_mults=[1, 2, 2], condition=True,
            backbone="Spark_Encoder_2D", version="resnet50", cond_dim=128, noisetype="simplex", noise_ensemble=True,
            test_timesteps=500, lr=1e-4, resizedEvaluation=True, erodeBrainmask=True, medianFiltering=True,
            saveOutputImages=False, evalSeg=True, threshold="auto")
    c.update(over)
    return c


def test_library_exports_every_declared_symbol():
    import ctypes

    from cddpm import _lib

    header = open(os.path.join(ROOT, "include", "cddpm_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(cddpm_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 40
    lib = _lib.lib()  # resolves every symbol in the ctypes table
    table = set(_lib._signatures(ctypes).keys())
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in table, f"{name} declared in the header but missing from the ctypes table"
    assert table <= set(declared)
    assert b"sm_100a" in lib.cddpm_version()


def test_no_oracle_import_in_product():
    """The product path must never route through the oracle (test infrastructure)."""
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(base, f)


def test_dropin_state_dict_layout_and_roundtrip():
    from oracle import diffusion_port, resnet_port, unet_port
    from oracle.weights import make_state_dict
    from src.models.DDPM_2D import DDPM_2D

    m = DDPM_2D(_cfg(), prefix="p/")
    exp = (["encoder.encoder." + k for k, _ in resnet_port.param_shapes(128)]
           + ["diffusion." + k for k in diffusion_port.BUFFER_NAMES]
           + ["diffusion.model." + k for k, _ in unet_port.param_shapes(unet_port.UNetSpec())])
    sd = m.state_dict()
    assert list(sd.keys()) == exp and len(exp) == 649
    shapes = dict(unet_port.param_shapes(unet_port.UNetSpec()))
    for k, s in shapes.items():
        assert tuple(sd["diffusion.model." + k].shape) == s
    # zero-initialised output convolutions, as in the reference (zero_module)
    assert float(sd["diffusion.model.out.2.weight"].abs().sum()) == 0.0
    assert float(sd["diffusion.model.middle_block.1.proj_out.weight"].abs().sum()) == 0.0
    full = {k: torch.randn_like(v) if v.dtype.is_floating_point else v for k, v in sd.items()}
    m.load_state_dict(full, strict=True)
    for k, v in m.state_dict().items():
        assert torch.equal(v, full[k])
    assert m.test_timesteps == 500 and m.prefix == "p/"
    m.update_prefix("q/")
    assert m.prefix == "q/"
    assert isinstance(m.configure_optimizers(), torch.optim.Adam)


def test_unconditioned_layout():
    from oracle import unet_port
    from src.models.DDPM_2D import DDPM_2D

    m = DDPM_2D(_cfg(condition=False), prefix=None)
    keys = [k for k in m.state_dict() if k.startswith("diffusion.model.")]
    assert keys == ["diffusion.model." + k for k, _ in unet_port.param_shapes(unet_port.UNetSpec(num_classes=None))]
    assert not hasattr(m, "encoder")
    assert m(torch.zeros(1, 1, 96, 96)) is None


def test_product_fails_loudly_without_cuda():
    from cddpm import CddpmError
    from src.models.DDPM_2D import DDPM_2D
    from src.utils.utils_eval import apply_3d_median_filter

    m = DDPM_2D(_cfg(), prefix="p/").eval()
    with torch.no_grad():
        with pytest.raises(CddpmError):
            m(torch.zeros(1, 1, 96, 96))
        with pytest.raises(CddpmError):
            m.diffusion(torch.zeros(1, 1, 96, 96), cond=torch.zeros(1, 128), t=10, noise=torch.zeros(1, 1, 96, 96))
    with pytest.raises((CddpmError, RuntimeError, AssertionError)):
        apply_3d_median_filter(torch.zeros(8, 8, 8))
    with pytest.raises(CddpmError):  # the training step has no CPU path either
        m.train()
        m.training_step({"vol": {"data": torch.zeros(1, 1, 96, 96, 1)}}, 0)


def test_schedule_buffers_are_reference_bit_exact():
    from cddpm.diffusion import GaussianDiffusion

    d = GaussianDiffusion(torch.nn.Identity(), image_size=(96, 96), timesteps=1000, objective="pred_x0", channels=1)
    g = np.load(os.path.join(GOLD, "schedule.npz"))
    sd = d.state_dict()
    assert list(sd.keys()) == list(g.keys())
    for k in sd:
        assert np.array_equal(sd[k].numpy(), g[k]), k
    with pytest.raises(ValueError):
        GaussianDiffusion(torch.nn.Identity(), image_size=(96, 96), beta_schedule="nope")
    with pytest.raises(AssertionError):
        GaussianDiffusion(torch.nn.Identity(), image_size=(96, 96), objective="pred_v")


def test_permutation_matches_oracle_and_rng_stream():
    from cddpm import noise
    from oracle import simplex_port

    for seed in (1, -5, 9999999999, -10000000000):
        assert list(noise.permutation(seed)) == simplex_port.permutation_from_seed(seed).tolist()
    # two draws from numpy's global RNG per gen_noise call, like the reference
    np.random.seed(3)
    a, b = noise._new_seed(), noise._new_seed()
    np.random.seed(3)
    assert (simplex_port.draw_seed(), simplex_port.draw_seed()) == (a, b)


def test_bisection_control_flow_matches_oracle():
    from cddpm import eval_tail
    from oracle import tail_port

    rng = np.random.default_rng(0)
    x = rng.random(4000).astype(np.float32)
    x[rng.random(4000) < 0.5] = 0
    y = rng.random(4000) < 0.2 + 0.5 * x

    def counts(qs):
        return [int(y.sum()), int((x > qs[0]).sum()), int(((x > qs[0]) & y).sum()), int((x > qs[1]).sum()),
                int(((x > qs[1]) & y).sum())]

    assert eval_tail._bisect(counts, (0, np.max(x)), 10) == tail_port.find_best_val(x, y, (0, np.max(x)), 10)
    ed = eval_tail.get_eval_dictionary()
    assert len(ed) == 108 and all(v == [] for v in ed.values())
    assert "AnomalyScoreRegPerVol" in ed and "DiceScorePerSlice" in ed


def test_host_side_small_rankings_match_sklearn():
    from sklearn.metrics import auc, average_precision_score, roc_curve

    from cddpm import eval_tail

    rng = np.random.default_rng(2)
    s = rng.random(96)
    s[::5] = 0.0
    lab = (rng.random(96) < 0.3).astype(int)
    a, fpr, tpr, thr = eval_tail.compute_roc(s, lab)
    f2, t2, th2 = roc_curve(lab, s, pos_label=1)
    assert np.allclose(fpr, f2) and np.allclose(tpr, t2) and abs(a - auc(f2, t2)) < 1e-12
    p, *_ = eval_tail.compute_prc(s, lab)
    assert abs(p - average_precision_score(lab, s)) < 1e-12


_GLOO_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path[:0] = [os.environ["CDDPM_ROOT"], os.path.join(os.environ["CDDPM_ROOT"], "conditioned-diffusion-models-uad_b200")]
from cddpm import eval_tail
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
rng = np.random.default_rng(0)
x = rng.random(8000).astype(np.float32); x[rng.random(8000) < 0.5] = 0
y = rng.random(8000) < 0.2 + 0.5 * x
lo, hi = (0, 4000) if rank == 0 else (4000, 8000)   # each rank owns half of the validation voxels
def counts(qs):
    xs, ys = x[lo:hi], y[lo:hi]
    c = torch.tensor([int(ys.sum()), int((xs > qs[0]).sum()), int(((xs > qs[0]) & ys).sum()),
                      int((xs > qs[1]).sum()), int(((xs > qs[1]) & ys).sum())], dtype=torch.int64)
    eval_tail._dist_sum(c)       # the all-reduce the sharded sweep uses (NCCL on GPUs, gloo here)
    return c.numpy()
top = torch.tensor([float(np.max(x[lo:hi]))]); eval_tail._dist_max(top)
got = eval_tail._bisect(counts, (0, np.float32(top.item())), 10)
from oracle import tail_port
want = tail_port.find_best_val(x, y, (0, np.max(x)), 10)
assert got == want, (got, want)
dist.destroy_process_group()
print("ok", rank)
"""


def test_global_threshold_world_size_2_gloo(tmp_path):
    """The sharded sweep's global Dice threshold (all-reduced counts) equals the serial search over all voxels."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, CDDPM_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29513", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


_GLOO_TRAIN_WORKER = r"""
import os, sys
import torch, torch.distributed as dist
sys.path[:0] = [os.environ["CDDPM_ROOT"], os.path.join(os.environ["CDDPM_ROOT"], "conditioned-diffusion-models-uad_b200")]
from cddpm.dist_train import sync_gradients
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.manual_seed(0)
m = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4), torch.nn.Linear(4, 2))
# the first two layers get gradients that are views of ONE flat buffer (what the UNet engine's backward returns),
# the last one a loose tensor (what torch autograd returns for the encoder)
flat = torch.arange(8 * 16 + 16 + 16 * 4 + 4, dtype=torch.float32) * (rank + 1)
off = 0
for p in list(m[0].parameters()) + list(m[1].parameters()):
    p.grad = flat[off:off + p.numel()].view(p.shape); off += p.numel()
for p in m[2].parameters():
    p.grad = torch.full_like(p, float(rank + 1))
calls = sync_gradients(m)
assert calls == 2, calls
want = torch.arange(flat.numel(), dtype=torch.float32) * 1.5   # mean of x1 and x2
assert torch.equal(flat, want)
assert torch.equal(m[0].weight.grad.flatten(), want[:128])
for p in m[2].parameters():
    assert torch.equal(p.grad, torch.full_like(p, 1.5))
dist.destroy_process_group()
print("ok", rank)
"""


def test_gradient_sync_world_size_2_gloo(tmp_path):
    """Data-parallel training: the engine's flat gradient buffer travels as one all-reduce, the rest as one more."""
    script = tmp_path / "train_worker.py"
    script.write_text(_GLOO_TRAIN_WORKER)
    env = dict(os.environ, CDDPM_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


_GLOO_SWEEP_WORKER = r"""
import os, sys, pickle
import torch, torch.distributed as dist
sys.path[:0] = [os.environ["CDDPM_ROOT"], os.path.join(os.environ["CDDPM_ROOT"], "conditioned-diffusion-models-uad_b200")]
from cddpm import sweep
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)

class Stub(torch.nn.Module):
    # the hooks of the LightningModule drop-in, with per-volume and per-slice lists like get_eval_dictionary()
    def __init__(self):
        super().__init__(); self.w = torch.nn.Parameter(torch.zeros(1)); self.threshold = {}
    def on_test_start(self):
        self.eval_dict = {"IDs": [], "DiceScorePerVol": [], "PerSlice": []}
    def test_step(self, batch, idx):
        self.stage = batch["stage"]
        self.eval_dict["IDs"].append(batch["ID"][0])
        self.eval_dict["DiceScorePerVol"].append(float(batch["vol"].sum()))
        self.eval_dict["PerSlice"].extend([idx] * (1 + idx % 2))
    def on_test_end(self):
        v = self.eval_dict["DiceScorePerVol"]
        self.eval_dict["DiceScorePerVolMean"] = sum(v) / len(v)
        if "val" in self.stage: self.threshold["total"] = max(v)

def loader(stage, n):
    return [{"ID": [f"{stage}{i}"], "vol": torch.full((2, 2), float(i)), "stage": stage} for i in range(n)]

m = Stub()
preds, logs = sweep.test_sweep(m, {"Datamodules_eval.Brats21": (loader("val", 5), loader("test", 7))}, fold=0,
                               log_dir=os.environ["OUT_DIR"])
ev = preds["test"]["Datamodules_eval.Brats21"]
assert ev["IDs"] == [f"test{i}" for i in range(7)], ev["IDs"]                       # loader order restored
assert ev["DiceScorePerVol"] == [4.0 * i for i in range(7)]
assert sorted(ev["PerSlice"]) == sorted(sum(([i] * (1 + i % 2) for i in range(7)), []))
assert abs(ev["DiceScorePerVolMean"] - 12.0) < 1e-9                                  # mean over ALL volumes
assert m.threshold["total"] == 16.0                                                  # val stage: 4 * max(0..4)
assert logs["1/Datamodules_eval.Brats21/test/DiceScorePerVolMean"] == ev["DiceScorePerVolMean"]
assert "1/Datamodules_eval.Brats21/val/DiceScorePerVolMean" in logs and not any(type(v) is list for v in logs.values())
dist.barrier()
if rank == 0:
    with open(os.path.join(os.environ["OUT_DIR"], "1_preds_dict.pkl"), "rb") as f:
        assert set(pickle.load(f)) == {"val", "test"}
dist.destroy_process_group()
print("ok", rank)
"""


def test_sharded_test_sweep_world_size_2_gloo(tmp_path):
    """train.py:182-237 without a Trainer: volumes dealt round-robin to 2 ranks, per-volume lists gathered back in
    loader order, means over all volumes, val -> test threshold hand-off, preds_dict.pkl on rank 0."""
    script = tmp_path / "sweep_worker.py"
    script.write_text(_GLOO_SWEEP_WORKER)
    env = dict(os.environ, CDDPM_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29519", WORLD_SIZE="2",
               OUT_DIR=str(tmp_path / "logs"))
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py's stdout contract: ONE JSON line on the process's stdout, everything else (library banners, notes) on
    stderr - main() keeps a private copy of the real stdout and points descriptor 1 at stderr, because NCCL writes its
    banner to descriptor 1 from native code.  Exercised here through the CPU-runnable reference arm (the oracle port
    timed on the host), which must also carry the keys the driver pairs with our arm's line."""
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "slices/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("configs[1]")


def test_compose_grid_port_layout_and_colour_tables():
    """oracle.tail_port.compose_grid_port (the checker of cddpm_compose_grid / log_images): panel order, rot90(., 3),
    gray normalisation to the panel's own range, inferno end points and the constant-image case."""
    import numpy as np

    from oracle.tail_port import compose_grid_port

    H, W = 4, 6
    panels = np.zeros((4, H, W), dtype=np.float32)
    panels[0, 0, 0] = 2.0          # top-left of the original -> after rot90(., 3): row 0, column H-1
    panels[2] = np.linspace(0, 1, H * W, dtype=np.float32).reshape(H, W)
    panels[3] = 0.5                # constant panel: bottom of the gray table
    ranges = np.array([[0, 2], [0, 0], [0, 1], [0.5, 0.5]], dtype=np.float32)
    img = compose_grid_port(panels, ranges)
    assert img.shape == (W, 4 * H, 3) and img.dtype == np.uint8
    assert tuple(img[0, H - 1]) == (255, 255, 255) and img[:, :H].sum() == 3 * 255
    assert img[:, H:2 * H].max() == 0 and img[:, 3 * H:].max() == 0
    rot = np.rot90(panels[2], 3)
    lo, hi = np.unravel_index(rot.argmin(), rot.shape), np.unravel_index(rot.argmax(), rot.shape)
    assert tuple(img[lo[0], 2 * H + lo[1]]) == (0, 0, 4) and tuple(img[hi[0], 2 * H + hi[1]]) == (252, 255, 164)
