"""GPU parity of the UNet engine (cddpm_unet_forward through the C ABI) against
  (1) the golden outputs of the live reference UNetModel (tests/golden/unet_*.npz, made by oracle/make_golden.py), and
  (2) the fp32 oracle port evaluated here on CPU, layer by layer (taps), so a failure names the first bad layer.

Tolerance: north_star allows max-abs <= 1e-2 on the reconstruction reco = (model_out + 1) / 2, i.e. 2e-2 on model_out,
for fp16 tensor-core operands (the reference's own AMP dtype, trainer/default.yaml:7) with fp32 accumulation against
the fp32 reference.  bf16 operands (selectable) keep 3 fewer mantissa bits in the residual stream (|h| ~ 8) and are held
to the looser 5e-2 here; see DESIGN.md "Numerics"."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

pytestmark = pytest.mark.gpu

TOL_MODEL_OUT = 2e-2


def _engine(spec, image, dtype=torch.float16):
    from cddpm.engine import UNetEngine

    return UNetEngine(image_size=(image, image), in_channels=1, model_channels=spec.model_channels, out_channels=1,
                      num_res_blocks=spec.num_res_blocks, attention_resolutions=spec.attention_resolutions,
                      channel_mult=spec.channel_mult, num_classes=spec.num_classes, num_head_channels=64, dtype=dtype)


def _load(eng, sd):
    eng.load_state_dict({k: v.cuda() for k, v in sd.items()})


def _layer_report(eng, taps, B):
    rows = []
    for name, ref in taps.items():
        if name == "emb":
            continue
        try:
            got = eng.tap(name, B).cpu()
        except Exception:
            continue
        err = (got - ref).abs().max().item()
        rows.append((name, err, ref.abs().max().item()))
    return rows


def test_engine_param_names_match_reference_layout():
    from oracle import unet_port

    spec = unet_port.UNetSpec()
    eng = _engine(spec, 96)
    names = eng.param_names()
    shapes = unet_port.param_shapes(spec)
    assert [n for n, _ in names] == [k for k, _ in shapes]
    assert [c for _, c in names] == [int(np.prod(s)) for _, s in shapes]


def test_small_unet_matches_golden_and_port():
    from oracle import unet_port
    from oracle.weights import make_state_dict

    spec = unet_port.UNetSpec(model_channels=64, channel_mult=(1, 2), num_res_blocks=1, num_classes=128)
    sd = make_state_dict(unet_port.param_shapes(spec), seed=2)
    g = np.load(os.path.join(GOLD, "unet_small_32.npz"))
    x, t, cond, y = (torch.from_numpy(g[k]) for k in ("x", "t", "cond", "y"))
    taps = {}
    with torch.no_grad():
        port = unet_port.unet_forward(sd, spec, x, t, cond, taps=taps)
    assert (port - y).abs().max().item() < 1e-4  # oracle port == live reference
    eng = _engine(spec, 32)
    _load(eng, sd)
    out = eng.forward(x.cuda(), t.cuda(), cond.cuda()).cpu()
    film_ref = torch.cat([torch.nn.functional.linear(torch.nn.functional.silu(taps["emb"]), sd[k], sd[k.replace("weight", "bias")])
                          for k in sd if k.endswith("emb_layers.1.weight")], dim=1)
    film_err = (eng.film(3).cpu() - film_ref).abs().max().item()
    rows = _layer_report(eng, taps, 3)
    err = (out - y).abs().max().item()
    report = "\n".join(f"  {n:32s} err {e:.4g} (ref max {m:.3g})" for n, e, m in rows)
    print(f"film err {film_err:.3g}\n{report}\nfinal err {err:.4g}")
    assert film_err < 1e-3, f"FiLM projection mismatch {film_err}"
    assert err <= TOL_MODEL_OUT, f"model_out max-abs {err:.4g} > {TOL_MODEL_OUT}\n{report}"


@pytest.mark.parametrize("tag", ["cond", "uncond"])
def test_full_unet_matches_reference_golden(tag):
    from oracle import unet_port
    from oracle.weights import make_state_dict

    spec = unet_port.UNetSpec(num_classes=128 if tag == "cond" else None)
    sd = make_state_dict(unet_port.param_shapes(spec), seed=1)
    g = np.load(os.path.join(GOLD, f"unet_{tag}_96.npz"))
    x, t, y = (torch.from_numpy(g[k]) for k in ("x", "t", "y"))
    cond = torch.from_numpy(g["cond"]).cuda() if tag == "cond" else None
    eng = _engine(spec, 96)
    _load(eng, sd)
    out = eng.forward(x.cuda(), t.cuda(), cond).cpu()
    err = (out - y).abs().max().item()
    print(f"unet {tag} 96x96: max-abs {err:.4g}, mean-abs {(out - y).abs().mean().item():.4g}, ref max {y.abs().max().item():.3g}")
    assert err <= TOL_MODEL_OUT
    # batch-size change re-plans (GroupNorm partial sums are chunked differently, so only fp32 summation order moves)
    out1 = eng.forward(x[:1].cuda(), t[:1].cuda(), cond[:1] if cond is not None else None).cpu()
    assert (out1 - out[:1]).abs().max().item() < 5e-3
    assert (out1 - y[:1]).abs().max().item() <= TOL_MODEL_OUT


def test_bf16_operands_selectable_and_within_looser_bound():
    from oracle import unet_port
    from oracle.weights import make_state_dict

    spec = unet_port.UNetSpec(model_channels=64, channel_mult=(1, 2), num_res_blocks=1, num_classes=128)
    sd = make_state_dict(unet_port.param_shapes(spec), seed=2)
    g = np.load(os.path.join(GOLD, "unet_small_32.npz"))
    x, t, cond, y = (torch.from_numpy(g[k]) for k in ("x", "t", "cond", "y"))
    eng = _engine(spec, 32, dtype=torch.bfloat16)
    _load(eng, sd)
    out = eng.forward(x.cuda(), t.cuda(), cond.cuda()).cpu()
    assert (out - y).abs().max().item() <= 5e-2


def test_missing_parameter_is_an_error():
    from cddpm import CddpmError
    from oracle import unet_port

    spec = unet_port.UNetSpec(model_channels=64, channel_mult=(1, 2), num_res_blocks=1, num_classes=128)
    eng = _engine(spec, 32)
    with pytest.raises(CddpmError):
        eng.forward(torch.zeros(1, 1, 32, 32, device="cuda"), torch.zeros(1, dtype=torch.long, device="cuda"),
                    torch.zeros(1, 128, device="cuda"))


def test_graph_replay_matches_direct_launches():
    """The first forward of a plan launches every kernel directly; later ones replay the captured CUDA graph through
    the engine's staging buffers.  Same kernels, same order: results agree to the order of the GroupNorm atomics."""
    from oracle import unet_port
    from oracle.weights import make_state_dict

    spec = unet_port.UNetSpec()
    sd = make_state_dict(unet_port.param_shapes(spec), seed=3)
    g = torch.Generator().manual_seed(11)
    xs = [torch.rand(2, 1, 96, 96, generator=g).cuda() for _ in range(2)]
    ts = [torch.tensor([499, 17]).cuda(), torch.tensor([3, 250]).cuda()]
    cs = [torch.randn(2, 128, generator=g).cuda() for _ in range(2)]
    eng = _engine(spec, 96)
    _load(eng, sd)
    direct0 = eng.forward(xs[0], ts[0], cs[0]).clone()  # direct launches
    graph1 = eng.forward(xs[1], ts[1], cs[1]).clone()   # capture + first replay
    graph0 = eng.forward(xs[0], ts[0], cs[0]).clone()   # replay with other inputs
    eng2 = _engine(spec, 96)
    _load(eng2, sd)
    direct1 = eng2.forward(xs[1], ts[1], cs[1]).clone()
    assert (graph0 - direct0).abs().max().item() < 1e-4
    assert (graph1 - direct1).abs().max().item() < 1e-4
    assert (graph1 - graph0).abs().max().item() > 1e-3  # the replay really consumed the new inputs


def test_fused_groupnorm_finish_matches_golden():
    """CDDPM_FUSE_GN=1 (experimental, off by default): the out_layers GroupNorm + FiLM + SiLU finished inside the
    producing convolution's epilogue (cross-CTA rendezvous per image).  Same golden vectors, same tolerance, fewer
    launches; a backward on such a plan is refused.  The switch is read once per process, hence the subprocess."""
    import subprocess
    import sys

    code = r'''
import os, sys
import numpy as np, torch
ROOT = sys.argv[1]
sys.path[:0] = [ROOT, os.path.join(ROOT, "conditioned-diffusion-models-uad_b200")]
from cddpm.engine import UNetEngine
from cddpm._lib import CddpmError
from oracle import unet_port
from oracle.weights import make_state_dict
spec = unet_port.UNetSpec(num_classes=128)
sd = make_state_dict(unet_port.param_shapes(spec), seed=1)
g = np.load(os.path.join(ROOT, "tests", "golden", "unet_cond_96.npz"))
x, t, y, cond = (torch.from_numpy(g[k]) for k in ("x", "t", "y", "cond"))
eng = UNetEngine(image_size=(96, 96), in_channels=1, model_channels=128, out_channels=1, num_res_blocks=3,
                 attention_resolutions=(3, 6, 12), channel_mult=(1, 2, 2), num_classes=128, dtype=torch.float16)
eng.load_state_dict({k: v.cuda() for k, v in sd.items()})
outs = [eng.forward(x.cuda(), t.cuda(), cond.cuda()).cpu() for _ in range(3)]  # direct launches, capture, replay
err = max((o - y).abs().max().item() for o in outs)
one = eng.forward(x[:1].cuda(), t[:1].cuda(), cond[:1].cuda()).cpu()
err1 = (one - y[:1]).abs().max().item()
refused = False
try:
    eng.backward(torch.zeros(1, 1, 96, 96, device="cuda"))
except CddpmError:
    refused = True
print("RESULT", err, err1, eng.launches_per_forward, refused)
'''
    env = dict(os.environ, CDDPM_FUSE_GN="1")
    r = subprocess.run([sys.executable, "-c", code, ROOT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][-1].split()
    err, err1, launches, refused = float(line[1]), float(line[2]), int(line[3]), line[4] == "True"
    print(f"fused GroupNorm finish: max-abs {err:.4g} (B=1: {err1:.4g}), {launches} launches per forward")
    assert err <= TOL_MODEL_OUT and err1 <= TOL_MODEL_OUT
    assert launches < 123 and refused
