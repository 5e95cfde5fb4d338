"""Bring-up diagnostics for the tcgen05 conv kernel (not a pytest file).  Runs cases of increasing complexity and
prints where errors sit (which pixel rows, which channel blocks), so one GPU round trip tells apart descriptor,
swizzle, pipeline and epilogue mistakes.   python tests/diag_conv.py"""
import os
import sys
import traceback

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "conditioned-diffusion-models-uad_b200"))
from cddpm import ops  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def case(name, B, H, W, cin, k, cout, out_f32=True):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, cin, H, W, device="cuda", generator=g)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
    xq, wq = x.bfloat16().float(), w.bfloat16().float()
    ref = F.conv2d(xq, wq, None, padding=k // 2)
    wp = ops.pack_conv_weight(wq, [cin])
    out = ops.conv_igemm([xq.permute(0, 2, 3, 1).contiguous().bfloat16()], [k * k], wp, None, None, out_f32=out_f32)
    torch.cuda.synchronize()
    got = out.float().permute(0, 3, 1, 2)
    err = (got - ref).abs()
    m = err.max().item()
    ok = m < 2e-2 * max(1.0, ref.abs().max().item()) * (1e-2 if out_f32 else 1.0)
    print(f"[{name}] B{B} {H}x{W} cin{cin} k{k} cout{cout}: max_err={m:.4g} ref_max={ref.abs().max().item():.3g} "
          f"{'OK' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        bad = err > 1e-3
        print(f"   bad fraction {bad.float().mean().item():.4f}; got zero fraction {(got == 0).float().mean().item():.4f}; "
              f"nan {torch.isnan(got).any().item()}")
        per_c = err.amax(dim=(0, 2, 3))
        print("   max err per 8-channel block:", [round(v, 3) for v in per_c.view(-1, 8).amax(1).tolist()][:32])
        per_y = err.amax(dim=(0, 1, 3))
        print("   max err per row y:", [round(v, 3) for v in per_y.tolist()][:32])
        per_x = err.amax(dim=(0, 1, 2))
        print("   max err per col x:", [round(v, 3) for v in per_x.tolist()][:32])
        per_n = err.amax(dim=(1, 2, 3))
        print("   max err per image:", [round(v, 3) for v in per_n.tolist()])
        print("   sample got/ref:", got[0, :4, 0, :4].tolist(), ref[0, :4, 0, :4].tolist())
    return ok


def main():
    print("device:", torch.cuda.get_device_name(0), flush=True)
    cases = [
        ("1x1 one kstep", 2, 8, 8, 64, 1, 32),
        ("1x1 two ksteps", 2, 8, 8, 128, 1, 32),
        ("1x1 N=128", 2, 8, 8, 128, 1, 128),
        ("1x1 box16x8", 1, 8, 16, 64, 1, 64),
        ("3x3 small", 2, 8, 8, 64, 3, 64),
        ("3x3 box16x8", 1, 16, 16, 64, 3, 64),
        ("3x3 128@48", 1, 48, 48, 128, 3, 128),
        ("3x3 256@24 bf16 out", 2, 24, 24, 256, 3, 256, False),
        ("3x3 256@96 many tiles", 4, 96, 96, 256, 3, 256),
    ]
    allok = True
    for c in cases:
        try:
            allok &= case(*c)
        except Exception:
            traceback.print_exc()
            allok = False
            break
    print("DIAG", "PASS" if allok else "FAIL", flush=True)
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
