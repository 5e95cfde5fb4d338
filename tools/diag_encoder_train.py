"""Diagnostic (development aid): training-mode encoder - hand-written engine vs the library path in fp32 / bf16 autocast."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "conditioned-diffusion-models-uad_b200")]
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_encoder_train_gpu import _cos, _encoder, _rel  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    os.environ["CDDPM_ENCODER_GRAPH"] = "0"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    encs = {m: _encoder(m) for m in ("b200", "fp32", "bf16", "tf32")}
    sd = encs["b200"].state_dict()
    for e in encs.values():
        e.load_state_dict(sd)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(B, 1, 96, 96, generator=g).cuda()
    w = torch.randn(B, 128, generator=g).cuda()
    outs = {}
    for m, e in encs.items():
        if m == "tf32":
            torch.backends.cudnn.allow_tf32 = True
        o = e(x)
        (o * w).sum().backward()
        outs[m] = o.detach()
        torch.backends.cudnn.allow_tf32 = False
    ref = outs["fp32"]
    for m in ("b200", "bf16", "tf32"):
        print(f"B={B} features {m} vs fp32: rel-L2 {_rel(outs[m], ref):.4g}")
    pr = dict(encs["fp32"].named_parameters())
    for m in ("b200", "bf16", "tf32"):
        pe = dict(encs[m].named_parameters())
        rows = sorted(((_rel(pe[n].grad, pr[n].grad), _cos(pe[n].grad, pr[n].grad), n) for n in pe), reverse=True)
        med = sorted(r for r, _, _ in rows)[len(rows) // 2]
        print(f"   grads {m}: median rel-L2 {med:.4g}; worst {[(round(r, 3), round(c, 4), n) for r, c, n in rows[:4]]}")


if __name__ == "__main__":
    main()
