"""GPU parity of the hand-written TRAINING-mode condition encoder (reference:
DDPM_2D.training_step -> self(input) with the module in train(), DDPM_2D.py:101-122, spark/resnet.py:13-46, and
loss.backward() through it) against torch autograd over the same module's library path in fp32
(cddpm/encoder.py:_encoder_train_eager - the round-1 path, itself pinned to the live reference's training golden).

The engine computes every convolution with bf16 operands (fp32 accumulation; raw outputs, BatchNorm statistics and
normalisation in fp32).  The comparison is against plain fp32 autograd, and the bound is the deviation of torch's OWN
bf16-autocast path on the same instance: a ResNet-50 with batch-statistics BatchNorm is ill-conditioned at random init
(TF32 convolutions alone move its gradients by 8-17 %), and a reference that rounds where the engine rounds agrees no
better than fp32 does (measured: the differences are amplified rounding noise, not rounding structure).  The
weight-gradient GEMM is additionally tested on its own against a matmul."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


class Cfg(dict):
    __getattr__ = dict.get


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def _cos(a, b):
    return (torch.dot(a.flatten(), b.flatten()) / (a.norm() * b.norm()).clamp_min(1e-30)).item()


def _encoder(mode, drop=0.0, seed=3, bn3=float(os.environ.get("CDDPM_TEST_BN3", "0.7"))):
    from cddpm.encoder import get_encoder

    cfg = Cfg(imageDim=[192, 192, 100], rescaleFactor=2, backbone="Spark_Encoder_2D", version="resnet50", cond_dim=128,
              encoder_train_dtype=mode, encoder_drop_path_rate=drop)
    torch.manual_seed(seed)
    enc, _ = get_encoder(cfg)
    with torch.no_grad():
        for n, p in enc.named_parameters():
            if n.endswith("bn3.weight"):
                p.fill_(bn3)  # timm's zero_init_last would silence every residual branch
            elif n.endswith(".bias") and "bn" in n:
                p.normal_(0, 0.1)
    return enc.cuda().train()


@pytest.mark.parametrize("M,Cout,K", [(576, 64, 64), (36864, 64, 576), (2304, 256, 1024), (18, 512, 4608), (1000, 2048, 512)])
def test_flat_wgrad_matches_matmul(M, Cout, K):
    from cddpm._lib import check, current_stream, lib, ptr

    g = torch.Generator(device="cuda").manual_seed(M + Cout)
    dy = torch.randn(M, Cout, device="cuda", generator=g).to(torch.bfloat16)
    x = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    dw = torch.zeros(Cout, K, device="cuda")
    check(lib().cddpm_flat_wgrad(ptr(dy), ptr(x), M, Cout, K, ptr(dw), current_stream()), "cddpm_flat_wgrad")
    ref = dy.float().t() @ x.float()
    err = (dw - ref).abs().max().item()
    print(f"flat wgrad M={M} Cout={Cout} K={K}: max-abs {err:.3g} (ref max {ref.abs().max().item():.3g})")
    assert err <= 2e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("bn3", [0.1, 0.3])
def test_training_encoder_forward_backward_vs_autograd(bn3):
    """Residual-branch gain bn3 sets the conditioning of the instance: at 0.1 TF32 convolutions move the gradients by
    8 %, at 0.3 by 17 % (measured, tools/diag_encoder_train.py); the engine must stay inside the deviation of torch's
    own bf16 autocast in both, and at 0.1 every gradient must still point the fp32 way (cosine >= 0.93)."""
    eng = _encoder("b200", bn3=bn3)
    ref = _encoder("fp32", bn3=bn3)
    ref.load_state_dict(eng.state_dict())
    os.environ["CDDPM_ENCODER_GRAPH"] = "0"
    try:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        g = torch.Generator().manual_seed(1)
        x = torch.rand(16, 1, 96, 96, generator=g).cuda()
        w = torch.randn(16, 128, generator=g).cuda()
        out_e = eng(x)
        out_r = ref(x)
        lib16 = _encoder("bf16", bn3=bn3)  # torch's own bf16 autocast over cuDNN: the yardstick for "bf16-level agreement"
        lib16.load_state_dict(eng.state_dict())
        out_l = lib16(x)
        rel, rel_l = _rel(out_e.detach(), out_r.detach()), _rel(out_l.detach(), out_r.detach())
        print(f"training-mode features vs fp32: engine rel-L2 {rel:.3g}, torch bf16 autocast {rel_l:.3g}")
        assert rel <= max(1e-2, 1.2 * rel_l)
        (out_e * w).sum().backward()
        (out_r * w).sum().backward()
        (out_l * w).sum().backward()
        torch.cuda.synchronize()
    finally:
        os.environ.pop("CDDPM_ENCODER_GRAPH", None)
    # running statistics: same update rule (momentum 0.1, unbiased variance), same counter
    sd_e, sd_r = eng.state_dict(), ref.state_dict()
    for k in sd_e:
        if "running_mean" in k or "running_var" in k:
            assert _rel(sd_e[k], sd_r[k]) <= 2e-2, k
        if "num_batches_tracked" in k:
            assert int(sd_e[k]) == int(sd_r[k]) == 1, k
    # gradients against fp32 autograd: a ResNet-50 with batch-statistics BatchNorm at random init is ill-conditioned
    # (TF32 convolutions alone move its gradients by 8-17 %), so the bound is the deviation of torch's own bf16 autocast
    pe, pr, pl = dict(eng.named_parameters()), dict(ref.named_parameters()), dict(lib16.named_parameters())
    rel_e = sorted(_rel(pe[n].grad, pr[n].grad) for n in pe)
    rel_l = sorted(_rel(pl[n].grad, pr[n].grad) for n in pe)
    med_e, med_l = rel_e[len(rel_e) // 2], rel_l[len(rel_l) // 2]
    print(f"gradients vs fp32 autograd, median / max rel-L2: engine {med_e:.3g} / {rel_e[-1]:.3g}, "
          f"torch bf16 autocast {med_l:.3g} / {rel_l[-1]:.3g}")
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in eng.parameters())
    assert med_e <= 1.2 * med_l and rel_e[-1] <= 1.2 * rel_l[-1]
    if bn3 <= 0.1:
        worst_cos = min(_cos(pe[n].grad, pr[n].grad) for n in pe)
        print(f"worst cosine vs fp32 autograd: {worst_cos:.4f}")
        assert worst_cos >= 0.93


def test_training_encoder_drop_path_and_eval_after_training():
    enc = _encoder("b200", drop=0.5)
    x = torch.rand(8, 1, 96, 96, device="cuda")
    a = enc(x).detach().clone()
    b = enc(x).detach().clone()
    assert (a - b).abs().max().item() > 1e-4  # fresh per-sample masks on every call
    enc.encoder.drop_path_rate = 0.0
    c, d = enc(x).detach().clone(), enc(x).detach().clone()
    assert (c - d).abs().max().item() <= 1e-5
    # gradients flow with masks on, and the eval engine sees the statistics the training engine updated in place
    enc.encoder.drop_path_rate = 0.5
    enc(x).square().mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in enc.parameters())
    enc.eval()
    from cddpm.encoder import get_encoder

    with torch.no_grad():
        y = enc(x).clone()
    fresh, _ = get_encoder(Cfg(imageDim=[192, 192, 100], rescaleFactor=2, backbone="Spark_Encoder_2D", version="resnet50",
                               cond_dim=128))
    fresh.load_state_dict(enc.state_dict(), strict=True)
    fresh = fresh.cuda().eval()
    with torch.no_grad():
        assert torch.equal(y, fresh(x))


def test_training_encoder_graph_replay_follows_the_callers_buffers():
    """The launch lists are replayed as CUDA graphs keyed by every pointer they bind (csrc/resnet_train.cu:run_list): a
    key is run eagerly when first seen, captured the second time, replayed from the third.  Two inputs held alive at
    once (two addresses) are alternated: for each, the eager, the capturing and the replayed call must return the same
    features and the same gradients (the forward and the BatchNorm reductions are order-fixed; the weight-gradient GEMM
    accumulates with atomics, hence a tolerance there)."""
    enc = _encoder("b200", drop=0.0)
    g = torch.Generator(device="cuda").manual_seed(11)
    xs = [torch.rand(8, 1, 96, 96, device="cuda", generator=g) for _ in range(2)]
    w = torch.randn(8, 128, device="cuda", generator=g)
    first = {}
    for rnd in range(4):
        for i, x in enumerate(xs):
            for p in enc.parameters():
                p.grad = None
            out = enc(x)
            (out * w).sum().backward()
            torch.cuda.synchronize()
            grads = torch.cat([p.grad.flatten() for p in enc.parameters()])
            if rnd == 0:
                first[i] = (out.detach().clone(), grads.clone())
                continue
            assert torch.equal(out.detach(), first[i][0]), f"input {i}, call {rnd}: features differ from the eager call"
            rel = ((grads - first[i][1]).norm() / first[i][1].norm()).item()
            assert rel <= 1e-4, f"input {i}, call {rnd}: gradients differ from the eager call (rel {rel:.3g})"
    assert not torch.equal(first[0][0], first[1][0])
