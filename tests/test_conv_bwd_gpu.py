"""GPU parity of the convolution backward kernels against torch autograd (fp32, TF32 off) on the same 16-bit-rounded
operands: weight gradient on tcgen05 (MN-major operands) and data gradient through the forward kernel over the
transposed panel.  Reference: torch autograd of nn.Conv2d in ResBlock / AttentionBlock (OpenAI_Unet.py:231,257,268,
367,375) as run by DDPM_2D.training_step (DDPM_2D.py:114-138)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _nhwc16(x, dtype):
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


@pytest.mark.parametrize("B,H,W,cins,ksize,cout", [
    (2, 16, 16, [128], 3, 128),
    (3, 24, 24, [256], 3, 256),
    (2, 48, 48, [256, 128], 3, 128),
    (2, 24, 24, [256], 1, 256),
    (1, 32, 32, [128, 128], 1, 256),
    (4, 96, 96, [128], 3, 128),
    # widths that are not a multiple of 16 run on image-interleaved tiles (8 pixels of two consecutive images per row)
    (5, 24, 24, [128, 128], 3, 128),
    (2, 40, 40, [128], 3, 128),
    (9, 24, 24, [128], 1, 128),
])
def test_wgrad_matches_autograd(B, H, W, cins, ksize, cout):
    from cddpm import ops

    _setup()
    dtype = torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H + cout)
    cin = sum(cins)
    x = torch.randn(B, cin, H, W, device="cuda", generator=g).to(dtype).float()
    dy = torch.randn(B, cout, H, W, device="cuda", generator=g).to(dtype).float()
    w = torch.zeros(cout, cin, ksize, ksize, device="cuda", requires_grad=True)
    F.conv2d(x, w, padding=ksize // 2).backward(dy)
    ref = w.grad
    srcs, off = [], 0
    for c in cins:
        srcs.append(_nhwc16(x[:, off:off + c], dtype))
        off += c
    dw = ops.conv_wgrad(srcs, [ksize * ksize] * len(cins), _nhwc16(dy, dtype))
    got = ops.unpack_conv_grad(dw, cout, cin, ksize, cins)
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    tol = 1e-4 * max(ref.abs().max().item(), 1.0) * (B * H * W) ** 0.5 / 16
    assert err <= tol, f"wgrad max err {err:.4g} > {tol:.4g} (ref max {ref.abs().max().item():.4g})"


@pytest.mark.parametrize("B,H,W,cin,ksize,cout", [(2, 16, 16, 128, 3, 256), (2, 24, 24, 256, 1, 128),
                                                    (1, 48, 48, 384, 3, 128)])
def test_dgrad_through_forward_kernel(B, H, W, cin, ksize, cout):
    from cddpm import ops

    _setup()
    dtype = torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(7)
    w = (torch.randn(cout, cin, ksize, ksize, device="cuda", generator=g) / (cout * ksize * ksize) ** 0.5).to(dtype).float()
    dy = torch.randn(B, cout, H, W, device="cuda", generator=g).to(dtype).float()
    x = torch.zeros(B, cin, H, W, device="cuda", requires_grad=True)
    F.conv2d(x, w, padding=ksize // 2).backward(dy)
    ref = x.grad
    wt = ops.pack_conv_weight_t(w, dtype)
    got = ops.conv_igemm([_nhwc16(dy, dtype)], [ksize * ksize], wt, None, None, out_f32=False)
    torch.cuda.synchronize()
    err = (got.float().permute(0, 3, 1, 2) - ref).abs().max().item()
    tol = 2.0 ** -8 * max(ref.abs().max().item(), 1.0) + 1e-4
    assert err <= tol, f"dgrad max err {err:.4g} > {tol:.4g}"
