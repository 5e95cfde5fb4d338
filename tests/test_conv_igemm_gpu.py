"""GPU parity of the tcgen05 implicit-GEMM convolution against torch fp32 convolution (TF32 off) on the same
16-bit-rounded operands.  Shape classes follow SURVEY.md Appendix C (one conditioned UNet forward).

Reference call sites: nn.Conv2d in ResBlock (OpenAI_Unet.py:231,257,268) and the concat feeding every
output block (OpenAI_Unet.py:948)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _nhwc16(x, dtype):
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


def _run_case(B, H, W, cins, ksize, cout, *, bias=True, residual=False, skip_c=None, dtype=torch.bfloat16,
              out_f32=False, seed=0):
    from cddpm import ops

    _setup()
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    cin = sum(cins)
    x = torch.randn(B, cin, H, W, device=dev, generator=g)
    w = torch.randn(cout, cin, ksize, ksize, device=dev, generator=g) / (cin * ksize * ksize) ** 0.5
    b = torch.randn(cout, device=dev, generator=g) if bias else None
    xq = x.to(dtype).float()
    wq = w.to(dtype).float()
    ref = F.conv2d(xq, wq, b, padding=ksize // 2)
    srcs = []
    off = 0
    for c in cins:
        srcs.append(_nhwc16(x[:, off:off + c], dtype))
        off += c
    taps = [ksize * ksize] * len(cins)
    wp = ops.pack_conv_weight(wq, cins, dtype)
    res = None
    if skip_c is not None:
        # fused 1x1 skip branch over a second tensor (ResBlock.skip_connection + out_layers conv in one accumulator)
        xs = torch.randn(B, sum(skip_c), H, W, device=dev, generator=g)
        ws = torch.randn(cout, sum(skip_c), 1, 1, device=dev, generator=g) / sum(skip_c) ** 0.5
        xsq, wsq = xs.to(dtype).float(), ws.to(dtype).float()
        ref = ref + F.conv2d(xsq, wsq)
        off = 0
        for c in skip_c:
            srcs.append(_nhwc16(xs[:, off:off + c], dtype))
            taps.append(1)
            off += c
        wp = torch.cat([wp, ops.pack_conv_weight(wsq, skip_c, dtype)], dim=1).contiguous()
    if residual:
        r = torch.randn(B, cout, H, W, device=dev, generator=g)
        rq = r.to(dtype).float()
        ref = ref + rq
        res = _nhwc16(r, dtype)
    out = ops.conv_igemm(srcs, taps, wp, b, res, out_f32=out_f32)
    torch.cuda.synchronize()
    got = out.float().permute(0, 3, 1, 2)
    err = (got - ref).abs()
    scale = ref.abs().max().item()
    # fp32 accumulation in a different order + one rounding to the 16-bit output type
    eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    tol = (1e-4 if out_f32 else eps) * max(scale, 1.0) + 1e-4
    assert err.max().item() <= tol, (
        f"max err {err.max().item():.4g} > tol {tol:.4g} (ref max {scale:.3g}) at "
        f"{[int(i) for i in (err == err.max()).nonzero()[0]]}"
    )


# (B, H, W, cins, ksize, cout) — the 3x3 and 1x1 GEMM shape classes of one UNet forward
SHAPES = [
    (2, 96, 96, [128], 3, 128),
    (2, 48, 48, [128], 3, 128),
    (2, 48, 48, [128], 3, 256),
    (2, 48, 48, [256], 3, 256),
    (2, 24, 24, [256], 3, 256),
    (1, 96, 96, [256], 3, 256),
    (2, 24, 24, [256, 256], 3, 256),
    (2, 48, 48, [256, 256], 3, 256),
    (1, 48, 48, [256, 128], 3, 256),
    (1, 96, 96, [256, 128], 3, 128),
    (1, 96, 96, [128, 128], 3, 128),
    (2, 48, 48, [128], 1, 256),
    (2, 24, 24, [256, 256], 1, 256),
    (3, 24, 24, [256], 1, 768),  # attention qkv as 1x1 (three N tiles), odd box count (27 boxes -> padded tile)
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"B{s[0]}_{s[1]}x{s[2]}_{'+'.join(map(str, s[3]))}_k{s[4]}_o{s[5]}")
def test_conv_shape_classes_f32_out(shape):
    B, H, W, cins, k, cout = shape
    _run_case(B, H, W, cins, k, cout, out_f32=True)


@pytest.mark.parametrize("shape", SHAPES[:6], ids=lambda s: f"B{s[0]}_{s[1]}x{s[2]}_{'+'.join(map(str, s[3]))}_k{s[4]}_o{s[5]}")
def test_conv_bf16_out(shape):
    B, H, W, cins, k, cout = shape
    _run_case(B, H, W, cins, k, cout)


def test_conv_residual_and_fused_skip():
    _run_case(2, 48, 48, [256], 3, 256, residual=True)
    _run_case(1, 48, 48, [256], 3, 256, skip_c=[256, 256])
    _run_case(1, 96, 96, [128], 3, 128, skip_c=[256, 128], out_f32=True)


def test_conv_fp16_and_no_bias():
    _run_case(1, 48, 48, [128], 3, 128, bias=False, dtype=torch.float16)


def test_conv_many_tiles_persistent():
    # more tiles than SMs: every CTA loops, accumulator stages and smem ring wrap several times
    _run_case(8, 96, 96, [128], 3, 128, out_f32=True, seed=3)
    # ... and the macro-tile kernel with 16-bit output (small batches are planned with 8 x 16 tiles instead)
    _run_case(8, 96, 96, [128], 3, 128, seed=3)
    _run_case(8, 48, 48, [256], 3, 256, skip_c=[256, 128], dtype=torch.float16, seed=4)


# Image-interleaved M tiles (conv_igemm2.cu, kIL): geometries whose height is not a multiple of 16 - the 24 x 24 level
# of the UNet (OpenAI_Unet.py: input_blocks 8-11, middle_block, output_blocks 0-3) - pair the rows of two consecutive
# images in one tile.  Odd batches (last pair has one image), every source mix the level sees, 8 x 8 and 40 x 40 images.
IL_CASES = [
    dict(B=2, H=24, W=24, cins=[256], ksize=3, cout=256),
    dict(B=3, H=24, W=24, cins=[256], ksize=3, cout=256),
    dict(B=5, H=24, W=24, cins=[256, 256], ksize=3, cout=256),
    dict(B=4, H=24, W=24, cins=[256], ksize=3, cout=256, residual=True),
    dict(B=3, H=24, W=24, cins=[256], ksize=3, cout=256, skip_c=[256, 256]),
    dict(B=3, H=24, W=24, cins=[256], ksize=1, cout=768),
    # more work items than CTA pairs: two interleaved tiles per CTA (each its own TMA box), odd tile count (153)
    dict(B=33, H=24, W=24, cins=[256], ksize=3, cout=256),
    dict(B=33, H=24, W=24, cins=[128], ksize=3, cout=128, skip_c=[256, 128]),
    dict(B=34, H=24, W=24, cins=[128, 128], ksize=3, cout=256, residual=True),
    dict(B=70, H=24, W=24, cins=[64], ksize=3, cout=128),  # ... and several rounds of them
    dict(B=4, H=8, W=8, cins=[64], ksize=3, cout=128),
    dict(B=3, H=40, W=40, cins=[64], ksize=3, cout=128),
    dict(B=2, H=24, W=48, cins=[128], ksize=3, cout=128),
]


@pytest.mark.parametrize("case", IL_CASES, ids=lambda c: "_".join(f"{k}{v}" for k, v in c.items()).replace(" ", ""))
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_conv_interleaved_tiles(case, dtype):
    _run_case(dtype=dtype, **case)


@pytest.mark.parametrize("B,H,W,cin,cout", [(2, 24, 24, 256, 256), (3, 24, 24, 256, 256), (33, 24, 24, 128, 256),
                                            (2, 48, 48, 128, 256), (8, 96, 96, 128, 128),
                                            (1, 96, 96, 128, 128), (5, 8, 8, 64, 128)])
def test_conv_epilogue_groupnorm_statistics(B, H, W, cin, cout):
    """(sum, sum of squares) per image and 4-channel bucket of the fp32 convolution result, emitted by the epilogue
    (GroupNorm32 input statistics, util.py:214-216) - plain and image-interleaved tiles."""
    from cddpm import ops

    _setup()
    dtype = torch.float16
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(B, cin, H, W, device="cuda", generator=g)
    w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (cin * 9) ** 0.5
    b = torch.randn(cout, device="cuda", generator=g)
    # per-image offsets make a pixel attributed to the wrong image of a pair visible in the sums
    x = x + torch.arange(B, device="cuda", dtype=torch.float32).view(B, 1, 1, 1) * 0.25
    xq, wq = x.to(dtype).float(), w.to(dtype).float()
    ref = F.conv2d(xq, wq, b, padding=1).double()
    out, stats = ops.conv_igemm_stats([_nhwc16(x, dtype)], [9], ops.pack_conv_weight(wq, [cin], dtype), b)
    torch.cuda.synchronize()
    rb = ref.view(B, cout // 4, 4, H * W)
    want = torch.stack([rb.sum(dim=(2, 3)), (rb * rb).sum(dim=(2, 3))], dim=-1)
    scale = want.abs().amax(dim=(1,), keepdim=True).clamp_min(1.0)
    err = ((stats - want).abs() / scale).max().item()
    assert err < 2e-4, f"statistics off by {err:.3g} (relative to the largest bucket of each image)"
    got = out.float().permute(0, 3, 1, 2)
    assert (got - ref.float()).abs().max().item() <= 2.0 ** -11 * max(ref.abs().max().item(), 1.0) + 1e-4
