"""GPU parity of the attention core (tcgen05 path for L % 64 == 0, L <= 576; SIMT path otherwise) against the
reference formula in fp32 on the same 16-bit-rounded qkv.

Reference: QKVAttention.forward, src/models/modules/OpenAI_Unet.py:457-476 (q, k scaled by ch**-0.25, softmax in fp32,
weights cast back to the activation dtype before the product with v)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(qkv, C):
    B, L, _ = qkv.shape
    H = C // 64
    q, k, v = qkv.float().split(C, dim=2)
    q = q.reshape(B, L, H, 64).permute(0, 2, 1, 3)
    k = k.reshape(B, L, H, 64).permute(0, 2, 1, 3)
    v = v.reshape(B, L, H, 64).permute(0, 2, 1, 3)
    scale = 64 ** -0.25
    w = torch.softmax((q * scale) @ (k * scale).transpose(2, 3), dim=-1)
    return (w @ v).permute(0, 2, 1, 3).reshape(B, L, C)


@pytest.mark.parametrize("B,L,C", [(3, 576, 256), (2, 64, 128), (2, 256, 64), (1, 384, 256), (2, 144, 128), (2, 100, 64)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_attention_matches_fp32_reference(B, L, C, dtype):
    from cddpm import ops

    g = torch.Generator(device="cuda").manual_seed(L + C)
    # scores with a spread of a few units so the softmax is neither flat nor one-hot
    qkv = (torch.randn(B, L, 3 * C, device="cuda", generator=g) * 1.5).to(dtype)
    out = ops.attention(qkv, C).float()
    ref = _reference(qkv, C)
    tol = 4e-3 if dtype == torch.float16 else 3e-2  # P and the output are rounded to the 16-bit type
    err = (out - ref).abs().max().item()
    assert err < tol, f"max abs err {err}"


def test_attention_peaked_softmax_is_stable():
    """Large logits: the row maximum must be subtracted before the exponential."""
    from cddpm import ops

    g = torch.Generator(device="cuda").manual_seed(7)
    qkv = (torch.randn(2, 576, 3 * 64, device="cuda", generator=g) * 6.0).to(torch.float16)
    out = ops.attention(qkv, 64).float()
    ref = _reference(qkv, 64)
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() < 3e-2


@pytest.mark.parametrize("B,I,O,silu_in,silu_out", [(50, 2048, 128, 0, 0), (32, 2048, 128, 0, 0), (3, 512, 130, 1, 1),
                                                    (9, 256, 5, 0, 1), (4, 100, 7, 1, 0), (1, 128, 512, 0, 1)])
def test_linear_matches_fp32_reference(B, I, O, silu_in, silu_out):
    """cddpm_linear (nn.Linear on fp32 rows with optional SiLU either side: time_embed / label_emb,
    OpenAI_Unet.py:583-602; the condition encoder's fc head): both the wide-K kernel (I >= 256, I % 4 == 0 - the
    2048 -> 128 head) and the narrow one, ragged O and B."""
    import ctypes

    from cddpm._lib import check, current_stream, lib, ptr

    g = torch.Generator(device="cuda").manual_seed(B * 1000 + I + O)
    x = torch.randn(B, I, device="cuda", generator=g)
    w = torch.randn(O, I, device="cuda", generator=g) / I ** 0.5
    bias = torch.randn(O, device="cuda", generator=g)
    out = torch.full((B, O), float("nan"), device="cuda")
    check(lib().cddpm_linear(ptr(x), I, ptr(w), ptr(bias), ptr(out), O, B, I, O, silu_in, silu_out, current_stream()),
          "cddpm_linear")
    xin = torch.nn.functional.silu(x.double()) if silu_in else x.double()
    ref = xin @ w.double().t() + bias.double()
    if silu_out:
        ref = torch.nn.functional.silu(ref)
    err = (out.double() - ref).abs().max().item()
    assert err < 2e-5, f"max abs err {err}"
