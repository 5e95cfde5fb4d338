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
