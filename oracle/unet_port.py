"""ORACLE (test infrastructure, never imported by the product path).

Plain-PyTorch fp32 functional restatement of the reference's conditioned UNet forward, driven by a state_dict with
the reference's key layout:
  UNetModel.__init__ / forward        src/models/modules/OpenAI_Unet.py:513-797, :823-1006
  ResBlock._forward                   :284-338   (FiLM scale/shift: :325-331; up/down: :287-293)
  AttentionBlock / QKVAttention       :386-394, :457-476
  timestep_embedding, GroupNorm32     src/models/LDM/modules/diffusionmodules/util.py:151-171, :214-216
Runs on CPU (the timed CPU baseline of bench.py) or on any device torch supports.  Pinned against the live
reference by tests/golden/unet_*.npz (oracle/make_golden.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


@dataclass
class UNetSpec:
    """Architecture knobs as DDPM_2D.__init__ passes them (src/models/DDPM_2D.py:37-59)."""

    in_channels: int = 1
    model_channels: int = 128
    out_channels: int = 1
    num_res_blocks: int = 3
    attention_resolutions: Tuple[int, ...] = (3, 6, 12)
    channel_mult: Tuple[int, ...] = (1, 2, 2)
    num_classes: Optional[int] = 128
    num_head_channels: int = 64
    groups: int = 32

    @property
    def emb_dim(self) -> int:
        return self.model_channels * 4 * (2 if self.num_classes is not None else 1)


# A block is a list of layers; a layer is ("conv", cin, cout) | ("res", cin, cout, mode) | ("attn", ch, heads)
def block_plan(spec: UNetSpec):
    """Layer plan of input_blocks / middle_block / output_blocks (OpenAI_Unet.py:605-797)."""
    mc = spec.model_channels
    inputs: List[list] = [[("conv", spec.in_channels, mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, mult in enumerate(spec.channel_mult):
        for _ in range(spec.num_res_blocks):
            layers = [("res", ch, mult * mc, "none")]
            ch = mult * mc
            if ds in spec.attention_resolutions:
                layers.append(("attn", ch, ch // spec.num_head_channels))
            inputs.append(layers)
            chans.append(ch)
        if level != len(spec.channel_mult) - 1:
            inputs.append([("res", ch, ch, "down")])
            chans.append(ch)
            ds *= 2
    middle = [("res", ch, ch, "none"), ("attn", ch, ch // spec.num_head_channels), ("res", ch, ch, "none")]
    outputs: List[list] = []
    for level, mult in list(enumerate(spec.channel_mult))[::-1]:
        for i in range(spec.num_res_blocks + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, mc * mult, "none")]
            ch = mc * mult
            if ds in spec.attention_resolutions:
                layers.append(("attn", ch, ch // spec.num_head_channels))
            if level and i == spec.num_res_blocks:
                layers.append(("res", ch, ch, "up"))
                ds //= 2
            outputs.append(layers)
    return inputs, middle, outputs, ch


def param_shapes(spec: UNetSpec) -> List[Tuple[str, Tuple[int, ...]]]:
    """(key, shape) of every UNet state_dict entry, in the reference's registration order."""
    out: List[Tuple[str, Tuple[int, ...]]] = []
    mc = spec.model_channels
    half = spec.emb_dim // (2 if spec.num_classes is not None else 1)

    def lin(p, i, o):
        out.extend([(p + ".weight", (o, i)), (p + ".bias", (o,))])

    def conv(p, i, o, k):
        out.extend([(p + ".weight", (o, i, k, k)), (p + ".bias", (o,))])

    def gn(p, c):
        out.extend([(p + ".weight", (c,)), (p + ".bias", (c,))])

    def layer(p, l):
        if l[0] == "conv":
            conv(p, l[1], l[2], 3)
        elif l[0] == "res":
            _, cin, cout, _mode = l
            gn(p + ".in_layers.0", cin)
            conv(p + ".in_layers.2", cin, cout, 3)
            lin(p + ".emb_layers.1", spec.emb_dim, 2 * cout)
            gn(p + ".out_layers.0", cout)
            conv(p + ".out_layers.3", cout, cout, 3)
            if cin != cout:
                conv(p + ".skip_connection", cin, cout, 1)
        else:
            _, c, _h = l
            gn(p + ".norm", c)
            out.extend([(p + ".qkv.weight", (3 * c, c, 1)), (p + ".qkv.bias", (3 * c,))])
            out.extend([(p + ".proj_out.weight", (c, c, 1)), (p + ".proj_out.bias", (c,))])

    if spec.num_classes is not None:
        lin("label_emb.0", spec.num_classes, half)
        lin("label_emb.2", half, half)
    lin("time_embed.0", mc, half)
    lin("time_embed.2", half, half)
    inputs, middle, outputs, ch = block_plan(spec)
    for bi, layers in enumerate(inputs):
        for li, l in enumerate(layers):
            layer(f"input_blocks.{bi}.{li}", l)
    for li, l in enumerate(middle):
        layer(f"middle_block.{li}", l)
    for bi, layers in enumerate(outputs):
        for li, l in enumerate(layers):
            layer(f"output_blocks.{bi}.{li}", l)
    gn("out.0", ch)
    conv("out.2", mc, spec.out_channels, 3)
    return out


def sinusoid(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn(x, sd, p, groups):
    return F.group_norm(x.float(), groups, sd[p + ".weight"], sd[p + ".bias"], 1e-5).type(x.dtype)


def _res(sd, p, x, emb, l, groups):
    _, cin, cout, mode = l
    h = F.silu(_gn(x, sd, p + ".in_layers.0", groups))
    if mode == "up":
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    elif mode == "down":
        h = F.avg_pool2d(h, 2)
        x = F.avg_pool2d(x, 2)
    h = F.conv2d(h, sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    scale, shift = e[:, :cout, None, None], e[:, cout:, None, None]
    h = _gn(h, sd, p + ".out_layers.0", groups) * (1 + scale) + shift
    h = F.conv2d(F.silu(h), sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if cin != cout:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def _attn(sd, p, x, l, groups):
    _, c, heads = l
    b, _, hh, ww = x.shape
    xf = x.reshape(b, c, hh * ww)
    qkv = F.conv1d(_gn(xf, sd, p + ".norm", groups), sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    q, k, v = qkv.chunk(3, dim=1)
    d = c // heads
    s = 1.0 / math.sqrt(math.sqrt(d))
    L = hh * ww
    q = (q * s).reshape(b * heads, d, L)
    k = (k * s).reshape(b * heads, d, L)
    v = v.reshape(b * heads, d, L)
    w = torch.softmax(torch.einsum("bct,bcs->bts", q, k).float(), dim=-1).type(q.dtype)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(b, c, L)
    o = F.conv1d(a, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (xf + o).reshape(b, c, hh, ww)


def _run_layers(sd, prefix, layers, h, emb, groups, taps=None):
    for li, l in enumerate(layers):
        p = f"{prefix}.{li}"
        if l[0] == "conv":
            h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
        elif l[0] == "res":
            h = _res(sd, p, h, emb, l, groups)
        else:
            h = _attn(sd, p, h, l, groups)
        if taps is not None:
            taps[p] = h
    return h


def embedding(sd: Dict[str, torch.Tensor], spec: UNetSpec, t: torch.Tensor, cond: Optional[torch.Tensor]):
    e = sinusoid(t, spec.model_channels)
    e = F.linear(F.silu(F.linear(e, sd["time_embed.0.weight"], sd["time_embed.0.bias"])),
                 sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    if spec.num_classes is not None:
        c = F.linear(F.silu(F.linear(cond, sd["label_emb.0.weight"], sd["label_emb.0.bias"])),
                     sd["label_emb.2.weight"], sd["label_emb.2.bias"])
        e = torch.cat([e, c], dim=1)
    return e


def unet_forward(sd: Dict[str, torch.Tensor], spec: UNetSpec, x: torch.Tensor, t: torch.Tensor,
                 cond: Optional[torch.Tensor] = None, taps: Optional[dict] = None) -> torch.Tensor:
    """model(x, t, cond) of the reference (OpenAI_Unet.py:823-1006) without the debug clones."""
    inputs, middle, outputs, _ = block_plan(spec)
    emb = embedding(sd, spec, t, cond)
    if taps is not None:
        taps["emb"] = emb
    hs = []
    h = x
    for bi, layers in enumerate(inputs):
        h = _run_layers(sd, f"input_blocks.{bi}", layers, h, emb, spec.groups, taps)
        hs.append(h)
    h = _run_layers(sd, "middle_block", middle, h, emb, spec.groups, taps)
    for bi, layers in enumerate(outputs):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_layers(sd, f"output_blocks.{bi}", layers, h, emb, spec.groups, taps)
    h = F.silu(_gn(h, sd, "out.0", spec.groups))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)
