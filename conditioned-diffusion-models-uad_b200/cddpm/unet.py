"""UNetModel — drop-in for the reference's src.models.modules.OpenAI_Unet.UNetModel (OpenAI_Unet.py:483-1006).

Same constructor signature, same module tree and therefore the same state_dict keys / shapes (fp32 nn.Parameters are
the masters and round-trip bit-exactly); `forward(x, timesteps, cond=None, context=None)` runs the whole network inside
the C++/CUDA engine (cddpm_unet_forward).  The sub-modules below are parameter holders only: they have no arithmetic
of their own and there is no eager/PyTorch fallback.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from ._lib import CddpmError
from .engine import UNetEngine


class GroupNorm32(nn.GroupNorm):
    """Parameter holder for normalization(channels) (util.py:199-216): 32 groups, eps 1e-5, affine."""


def _no_eager(name):
    def forward(self, *a, **k):
        raise CddpmError(f"{name} has no stand-alone forward: it executes inside the CUDA UNet engine "
                         "(call UNetModel.forward on CUDA tensors)")

    return forward


def _zero(module: nn.Module) -> nn.Module:
    for p in module.parameters():
        p.detach().zero_()
    return module


class TimestepEmbedSequential(nn.Sequential):
    forward = _no_eager("TimestepEmbedSequential")


class Upsample(nn.Module):
    """Nearest x2 (no parameters when use_conv is False, which is the only use in this model)."""

    def __init__(self, channels, use_conv=False, dims=2, out_channels=None):
        super().__init__()
        assert not use_conv and dims == 2
        self.channels = channels

    forward = _no_eager("Upsample")


class Downsample(nn.Module):
    """2x2 average pooling (no parameters when use_conv is False)."""

    def __init__(self, channels, use_conv=False, dims=2, out_channels=None):
        super().__init__()
        assert not use_conv and dims == 2
        self.channels = channels

    forward = _no_eager("Downsample")


class ResBlock(nn.Module):
    """FiLM residual block (OpenAI_Unet.py:182-338) with use_scale_shift_norm=True."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, up=False, down=False):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.up, self.down = up, down
        self.in_layers = nn.Sequential(GroupNorm32(32, channels), nn.SiLU(),
                                       nn.Conv2d(channels, self.out_channels, 3, padding=1))
        if up:
            self.h_upd, self.x_upd = Upsample(channels), Upsample(channels)
        elif down:
            self.h_upd, self.x_upd = Downsample(channels), Downsample(channels)
        else:
            self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, 2 * self.out_channels))
        self.out_layers = nn.Sequential(GroupNorm32(32, self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        _zero(nn.Conv2d(self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 1)

    forward = _no_eager("ResBlock")


class AttentionBlock(nn.Module):
    """Spatial self-attention (OpenAI_Unet.py:341-394), new attention order, head dim 64."""

    def __init__(self, channels, num_heads=1, num_head_channels=-1):
        super().__init__()
        self.channels = channels
        self.num_heads = num_heads if num_head_channels == -1 else channels // num_head_channels
        self.norm = GroupNorm32(32, channels)
        self.qkv = nn.Conv1d(channels, channels * 3, 1)
        self.proj_out = _zero(nn.Conv1d(channels, channels, 1))

    forward = _no_eager("AttentionBlock")


class UNetModel(nn.Module):
    def __init__(
        self,
        image_size,
        in_channels,
        model_channels,
        out_channels,
        num_res_blocks,
        attention_resolutions,
        dropout=0,
        channel_mult=(1, 2, 4, 8),
        conv_resample=True,
        dims=2,
        num_classes=None,
        use_checkpoint=False,
        use_fp16=True,
        num_heads=1,
        num_head_channels=-1,
        num_heads_upsample=-1,
        use_scale_shift_norm=False,
        resblock_updown=False,
        use_new_attention_order=False,
        use_spatial_transformer=False,
        transformer_depth=1,
        context_dim=None,
        legacy=True,
        num_mem_kv=0,
        engine_dtype=torch.float16,
    ):
        super().__init__()
        if use_spatial_transformer:
            assert context_dim is not None, "spatial transformer needs context_dim"
            raise NotImplementedError("use_spatial_transformer=True is a dead branch in the reference cDDPM configs")
        assert context_dim is None, "context_dim requires use_spatial_transformer"
        if dims != 2 or not use_scale_shift_norm or not resblock_updown or not use_new_attention_order:
            raise NotImplementedError("the CUDA engine implements the cDDPM configuration: dims=2, "
                                      "use_scale_shift_norm, resblock_updown, use_new_attention_order")
        if dropout:
            raise NotImplementedError("dropout_unet > 0 is a training-only feature (not in this build)")
        if num_head_channels != 64:
            raise NotImplementedError("attention head width must be 64 (DDPM_2D.py:52)")
        if isinstance(image_size, int):
            image_size = (image_size, image_size)
        self.features_info = {}  # kept for API compatibility; the reference's debug collector stays empty
        self.image_size = tuple(int(s) for s in image_size)
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = tuple(attention_resolutions)
        self.dropout = dropout
        self.channel_mult = tuple(int(m) for m in channel_mult)
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.dtype = torch.float16 if use_fp16 else torch.float32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads if num_heads_upsample == -1 else num_heads_upsample
        self.num_mem_kv = num_mem_kv
        self.engine_dtype = engine_dtype

        mc = model_channels
        if num_classes is not None:
            emb = mc * 4 * 2
            self.label_emb = nn.Sequential(nn.Linear(num_classes, emb // 2), nn.SiLU(), nn.Linear(emb // 2, emb // 2))
            self.time_embed = nn.Sequential(nn.Linear(mc, emb // 2), nn.SiLU(), nn.Linear(emb // 2, emb // 2))
        else:
            emb = mc * 4
            self.time_embed = nn.Sequential(nn.Linear(mc, emb), nn.SiLU(), nn.Linear(emb, emb))

        def attn(ch):
            return AttentionBlock(ch, num_heads=ch // num_head_channels, num_head_channels=num_head_channels)

        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(in_channels, mc, 3, padding=1))])
        chans: List[int] = [mc]
        ch, ds = mc, 1
        for level, mult in enumerate(self.channel_mult):
            for _ in range(num_res_blocks):
                layers: List[nn.Module] = [ResBlock(ch, emb, dropout, out_channels=mult * mc)]
                ch = mult * mc
                if ds in self.attention_resolutions:
                    layers.append(attn(ch))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                chans.append(ch)
            if level != len(self.channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(ResBlock(ch, emb, dropout, out_channels=ch, down=True)))
                chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(ResBlock(ch, emb, dropout), attn(ch), ResBlock(ch, emb, dropout))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(self.channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = chans.pop()
                layers = [ResBlock(ch + ich, emb, dropout, out_channels=mc * mult)]
                ch = mc * mult
                if ds in self.attention_resolutions:
                    layers.append(attn(ch))
                if level and i == num_res_blocks:
                    layers.append(ResBlock(ch, emb, dropout, out_channels=ch, up=True))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(GroupNorm32(32, ch), nn.SiLU(), _zero(nn.Conv2d(mc, out_channels, 3, padding=1)))

        self._engine: Optional[UNetEngine] = None
        self._engine_versions = None
        self._train_engine: Optional[UNetEngine] = None
        self._train_versions = None
        self._items = None

    # ------------------------------------------------------------------ reference API
    def convert_to_fp16(self):
        """No-op, as in the reference (OpenAI_Unet.py:23-28, :799-805); precision is the engine's operand dtype."""

    def convert_to_fp32(self):
        """No-op, as in the reference."""

    def forward_with_cond_scale(self, *args, cond_scale=2.0, **kwargs):
        return self.forward(*args, **kwargs)

    # ------------------------------------------------------------------ engine plumbing
    def _param_items(self):
        if self._items is None:
            self._items = list(self.state_dict(keep_vars=True).items())
        return self._items

    def _apply(self, fn, *args, **kwargs):
        # .to()/.cuda()/.float() may re-seat parameter storage: drop the cached view and re-push everything
        self._items = None
        self._engine_versions = None if self._engine is None else [None] * len(self._engine_versions)
        self._train_versions = None if self._train_engine is None else [None] * len(self._train_versions)
        return super()._apply(fn, *args, **kwargs)

    def engine(self) -> UNetEngine:
        """The CUDA engine with the current parameter values (re-pushed when a parameter tensor changed)."""
        items = self._param_items()
        dev = items[0][1].device
        if dev.type != "cuda":
            raise CddpmError("UNetModel is on the CPU; the cDDPM engine has no CPU path — move the module to CUDA")
        if self._engine is None:
            self._engine = UNetEngine(
                image_size=self.image_size, in_channels=self.in_channels, model_channels=self.model_channels,
                out_channels=self.out_channels, num_res_blocks=self.num_res_blocks,
                attention_resolutions=self.attention_resolutions, channel_mult=self.channel_mult,
                num_classes=self.num_classes, num_head_channels=self.num_head_channels, dtype=self.engine_dtype)
            names = [n for n, _ in self._engine.param_names()]
            if names != [n for n, _ in items]:
                raise CddpmError("engine parameter list differs from the module's state_dict layout")
            self._engine_versions = [None] * len(items)
        self._push(self._engine, self._engine_versions, items)
        return self._engine

    def train(self, mode: bool = True):
        # leaving / entering training: the next engine call re-pushes every parameter (an optimizer may have changed
        # them without advancing the version counters - torch's fused optimizers do)
        if mode != self.training:
            if self._engine_versions is not None:
                self._engine_versions = [None] * len(self._engine_versions)
            if self._train_versions is not None:
                self._train_versions = [None] * len(self._train_versions)
        return super().train(mode)

    def train_engine(self) -> UNetEngine:
        """The bf16 engine the training step runs on (forward + backward kernels); shares nothing with the inference
        engine when that one is fp16.  Parameters are re-pushed when the optimizer changed them."""
        if self.engine_dtype == torch.bfloat16:
            if self._engine_versions is not None:
                for i in range(len(self._engine_versions)):
                    self._engine_versions[i] = None  # see below: always push in training
            eng = self.engine()
            # the shared bf16 engine must keep every intermediate the backward pass reads: no inference-only fusions
            # (the head's GroupNorm inside conv_out) once it has served a training step
            eng.set_training(True)
            return eng
        items = self._param_items()
        if items[0][1].device.type != "cuda":
            raise CddpmError("UNetModel is on the CPU; the cDDPM engine has no CPU path — move the module to CUDA")
        if self._train_engine is None:
            self._train_engine = UNetEngine(
                image_size=self.image_size, in_channels=self.in_channels, model_channels=self.model_channels,
                out_channels=self.out_channels, num_res_blocks=self.num_res_blocks,
                attention_resolutions=self.attention_resolutions, channel_mult=self.channel_mult,
                num_classes=self.num_classes, num_head_channels=self.num_head_channels, dtype=torch.bfloat16,
                training=True)
            self._train_versions = [None] * len(items)
        # training: push everything on every call.  Version counters are not a reliable change signal here - torch's
        # fused optimizers update parameters without advancing them - and the bulk push replays one CUDA graph.
        for i in range(len(self._train_versions)):
            self._train_versions[i] = None
        self._push(self._train_engine, self._train_versions, items)
        return self._train_engine

    @staticmethod
    def _push(engine, seen, items):
        """Re-push the parameters whose storage or version changed since the last call (all of them after an
        optimizer step) with one bulk call."""
        changed = []
        any_changed = False
        for i, (_, p) in enumerate(items):
            v = (p.data_ptr(), p._version)
            if seen[i] != v:
                changed.append(p)
                seen[i] = v
                any_changed = True
            else:
                changed.append(None)
        if any_changed:
            engine.set_params(changed)

    def forward(self, x, timesteps, cond=None, context=None):
        """model(x, t, cond): x [B,1,H,W], timesteps [B], cond [B,num_classes] -> [B,1,H,W] fp32.
        In train() mode with autograd enabled and parameters that require gradients, the call is recorded as one autograd node whose
        backward runs the engine's backward kernels (DDPM_2D.training_step -> loss.backward())."""
        self.features_info.clear()
        if self.num_classes is None:
            cond = None
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            params = [p for _, p in self._param_items()]
            return _UNetTrainFunction.apply(self, x, timesteps, cond, *params)
        return self.engine().forward(x, timesteps, cond)


class _UNetTrainFunction(torch.autograd.Function):
    """UNetModel.forward as ONE autograd node: forward = cddpm_unet_forward on the bf16 engine (which keeps every
    layer output), backward = cddpm_unet_backward.  Parameter gradients come back as views of one flat fp32 buffer in
    the reference's parameter layouts; d cond flows on to the condition encoder."""

    @staticmethod
    def forward(ctx, module, x, t, cond, *params):
        eng = module.train_engine()
        x = x.detach().float().contiguous()
        t = t.to(device=x.device, dtype=torch.int64).contiguous()
        c = cond.detach().float().contiguous() if cond is not None else None
        out = eng.forward(x, t, c)
        ctx.engine = eng
        ctx.serial = eng.forward_serial
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.numels = [p.numel() for p in params]
        ctx.needs = [p.requires_grad for p in params]
        ctx.want_dcond = cond is not None and cond.requires_grad
        ctx.save_for_backward(x, t, *([c] if c is not None else []))  # the engine reads them again in backward
        return out

    @staticmethod
    def backward(ctx, dout):
        eng = ctx.engine
        if eng.forward_serial != ctx.serial:
            raise CddpmError("the UNet engine keeps the activations of its LAST forward only: call backward() before "
                             "running the model again")
        flat, dcond = eng.backward(dout, want_dcond=ctx.want_dcond)
        _, offs = eng.grad_layout()
        grads = [flat[o:o + n].view(s) if need else None
                 for o, n, s, need in zip(offs, ctx.numels, ctx.shapes, ctx.needs)]
        return (None, None, None, dcond, *grads)
