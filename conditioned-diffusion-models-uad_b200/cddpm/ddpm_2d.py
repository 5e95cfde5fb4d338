"""DDPM_2D — drop-in for the reference LightningModule src.models.DDPM_2D.DDPM_2D (DDPM_2D.py:17-308).

Same constructor (`DDPM_2D(cfg, prefix=None)`, the Hydra `_target_` of configs/model/DDPM_2D.yaml), same sub-module
names (`encoder`, `diffusion`, `diffusion.model`) and therefore the same 649-entry state_dict, same hooks
(`forward`, `training_step`, `validation_step`, `on_test_start`, `test_step`, `on_test_end`, `configure_optimizers`,
`update_prefix`).  The numerical work of test_step — encoder, q_sample, UNet, reconstruction, ensemble mean, residual,
mask erosion, median filter, threshold search, metrics — runs in libcddpm_b200 (sm_100a CUDA); nothing falls back to
PyTorch eager or the CPU.

Differences on purpose: `num_eval_slices` is honoured when the cfg carries `force_num_eval_slices:
False` (the reference fork hard-codes 4 centre slices, DDPM_2D.py:193; that stays the default); image logging is off.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any

import numpy as np
import torch
import torch.optim as optim

from ._lib import CddpmError
from .diffusion import GaussianDiffusion
from .encoder import get_encoder
from .eval_tail import _test_end, _test_step, get_eval_dictionary
from .noise import gen_noise
from .unet import UNetModel as OpenAI_UNet

try:  # Lightning is optional: without it the module is a plain nn.Module with the same hooks
    from pytorch_lightning.core.lightning import LightningModule  # type: ignore
except Exception:  # pragma: no cover - depends on the environment
    try:
        from pytorch_lightning import LightningModule  # type: ignore
    except Exception:
        class LightningModule(torch.nn.Module):  # type: ignore
            def save_hyperparameters(self, *a, **k):
                pass

            def log(self, *a, **k):
                pass

            @property
            def device(self):
                return next(self.parameters()).device

try:
    from omegaconf import open_dict  # type: ignore
except Exception:  # pragma: no cover
    from contextlib import contextmanager

    @contextmanager
    def open_dict(cfg):
        yield cfg

DATA = "data"  # torchio.DATA


def _engine_dtype(cfg):
    return {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16}.get(str(cfg.get("engine_dtype", "fp16")), torch.float16)


class DDPM_2D(LightningModule):
    def __init__(self, cfg, prefix=None):
        super().__init__()
        self.cfg = cfg
        if cfg.get("condition", True):
            with open_dict(self.cfg):
                self.cfg["cond_dim"] = cfg.get("unet_dim", 128)
            self.encoder, out_features = get_encoder(cfg)
        else:
            out_features = None
        size = (int(cfg.imageDim[0] / cfg.rescaleFactor), int(cfg.imageDim[1] / cfg.rescaleFactor))
        model = OpenAI_UNet(
            image_size=size, in_channels=1, model_channels=cfg.get("unet_dim", 64), out_channels=1,
            num_res_blocks=cfg.get("num_res_blocks", 3), attention_resolutions=tuple(cfg.get("att_res", [3, 6, 12])),
            dropout=cfg.get("dropout_unet", 0), channel_mult=cfg.get("dim_mults", [1, 2, 4, 8]), conv_resample=True,
            dims=2, num_classes=out_features, use_checkpoint=False, use_fp16=True, num_heads=1, num_head_channels=64,
            num_heads_upsample=-1, use_scale_shift_norm=True, resblock_updown=True, use_new_attention_order=True,
            use_spatial_transformer=cfg.get("spatial_transformer", False), transformer_depth=1,
            engine_dtype=_engine_dtype(cfg))
        model.convert_to_fp16()
        timesteps = cfg.get("timesteps", 1000)
        sampling_timesteps = cfg.get("sampling_timesteps", timesteps)
        self.test_timesteps = cfg.get("test_timesteps", 150)
        self.diffusion = GaussianDiffusion(
            model, image_size=size, timesteps=timesteps, sampling_timesteps=sampling_timesteps,
            objective=cfg.get("objective", "pred_x0"), channels=1, loss_type=cfg.get("loss", "l1"),
            p2_loss_weight_gamma=cfg.get("p2_gamma", 0), cfg=cfg)
        if cfg.get("pretrained_encoder", False):
            assert cfg.get("encoder_path", None) is not None
            pre = torch.load(cfg.get("encoder_path", None))["state_dict"]
            remapped = OrderedDict()
            for key, val in pre.items():
                if "slice_encoder" in key:
                    remapped["slice_encoder" + key.split("encoder")[-1]] = val
                elif "sparse_encoder" in key:
                    if "fc.weight" not in key and "fc.bias" not in key:
                        remapped["encoder" + key.split("sp_cnn")[-1]] = val
                else:
                    remapped[key] = val
            self.encoder.load_state_dict(remapped, strict=False)
        self.prefix = prefix
        self.save_hyperparameters()

    # ------------------------------------------------------------------ encoder
    def forward(self, x):
        return self.encoder(x) if self.cfg.get("condition", True) else None

    # ------------------------------------------------------------------ train / val
    def _loss_step(self, batch):
        input = batch["vol"][DATA].squeeze(-1)
        features = self(input)
        noise = gen_noise(self.cfg, input.shape, device=input.device) if self.cfg.get("noisetype") is not None else None
        loss, _ = self.diffusion(input, cond=features, noise=noise)
        return loss, input.shape[0]

    def training_step(self, batch, batch_idx: int):
        """DDPM_2D.py:114-138.  The returned loss carries the autograd tape: loss.backward() (Lightning's automatic
        optimisation) runs the UNet engine's backward kernels and, for the condition encoder, torch autograd."""
        loss, n = self._loss_step(batch)
        self.log(f"{self.prefix}train/Loss", loss, prog_bar=False, on_step=False, on_epoch=True, batch_size=n,
                 sync_dist=True)
        return {"loss": loss}

    def validation_step(self, batch: Any, batch_idx: int):
        with torch.no_grad():
            loss, n = self._loss_step(batch)
        self.log(f"{self.prefix}val/Loss_comb", loss, prog_bar=False, on_step=False, on_epoch=True, batch_size=n,
                 sync_dist=True)
        return {"loss": loss}

    # ------------------------------------------------------------------ test
    def on_test_start(self):
        self.eval_dict = get_eval_dictionary()
        self.inds = []
        self.latentSpace_slice = []
        self.new_size = [160, 190, 160]
        self.diffs_list = []
        self.seg_list = []
        if not hasattr(self, "threshold"):
            self.threshold = {}

    @torch.no_grad()
    def reconstruct_slices(self, input, features=None):
        """[D,1,H,W] slices -> (reco [D,1,H,W], loss of the last ensemble member) — DDPM_2D.py:214-247."""
        if features is None:
            features = self(input)
        if self.cfg.get("noise_ensemble", False):
            timesteps = self.cfg.get("step_ensemble", [250, 500, 750])
            if self.cfg.get("stack_ensemble", True) and not torch.is_grad_enabled():
                # the k members as one UNet forward over k x D stacked slices (same noise draws, in the same order)
                noises = [gen_noise(self.cfg, input.shape, device=input.device) if self.cfg.get("noisetype") is not None
                          else None for _ in timesteps]
                loss_diff, reco = self.diffusion.ensemble_reconstruct(input, [t - 1 for t in timesteps], cond=features,
                                                                      noises=noises)
                return reco, loss_diff, features
            reco = torch.empty_like(input, dtype=torch.float32)
            k = len(timesteps)
            for i, t in enumerate(timesteps):
                noise = gen_noise(self.cfg, input.shape, device=input.device) if self.cfg.get("noisetype") is not None else None
                # the ensemble mean is accumulated inside the reconstruction kernel: reco = reco * beta + r / k
                loss_diff, _ = self.diffusion(input, cond=features, t=t - 1, noise=noise, _reco_out=reco,
                                              _reco_alpha=1.0 / k, _reco_beta=0.0 if i == 0 else 1.0)
        else:
            noise = gen_noise(self.cfg, input.shape, device=input.device) if self.cfg.get("noisetype") is not None else None
            loss_diff, reco = self.diffusion(input, cond=features, t=self.test_timesteps - 1, noise=noise)
        return reco, loss_diff, features

    @torch.no_grad()
    def test_step(self, batch: Any, batch_idx: int):
        """DDPM_2D.test_step (DDPM_2D.py:171-286): reconstruction of every slice, then the model-independent tail."""
        return self.test_step_finish(self.test_step_reconstruct(batch), batch_idx)

    @torch.no_grad()
    def test_step_reconstruct(self, batch: Any):
        """First half of test_step: everything that only ENQUEUES GPU work (slice crop, encoder, noise ensemble).  No
        host read of a device value happens here, so a sweep can enqueue volume i+1 before it scores volume i
        (cddpm/sweep.py:run_stage).  Returns the state test_step_finish needs."""
        self.dataset = batch["Dataset"]
        input = batch["vol"][DATA]
        data_orig = batch["vol_orig"][DATA]
        data_seg = batch["seg_orig"][DATA] if batch["seg_available"] else torch.zeros_like(data_orig)
        data_mask = batch["mask_orig"][DATA]
        dev = self.device
        if dev.type != "cuda":
            raise CddpmError("DDPM_2D.test_step needs the module on a CUDA device (there is no CPU path)")
        input, data_orig, data_seg, data_mask = (t.to(dev, non_blocking=True) for t in (input, data_orig, data_seg, data_mask))

        if self.cfg.get("force_num_eval_slices", True):
            self.cfg["num_eval_slices"] = 4  # the fork's hard-coded value (DDPM_2D.py:193)
        if self.cfg.get("num_eval_slices", input.size(4)) != input.size(4):
            num_slices = self.cfg.get("num_eval_slices", input.size(4))
            start = int((input.size(4) - num_slices) / 2)
            sl = slice(start, start + num_slices)
            input, data_orig, data_seg, data_mask = input[..., sl], data_orig[..., sl], data_seg[..., sl], data_mask[..., sl]

        assert input.shape[0] == 1, "Batch size must be 1"
        input = input.squeeze(0).permute(3, 0, 1, 2).contiguous()  # [1,C,H,W,D] -> [D,C,H,W]
        reco, loss_diff, features = self.reconstruct_slices(input)
        latent = features.mean(0).squeeze() if self.cfg.condition else None
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(dev))
        return {"reco": reco, "loss_diff": loss_diff, "latent": latent, "n_slices": input.shape[0], "orig": data_orig,
                "seg": data_seg, "mask": data_mask, "ID": batch["ID"], "stage": batch["stage"], "label": batch["label"],
                "dataset": batch["Dataset"], "done": done}

    @torch.no_grad()
    def test_step_finish(self, st: Any, batch_idx: int):
        """Second half of test_step: the host reads (latent, loss) and utils_eval._test_step.  Runs on the CURRENT
        stream, which is made to wait for the reconstruction's event first."""
        torch.cuda.current_stream(st["reco"].device).wait_event(st["done"])
        self.dataset = st["dataset"]
        self.stage = st["stage"]
        reco, loss_diff = st["reco"], st["loss_diff"]
        latent = st["latent"].detach().cpu() if self.cfg.condition else torch.tensor([0], dtype=float).repeat(st["n_slices"])
        self.latentSpace_slice.extend([latent])
        self.eval_dict["latentSpace"].append(torch.mean(torch.stack([latent]), 0))
        score = np.mean([loss_diff.cpu()])
        self.eval_dict["AnomalyScoreRegPerVol"].append(score)
        if not self.cfg.get("use_postprocessed_score", True):
            self.eval_dict["AnomalyScoreRecoPerVol"].append(score)
            self.eval_dict["AnomalyScoreCombPerVol"].append(score)
            self.eval_dict["AnomalyScoreCombiPerVol"].append(score * score)
            self.eval_dict["AnomalyScoreCombPriorPerVol"].append(score + self.cfg.beta * 0)
            self.eval_dict["AnomalyScoreCombiPriorPerVol"].append(score * 0)
        # [D,1,H,W] -> logical [1,1,H,W,D] without a copy: the tail kernels read the strided view in place
        final_volume = reco.squeeze(1).permute(1, 2, 0).unsqueeze(0).unsqueeze(0)
        _test_step(self, final_volume, st["orig"], st["seg"], st["mask"], batch_idx, st["ID"], st["label"])
        return final_volume

    def on_test_end(self):
        _test_end(self)

    def configure_optimizers(self):
        # the reference's optimizer (DDPM_2D.py:305-306: Adam(lr), default betas / eps).  On CUDA: the same update rule
        # and state layout as ONE kernel launch over all 636 tensors (cddpm/optim.py); `optimizer: torch` in the
        # config selects torch.optim.Adam itself.
        params = list(self.parameters())
        if all(p.is_cuda for p in params) and str(self.cfg.get("optimizer", "cddpm")) != "torch":
            from .optim import Adam

            return Adam(params, lr=self.cfg.lr)
        return optim.Adam(params, lr=self.cfg.lr)

    def update_prefix(self, prefix):
        self.prefix = prefix
