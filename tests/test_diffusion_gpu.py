"""GPU parity of the DDPM arithmetic (GaussianDiffusion mirror + fused kernels + GPU simplex noise) against golden
vectors produced by the live reference (oracle/make_golden.py) and against the oracle port.

Reference: src/models/modules/cond_DDPM.py:548-554 (q_sample), :565-655 (forward/p_losses), :432-464 (reverse loop);
src/utils/generate_noise.py:8-52 (gen_noise)."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

pytestmark = pytest.mark.gpu


class Cfg(dict):
    __getattr__ = dict.get


def _small_model():
    from cddpm.unet import UNetModel
    from oracle import unet_port
    from oracle.weights import make_state_dict

    spec = unet_port.UNetSpec(model_channels=64, channel_mult=(1, 2), num_res_blocks=1, num_classes=128)
    m = UNetModel(image_size=(32, 32), in_channels=1, model_channels=64, out_channels=1, num_res_blocks=1,
                  attention_resolutions=(3, 6, 12), dropout=0, channel_mult=[1, 2], conv_resample=True, dims=2,
                  num_classes=128, use_checkpoint=False, use_fp16=True, num_heads=1, num_head_channels=64,
                  num_heads_upsample=-1, use_scale_shift_norm=True, resblock_updown=True,
                  use_new_attention_order=True, use_spatial_transformer=False, transformer_depth=1)
    sd = make_state_dict(unet_port.param_shapes(spec), seed=2)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd, spec


def test_simplex_noise_bit_exact_vs_reference_golden():
    from cddpm.noise import gen_noise

    g = np.load(os.path.join(GOLD, "simplex.npz"))
    for seed in (0, 7):
        np.random.seed(seed)
        n = gen_noise(Cfg(noisetype="simplex"), (2, 1, 96, 96))
        assert n.dtype == torch.float16 and n.is_cuda and tuple(n.shape) == (2, 1, 96, 96)
        ref = torch.from_numpy(g[f"field_seed{seed}"])
        assert torch.equal(n[0, 0].cpu(), ref) and torch.equal(n[1, 0].cpu(), ref)
    with pytest.raises(ValueError):
        gen_noise(Cfg(noisetype="gauss"), (1, 1, 96, 96))


def test_simplex_matches_oracle_port_other_sizes():
    from cddpm.noise import simplex_field
    from oracle.simplex_port import fractal_field, permutation_from_seed

    for size, seed in ((32, 12345), (64, -987654321)):
        got = simplex_field(seed, (1, 1, size, size))[0, 0].cpu()
        ref = torch.from_numpy(fractal_field((size, size), permutation_from_seed(seed))).half()
        assert torch.equal(got, ref)


def test_q_sample_bit_exact():
    from cddpm.diffusion import GaussianDiffusion
    from oracle import diffusion_port

    d = GaussianDiffusion(torch.nn.Identity(), image_size=(96, 96), timesteps=1000, objective="pred_x0", channels=1).cuda()
    sched = diffusion_port.schedule_buffers()
    g = torch.Generator().manual_seed(0)
    x0 = torch.rand(5, 1, 96, 96, generator=g) * 2 - 1
    noise = torch.randn(5, 1, 96, 96, generator=g).half()
    t = torch.tensor([0, 249, 499, 749, 999])
    ref = diffusion_port.q_sample(sched, x0, t, noise)
    got = d.q_sample(x0.cuda(), t.cuda(), noise.cuda()).cpu()
    assert torch.equal(got, ref)
    # shared single-index form used by p_sample_loop (cond_DDPM.py:452)
    ref1 = diffusion_port.q_sample(sched, x0, torch.tensor([500]), noise)
    got1 = d.q_sample(x0.cuda(), torch.tensor([500]).cuda(), noise.cuda()).cpu()
    assert torch.equal(got1, ref1)


def test_schedule_buffers_are_state_dict_entries():
    from cddpm.diffusion import GaussianDiffusion
    from oracle import diffusion_port

    d = GaussianDiffusion(torch.nn.Identity(), image_size=(96, 96), timesteps=1000, objective="pred_x0", channels=1)
    g = np.load(os.path.join(GOLD, "schedule.npz"))
    sd = d.state_dict()
    assert list(sd.keys()) == diffusion_port.BUFFER_NAMES
    for k in sd:
        assert np.array_equal(sd[k].numpy(), g[k]), k


@pytest.mark.parametrize("objective", ["pred_x0", "pred_noise"])
def test_single_step_reconstruction_vs_reference(objective):
    from cddpm.diffusion import GaussianDiffusion

    m, _, _ = _small_model()
    g = np.load(os.path.join(GOLD, "diffusion_small_32.npz"))
    img, cond, noise = (torch.from_numpy(g[k]).cuda() for k in ("img", "cond", "noise"))
    d = GaussianDiffusion(m, image_size=(32, 32), timesteps=1000, sampling_timesteps=1000, objective=objective,
                          channels=1, loss_type="l1", p2_loss_weight_gamma=0, cfg=Cfg(noisetype="simplex")).cuda()
    with torch.no_grad():
        loss, reco = d(img, cond=cond, t=499, noise=noise)
    ref = torch.from_numpy(g[f"reco_{objective}"])
    err = (reco.cpu() - ref).abs().max().item()
    lerr = abs(loss.item() - float(g[f"loss_{objective}"]))
    print(f"{objective}: reco max-abs {err:.4g}, loss err {lerr:.3g}")
    assert err <= 1e-2 and lerr <= 1e-3


def test_reverse_loop_vs_reference(monkeypatch):
    import cddpm.diffusion as dmod
    from cddpm.diffusion import GaussianDiffusion

    m, _, _ = _small_model()
    g = np.load(os.path.join(GOLD, "diffusion_small_32.npz"))
    img, cond = torch.from_numpy(g["img"]).cuda(), torch.from_numpy(g["cond"]).cuda()
    noises = [n.cuda() for n in torch.from_numpy(g["reverse_noises"])]
    T0 = int(g["reverse_T0"])
    d = GaussianDiffusion(m, image_size=(32, 32), timesteps=1000, sampling_timesteps=1000, objective="pred_x0",
                          channels=1, loss_type="l1", p2_loss_weight_gamma=0, cfg=Cfg(noisetype="simplex")).cuda()
    feed = list(noises)
    monkeypatch.setattr(dmod, "gen_noise", lambda cfg, shape, device=None: feed.pop(0))
    rec = d.sample(cond=cond, x_start=img * 2 - 1, start_t=T0, noise=True)
    assert not feed  # one draw for the initial q_sample + one per step, like the reference
    err = (rec.cpu() - torch.from_numpy(g["reverse_out"])).abs().max().item()
    print(f"reverse loop T0={T0}: max-abs {err:.4g}")
    assert err <= 1e-2


@pytest.mark.parametrize("objective", ["pred_x0", "pred_noise"])
def test_ddim_sample_vs_reference(objective):
    """sample() with sampling_timesteps < timesteps -> ddim_sample (cond_DDPM.py:466-515), start_t = 300 with the
    simplex flag: golden of the live reference (oracle/make_golden.py ddim), noise regenerated from the numpy seed."""
    from cddpm.diffusion import GaussianDiffusion

    m, _, _ = _small_model()
    g = np.load(os.path.join(GOLD, "ddim_small_32.npz"))
    img, cond = torch.from_numpy(g["img"]).cuda(), torch.from_numpy(g["cond"]).cuda()
    d = GaussianDiffusion(m, image_size=(32, 32), timesteps=1000, sampling_timesteps=6, objective=objective, channels=1,
                          loss_type="l1", p2_loss_weight_gamma=0, ddim_sampling_eta=0.7,
                          cfg=Cfg(noisetype="simplex")).cuda()
    assert d.is_ddim_sampling
    np.random.seed(5)
    rec = d.sample(cond=cond, x_start=img * 2 - 1, start_t=300, noise=True)
    err = (rec.cpu() - torch.from_numpy(g[f"ddim_simplex_{objective}"])).abs().max().item()
    # An iterated map: six UNet evaluations feed each other, and pred_x0 divides the prediction error by
    # sqrt(1/acp - 1) (cond_DDPM.py:385-389).  The yardstick stored with the golden is the deviation of the SAME
    # reference code under its configured precision (fp16 autocast, configs/trainer/default.yaml:7) from its fp32 path.
    amp = float(g[f"ddim_simplex_{objective}_amp16_dev"])
    print(f"ddim {objective}: max-abs {err:.4g} (reference fp16-autocast vs its fp32: {amp:.4g})")
    assert err <= max(1e-2, 1.5 * amp)


def test_ddim_sample_from_gaussian_start_vs_reference(monkeypatch):
    import cddpm.diffusion as dmod
    from cddpm.diffusion import GaussianDiffusion

    m, _, _ = _small_model()
    g = np.load(os.path.join(GOLD, "ddim_small_32.npz"))
    img, cond = torch.from_numpy(g["img"]).cuda(), torch.from_numpy(g["cond"]).cuda()
    d = GaussianDiffusion(m, image_size=(32, 32), timesteps=1000, sampling_timesteps=4, objective="pred_x0", channels=1,
                          loss_type="l1", p2_loss_weight_gamma=0, ddim_sampling_eta=1.0,
                          cfg=Cfg(noisetype="simplex")).cuda()
    draws = [torch.zeros(2, 1, 32, 32).cuda(), torch.from_numpy(g["ddim_xT"]).cuda()]  # first draw is discarded
    monkeypatch.setattr(dmod, "_randn", lambda shape, device: draws.pop(0))
    np.random.seed(6)
    rec = d.sample(cond=cond, x_start=img * 2 - 1, start_t=0, noise=None)
    assert not draws
    err = (rec.cpu() - torch.from_numpy(g["ddim_gauss_start"])).abs().max().item()
    print(f"ddim from x_T: max-abs {err:.4g}")
    assert err <= 1e-2


def test_p_sample_clip_denoised_flag_vs_reference(monkeypatch):
    import cddpm.diffusion as dmod
    from cddpm.diffusion import GaussianDiffusion

    m, _, _ = _small_model()
    g = np.load(os.path.join(GOLD, "ddim_small_32.npz"))
    cond = torch.from_numpy(g["cond"]).cuda()
    x, nz = torch.from_numpy(g["noclip_x"]).cuda(), torch.from_numpy(g["noclip_noise"]).cuda()
    d = GaussianDiffusion(m, image_size=(32, 32), timesteps=1000, sampling_timesteps=1000, objective="pred_x0",
                          channels=1, loss_type="l1", p2_loss_weight_gamma=0, cfg=Cfg(noisetype="simplex")).cuda()
    monkeypatch.setattr(dmod, "_randn", lambda shape, device: nz)
    for key, clip in (("noclip_out", False), ("clip_out", True)):
        out = d.p_sample(x.clone(), 400, clip_denoised=clip, cond=cond, noise=None)
        err = (out.cpu() - torch.from_numpy(g[key])).abs().max().item()
        print(f"p_sample clip_denoised={clip}: max-abs {err:.4g}")
        assert err <= 1e-3  # coef1[400] = 3.7e-3 scales the model error


def test_out_of_range_timesteps_raise_like_the_reference():
    """The reference indexes its schedule buffers with t (IndexError outside [0, T)); the kernels read raw pointers, so
    the host checks."""
    from cddpm.diffusion import GaussianDiffusion

    m, _, _ = _small_model()
    d = GaussianDiffusion(m, image_size=(32, 32), timesteps=1000, objective="pred_x0", channels=1,
                          cfg=Cfg(noisetype="simplex")).cuda()
    img = torch.rand(2, 1, 32, 32).cuda()
    cond = torch.randn(2, 128).cuda()
    with pytest.raises(IndexError):
        d(img, cond=cond, t=1000)
    with pytest.raises(IndexError):
        d(img, cond=cond, t=-1)
    with pytest.raises(IndexError):  # start_t=0 -> q_sample at index num_timesteps (cond_DDPM.py:452)
        d.sample(cond=cond, x_start=img * 2 - 1, start_t=0, noise=True)
    # broadcast-shaped noise is expanded once for q_sample and the finishing kernels (pred_noise reads it again)
    dn = GaussianDiffusion(m, image_size=(32, 32), timesteps=1000, objective="pred_noise", channels=1,
                           cfg=Cfg(noisetype="simplex")).cuda()
    nz = torch.randn(1, 1, 32, 32).cuda()
    with torch.no_grad():
        l1, r1 = dn(img, cond=cond, t=100, noise=nz)
        l2, r2 = dn(img, cond=cond, t=100, noise=nz.expand(2, 1, 32, 32).contiguous())
    # (the UNet's GroupNorm statistics are summed with atomics: two forwards agree to rounding, not bit for bit)
    assert (r1 - r2).abs().max().item() <= 2e-3 and abs(float(l1) - float(l2)) <= 1e-4
