"""Drop-in overlay with the reference's dotted module paths (Hydra `_target_: src.models.DDPM_2D.DDPM_2D`, ...).
Put `conditioned-diffusion-models-uad_b200/` ahead of the reference checkout on PYTHONPATH (or copy these files over
the reference's) and the cDDPM hot path resolves to the B200 implementation; see INTEGRATION.md."""
