"""src.models.modules.cond_DDPM — B200 drop-in (reference: src/models/modules/cond_DDPM.py:289-655)."""
from cddpm.diffusion import (GaussianDiffusion, ModelPrediction, cosine_beta_schedule,  # noqa: F401
                             linear_beta_schedule)
from cddpm.noise import gen_noise  # noqa: F401
