"""src.models.DDPM_2D — B200 drop-in (reference: src/models/DDPM_2D.py:17-308)."""
from cddpm.ddpm_2d import DDPM_2D  # noqa: F401
