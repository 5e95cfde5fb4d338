"""src.models.modules.DDPM_encoder — B200 drop-in (reference: src/models/modules/DDPM_encoder.py:6-29)."""
from cddpm.encoder import get_encoder  # noqa: F401
