"""src.utils.generate_noise — B200 drop-in (reference: src/utils/generate_noise.py:8-15)."""
from cddpm.noise import gen_noise  # noqa: F401
