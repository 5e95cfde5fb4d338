"""src.models.modules.OpenAI_Unet — B200 drop-in (reference: src/models/modules/OpenAI_Unet.py:483-1006)."""
from cddpm.unet import (AttentionBlock, Downsample, GroupNorm32, ResBlock, TimestepEmbedSequential,  # noqa: F401
                        UNetModel, Upsample)
