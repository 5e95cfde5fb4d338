"""src.models.modules.spark.Spark_2D — only the encoder wrapper used by cDDPM (reference: spark/Spark_2D.py:268-290)."""
from cddpm.encoder import SparK_2D_encoder  # noqa: F401
