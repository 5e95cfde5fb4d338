from pytorch_lightning import LightningModule  # noqa: F401
