"""Stub: open_dict is a no-op context manager (DDPM_2D.py:13,29)."""
from contextlib import contextmanager


@contextmanager
def open_dict(cfg):
    yield cfg


class DictConfig(dict):
    pass


class OmegaConf:
    pass
