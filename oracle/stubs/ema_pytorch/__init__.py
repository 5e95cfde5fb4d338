"""Stub: cond_DDPM.py:21 imports EMA but never uses it."""
class EMA:  # pragma: no cover
    pass
