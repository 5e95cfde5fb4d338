"""Stub: only the dictionary key constant is used (DDPM_2D.py:10,119,177)."""
DATA = "data"
