"""Repository package root of the B200-native cDDPM path.

Put this directory on sys.path to get
  * `cddpm`  — the host package (ctypes binding + engine + reference-interface mirrors), and
  * `src`    — a drop-in overlay with the reference's dotted paths (src.models.DDPM_2D, ...).
"""
