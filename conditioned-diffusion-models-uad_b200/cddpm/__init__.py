"""cddpm — host-side Python of the B200-native cDDPM reconstruction + anomaly-scoring path.

Everything numerical runs in libcddpm_b200.so (hand-written sm_100a CUDA, C ABI in include/cddpm_b200.h); this
package mirrors the reference's module interfaces on top of it.
"""
from ._lib import CddpmError, lib  # noqa: F401

__version__ = "0.1"
