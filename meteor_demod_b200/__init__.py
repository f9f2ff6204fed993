"""meteor_demod_b200 -- B200-native LRPT demodulator hot path.

The product is liblrpt_b200.so (hand-written sm_100a CUDA behind the C ABI in
include/lrpt_b200.h). This package is the thin host-side mirror of the
reference's demodulator interface used by tests and bench.
"""
from .demod import Demod, describe, make_params, symbol_capacity  # noqa: F401
from ._lib import LrptError  # noqa: F401

__all__ = ["Demod", "describe", "make_params", "symbol_capacity", "LrptError"]
