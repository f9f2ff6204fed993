import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a CUDA device skips the GPU tests instead of failing them
    (the product has no CPU fallback, so they cannot pass there)."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_mod():
    """oracle.pyoracle with the port built (and the reference when its sources or a prebuilt copy exist)."""
    from oracle import pyoracle
    if not os.path.exists(pyoracle.PORT_SO):
        pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def lib():
    """liblrpt_b200.so, built in-tree if missing (nvcc cross-compiles without a GPU)."""
    from meteor_demod_b200 import _lib, build
    if not os.path.exists(_lib.SO_PATH):
        build.build()
    return _lib.load()


def bits(a):
    """Bit pattern view of a float32 array, for exact comparison (distinguishes -0.0, NaN payloads)."""
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# the configurations BASELINE.json names (C1, C2, C3 flags) + two sweep corners (C5)
CONFIGS = {
    "C1_qpsk72k_s16_o32_L5": dict(symrate=72000, oqpsk=0, bps=16, order=32, interp=5),
    "C2_oqpsk80k_u8_o32_L5": dict(symrate=80000, oqpsk=1, bps=8, order=32, interp=5),
    "C3_qpsk72k_s16_o64_L8": dict(symrate=72000, oqpsk=0, bps=16, order=64, interp=8),
    "C5_qpsk72k_f32_o16_L3": dict(symrate=72000, oqpsk=0, bps=32, order=16, interp=3),
    "C5_oqpsk80k_s16_o128_L8": dict(symrate=80000, oqpsk=1, bps=16, order=128, interp=8),
}


def make_case(name, nsamples, seed=1, cfo_hz=700.0, rms=6000.0):
    from meteor_demod_b200 import synth
    c = CONFIGS[name]
    return synth.make_raw(nsamples, symrate=c["symrate"], oqpsk=bool(c["oqpsk"]), bps=c["bps"], seed=seed,
                          cfo_hz=cfo_hz, rms=rms)
