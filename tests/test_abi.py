"""The C-ABI shared library: loads, exports every symbol include/lrpt_b200.h declares, and its
host-only entry points behave. No GPU compute here."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "lrpt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lrpt_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from meteor_demod_b200 import _lib
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(_lib.SYMBOLS) == names            # the binding covers exactly the header


def test_abi_version_and_strerror(lib):
    assert lib.lrpt_abi_version() == 1
    assert lib.lrpt_strerror(0) == b"ok"
    assert b"CUDA" in lib.lrpt_strerror(-2)


def test_struct_layouts_match_header(lib):
    from meteor_demod_b200._lib import Params, State, Status
    assert C.sizeof(Params) == 12 * 4
    assert C.sizeof(State) == 16 * 4 + 3 * 8
    assert C.sizeof(Status) == 5 * 4 + 4 + 3 * 8     # 4 bytes padding before the int64 block


def test_describe_rejects_bad_configs(lib):
    from meteor_demod_b200 import LrptError, describe
    import pytest
    for bad in (dict(bps=12), dict(interp_factor=0), dict(interp_factor=99), dict(rrc_order=-1), dict(samplerate=0),
                dict(symrate=0)):
        with pytest.raises(LrptError):
            describe(**bad)


def test_freq_delta_conversion(lib, oracle_mod):
    o = oracle_mod.Oracle.lib()
    for hz, sr in ((3500.0, 72000.0), (-1.0, 72000.0), (100.0, 80000.0)):
        assert lib.lrpt_freq_delta_from_hz(hz, sr) == o.lrpt_oracle_freq_delta(hz, sr)
    assert lib.lrpt_freq_delta_from_hz(-1.0, 72000.0) < 0       # stays negative => FREQ_MAX default (pll.c:30)


def test_create_without_gpu_fails_loudly(lib):
    """No CPU fallback: without a device lrpt_create returns LRPT_ERR_CUDA (skipped where a GPU exists)."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from meteor_demod_b200 import Demod, LrptError
    with pytest.raises(LrptError) as e:
        Demod()
    assert e.value.code == -2


def test_egress_gating_rules():
    from meteor_demod_b200 import egress
    soft = (np.arange(2 * 1500) % 251 - 125).astype(np.int8).reshape(1500, 2)
    # never locked: only the partial tail block is flushed (main.c:321 is not gated)
    assert egress.gate(soft, -1) == soft[1024:].tobytes()
    # locked inside block 1 => output starts at symbol 512
    assert egress.gate(soft, 700) == soft[512:].tobytes()
    assert egress.gate(soft, 0) == soft.tobytes()
    # locked only in the tail block: no full block qualifies
    assert egress.gate(soft, 1100) == soft[1024:].tobytes()
    # reference-compatible tail: 2*ring_idx bytes, valid then stale ring content
    out = egress.gate(soft, 0, ref_compatible_tail=True)
    tail = soft[1024:].reshape(-1)
    assert len(out) == 2 * 1024 + 2 * tail.size
    stale = np.frombuffer(out[-tail.size:], np.int8)
    prev = soft[512:1024].reshape(-1)
    assert np.array_equal(stale[: 1024 - tail.size], prev[tail.size:1024][: stale.size])
    assert egress.consumed_samples(100000, 16) == 3 * 32768 // 4


def test_sharded_entry_point_checks_arguments_and_needs_a_gpu(lib):
    """lrpt_sharded_process: argument errors are reported before any CUDA call; without a device the call fails
    with LRPT_ERR_CUDA (no CPU path). Skipped parts aside, no compute happens here."""
    import pytest
    import torch
    from meteor_demod_b200 import LrptError, sharded
    raw = np.zeros(2 * 4096, np.int16)
    for bad in (dict(chunk=1001), dict(warm=12), dict(overlap=0), dict(bps=12)):
        kw = dict(chunk=1024, warm=512, overlap=64, symrate=72000, bps=16)
        kw.update(bad)
        with pytest.raises(LrptError) as e:
            sharded.process_host(raw, **kw)
        assert e.value.code == -1, bad
    if not torch.cuda.is_available():
        with pytest.raises(LrptError) as e:
            sharded.process_host(raw, chunk=1024, warm=512, overlap=64, symrate=72000, bps=16)
        assert e.value.code == -2
