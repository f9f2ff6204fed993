"""The stand-alone feed-forward stage (csrc/fir_stage.cu, lrpt_fir_stage_device): ingest + all-phase RRC polyphase
FIR, against the oracle's filter_get at every (sample, sub-step). mode 0 is bit-exact; mode 1 (fused multiply-add)
is within a few ulp of the accumulated magnitude."""
import ctypes as C

import numpy as np
import pytest

from conftest import bits

CASES = [dict(symrate=72000, oqpsk=0, bps=16, order=32, interp=5),      # C1 / C2 filter
         dict(symrate=80000, oqpsk=1, bps=8, order=32, interp=5),
         dict(symrate=72000, oqpsk=0, bps=16, order=64, interp=8),      # C3
         dict(symrate=72000, oqpsk=0, bps=32, order=16, interp=3),
         dict(symrate=80000, oqpsk=1, bps=16, order=128, interp=8),
         dict(symrate=72000, oqpsk=0, bps=8, order=2, interp=1)]


def fir_stage(raw2d, cfg, mode):
    import torch
    from meteor_demod_b200 import _lib
    from meteor_demod_b200.demod import make_params
    lib = _lib.load()
    p = make_params(symrate=cfg["symrate"], oqpsk=cfg["oqpsk"], bps=cfg["bps"], rrc_order=cfg["order"], interp_factor=cfg["interp"])
    rows, n = raw2d.shape[0], raw2d.shape[1] // 2
    t_raw = torch.from_numpy(raw2d).cuda()
    out = torch.full((rows, n * cfg["interp"] * 2), float("nan"), dtype=torch.float32, device="cuda")
    rc = lib.lrpt_fir_stage_device(C.byref(p), t_raw.data_ptr(), t_raw.stride(0) * t_raw.element_size(), rows, n,
                                   out.data_ptr(), out.stride(0) * 4, mode, None)
    assert rc == 0, rc
    torch.cuda.synchronize()
    return out.cpu().numpy().reshape(rows, n, cfg["interp"], 2)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", CASES, ids=["%s%d_o%d_L%d" % ("oq" if c["oqpsk"] else "q", c["bps"], c["order"], c["interp"]) for c in CASES])
def test_fir_stage_equals_filter_get(cfg, oracle_mod, lib):
    from meteor_demod_b200 import synth
    per16 = 16 // (cfg["bps"] // 4)
    for n in (3000, 1024, 1031 // per16 * per16 + per16, 8):              # several tiles, one tile, ragged tail, tiny
        raw = np.stack([synth.make_raw(n, symrate=cfg["symrate"], oqpsk=bool(cfg["oqpsk"]), bps=cfg["bps"], seed=60 + s,
                                       cfo_hz=100.0 * s) for s in range(3)])
        got = fir_stage(raw, cfg, 0)
        fma = fir_stage(raw, cfg, 1)
        # bits 1-2 of mode force one or two output samples per thread: a tuning knob, never a different result
        for ns in (1, 2):
            assert np.array_equal(bits(fir_stage(raw, cfg, ns << 1)), bits(got)), (n, ns)
            assert np.array_equal(bits(fir_stage(raw, cfg, (ns << 1) | 1)), bits(fma)), (n, ns)
        for s in range(3):
            want = oracle_mod.Oracle(**cfg).fir_all(raw[s])
            assert np.array_equal(bits(got[s]), bits(want)), (n, s)
            scale = np.abs(want).max() + 1.0
            assert np.abs(fma[s] - want).max() <= 3e-6 * scale * cfg["order"], (n, s)


@pytest.mark.gpu
def test_fir_stage_rejects_bad_arguments(lib):
    import torch
    from meteor_demod_b200 import _lib
    from meteor_demod_b200.demod import make_params
    L = _lib.load()
    raw = torch.zeros((1, 64), dtype=torch.int16, device="cuda")
    out = torch.zeros((1, 32 * 5 * 2), dtype=torch.float32, device="cuda")
    p = make_params()
    assert L.lrpt_fir_stage_device(C.byref(p), raw.data_ptr(), 128, 1, 32, out.data_ptr(), 32 * 40 - 16, 0, None) == _lib.LRPT_ERR_ARG
    assert L.lrpt_fir_stage_device(C.byref(p), raw.data_ptr() + 2, 128, 1, 16, out.data_ptr(), 32 * 40, 0, None) == _lib.LRPT_ERR_ARG
    p9 = make_params(interp_factor=9)
    assert L.lrpt_fir_stage_device(C.byref(p9), raw.data_ptr(), 128, 1, 32, out.data_ptr(), 32 * 72, 0, None) == _lib.LRPT_ERR_ARG
