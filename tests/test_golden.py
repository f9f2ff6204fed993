"""Oracle port (and the host-side derivations of the product) against the committed golden
vectors, which were produced by the compiled reference itself (tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

from conftest import bits

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def load(path):
    g = np.load(path)
    cfg = dict(zip(g["cfg_names"].tolist(), g["cfg_vals"].tolist()))
    state = dict(zip(g["state_names"].tolist(), g["state_bits"].tolist()))
    return g, cfg, state


def test_golden_present():
    assert len(GOLDEN) >= 4


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_golden(path, oracle_mod):
    g, cfg, state = load(path)
    o = oracle_mod.Oracle(**cfg)
    assert np.array_equal(bits(o.taps()), g["taps_bits"])
    out = o.process(g["raw"])
    assert out.nsym == g["soft"].shape[0]
    assert np.array_equal(bits(out.sym), g["sym_bits"])           # bit-exact float symbols
    assert np.array_equal(out.soft, g["soft"])                    # int8 soft symbols
    assert np.array_equal(out.sample_idx, g["sample_idx"])
    assert np.array_equal(out.lock_once, g["lock_once"])
    assert np.array_equal(bits(o.history()), g["history_bits"])
    st = o.state()
    for k in ("t_prev", "t_phase", "t_freq", "agc_gain", "agc_bias_re", "agc_bias_im", "p_freq", "p_phase", "p_err"):
        assert int(np.float32(st[k]).view(np.uint32)) == state[k], k
    assert st["p_locked"] == state["p_locked"] and st["p_locked_once"] == state["p_locked_once"]


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_chunked_equals_one_shot(path, oracle_mod):
    """Block boundaries must be invisible: ragged pushes (1, 7, 4096, ... samples) give the same stream."""
    g, cfg, _ = load(path)
    raw = g["raw"]
    o = oracle_mod.Oracle(**cfg)
    n = raw.size // 2
    cuts = [0, 1, 8, 9, 64, 65, 4096, 4099, 20000, n]
    sym, soft = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        out = o.process(raw[2 * a: 2 * b])
        sym.append(out.sym)
        soft.append(out.soft)
    assert np.array_equal(bits(np.concatenate(sym)), g["sym_bits"])
    assert np.array_equal(np.concatenate(soft), g["soft"])


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_product_host_derivation_matches_golden(path, lib):
    """lrpt_describe (host only, no GPU): tap banks bit-identical to the reference's filter_init_rrc."""
    from meteor_demod_b200 import describe
    g, cfg, state = load(path)
    d = describe(symrate=cfg["symrate"], oqpsk=cfg["oqpsk"], bps=cfg["bps"], rrc_order=cfg["order"],
                 interp_factor=cfg["interp"])
    assert np.array_equal(bits(d["taps"]), g["taps_bits"])
    for k in ("t_alpha", "t_beta", "p_alpha", "p_beta", "p_fmax", "t_center", "t_maxdev"):
        assert int(np.float32(d["consts"][k]).view(np.uint32)) == state[k], k
