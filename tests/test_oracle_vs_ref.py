"""Pin the oracle port to the compiled reference (oracle/_ref, built from /root/reference by
oracle/Makefile) and to the known-answer tripwires of SURVEY.md A.3. The _ref comparisons skip
when neither /root/reference nor a prebuilt oracle/_ref is available."""
import numpy as np
import pytest

from conftest import CONFIGS, bits, make_case


def need_ref(oracle_mod):
    if not oracle_mod.have_ref("strict"):
        try:
            oracle_mod.build()
        except Exception:
            pass
    if not oracle_mod.have_ref("strict"):
        pytest.skip("compiled reference (oracle/_ref) not available")


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_port_equals_reference_strict(name, oracle_mod):
    need_ref(oracle_mod)
    cfg = CONFIGS[name]
    raw = make_case(name, 400_000)
    o = oracle_mod.Oracle(**cfg)
    r = oracle_mod.Ref(kind="strict", **cfg)
    a, b = o.process(raw), r.process(raw)
    assert a.nsym == b.nsym > 100_000
    assert np.array_equal(bits(a.sym), bits(b.sym))
    assert np.array_equal(a.soft, b.soft)
    assert np.array_equal(a.sample_idx, b.sample_idx)
    assert np.array_equal(a.lock_once, b.lock_once)
    assert np.array_equal(bits(o.taps()), bits(r.taps()))
    assert np.array_equal(bits(o.history()), bits(r.history()))
    so, sr = o.state(), r.state()
    for k in ("t_prev", "t_phase", "t_freq", "agc_gain", "agc_bias_re", "agc_bias_im", "p_freq", "p_phase", "p_err"):
        assert np.float32(so[k]).tobytes() == np.float32(sr[k]).tobytes(), k
    assert so["p_locked"] == sr["p_locked"] and so["p_locked_once"] == sr["p_locked_once"]


def test_port_locks_like_reference(oracle_mod):
    """C1 at +700 Hz: lock happens and the carrier / clock estimates are the programmed ones."""
    need_ref(oracle_mod)
    cfg = CONFIGS["C1_qpsk72k_s16_o32_L5"]
    raw = make_case("C1_qpsk72k_s16_o32_L5", 600_000)
    o = oracle_mod.Oracle(**cfg)
    out = o.process(raw)
    st = o.state()
    assert st["p_locked"] == 1 and 30_000 < st["first_lock_symbol"] < 60_000
    assert abs(st["p_freq"] * 72000 / (2 * np.pi) - 700.0) < 2.0
    assert abs(st["t_freq"] * 230000 * 5 / (2 * np.pi) - 72000.0) < 1.0
    assert int(np.argmax(out.lock_once)) == st["first_lock_symbol"]


def test_scalar_blocks_equal_reference(oracle_mod):
    need_ref(oracle_mod)
    r = oracle_mod.Ref(kind="strict")
    L = oracle_mod.Oracle.lib()
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.uniform(-8, 8, 20000), np.linspace(-7.5, 7.5, 3001),
                         [0.0, -0.0, 6.2831855, -6.2831855, 3.1415927, 1e-8, -1e-8]]).astype(np.float32)
    for x in xs.tolist():
        assert L.lrpt_oracle_fast_sin(x) == r.L.ref_fast_sin(x)
        assert L.lrpt_oracle_fast_cos(x) == r.L.ref_fast_cos(x)
    a = (rng.standard_normal((20000, 2)) * rng.choice([1e-3, 1.0, 200.0, 3e4], (20000, 1))).astype(np.float32)
    for re, im in a.tolist():
        assert L.lrpt_oracle_cabsf(re, im) == r.L.ref_cabsf(re, im)          # libm cabsf model
    for order, interp in ((16, 3), (32, 5), (64, 8)):
        taps = (2 * order + 1) * interp
        osf = np.float32(np.float32(230000) / 72000) * np.float32(interp)
        for n in range(0, taps, 7):
            assert L.lrpt_oracle_rrc_coeff(n, taps, osf, 0.6) == r.L.ref_rrc_coeff(n, taps, osf, 0.6)


# ---- known answers (extracted from the strict reference build) -------------------

TAP_HASHES_72K = {(16, 3): 0xdac8a427, (16, 5): 0x78da6ae2, (16, 8): 0xbdeea03a, (32, 3): 0x5d5759d0,
                  (32, 5): 0x54ab68c0, (32, 8): 0x1e715cfe, (64, 3): 0xeace77b9, (64, 5): 0xac64e76f,
                  (64, 8): 0x2a16f80e, (128, 3): 0x8211bdd0, (128, 5): 0x87749f39, (128, 8): 0xc215eba1}
TAP_HASHES_80K = {(16, 3): 0xf715820a, (16, 5): 0x7c660e85, (16, 8): 0x131a8124, (32, 3): 0x0c8a4285,
                  (32, 5): 0xb182bd4e, (32, 8): 0xc6a2563b, (64, 3): 0x3985c458, (64, 5): 0x8087252e,
                  (64, 8): 0x68de0782, (128, 3): 0xe60fdd12, (128, 5): 0x98d5eb85, (128, 8): 0x4086eede}


@pytest.mark.parametrize("symrate,table", [(72000, TAP_HASHES_72K), (80000, TAP_HASHES_80K)])
def test_tap_bank_hashes(symrate, table, oracle_mod, lib):
    from meteor_demod_b200 import describe
    for (order, interp), want in table.items():
        o = oracle_mod.Oracle(symrate=symrate, order=order, interp=interp)
        assert oracle_mod.fnv1a32(o.taps()) == want, (order, interp)
        d = describe(symrate=symrate, rrc_order=order, interp_factor=interp)
        assert oracle_mod.fnv1a32(d["taps"]) == want, ("product", order, interp)
        assert np.all(np.isfinite(d["taps"]))


def test_known_scalars(oracle_mod):
    L = oracle_mod.Oracle.lib()
    f = float.fromhex
    sin = {0.0: 0.0, -0.5: -f("0x1.ed9p-2"), -1.0: -f("0x1.afd8p-1"), -4.0: f("0x1.84cp-1"),
           -6.283185: f("0x1p-13"), 1.0: f("0x1.afd8p-1"), 2.5: f("0x1.33e8p-1"), 7.0: f("0x1.51c8p-1"),
           -7.0: -f("0x1.51c8p-1")}
    for x, want in sin.items():
        assert L.lrpt_oracle_fast_sin(x) == want, x
    cos = {-0.5: f("0x1.c22p-1"), -1.0: f("0x1.1608p-1"), 1.0: f("0x1.161p-1"), -4.0: -f("0x1.502p-1"),
           2.5: -f("0x1.9b5p-1")}
    for x, want in cos.items():
        assert L.lrpt_oracle_fast_cos(x) == want, x
    o = oracle_mod.Oracle()
    lut = np.array(o.s.lut_tanh[:], np.float32)
    assert lut[16] == 0 and lut[0] == -1 and lut[31] == 1
    assert float(lut[17]) == f("0x1.85efacp-1") and float(lut[18]) == f("0x1.ed9506p-1")
    assert float(lut[15]) == -f("0x1.85efacp-1")
    # loop constants of C1
    assert o.s.t_center == f("0x1.92d2bep-2") and o.s.t_alpha == f("0x1.4f89ap-15")
    assert o.s.t_beta == f("0x1.b7cbbap-32") and o.s.t_maxdev == f("0x1.92d2bep-14")
    assert o.s.p_alpha == f("0x1.02ccfap-13") and o.s.p_beta == f("0x1.05a5f2p-27")
    o2 = oracle_mod.Oracle(symrate=80000, oqpsk=1, bps=8)
    assert o2.s.p_alpha == f("0x1.d1d17ap-13") and o2.s.p_beta == f("0x1.a7d964p-26")
    assert abs(o2.s.p_fmax - 0.15) < 1e-7 and o2.s.t_center == f("0x1.bf94d2p-2")


def test_quantiser(oracle_mod):
    L = oracle_mod.Oracle.lib()
    for v, want in ((0.0, 0), (1.9, 0), (2.0, 1), (-1.9, 0), (-2.0, -1), (253.9, 126), (254.0, 127), (1e9, 127),
                    (-254.0, -127), (-1e9, -127), (3.99, 1), (-3.99, -1)):
        assert L.lrpt_oracle_quantise(v) == want, v


def _same_run(o, r, raw):
    a, b = o.process(raw), r.process(raw)
    assert a.nsym == b.nsym
    assert np.array_equal(bits(a.sym), bits(b.sym)) and np.array_equal(a.soft, b.soft)
    assert np.array_equal(a.sample_idx, b.sample_idx) and np.array_equal(a.lock_once, b.lock_once)
    assert np.array_equal(bits(o.history()), bits(r.history()))
    so, sr = o.state(), r.state()
    for k in ("t_prev", "t_phase", "t_freq", "agc_gain", "agc_bias_re", "agc_bias_im", "p_freq", "p_phase", "p_err"):
        assert np.float32(so[k]).tobytes() == np.float32(sr[k]).tobytes(), k
    return a.nsym


def test_port_equals_reference_on_the_config5_grid(oracle_mod):
    """BASELINE config 5 in full -- RRC order {16,32,64,128} x oversampling {3,5,8}, QPSK 72k s16 and OQPSK 80k
    u8 -- port against the compiled reference on 120k samples per corner (taps, symbols, state, delay line)."""
    need_ref(oracle_mod)
    from meteor_demod_b200 import synth
    for oq, symrate, bps in ((0, 72000, 16), (1, 80000, 8)):
        raw = synth.make_raw(120_000, symrate=symrate, oqpsk=bool(oq), bps=bps, seed=8, cfo_hz=-333.0)
        for order in (16, 32, 64, 128):
            for interp in (3, 5, 8):
                cfg = dict(symrate=symrate, oqpsk=oq, bps=bps, order=order, interp=interp)
                o, r = oracle_mod.Oracle(**cfg), oracle_mod.Ref(kind="strict", **cfg)
                assert np.array_equal(bits(o.taps()), bits(r.taps())), cfg
                assert _same_run(o, r, raw) > 30_000, cfg


def test_port_equals_reference_on_random_parameters(oracle_mod):
    """Seeded fuzz over everything demod_init takes (demod.h:29): sample and symbol rates, order, oversampling,
    PLL bandwidth, frequency limit, mode, sample format; ragged pushes on the port side."""
    need_ref(oracle_mod)
    from meteor_demod_b200 import synth
    rng = np.random.default_rng(2024)
    for trial in range(16):
        oq = int(rng.integers(0, 2))
        bps = int(rng.choice([8, 16, 32]))
        symrate = int(rng.choice([72000, 80000, 64000]))
        fs = int(rng.choice([230000, 250000, 1_000_000 // 4]))
        cfg = dict(samplerate=fs, symrate=symrate, oqpsk=oq, bps=bps, order=int(rng.integers(4, 90)),
                   interp=int(rng.integers(1, 9)), pll_bw=float(rng.choice([0.5, 1.0, 2.0, 4.0])),
                   freq_max=float(rng.choice([-1.0, 0.05, 0.2, 0.6])))
        raw = synth.make_raw(60_000, symrate=symrate, fs=fs, oqpsk=bool(oq), bps=bps, seed=100 + trial,
                             cfo_hz=float(rng.uniform(-1500, 1500)))
        o, r = oracle_mod.Oracle(**cfg), oracle_mod.Ref(kind="strict", **cfg)
        w = r.process(raw)
        cuts = np.sort(rng.integers(0, 60_000, 5))
        parts, prev = [], 0
        for c in list(cuts) + [60_000]:
            parts.append(o.process(raw[2 * prev: 2 * int(c)]).soft)
            prev = int(c)
        got = np.concatenate(parts)
        assert got.shape[0] == w.nsym and np.array_equal(got, w.soft), (trial, cfg)
        assert np.array_equal(bits(o.history()), bits(r.history())), (trial, cfg)


def test_reference_power_on_image_restores_a_fresh_process(oracle_mod):
    """bench.py's reference arm runs many power-on streams in ONE process: pyoracle.Ref(in_place=True).power_on() puts
    the library's writable data back as it was right after loading (oracle/ref_driver.c). A stream demodulated after
    power_on() must equal the same stream in a fresh private copy of the library, function-scope statics included
    (OQPSK: timing.c:43, demod.c:54; the sweep direction: pll.c:112)."""
    if not oracle_mod.have_ref("fma"):
        pytest.skip("oracle/_ref not built")
    from meteor_demod_b200 import synth
    for cfg in (dict(symrate=72000, oqpsk=0, bps=16, order=32, interp=5), dict(symrate=80000, oqpsk=1, bps=8, order=32, interp=5)):
        kw = dict(symrate=cfg["symrate"], oqpsk=bool(cfg["oqpsk"]), bps=cfg["bps"])
        a = synth.make_raw(50_000, seed=5, cfo_hz=300.0, **kw)
        b = synth.make_raw(40_000, seed=6, cfo_hz=-900.0, **kw)
        for kind in ("strict", "fma"):
            fresh = oracle_mod.Ref(kind=kind, **cfg).process(b)
            r = oracle_mod.Ref(kind=kind, in_place=True, **cfg)
            r.process(a)
            r.power_on()
            again = r.process(b)
            assert again.nsym == fresh.nsym
            assert np.array_equal(again.sym.view(np.uint32), fresh.sym.view(np.uint32)), (cfg, kind)
            assert np.array_equal(again.lock_once, fresh.lock_once)
