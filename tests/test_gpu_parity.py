"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle and the
golden vectors produced by the reference itself. Bit-exact everywhere: float symbols, int8 soft
symbols, symbol counts and the complete loop state (the contract is integer-like: the strict-IEEE
reference is reproduced operation by operation)."""
import glob
import os

import numpy as np
import pytest

from conftest import CONFIGS, bits, make_case

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
KERNELS = ["simple", "ws", "spec", "lane"]
STATE_KEYS = ("t_phase", "t_freq", "t_prev", "agc_gain", "agc_bias_re", "agc_bias_im", "p_phase", "p_freq", "p_err")
INT_KEYS = ("p_locked", "p_locked_once", "p_updown", "t_dual_state", "nsamples", "nsymbols", "first_lock_symbol")


def demod_for(cfg, kernel, nstreams=1):
    from meteor_demod_b200 import Demod, LrptError
    try:
        return Demod(symrate=cfg["symrate"], oqpsk=cfg["oqpsk"], bps=cfg["bps"], rrc_order=cfg["order"],
                     interp_factor=cfg["interp"], nstreams=nstreams, kernel=kernel)
    except LrptError as e:
        if kernel in ("ws", "spec", "lane") and e.code == -1:
            pytest.skip("configuration not covered by the %s kernel" % kernel)
        raise


def assert_state_equal(gpu_state, oracle):
    so = oracle.state()
    for k in STATE_KEYS + ("oq_inphase",):
        assert np.float32(gpu_state[k]).tobytes() == np.float32(so[k]).tobytes(), k
    for k in INT_KEYS:
        assert gpu_state[k] == so[k], k
    assert np.array_equal(bits(gpu_state["history"]), bits(oracle.history()))


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_matches_reference_golden(path, kernel, lib):
    g = np.load(path)
    cfg = dict(zip(g["cfg_names"].tolist(), g["cfg_vals"].tolist()))
    d = demod_for(cfg, kernel)
    assert np.array_equal(bits(d.taps()), g["taps_bits"])
    raw = g["raw"].reshape(1, -1)
    soft, counts, symf = d.process_batch(raw, want_float=True)
    n = int(counts[0])
    assert n == g["soft"].shape[0]
    assert np.array_equal(bits(symf[0, :n]), g["sym_bits"])
    assert np.array_equal(soft[0, :n], g["soft"])
    st = d.state()
    want = dict(zip(g["state_names"].tolist(), g["state_bits"].tolist()))
    for k in STATE_KEYS:
        assert int(np.float32(st[k]).view(np.uint32)) == want[k], k
    assert st["p_locked"] == want["p_locked"] and st["p_locked_once"] == want["p_locked_once"]
    assert np.array_equal(bits(st["history"]), g["history_bits"])
    lock = g["lock_once"]
    assert st["first_lock_symbol"] == (int(np.argmax(lock)) if lock.any() else -1)
    assert d.launch_count() >= 1 and d.kernel_name() == kernel


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_cuda_matches_oracle(name, kernel, oracle_mod, lib):
    """Seeded synthetic input per BASELINE config, 300k samples (lock is reached for C1 at +700 Hz)."""
    cfg = CONFIGS[name]
    raw = make_case(name, 300_000, seed=5)
    o = oracle_mod.Oracle(**cfg)
    want = o.process(raw)
    d = demod_for(cfg, kernel)
    soft, counts, symf = d.process_batch(raw.reshape(1, -1), want_float=True)
    n = int(counts[0])
    assert n == want.nsym
    assert np.array_equal(bits(symf[0, :n]), bits(want.sym))
    assert np.array_equal(soft[0, :n], want.soft)
    assert_state_equal(d.state(), o)


@pytest.mark.parametrize("kernel", KERNELS)
def test_ragged_pushes_and_state_roundtrip(kernel, oracle_mod, lib):
    """Ragged block sizes (empty, 1 sample, < taps, odd) through lrpt_process; then export the state,
    import it into a second handle and continue there: the stream must not notice."""
    name = "C1_qpsk72k_s16_o32_L5"
    cfg = CONFIGS[name]
    raw = make_case(name, 120_000, seed=9, cfo_hz=30.0)
    want = oracle_mod.Oracle(**cfg).process(raw)
    d1 = demod_for(cfg, kernel)
    cuts = [0, 0, 1, 2, 9, 64, 65, 129, 1000, 4097, 50_001, 80_000]
    got = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        s, _ = d1.process(raw[2 * a: 2 * b])
        got.append(s)
    blob = d1.export_state()
    assert len(blob) == d1.state_size()
    d2 = demod_for(cfg, kernel)
    d2.import_state(blob)
    s, first = d2.process(raw[2 * 80_000:])
    got.append(s)
    got = np.concatenate(got)
    assert got.shape[0] == want.nsym
    assert np.array_equal(got, want.soft)
    assert first == (int(np.argmax(want.lock_once)) if want.lock_once.any() else -1)
    with pytest.raises(Exception):
        d2.import_state(blob[:-8])
    bad = bytearray(blob)
    bad[-4:] = np.float32(0.5).tobytes()                   # no 16-bit input leaves 0.5 in the delay line
    with pytest.raises(Exception):
        d2.import_state(bytes(bad))


@pytest.mark.parametrize("kernel", KERNELS)
def test_batch_of_streams_is_independent(kernel, oracle_mod, lib):
    """37 streams with different signals in one launch == 37 separate oracle runs; also checks the
    device-buffer entry point against the host-buffer one."""
    import torch
    name = "C2_oqpsk80k_u8_o32_L5"
    cfg = CONFIGS[name]
    ns, n = 37, 60_000
    raw = np.stack([make_case(name, n, seed=100 + s, cfo_hz=-200.0 + 13 * s) for s in range(ns)])
    d = demod_for(cfg, kernel, nstreams=ns)
    soft, counts, symf = d.process_batch(raw, want_float=True)
    for s in range(ns):
        o = oracle_mod.Oracle(**cfg)
        w = o.process(raw[s])
        assert counts[s] == w.nsym, s
        assert np.array_equal(bits(symf[s, :w.nsym]), bits(w.sym)), s
        assert np.array_equal(soft[s, :w.nsym], w.soft), s
        assert_state_equal(d.state(s), o)
    d.reset()
    cap = d.capacity(n)
    cap16 = (cap + 7) // 8 * 8
    t_raw = torch.from_numpy(raw).cuda()
    t_soft = torch.zeros((ns, 2 * cap16), dtype=torch.int8, device="cuda")
    t_n = torch.zeros(ns, dtype=torch.int32, device="cuda")
    d.process_device(t_raw, t_soft, nsym=t_n)
    d.sync()
    assert np.array_equal(t_n.cpu().numpy().astype(np.uint32), counts)
    assert np.array_equal(d.counts(), counts)
    dev = t_soft.cpu().numpy().reshape(ns, cap16, 2)
    for s in range(ns):
        assert np.array_equal(dev[s, :counts[s]], soft[s, :counts[s]]), s


@pytest.mark.parametrize("kernel", KERNELS)
def test_capacity_overflow_is_reported(kernel, lib):
    from meteor_demod_b200 import _lib
    name = "C1_qpsk72k_s16_o32_L5"
    raw = make_case(name, 20_000)
    d = demod_for(CONFIGS[name], kernel)
    soft, _ = d.process(raw, cap=100)
    assert d.last_rc == _lib.LRPT_ERR_CAP and soft.shape[0] == 100
    assert d.status()["nsymbols"] > 6000          # state advanced over the whole block


@pytest.mark.parametrize("kernel", KERNELS)
def test_edge_inputs(kernel, oracle_mod, lib):
    """All-zero input, full-scale input (AGC limit cycle region) and a DC-only input."""
    name = "C1_qpsk72k_s16_o32_L5"
    cfg = CONFIGS[name]
    rng = np.random.default_rng(0)
    cases = [np.zeros(2 * 30_000, np.int16),
             rng.choice(np.array([-32768, 32767], np.int16), 2 * 30_000),
             np.full(2 * 30_000, 1234, np.int16),
             make_case(name, 30_000, rms=20000.0)]
    for raw in cases:
        o = oracle_mod.Oracle(**cfg)
        w = o.process(raw)
        d = demod_for(cfg, kernel)
        soft, counts, symf = d.process_batch(raw.reshape(1, -1), want_float=True)
        assert counts[0] == w.nsym
        assert np.array_equal(bits(symf[0, :w.nsym]), bits(w.sym))
        assert np.array_equal(soft[0, :w.nsym], w.soft)
        assert_state_equal(d.state(), o)


def test_configurations_outside_the_fast_kernel_fall_back_exactly(oracle_mod, lib):
    """taps > 257 are served by the lane kernel and interp > 8 by the simple kernel under LRPT_KERNEL_AUTO;
    asking for the warp-specialised kernel explicitly is refused. Results stay bit-exact."""
    from meteor_demod_b200 import Demod, LrptError
    for cfg, served_by in ((dict(symrate=72000, oqpsk=0, bps=16, order=140, interp=5), "lane"),
                           (dict(symrate=72000, oqpsk=1, bps=8, order=20, interp=11), "simple")):
        raw = make_case("C2_oqpsk80k_u8_o32_L5" if cfg["bps"] == 8 else "C1_qpsk72k_s16_o32_L5", 20_000, seed=3)
        d = Demod(symrate=cfg["symrate"], oqpsk=cfg["oqpsk"], bps=cfg["bps"], rrc_order=cfg["order"],
                  interp_factor=cfg["interp"], kernel="auto")
        assert d.kernel_name() == served_by
        soft, counts, symf = d.process_batch(raw.reshape(1, -1), want_float=True)
        o = oracle_mod.Oracle(**cfg)
        w = o.process(raw)
        assert counts[0] == w.nsym and np.array_equal(bits(symf[0, :w.nsym]), bits(w.sym))
        assert_state_equal(d.state(), o)
        with pytest.raises(LrptError):
            Demod(symrate=cfg["symrate"], oqpsk=cfg["oqpsk"], bps=cfg["bps"], rrc_order=cfg["order"],
                  interp_factor=cfg["interp"], kernel="ws")


def test_longest_filters_pick_a_kernel_that_fits(oracle_mod, lib):
    """rrc_order 512 (1025 taps): the lane kernel's shared-memory delay lines fit for 8-bit input but not
    for float input; AUTO has to notice and stay exact either way."""
    from meteor_demod_b200 import Demod, synth
    for bps, served_by in ((8, "lane"), (32, "simple")):
        cfg = dict(symrate=72000, oqpsk=0, bps=bps, order=512, interp=4)
        raw = synth.make_raw(6000, bps=bps, seed=4)
        d = Demod(symrate=72000, oqpsk=0, bps=bps, rrc_order=512, interp_factor=4, kernel="auto")
        assert d.kernel_name() == served_by
        soft, n = d.process(raw)
        o = oracle_mod.Oracle(**cfg)
        w = o.process(raw)
        assert n is not None and soft.shape[0] == w.nsym and np.array_equal(soft, w.soft)
        assert_state_equal(d.state(), o)
        d.close()


def test_long_push_is_split_transparently(lib):
    """More samples than one launch addresses with int32 sub-step indices (2^26): the library cuts the push
    into several launches; the result must equal two half-size pushes (state + append cursor carried)."""
    import torch
    from meteor_demod_b200 import Demod, synth
    n = (1 << 26) + 300_000
    per = synth.baseband(230000, periodic=True, seed=3).astype(np.complex64)
    raw = synth.device_long_stream(per, n, cfo_hz=90.0).view(1, -1)
    d1 = Demod(nstreams=1)
    cap = (d1.capacity(n) + 7) // 8 * 8
    s1 = torch.zeros((1, 2 * cap), dtype=torch.int8, device="cuda")
    l0 = d1.launch_count()
    d1.process_device(raw, s1)
    d1.sync()
    assert d1.launch_count() - l0 == 2
    n1 = int(d1.counts()[0])
    d2 = Demod(nstreams=1)
    half = (n // 2) // 8 * 8
    sa = torch.zeros((1, 2 * cap), dtype=torch.int8, device="cuda")
    sb = torch.zeros((1, 2 * cap), dtype=torch.int8, device="cuda")
    d2.process_device(raw[:, : 2 * half].contiguous(), sa)
    d2.sync()
    na = int(d2.counts()[0])
    d2.process_device(raw[:, 2 * half:].contiguous(), sb)
    d2.sync()
    nb = int(d2.counts()[0])
    assert n1 == na + nb
    assert torch.equal(s1[0, : 2 * na], sa[0, : 2 * na]) and torch.equal(s1[0, 2 * na: 2 * n1], sb[0, : 2 * nb])
    assert d1.export_state() == d2.export_state()


@pytest.mark.parametrize("name", ["C1_qpsk72k_s16_o32_L5", "C2_oqpsk80k_u8_o32_L5"])
def test_lane_kernel_many_warps_per_sm(name, oracle_mod, lib):
    """10 000 streams = 313 warps over 148 SMs (3 warps per CTA, idle lanes in the last warp): streams
    sampled across warps, CTAs and the tail are compared with the oracle, all others with their twins
    (stream s and s + 5000 carry the same input)."""
    import torch
    cfg = CONFIGS[name]
    half, n = 5000, 6000
    base = [make_case(name, n + 1024, seed=200 + k, cfo_hz=-300.0 + 40 * k) for k in range(16)]
    rows = np.stack([base[s % 16][2 * ((s * 37) % 997): 2 * ((s * 37) % 997) + 2 * n] for s in range(half)])
    raw = np.concatenate([rows, rows])
    d = demod_for(cfg, "lane", nstreams=2 * half)
    cap16 = (d.capacity(n) + 7) // 8 * 8
    t_raw = torch.from_numpy(raw).cuda()
    t_soft = torch.zeros((2 * half, 2 * cap16), dtype=torch.int8, device="cuda")
    d.process_device(t_raw, t_soft)
    d.sync()
    counts = d.counts().astype(np.int64)
    soft = t_soft.cpu().numpy().reshape(2 * half, cap16, 2)
    assert np.array_equal(counts[:half], counts[half:])
    for s in range(half):
        assert np.array_equal(soft[s, :counts[s]], soft[s + half, :counts[s]]), s
    for s in [0, 1, 31, 32, 33, 95, 96, 4735, 4736, 4999, 5000, 7777, 9967, 9968, 9999]:
        o = oracle_mod.Oracle(**cfg)
        w = o.process(raw[s])
        assert counts[s] == w.nsym, s
        assert np.array_equal(soft[s, :w.nsym], w.soft), s
        assert_state_equal(d.state(s), o)


@pytest.mark.parametrize("wf", ["0", "1", "2"])
@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_lane_window_formats(name, wf, oracle_mod, lib, monkeypatch):
    """The lane kernel's delay-line window in the input's own type (LRPT_LANE_WF=0), as float pairs (=1) and,
    for 8-bit input, as bfloat16 pairs (=2) -- the launch picks the widest that costs no warp: all bit-exact,
    ragged launches, 70 streams (3 warps, idle lanes in the last), float symbols and state included."""
    cfg = CONFIGS[name]
    if wf == "2" and cfg["bps"] != 8:
        pytest.skip("bfloat16 pairs hold 8-bit samples only")
    monkeypatch.setenv("LRPT_LANE_WF", wf)
    n = 40_000
    raw = np.stack([make_case(name, n, seed=400 + s, cfo_hz=-600.0 + 17 * s) for s in range(7)])
    raw = np.concatenate([raw] * 10)
    d = demod_for(cfg, "lane", nstreams=70)
    parts, total = [], np.zeros(70, np.int64)
    for a, b in ((0, 17), (17, 18), (18, 30_001), (30_001, n)):
        soft, counts, symf = d.process_batch(np.ascontiguousarray(raw[:, 2 * a: 2 * b]), want_float=True)
        parts.append((soft, counts.astype(np.int64), symf))
    for s in (0, 3, 31, 32, 63, 64, 69):
        o = oracle_mod.Oracle(**cfg)
        w = o.process(raw[s])
        got = np.concatenate([p[0][s, :p[1][s]] for p in parts])
        gotf = np.concatenate([p[2][s, :p[1][s]] for p in parts])
        assert got.shape[0] == w.nsym, s
        assert np.array_equal(got, w.soft), s
        assert np.array_equal(bits(gotf), bits(w.sym)), s
        assert_state_equal(d.state(s), o)


def test_full_size_batch_twins_and_samples(oracle_mod, lib):
    """BASELINE-size batch (75776 streams = 16 warps on every SM): a size-independent property --
    streams with identical input give identical output, checked over the WHOLE batch -- plus sampled
    streams against the oracle."""
    import torch
    name = "C1_qpsk72k_s16_o32_L5"
    cfg = CONFIGS[name]
    half, n = 37888, 4096
    base = [make_case(name, n + 2048, seed=300 + k, cfo_hz=-500.0 + 60 * k) for k in range(16)]
    b = torch.from_numpy(np.stack(base)).cuda()                                   # [16, 2*(n+2048)]
    s = torch.arange(half, device="cuda")
    off = 2 * ((s * 53) % 1999)
    idx = off[:, None] + torch.arange(2 * n, device="cuda")[None, :]
    rows = b[(s % 16)[:, None], idx]
    t_raw = torch.cat([rows, rows]).contiguous()
    d = demod_for(cfg, "auto", nstreams=2 * half)
    assert d.kernel_name() == "lane"
    cap16 = (d.capacity(n) + 7) // 8 * 8
    t_soft = torch.zeros((2 * half, 2 * cap16), dtype=torch.int8, device="cuda")
    d.process_device(t_raw, t_soft)
    d.sync()
    counts = torch.from_numpy(d.counts().astype(np.int64)).cuda()
    assert torch.equal(counts[:half], counts[half:])
    valid = torch.arange(cap16, device="cuda")[None, :] < counts[:half, None]
    a3, b3 = t_soft[:half].view(half, cap16, 2), t_soft[half:].view(half, cap16, 2)
    assert bool(((a3 == b3).all(dim=2) | ~valid).all())
    assert int(counts.min()) > 0.3 * n and int(counts.max()) < 0.33 * n
    for i in [0, 511, 512, 20000, 37887, 37888, 75775]:
        w = oracle_mod.Oracle(**cfg).process(t_raw[i].cpu().numpy())
        got = t_soft[i, : 2 * w.nsym].cpu().numpy().reshape(-1, 2)
        assert int(counts[i]) == w.nsym and np.array_equal(got, w.soft), i


@pytest.mark.parametrize("kernel", ["lane", "simple"])
def test_random_configurations(kernel, oracle_mod, lib):
    """Seeded fuzz over the parameter space (order 0..140, oversampling 1..8, all three input types, both
    modes, odd symbol rates, ragged two-part pushes, stream counts that leave idle lanes): soft symbols
    and the complete state against the oracle."""
    from meteor_demod_b200 import Demod, synth
    rng = np.random.default_rng(20261017)
    for case in range(28):
        order = int(rng.choice([0, 1, 2, 3, 7, 8, 15, 16, 31, 33, 40, 63, 64, 100, 140]))
        interp = int(rng.integers(1, 9))
        bps = int(rng.choice([8, 16, 32]))
        oq = int(rng.integers(0, 2))
        symrate = int(rng.choice([72000, 80000, 57500, 91000, 33000]))
        ns = int(rng.choice([1, 3, 33, 65]))
        n = int(rng.integers(40, 7000))
        cut = int(rng.integers(0, n + 1))
        cfg = dict(symrate=symrate, oqpsk=oq, bps=bps, order=order, interp=interp)
        raw = np.stack([synth.make_raw(n, symrate=symrate, oqpsk=bool(oq), bps=bps, seed=1000 + 7 * case + s,
                                       cfo_hz=float(rng.integers(-900, 900))) for s in range(ns)])
        d = Demod(symrate=symrate, oqpsk=oq, bps=bps, rrc_order=order, interp_factor=interp, nstreams=ns, kernel=kernel)
        s1, c1 = d.process_batch(np.ascontiguousarray(raw[:, : 2 * cut]))
        s2, c2 = d.process_batch(np.ascontiguousarray(raw[:, 2 * cut:]))
        for s in range(ns):
            o = oracle_mod.Oracle(**cfg)
            w = o.process(raw[s])
            got = np.concatenate([s1[s, : c1[s]], s2[s, : c2[s]]])
            assert got.shape[0] == w.nsym, (case, cfg, ns, n, cut, s)
            assert np.array_equal(got, w.soft), (case, cfg, ns, n, cut, s)
            assert_state_equal(d.state(s), o)
        d.close()


C5_GRID = [(oq, order, L) for oq in (0, 1) for order in (16, 32, 64, 128) for L in (3, 5, 8)]


@pytest.mark.parametrize("oq,order,L", C5_GRID, ids=["%s_o%d_L%d" % ("oqpsk80k_u8" if g[0] else "qpsk72k_s16", g[1], g[2]) for g in C5_GRID])
def test_config5_whole_grid(oq, order, L, oracle_mod, lib):
    """BASELINE config 5, every corner: RRC order {16,32,64,128} x oversampling {3,5,8} x {QPSK 72k s16, OQPSK 80k u8}.
    33 streams through the lane kernel (stream 31 = last lane of a warp, stream 32 alone in the next warp) in two
    ragged pushes, streams 0 / 31 / 32 bit for bit against the oracle (symbols, float symbols, complete state), and
    stream 0 again through whatever kernel AUTO picks for a single stream."""
    from meteor_demod_b200 import Demod, synth
    symrate, bps = (80000, 8) if oq else (72000, 16)
    cfg = dict(symrate=symrate, oqpsk=oq, bps=bps, order=order, interp=L)
    n, cut = 36_000, 12_345
    sig = [synth.make_raw(n, symrate=symrate, oqpsk=bool(oq), bps=bps, seed=700 + 3 * order + L + k, cfo_hz=-350.0 + 300.0 * k)
           for k in range(3)]
    raw = np.stack([sig[s % 3] if s < 31 else sig[s - 30] for s in range(33)])       # 31 -> sig[1], 32 -> sig[2]
    d = demod_for(cfg, "lane", nstreams=33)
    s1, c1, f1 = d.process_batch(np.ascontiguousarray(raw[:, : 2 * cut]), want_float=True)
    s2, c2, f2 = d.process_batch(np.ascontiguousarray(raw[:, 2 * cut:]), want_float=True)
    want = {}
    for s in (0, 31, 32):
        o = oracle_mod.Oracle(**cfg)
        w = want[s] = o.process(raw[s])
        got = np.concatenate([s1[s, : c1[s]], s2[s, : c2[s]]])
        gotf = np.concatenate([f1[s, : c1[s]], f2[s, : c2[s]]])
        assert got.shape[0] == w.nsym, s
        assert np.array_equal(got, w.soft), s
        assert np.array_equal(bits(gotf), bits(w.sym)), s
        assert_state_equal(d.state(s), o)
    d.close()
    d1 = Demod(symrate=symrate, oqpsk=oq, bps=bps, rrc_order=order, interp_factor=L, kernel="auto")
    assert d1.kernel_name() in ("ws", "spec")
    soft, counts, symf = d1.process_batch(raw[:1], want_float=True)
    w = want[0]
    assert counts[0] == w.nsym and np.array_equal(soft[0, : w.nsym], w.soft) and np.array_equal(bits(symf[0, : w.nsym]), bits(w.sym))
    d1.close()


def test_handle_stream_is_ordered_after_torch_stream(lib):
    """The handle's own stream is non-blocking; process_device orders it after torch's current stream, so buffers
    torch is still filling (a device copy of the input, the clear of the output) are complete when the kernel runs."""
    import torch
    from meteor_demod_b200 import Demod, synth
    ns, n = 2048, 1 << 16
    per = synth.baseband(23000, periodic=True, seed=5).astype(np.complex64)
    src = synth.device_streams(per, ns, n, bps=16, seed=11)
    want = torch.empty((ns, 2 * 24000), dtype=torch.int8, device="cuda")
    with Demod(nstreams=ns) as d:
        d.process_device(src, want)
        d.sync()
        cw = d.counts().copy()
    torch.cuda.synchronize()
    with Demod(nstreams=ns) as d:
        junk = torch.full_like(src, 17)
        for _ in range(3):
            raw = junk.clone()                               # queue some work in front ...
        raw = src.clone()                                    # ... of the copy the kernel must wait for
        got = torch.full((ns, 2 * 24000), 99, dtype=torch.int8, device="cuda")
        got.zero_()
        want_clear = torch.zeros_like(want)
        d.process_device(raw, got)
        d.sync()
        cg = d.counts().copy()
        # ... and the reverse: a torch stream that reads the symbols without a host synchronisation in between
        lib_ = d.lib
        side = torch.cuda.Stream()
        d.reset()
        d.process_device(raw, got)
        assert lib_.lrpt_stream_release(d.h, side.cuda_stream) == 0
        with torch.cuda.stream(side):
            copy = got.clone()
        side.synchronize()
        d.sync()
        assert torch.equal(copy, got)
    assert np.array_equal(cw, cg)
    m = torch.arange(24000, device="cuda")[None, :] < torch.as_tensor(cw.astype(np.int64), device="cuda")[:, None]
    m2 = m.repeat_interleave(2, dim=1)
    assert torch.equal(want[m2], got[m2]) and int((got[~m2] != 0).sum()) == 0 and want_clear.sum() == 0
