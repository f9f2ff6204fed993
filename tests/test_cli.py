"""The C host (host/lrpt_demod, reference-compatible command line over the C ABI) against the
reference CLI itself (oracle/_ref/meteor_demod_ref_strict, when it was built) and against the
oracle + the egress rules. Byte-exact output files."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

HOST = os.path.join(ROOT, "host", "lrpt_demod")
REF = os.path.join(ROOT, "oracle", "_ref", "meteor_demod_ref_strict")


def run(cmd, stdin=None):
    return subprocess.run(cmd, input=stdin, capture_output=True, check=True)


@pytest.fixture(scope="module")
def host(lib):
    from meteor_demod_b200 import build
    build.build_host()
    assert os.path.exists(HOST)
    return HOST


def expected(raw_bytes, bps, oracle_mod, ref_tail, **cfg):
    from meteor_demod_b200 import egress
    dt = {8: np.uint8, 16: np.int16, 32: np.float32}[bps]
    n = egress.consumed_samples(len(raw_bytes), bps)
    raw = np.frombuffer(raw_bytes[: n * (bps // 4)], dt)
    w = oracle_mod.Oracle(bps=bps, **cfg).process(raw)
    fl = int(np.argmax(w.lock_once)) if w.lock_once.any() else -1
    return egress.gate(w.soft, fl, ref_compatible_tail=ref_tail), w.nsym


def test_wav_file_defaults(host, oracle_mod, tmp_path):
    """C1: 16-bit .wav, reference defaults (-B). The trailing partial 32 KiB block is not consumed."""
    from meteor_demod_b200 import synth
    raw = synth.make_raw(700_123, cfo_hz=55.0, seed=21)
    wav = tmp_path / "in.wav"
    wav.write_bytes(synth.wav_header(raw.nbytes) + raw.tobytes())
    out = tmp_path / "out.s"
    r = run([host, "-B", "-R", "1", "-o", str(out), str(wav)])
    assert b"Locked: Yes" in r.stdout and b"Carrier:" in r.stdout
    got = out.read_bytes()
    want, nsym = expected(raw.tobytes(), 16, oracle_mod, False)
    assert got == want and len(got) > 300_000
    out2 = tmp_path / "out2.s"
    run([host, "-B", "-q", "--ref-compatible-tail", "-o", str(out2), str(wav)])
    want2, _ = expected(raw.tobytes(), 16, oracle_mod, True)
    assert out2.read_bytes() == want2
    if os.path.exists(REF):
        ref_out = tmp_path / "ref.s"
        run([REF, "-B", "-q", "-o", str(ref_out), str(wav)])
        ref = ref_out.read_bytes()
        tail = (nsym % 512) * 2
        defined = len(want2) - tail + min(tail, 1024 - tail)      # beyond that the reference reads out of bounds
        assert len(ref) == len(want2)
        assert ref[:defined] == out2.read_bytes()[:defined]


def test_raw_stdin_oqpsk_u8(host, oracle_mod, tmp_path):
    """C2: OQPSK 80 ksym/s, 8-bit raw on stdin. A pipe cannot be rewound after the failed WAV probe,
    so the first 44 bytes are lost -- in the reference and here alike."""
    from meteor_demod_b200 import synth
    raw = synth.make_raw(500_000, symrate=80000, oqpsk=True, bps=8, cfo_hz=-80.0, seed=22).tobytes()
    args = ["-m", "oqpsk", "-r", "80000", "--bps", "8", "-s", "230000", "-B", "-q"]
    out = tmp_path / "o.s"
    run([host] + args + ["-o", str(out), "-"], stdin=raw)
    want, _ = expected(raw[44:], 8, oracle_mod, False, symrate=80000, oqpsk=1)
    assert out.read_bytes() == want
    if os.path.exists(REF):
        ref_out = tmp_path / "r.s"
        run([REF] + args + ["-o", str(ref_out), "-"], stdin=raw)
        n_valid = len(want) - (len(want) % 1024)
        assert ref_out.read_bytes()[:n_valid] == want[:n_valid]


def test_flags_order_oversamp_stdout(host, oracle_mod, tmp_path):
    """-f / -O / -b / -d and --stdout (C3 flags on a short float32 raw file)."""
    from meteor_demod_b200 import synth
    raw = synth.make_raw(300_000, bps=32, cfo_hz=20.0, seed=23)
    f = tmp_path / "in.raw"
    f.write_bytes(raw.tobytes())
    r = run([host, "-f", "64", "-O", "8", "-b", "2", "-d", "2000", "--bps", "32", "-s", "230000", "--stdout", str(f)])
    lib = oracle_mod.Oracle.lib()
    fmax = lib.lrpt_oracle_freq_delta(2000.0, 72000.0)
    want, _ = expected(raw.tobytes(), 32, oracle_mod, False, order=64, interp=8, pll_bw=2.0, freq_max=fmax)
    assert r.stdout == want


def test_usage_and_errors(host, tmp_path):
    r = subprocess.run([host], capture_output=True)
    assert r.returncode == 1 and b"Usage:" in r.stderr
    r = subprocess.run([host, "-B", str(tmp_path / "missing.wav")], capture_output=True)
    assert r.returncode == 1 and b"Could not open input file" in r.stderr
    raw = tmp_path / "x.raw"
    raw.write_bytes(b"\0" * 70000)
    r = subprocess.run([host, "-B", "-q", "-o", str(tmp_path / "o"), str(raw)], capture_output=True)
    assert r.returncode == 1 and b"Could not auto-detect sample rate" in r.stderr


def test_several_recordings_as_one_batch(host, oracle_mod, tmp_path):
    """Beyond the reference's command line: several inputs on one command line run as the streams of ONE batch
    (lrpt_process_batch). Every output file must be byte-identical to what a separate run writes -- the egress
    rules per file (whole 32 KiB blocks, lock gating, final flush) and different lengths included."""
    from meteor_demod_b200 import synth
    lens = (700_123, 65_536, 401_000, 0, 1_300_000)
    wavs, raws = [], []
    for i, n in enumerate(lens):
        raw = synth.make_raw(max(n, 1), cfo_hz=-120.0 + 90 * i, seed=40 + i)[: 2 * n]
        w = tmp_path / ("pass%d.wav" % i)
        w.write_bytes(synth.wav_header(raw.nbytes) + raw.tobytes())
        wavs.append(w)
        raws.append(raw)
    r = run([host, "-B"] + [str(w) for w in wavs])
    assert r.stdout.count(b"symbols") == len(lens)
    for w, raw in zip(wavs, raws):
        want, nsym = expected(raw.tobytes(), 16, oracle_mod, False)
        got = (tmp_path / (w.name + ".s")).read_bytes()
        assert got == want, w.name
    single = tmp_path / "single.s"
    run([host, "-B", "-q", "-o", str(single), str(wavs[4])])
    assert single.read_bytes() == (tmp_path / "pass4.wav.s").read_bytes()
    # one batch needs one sample format
    odd = tmp_path / "odd.wav"
    odd.write_bytes(synth.wav_header(8, bps=8) + b"\x80" * 8)
    r = subprocess.run([host, "-B", "-q", str(wavs[0]), str(odd)], capture_output=True)
    assert r.returncode == 1 and b"one batch needs one format" in r.stderr


def test_shard_flag_one_long_recording(host, oracle_mod, tmp_path):
    """--shard: the recording is cut into chunks that run side by side (lrpt_sharded_process). Output = the library
    call's symbols under the reference's egress rules; against the sequential reference run it has the same
    number of symbols, an exact head (chunks 0 and 1) and a small share of symbols off by more than one LSB."""
    from meteor_demod_b200 import egress, sharded, synth
    raw = synth.make_raw(1_700_123, cfo_hz=45.0, seed=77)
    wav = tmp_path / "long.wav"
    wav.write_bytes(synth.wav_header(raw.nbytes) + raw.tobytes())
    out = tmp_path / "long.s"
    r = run([host, "-B", "--shard", "400k", "-o", str(out), str(wav)])
    assert b"4 chunks" in r.stdout and b"Locked: Yes" in r.stdout
    n = egress.consumed_samples(raw.nbytes, 16)
    soft, rep = sharded.process_host(raw[: 2 * n], chunk=400_000, warm=150_000, overlap=8192, symrate=72000, bps=16)
    got = out.read_bytes()
    assert got == egress.gate(soft, rep["first_lock_symbol"])
    w = oracle_mod.Oracle(bps=16).process(raw[: 2 * n])
    assert soft.shape[0] == w.nsym and rep["first_lock_symbol"] == int(np.argmax(w.lock_once))
    head = int((150_000 + 2 * 400_000) * 72000 / 230000) - 16
    assert np.array_equal(soft[:head], w.soft[:head])
    d = np.abs(soft.astype(np.int16) - w.soft.astype(np.int16)).max(axis=1)
    assert (d > 1).mean() < 0.01
    # --seed: chunks after the first start at the coarse carrier estimate and warm up over 32 Ki samples
    out2 = tmp_path / "long_seeded.s"
    r = run([host, "-B", "--shard", "400k", "--seed", "4096", "-o", str(out2), str(wav)])
    assert b"Locked: Yes" in r.stdout
    soft2, rep2 = sharded.process_host(raw[: 2 * n], chunk=400_000, warm=32768, overlap=8192, seed_nfft=4096, symrate=72000, bps=16)
    assert out2.read_bytes() == egress.gate(soft2, rep2["first_lock_symbol"])
    d2 = np.abs(soft2.astype(np.int16) - w.soft.astype(np.int16)).max(axis=1)
    assert soft2.shape[0] == w.nsym and (d2 > 1).mean() < 0.01


def test_live_pipe_is_demodulated_as_it_arrives(host, oracle_mod, tmp_path):
    """A live source (rtl_sdr | lrpt_demod -): blocks are taken as they arrive instead of waiting for a full
    16 MiB slab. Bursts with pauses, odd burst sizes: the output must be what one uninterrupted run gives,
    and the first symbols must be on disk while the source is still open."""
    import time
    from meteor_demod_b200 import synth
    raw = synth.make_raw(600_000, symrate=80000, oqpsk=True, bps=8, cfo_hz=60.0, seed=31).tobytes()
    args = ["-m", "oqpsk", "-r", "80000", "--bps", "8", "-s", "230000", "-B", "-q"]
    out = tmp_path / "live.s"
    p = subprocess.Popen([host] + args + ["-o", str(out), "-"], stdin=subprocess.PIPE)
    cuts = [0, 100_000, 100_001, 500_000, 900_037, len(raw)]
    seen_early = False
    for a, b in zip(cuts[:-1], cuts[1:]):
        p.stdin.write(raw[a:b])
        p.stdin.flush()
        time.sleep(0.4)
        seen_early = seen_early or (out.exists() and out.stat().st_size > 0)
    for _ in range(50):                                 # a slow box: give the host up to 5 more seconds
        if seen_early:
            break
        time.sleep(0.1)
        seen_early = out.exists() and out.stat().st_size > 0
    assert seen_early                                   # symbols were written before the pipe was closed
    p.stdin.close()
    assert p.wait(timeout=60) == 0
    want, _ = expected(raw[44:], 8, oracle_mod, False, symrate=80000, oqpsk=1)
    assert out.read_bytes() == want


def test_wav_variants_are_walked_chunk_by_chunk(host, oracle_mod, tmp_path):
    """an 18-byte `fmt ` chunk and a `LIST` chunk before `data`. The reference takes any RIFF/WAVE
    file for canonical (wavfile.c:34-49) and would demodulate the rest of such a header as samples; the host walks
    the chunks and starts at the `data` payload: same bytes out as for the canonical file with the same samples."""
    import struct
    from meteor_demod_b200 import synth
    raw = synth.make_raw(400_000, cfo_hz=40.0, seed=23).tobytes()
    canon = tmp_path / "canon.wav"
    canon.write_bytes(synth.wav_header(len(raw)) + raw)
    fmt18 = struct.pack("<HHIIHHH", 1, 2, 230000, 230000 * 4, 4, 16, 0)
    listck = b"LIST" + struct.pack("<I", 26) + b"INFOISFT" + struct.pack("<I", 14) + b"lrpt test file"
    body = b"WAVE" + b"fmt " + struct.pack("<I", 18) + fmt18 + listck + b"data" + struct.pack("<I", len(raw)) + raw
    ext = tmp_path / "ext.wav"
    ext.write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)
    a, b = tmp_path / "a.s", tmp_path / "b.s"
    run([host, "-B", "-q", "-o", str(a), str(canon)])
    run([host, "-B", "-q", "-o", str(b), str(ext)])
    assert a.read_bytes() == b.read_bytes() and a.stat().st_size > 100_000
    # the same file through a pipe (no seeking in the chunk walk)
    c = tmp_path / "c.s"
    run([host, "-B", "-q", "-o", str(c), "-"], stdin=ext.read_bytes())
    assert c.read_bytes() == a.read_bytes()
    # a mono file is not an I/Q recording: treated like the reference treats it (raw, needs -s)
    mono = tmp_path / "mono.wav"
    mono.write_bytes(synth.wav_header(len(raw), channels=1) + raw)
    r = subprocess.run([host, "-B", "-q", "-o", str(tmp_path / "m.s"), str(mono)], capture_output=True)
    assert r.returncode != 0 and b"sample rate" in r.stderr


def test_status_panes(host, tmp_path):
    """the interactive panes of tui.c:139-247 (input position, bytes out, lock / gain / carrier / symbol
    rate, constellation of the last 512 symbols) from the per-block snapshots; --tui forces them onto a pipe (plain
    text frames). The symbols written are the ones -B writes."""
    from meteor_demod_b200 import synth
    raw = synth.make_raw(1_200_000, cfo_hz=30.0, seed=24)
    wav = tmp_path / "in.wav"
    wav.write_bytes(synth.wav_header(raw.nbytes) + raw.tobytes())
    a, b = tmp_path / "a.s", tmp_path / "b.s"
    run([host, "-B", "-q", "-o", str(a), str(wav)])
    r = run([host, "--tui", "-R", "0", "-o", str(b), str(wav)])
    assert a.read_bytes() == b.read_bytes()
    text = r.stdout.decode()
    assert "File in" in text and "Data out" in text and "PLL" in text and "Constellation" in text
    frame = text[text.rindex("+- File in"):]
    assert "Locked" in frame and "Carrier freq" in frame and "00:00:05" in frame          # 1.2 M samples at 230 kS/s
    plot = frame[frame.index("+- Constellation"):].splitlines()[1:22]
    assert len(plot) == 21 and all(len(ln) <= 45 for ln in plot)
    marks = sum(ln.count("#") + ln.count(".") for ln in plot)
    assert marks > 20                                            # four clusters of a locked QPSK constellation
    mid = plot[10]
    assert "-" in mid and "|" in plot[2]                         # axes through the middle (iq_draw_quadrants)
    # the clusters sit in the four quadrants, none on the axes' crossing
    quad = [sum(ln[2:23].count("#") for ln in plot[:10]), sum(ln[24:].count("#") for ln in plot[:10]),
            sum(ln[2:23].count("#") for ln in plot[11:]), sum(ln[24:].count("#") for ln in plot[11:])]
    assert all(q > 0 for q in quad), quad


def test_shard_reports_the_first_lock_of_a_late_signal(host, tmp_path):
    """A recording that starts before the signal is up (the normal case for a satellite pass): chunk 0 sees only
    noise and never locks, so --shard must take the stream-wide first lock from the later chunks instead of
    dropping every block (main.c:312 gates the output on pll_did_lock_once()). The instant itself is Tier-S: every
    chunk runs the reference's acquisition sweep (pll.c:126) from its own start, so it finds a carrier at +700 Hz
    sooner than the sequential run, whose sweep has passed that frequency while it was still listening to noise."""
    from meteor_demod_b200 import sharded, synth
    rng = np.random.default_rng(5)
    n_noise, n_sig = 450_000, 1_800_000
    noise = np.clip(np.rint(rng.normal(0, 1500, 2 * n_noise)), -32768, 32767).astype(np.int16)
    sig = synth.make_raw(n_sig, cfo_hz=700.0, seed=25)
    raw = np.concatenate([noise, sig])
    got, rep = sharded.process_host(raw, chunk=262144, warm=150000, overlap=8192)
    sym_per_sample = 72000 / 230000
    assert rep["nchunks"] >= 8
    assert rep["first_lock_symbol"] >= int(0.95 * n_noise * sym_per_sample), rep       # not on noise
    assert rep["first_lock_symbol"] <= int((n_noise + 700_000) * sym_per_sample), rep   # within two chunks of the signal
    wav = tmp_path / "late.wav"
    wav.write_bytes(synth.wav_header(raw.nbytes) + raw.tobytes())
    out = tmp_path / "late.s"
    run([host, "-B", "-q", "--shard", "262144", "-o", str(out), str(wav)])
    nbytes = out.stat().st_size
    nsym = int((raw.size // 2 // 8192 * 8192) * sym_per_sample)
    assert nbytes >= 2 * (nsym - rep["first_lock_symbol"] - 2048), (nbytes, nsym, rep)   # everything from the lock block on
    assert nbytes <= 2 * (nsym - int(0.9 * n_noise * sym_per_sample) + 2048)
