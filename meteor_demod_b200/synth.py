"""Synthetic LRPT-like I/Q for tests and bench (not part of the demodulator path).

Recipe (validated there against the reference: it locks at the
programmed carrier offset and symbol rate): uniform random QPSK symbols, RRC alpha=0.6
pulse spanning +-8 symbols at 16x the symbol rate, OQPSK = Q delayed by half a symbol,
rational resampling to fs (x115/576 for 72 ksym/s, x23/128 for 80 ksym/s at fs=230 kS/s),
unit rms, carrier offset + phase, complex AWGN at the requested Es/N0, then the raw
formats wavfile.c:58-69 ingests (u8 offset-128, s16, f32 with s16-scaled values).
"""
import struct
from fractions import Fraction

import numpy as np

UP = 16  # samples per symbol of the intermediate signal


def rrc_pulse(alpha=0.6, span=8, sps=UP):
    t = np.arange(-span * sps, span * sps + 1, dtype=np.float64) / sps
    h = np.empty_like(t)
    for i, x in enumerate(t):
        if abs(x) < 1e-12:
            h[i] = 1 - alpha + 4 * alpha / np.pi
        elif abs(abs(4 * alpha * x) - 1) < 1e-9:
            h[i] = alpha / np.sqrt(2) * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * alpha))
                                         + (1 - 2 / np.pi) * np.cos(np.pi / (4 * alpha)))
        else:
            h[i] = (np.sin(np.pi * x * (1 - alpha)) + 4 * alpha * x * np.cos(np.pi * x * (1 + alpha))) / \
                   (np.pi * x * (1 - (4 * alpha * x) ** 2))
    return h / np.sqrt(np.sum(h ** 2))


def _ratio(fs, symrate):
    fr = Fraction(int(fs), int(symrate) * UP)
    return fr.numerator, fr.denominator


def baseband(nsamples, symrate=72000, fs=230000, oqpsk=False, seed=1, periodic=False):
    """Noise-free, carrier-free unit-rms complex baseband at fs, `nsamples` long.

    periodic=True builds a seamlessly tileable period (circular pulse shaping and
    resampling); nsamples*symrate/fs must then be an integer.
    """
    from scipy.signal import resample_poly

    up, down = _ratio(fs, symrate)
    rng = np.random.Generator(np.random.PCG64(seed))
    if periodic:
        nsym_f = Fraction(nsamples * int(symrate), int(fs))
        if nsym_f.denominator != 1:
            raise ValueError("periodic baseband needs nsamples*symrate/fs integral")
        nsym = int(nsym_f)
    else:
        nsym = int(np.ceil(nsamples * symrate / fs)) + 64
    bits = rng.integers(0, 2, size=(nsym, 2))
    sym = (2.0 * bits[:, 0] - 1) + 1j * (2.0 * bits[:, 1] - 1)
    x = np.zeros(nsym * UP, np.complex128)
    x[::UP] = sym
    h = rrc_pulse()
    if periodic:
        H = np.fft.fft(np.concatenate([h, np.zeros(x.size - h.size)]))
        y = np.fft.ifft(np.fft.fft(x) * H)
        y = np.roll(y, -(h.size // 2))
    else:
        y = np.convolve(x, h)[h.size // 2:][: x.size]
    if oqpsk:
        y = y.real + 1j * np.roll(y.imag, UP // 2)
    if periodic:
        pad = down * 64
        ypad = np.concatenate([y[-pad:], y, y[:pad]])
        z = resample_poly(ypad, up, down)
        off = pad * up // down
        z = z[off: off + nsamples]
    else:
        z = resample_poly(y, up, down)[32: 32 + nsamples]
    assert z.size == nsamples, (z.size, nsamples)
    return z / np.sqrt(np.mean(np.abs(z) ** 2))


def modulate(symbols, symrate=72000, fs=230000, cfo_hz=700.0, phase=0.7, esn0_db=12.0, bps=16, rms=6000.0, seed=1):
    """GIVEN complex QPSK symbols (+-1 +-1j, one per symbol period) -> raw I/Q as make_raw builds it: RRC pulse at 16x,
    rational resampling to fs, unit rms, carrier offset + phase, AWGN, the raw sample format. For tests that need to
    know what was sent (the decoder front-end behind the demodulator)."""
    from scipy.signal import resample_poly

    up, down = _ratio(fs, symrate)
    sym = np.asarray(symbols, np.complex128)
    x = np.zeros(sym.size * UP, np.complex128)
    x[::UP] = sym
    h = rrc_pulse()
    y = np.convolve(x, h)[h.size // 2:][: x.size]
    z = resample_poly(y, up, down)
    z = z / np.sqrt(np.mean(np.abs(z) ** 2))
    return to_raw(impair(z, fs, cfo_hz, phase, esn0_db, sps=fs / symrate, seed=seed + 1000), bps, rms)


def impair(z, fs=230000, cfo_hz=700.0, phase=0.7, esn0_db=12.0, sps=None, seed=2, n0=0):
    """Carrier offset/phase + AWGN. Es/N0 refers to symbol energy: noise var = sps/EsN0 per sample."""
    n = np.arange(n0, n0 + z.size, dtype=np.float64)
    y = z * np.exp(1j * (2 * np.pi * cfo_hz * n / fs + phase))
    if esn0_db is not None:
        rng = np.random.Generator(np.random.PCG64(seed))
        sps = sps if sps is not None else 230000 / 72000
        sigma = np.sqrt(sps / (10 ** (esn0_db / 10)) / 2)
        y = y + sigma * (rng.standard_normal(z.size) + 1j * rng.standard_normal(z.size))
    return y


def to_raw(y, bps=16, rms=6000.0, dc=(30.0, -20.0)):
    """Complex signal -> interleaved raw array in the format `bps` names (wavfile.c:58-69)."""
    if bps == 16:
        v = y * rms + (dc[0] + 1j * dc[1])
        out = np.empty(2 * y.size, np.int16)
        out[0::2] = np.clip(np.rint(v.real), -32768, 32767)
        out[1::2] = np.clip(np.rint(v.imag), -32768, 32767)
        return out
    if bps == 8:
        peak = np.max(np.abs(np.concatenate([y.real, y.imag])))
        v = y * (64.0 / peak)
        out = np.empty(2 * y.size, np.uint8)
        out[0::2] = np.clip(np.rint(v.real) + 128, 0, 255)
        out[1::2] = np.clip(np.rint(v.imag) + 128, 0, 255)
        return out
    if bps == 32:
        v = y * rms + (dc[0] + 1j * dc[1])
        out = np.empty(2 * y.size, np.float32)
        out[0::2] = np.rint(v.real)
        out[1::2] = np.rint(v.imag)
        return out
    raise ValueError("bps must be 8, 16 or 32")


def make_raw(nsamples, symrate=72000, fs=230000, oqpsk=False, bps=16, seed=1, cfo_hz=700.0,
             phase=0.7, esn0_db=12.0, rms=6000.0):
    z = baseband(nsamples, symrate, fs, oqpsk, seed)
    y = impair(z, fs, cfo_hz, phase, esn0_db, sps=fs / symrate, seed=seed + 1000)
    return to_raw(y, bps, rms)


def wav_header(data_bytes, fs=230000, bps=16, channels=2):
    """Canonical 44-byte RIFF/WAVE header, the only form wavfile.c:16-49 accepts."""
    fmt = 3 if bps == 32 else 1
    return struct.pack("<4sI4s4sIHHIIHH4sI", b"RIFF", 36 + data_bytes, b"WAVE", b"fmt ", 16, fmt, channels,
                       fs, fs * channels * bps // 8, channels * bps // 8, bps, b"data", data_bytes)


def device_streams(period, nstreams, nsamples, bps=16, fs=230000, sps=230000 / 72000, seed=7,
                   esn0_db=12.0, rms=6000.0, device="cuda", out=None, group=64, cfo_max_hz=1500.0, cfo_min_hz=None,
                   row_items=None):
    """Build `nstreams` distinct raw streams on the device from one tileable baseband period.

    Stream b = period rolled by a per-stream shift, tiled to nsamples, mixed with a per-stream
    carrier (multiple of fs/len(period) so tiling stays seamless; offset in [cfo_min_hz, cfo_max_hz],
    cfo_min_hz = -cfo_max_hz by default), own phase,
    amplitude and noise. Returns a torch tensor [nstreams, 2*nsamples] of the raw dtype (a view of rows
    `row_items` items long when given: rows whose byte length is not a multiple of 16 need a padded
    stride for the library's vector loads). Plumbing only (torch ops, `group` streams per batch of ops).
    With nsamples == len(period) a row is exactly one period of a periodic signal: replaying it again
    and again is ONE continuous stream (bench.py's locked pass).
    """
    import torch

    dt = {8: torch.uint8, 16: torch.int16, 32: torch.float32}[bps]
    P = int(period.size)
    base = torch.from_numpy(np.ascontiguousarray(period.astype(np.complex64))).to(device)
    if out is None:
        if row_items is None:
            out = torch.empty((nstreams, 2 * nsamples), dtype=dt, device=device)
        else:
            out = torch.zeros((nstreams, row_items), dtype=dt, device=device)[:, : 2 * nsamples]
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rs = np.random.Generator(np.random.PCG64(seed))
    step = fs / P
    kmax = int(cfo_max_hz / step)
    shift = torch.from_numpy(rs.integers(0, P, nstreams)).to(device)
    kmin = -kmax if cfo_min_hz is None else int(np.ceil(cfo_min_hz / step))
    cfo = torch.from_numpy(step * rs.integers(kmin, kmax + 1, nstreams).astype(np.float64)).to(device)
    ph = torch.from_numpy(rs.uniform(0, 2 * np.pi, nstreams)).to(device)
    amp = torch.from_numpy(rs.uniform(0.6, 1.2, nstreams).astype(np.float32)).to(device)
    n = torch.arange(nsamples, device=device)
    sigma = float(np.sqrt(sps / (10 ** (esn0_db / 10)) / 2)) if esn0_db is not None else 0.0
    dc = torch.tensor([30.0, -20.0], device=device)
    for b0 in range(0, nstreams, group):
        b1 = min(nstreams, b0 + group)
        idx = (n[None, :] + shift[b0:b1, None]) % P
        # carrier phase reduced mod 2*pi in float64 before going to float32
        arg = torch.remainder(2 * np.pi / fs * cfo[b0:b1, None] * n[None, :].to(torch.float64) + ph[b0:b1, None],
                              2 * np.pi).to(torch.float32)
        y = base[idx] * torch.polar(torch.ones_like(arg), arg)
        if sigma:
            y = y + sigma * torch.complex(torch.randn(y.shape, device=device, generator=g),
                                          torch.randn(y.shape, device=device, generator=g))
        v = torch.view_as_real(y)
        if bps == 8:
            v = (v * (64.0 / 3.0 * amp[b0:b1, None, None])).round() + 128
            out[b0:b1] = v.clamp(0, 255).reshape(b1 - b0, -1).to(dt)
        else:
            v = ((v * (rms * amp[b0:b1, None, None])) + dc).round().clamp(-32768, 32767)
            out[b0:b1] = v.reshape(b1 - b0, -1).to(dt)
    return out


def device_long_stream(period, nsamples, total=None, bps=16, fs=230000, sps=230000 / 72000, seed=11, cfo_hz=700.0,
                       phase=0.7, esn0_db=12.0, rms=6000.0, device="cuda", block=1 << 25, first=0):
    """ONE long raw stream on the device (time-sharding workload): the tileable period repeated, a carrier
    offset (rounded to a multiple of fs/len(period)), noise, quantisation. Returns a 1-D tensor with
    2*total items holding samples [first, first + total) of the stream (default: all of it, total >=
    nsamples); samples at or beyond nsamples are zero padding. The noise is seeded per `block` of absolute
    sample indices, so any window of the stream comes out the same whichever rank generates it."""
    import torch

    dt = {8: torch.uint8, 16: torch.int16, 32: torch.float32}[bps]
    total = nsamples - first if total is None else total
    P = int(period.size)
    base = torch.from_numpy(np.ascontiguousarray(period.astype(np.complex64))).to(device)
    out = torch.zeros(2 * total, dtype=dt, device=device)
    g = torch.Generator(device=device)
    step = fs / P
    cfo = step * round(cfo_hz / step)
    sigma = float(np.sqrt(sps / (10 ** (esn0_db / 10)) / 2)) if esn0_db is not None else 0.0
    dc = torch.tensor([30.0, -20.0], device=device)
    end = min(nsamples, first + total)
    for b in range(first // block, (max(end, first + 1) - 1) // block + 1):
        n0, n1 = max(first, b * block), min(end, (b + 1) * block)
        if n1 <= n0:
            continue
        n = torch.arange(n0, n1, device=device)
        arg = torch.remainder(2 * np.pi / fs * cfo * n.to(torch.float64) + phase, 2 * np.pi).to(torch.float32)
        y = base[n % P] * torch.polar(torch.ones_like(arg), arg)
        if sigma:
            g.manual_seed(seed * 1000003 + b)                # the whole block's noise, then this window's part of it
            nz = torch.complex(torch.randn(block, device=device, generator=g), torch.randn(block, device=device, generator=g))
            y = y + sigma * nz[n0 - b * block: n1 - b * block]
            del nz
        v = torch.view_as_real(y)
        if bps == 8:
            v = ((v * (64.0 / 3.0)).round() + 128).clamp(0, 255)
        else:
            v = ((v * rms) + dc).round().clamp(-32768, 32767)
        out[2 * (n0 - first): 2 * (n1 - first)] = v.reshape(-1).to(dt)
    return out
