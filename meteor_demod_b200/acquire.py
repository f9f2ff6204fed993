"""Coarse carrier estimate for the warm-up of time-sharded chunks (an acquisition
accelerator, strictly opt-in because it is NOT what the reference does).

The reference finds the carrier by sweeping its Costas NCO at 1e-6 rad/symbol^2 until the lock detector fires
(pll.c:117-128); for OQPSK the detector fires early and the loop then pulls in the rest at its own pace --
measured on the oracle: 100 k samples at 333 Hz, 250 k at 700 Hz, 850 k at 1.2 kHz. A time-sharded chunk that
starts cold has to repeat that before its symbols are worth anything, which is what its warm-up is for. If the
chunk's Costas NCO starts AT the carrier instead (p_freq seeded, everything else power-on), 160 k samples of
warm-up are enough at any offset the reference accepts. Chunk 0 is never seeded: the head of the stream stays
the sequential run itself, with the reference's own acquisition time.

Estimator: the modulation is removed by a power law and the remaining spectral line located by an FFT --
  QPSK:   x^4 has a line at 4*f_c;
  OQPSK:  x^2 has two lines at 2*f_c -+ symrate (the I and Q pulse trains are half a symbol apart, so the
          cyclostationary parts of I^2 and Q^2 add up at the symbol rate instead of cancelling).
Bin spacing fs/nfft (1.75 Hz at 131072 points), refined by a three-point parabola around the peak; far inside the
loop's pull-in range even for short transforms. torch.fft on whatever
device the samples live on; plumbing, not the hot path.
"""
import math

import numpy as np
import torch


def to_complex(raw, bps):
    """[..., 2*n] interleaved raw I,Q (uint8 offset-128 / int16 / float32, wavfile.c:58-69) -> complex64 [..., n]."""
    t = raw if isinstance(raw, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(raw))
    v = t.to(torch.float32)
    if bps == 8:
        v = v - 128.0
    return torch.view_as_complex(v.contiguous().reshape(t.shape[:-1] + (t.shape[-1] // 2, 2)))


def estimate_cfo(x, fs, symrate, oqpsk, fmax=4000.0):
    """Carrier offset in Hz of every row of x (complex [M, nfft] or [nfft]), searched within +-fmax."""
    one = x.dim() == 1
    x = x.reshape(1, -1) if one else x
    n = x.shape[-1]
    x = x - x.mean(dim=-1, keepdim=True)                               # the DC term would survive the power law
    y = (x * x) if oqpsk else (x * x) * (x * x)
    win = torch.hann_window(n, periodic=False, device=x.device, dtype=torch.float32)
    S = torch.fft.fft(y * win, dim=-1).abs()
    df = fs / n
    order = 2 if oqpsk else 4
    kmax = int(order * fmax / df)
    ks = torch.arange(-kmax, kmax + 1, device=x.device)                # candidate bins of order*f_c
    if oqpsk:
        shift = int(round(symrate / df))
        score = S[:, (ks - shift) % n] + S[:, (ks + shift) % n]
    else:
        score = S[:, ks % n]
    j = score.argmax(dim=-1)
    # three-point parabola through the log magnitudes around the peak (exact for a Gaussian main lobe, within a few
    # percent of a bin for the Hann window): short transforms stay far inside the loop's pull-in range
    jc = j.clamp(1, score.shape[-1] - 2)
    rows = torch.arange(score.shape[0], device=x.device)
    la, lb, lc = (torch.log(score[rows, jc + d].to(torch.float64) + 1e-30) for d in (-1, 0, 1))
    den = la - 2.0 * lb + lc
    frac = torch.where(den.abs() > 1e-12, 0.5 * (la - lc) / den, torch.zeros_like(den)).clamp(-0.5, 0.5)
    frac = torch.where(jc == j, frac, torch.zeros_like(frac))
    best = (ks[j].to(torch.float64) + frac) * df / order
    return best[0] if one else best


def estimate_cfo_device(rows, params, nfft, fmax=4000.0):
    """The same estimate by the library's own kernel (csrc/acquire.cu, lrpt_carrier_estimate_device): rows is a CUDA
    tensor [M, >= 2*nfft] of the raw dtype (any row stride), params the lrpt_params_t of the stream; nfft a power of two
    in 256..16384. Returns float64 [M] on the device, asynchronous on torch's current stream."""
    import ctypes as C

    from . import _lib
    lib = _lib.load()
    out = torch.empty(rows.shape[0], dtype=torch.float64, device=rows.device)
    with torch.cuda.device(rows.device):
        rc = lib.lrpt_carrier_estimate_device(C.byref(params), rows.data_ptr(), rows.stride(0) * rows.element_size(), rows.shape[0],
                                              int(nfft), float(fmax), out.data_ptr(),
                                              C.c_void_p(torch.cuda.current_stream(rows.device).cuda_stream))
    if rc:
        raise _lib.LrptError(rc, "lrpt_carrier_estimate_device")
    return out


def p_freq_for(cfo_hz, symrate, oqpsk):
    """Costas NCO step (radians per loop update, float32) that tracks a carrier offset: the inverse of the status
    line's conversion freq_hz = pll_freq*symrate/(2*pi)*(oqpsk ? 2 : 1) (main.c:250, pll.c:38)."""
    return (2.0 * math.pi * cfo_hz / (symrate * (2 if oqpsk else 1))).to(torch.float32) if isinstance(cfo_hz, torch.Tensor) \
        else np.float32(2.0 * math.pi * cfo_hz / (symrate * (2 if oqpsk else 1)))
