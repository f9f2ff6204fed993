"""Host-side mirror of the reference's demodulator interface, over the C ABI.

The reference exposes demod_init / demod_qpsk / demod_oqpsk (demod.h:29-50) and the
getters pll_get_freq, pll_get_locked, pll_did_lock_once (pll.h:20-34), mm_omega
(timing.h:32), agc_get_gain (agc.h:18). `Demod` keeps those names and argument
meanings; the per-sample push becomes a block push because the work happens on
the GPU. All computation is in liblrpt_b200.so; nothing here touches samples.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import LrptError, Params, State, Status

RAW_DTYPES = {8: np.uint8, 16: np.int16, 32: np.float32}

# reference defaults, demod.h:8-15
RRC_ORDER, INTERP_FACTOR, SYM_BW, PLL_BW, SYM_RATE = 32, 5, 0.00005, 1.0, 72000


def symbol_capacity(nsamples, samplerate, symrate):
    """Symbols `nsamples` inputs can produce in steady state, with slack for the
    +-2^-12 clock range (timing.c:7) and start-up transients."""
    return int(nsamples * (symrate / samplerate) * 1.02) + 64


def make_params(samplerate=230000, symrate=SYM_RATE, interp_factor=INTERP_FACTOR, rrc_order=RRC_ORDER,
                oqpsk=False, bps=16, pll_bw=PLL_BW, sym_bw=SYM_BW, freq_max=-1.0, nstreams=1, device=0,
                kernel="auto"):
    return Params(pll_bw=pll_bw, sym_bw=sym_bw, freq_max=freq_max, samplerate=int(samplerate),
                  symrate=int(symrate), interp_factor=int(interp_factor), rrc_order=int(rrc_order),
                  oqpsk=int(bool(oqpsk)), bps=int(bps), device=int(device), nstreams=int(nstreams),
                  kernel=_lib.KERNELS[kernel])


def describe(**kw):
    """Host-only: taps, power-on state, loop constants, tanh table (needs no GPU)."""
    lib = _lib.load()
    p = make_params(**kw)
    n = max((2 * p.rrc_order + 1) * p.interp_factor, 1)
    taps = np.empty(n, np.float32)
    consts = np.empty(7, np.float32)
    lut = np.empty(32, np.float32)
    s0 = State()
    rc = lib.lrpt_describe(C.byref(p), C.byref(s0), taps.ctypes.data, n, consts.ctypes.data, lut.ctypes.data)
    if rc < 0:
        raise LrptError(rc)
    names = ("t_center", "t_maxdev", "t_alpha", "t_beta", "p_alpha", "p_beta", "p_fmax")
    return dict(taps=taps, state=state_to_dict(s0), consts=dict(zip(names, consts.tolist())), lut=lut)


def state_to_dict(s):
    return {k: getattr(s, k) for k, _ in State._fields_}


class Demod:
    """`nstreams` independent demodulators on one GPU (demod_init, demod.h:29)."""

    def __init__(self, samplerate=230000, symrate=SYM_RATE, interp_factor=INTERP_FACTOR, rrc_order=RRC_ORDER,
                 oqpsk=False, bps=16, pll_bw=PLL_BW, sym_bw=SYM_BW, freq_max=-1.0, nstreams=1, device=0,
                 kernel="auto"):
        self.lib = _lib.load()
        self.p = make_params(samplerate, symrate, interp_factor, rrc_order, oqpsk, bps, pll_bw, sym_bw,
                             freq_max, nstreams, device, kernel)
        self.h = C.c_void_p()
        rc = self.lib.lrpt_create(C.byref(self.h), C.byref(self.p))
        if rc:
            self.h = None
            raise LrptError(rc, "lrpt_create")
        self.nstreams = int(nstreams)
        self.bps = int(bps)

    # -- lifecycle ---------------------------------------------------------
    def close(self):                                   # demod_deinit, demod.h:34
        if getattr(self, "h", None):
            self.lib.lrpt_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what, allow=()):
        if rc and rc not in allow:
            raise LrptError(rc, what + ": " + self.lib.lrpt_last_error(self.h).decode())
        return rc

    def reset(self, stream=None, asynchronous=False):
        """Back to power-on state. asynchronous=True enqueues on `stream` (torch.cuda.Stream or None =
        the handle's stream) without synchronising the host."""
        if asynchronous or stream is not None:
            st = C.c_void_p(stream.cuda_stream) if stream is not None else None
            self._check(self.lib.lrpt_reset_async(self.h, st), "reset_async")
        else:
            self._check(self.lib.lrpt_reset(self.h), "reset")

    def capacity(self, nsamples):
        return symbol_capacity(nsamples, self.p.samplerate, self.p.symrate)

    # -- hot path ----------------------------------------------------------
    def _raw(self, raw):
        a = np.ascontiguousarray(raw)
        if a.dtype != RAW_DTYPES[self.bps]:
            raise TypeError("raw IQ dtype %s does not match bps=%d" % (a.dtype, self.bps))
        return a

    def process(self, raw, cap=None):
        """Single stream (stream 0), host buffers. Returns (soft[nsym,2] int8, first_lock_symbol)."""
        a = self._raw(raw).reshape(-1)
        n = a.size // 2
        cap = self.capacity(n) if cap is None else cap
        soft = np.empty((cap, 2), np.int8)
        nsym = C.c_size_t(0)
        first = C.c_longlong(-1)
        rc = self._check(self.lib.lrpt_process(self.h, a.ctypes.data, n, soft.ctypes.data, cap,
                                               C.byref(nsym), C.byref(first)), "process", allow=(_lib.LRPT_ERR_CAP,))
        self.last_rc = rc
        return soft[:min(nsym.value, cap)], first.value

    def process_batch(self, raw, cap=None, want_float=False):
        """All streams, host buffers. raw: [nstreams, 2*nsamples]. Returns (soft[nstreams,cap,2], counts[, symf])."""
        a = self._raw(raw)
        if a.ndim != 2 or a.shape[0] != self.nstreams:
            raise ValueError("raw must be [nstreams, 2*nsamples]")
        n = a.shape[1] // 2
        cap = self.capacity(n) if cap is None else cap
        soft = np.zeros((self.nstreams, cap, 2), np.int8)
        counts = np.zeros(self.nstreams, np.uint32)
        symf = np.zeros((self.nstreams, cap, 2), np.float32) if want_float else None
        rc = self._check(self.lib.lrpt_process_batch(
            self.h, a.ctypes.data, a.strides[0], n, soft.ctypes.data, soft.strides[0], cap, counts.ctypes.data,
            symf.ctypes.data if want_float else None, symf.strides[0] if want_float else 0),
            "process_batch", allow=(_lib.LRPT_ERR_CAP,))
        self.last_rc = rc
        return (soft, counts, symf) if want_float else (soft, counts)

    def _after_torch(self, tensor, stream):
        """The handle's own stream is non-blocking: order it after whatever torch has enqueued on ITS current
        stream for `tensor`'s device (the copy / fill that produced the buffers), without a host sync."""
        if stream is None and getattr(tensor, "is_cuda", False):
            import torch
            cur = torch.cuda.current_stream(tensor.device).cuda_stream
            self._check(self.lib.lrpt_stream_wait(self.h, C.c_void_p(cur) if cur else None), "stream_wait")

    def process_device(self, raw, soft, nsym=None, symf=None, stream=None, nsamples=None):
        """All streams, device buffers (torch tensors). Asynchronous on `stream` (torch.cuda.Stream
        or None = the handle's own stream, ordered after torch's current stream). raw: [nstreams, 2*nsamples] of
        the raw dtype; soft: int8 [nstreams, 2*cap]; nsym: optional uint32/int32 [nstreams]; symf: optional float32
        [nstreams, 2*cap]. Reading the results with torch ops needs sync() (or the same `stream`) first."""
        self._after_torch(raw, stream)
        if raw.dim() != 2 or raw.shape[0] != self.nstreams or soft.shape[0] != self.nstreams:
            raise ValueError("raw/soft must be [nstreams, ...]")
        n = raw.shape[1] // 2 if nsamples is None else int(nsamples)
        cap = soft.shape[1] // 2
        if symf is not None:
            cap = min(cap, symf.shape[1] // 2)
        st = C.c_void_p(stream.cuda_stream) if stream is not None else None
        self._check(self.lib.lrpt_process_batch_device(
            self.h, raw.data_ptr(), raw.stride(0) * raw.element_size(), n,
            soft.data_ptr(), soft.stride(0) * soft.element_size(), cap,
            nsym.data_ptr() if nsym is not None else None,
            symf.data_ptr() if symf is not None else None,
            symf.stride(0) * symf.element_size() if symf is not None else 0, st), "process_batch_device")

    def set_symbol_index_output(self, idx=None):
        """idx: uint32/int32 device tensor [nstreams, cap] receiving, per symbol, sample*interp + sub-step
        (None = off). Applies to subsequent process_device calls."""
        if idx is None:
            self._check(self.lib.lrpt_set_symbol_index_output(self.h, None, 0), "set_symbol_index_output")
        else:
            self._check(self.lib.lrpt_set_symbol_index_output(self.h, idx.data_ptr(), idx.stride(0) * idx.element_size()),
                        "set_symbol_index_output")
        self._symq = idx

    def sync(self, stream=None):
        st = C.c_void_p(stream.cuda_stream) if stream is not None else None
        self._check(self.lib.lrpt_sync(self.h, st), "sync")

    def counts(self):
        c = np.zeros(self.nstreams, np.uint32)
        self._check(self.lib.lrpt_get_counts(self.h, c.ctypes.data, self.nstreams), "get_counts")
        return c

    # -- status: the reference's getters -------------------------------------
    def status(self, stream=0):
        st = Status()
        self._check(self.lib.lrpt_status(self.h, stream, C.byref(st)), "status")
        return {k: getattr(st, k) for k, _ in Status._fields_}

    def pll_get_freq(self, stream=0):          # pll.h:20
        return self.status(stream)["pll_freq"]

    def pll_get_locked(self, stream=0):        # pll.h:27
        return self.status(stream)["locked"]

    def pll_did_lock_once(self, stream=0):     # pll.h:34
        return self.status(stream)["locked_once"]

    def mm_omega(self, stream=0):              # timing.h:32
        return self.status(stream)["mm_omega"]

    def agc_get_gain(self, stream=0):          # agc.h:18
        return self.status(stream)["agc_gain"]

    def carrier_hz(self, stream=0):            # main.c:250
        return self.pll_get_freq(stream) * self.p.symrate / (2 * np.pi) * (2 if self.p.oqpsk else 1)

    def symbol_rate_hz(self, stream=0):        # main.c:251
        return self.mm_omega(stream) * (self.p.samplerate * self.p.interp_factor) / (2 * np.pi)

    # -- state hand-off ------------------------------------------------------
    def state_size(self):
        return self.lib.lrpt_state_size(self.h)

    def export_state(self, stream=0):
        n = C.c_size_t(self.state_size())
        buf = C.create_string_buffer(n.value)
        self._check(self.lib.lrpt_export_state(self.h, stream, buf, C.byref(n)), "export_state")
        return buf.raw[:n.value]

    def import_state(self, blob, stream=0):
        self._check(self.lib.lrpt_import_state(self.h, stream, blob, len(blob)), "import_state")

    def state(self, stream=0):
        blob = self.export_state(stream)
        s = State.from_buffer_copy(blob[:C.sizeof(State)])
        d = state_to_dict(s)
        d["history"] = np.frombuffer(blob[C.sizeof(State):], np.float32).reshape(-1, 2).copy()
        return d

    def states_size(self):
        return self.lib.lrpt_states_size(self.h)

    def export_states_device(self, buf, stream=None):
        """All streams' states into a uint8 device tensor of states_size() bytes (async on `stream`)."""
        st = C.c_void_p(stream.cuda_stream) if stream is not None else None
        self._check(self.lib.lrpt_export_states_device(self.h, buf.data_ptr(), buf.numel() * buf.element_size(), st),
                    "export_states_device")

    def import_states_device(self, buf, check=True, stream=None):
        self._after_torch(buf, stream)
        st = C.c_void_p(stream.cuda_stream) if stream is not None else None
        self._check(self.lib.lrpt_import_states_device(self.h, buf.data_ptr(), buf.numel() * buf.element_size(),
                                                       int(check), st), "import_states_device")

    def snapshot(self):
        self._check(self.lib.lrpt_snapshot(self.h), "snapshot")

    def restore(self, quarter_turns=None):
        """Back to the snapshot; quarter_turns (int per stream) turns each Costas NCO back by k*pi/2."""
        if quarter_turns is None:
            self._check(self.lib.lrpt_restore(self.h, None), "restore")
        else:
            t = np.ascontiguousarray(quarter_turns, np.int32)
            assert t.size == self.nstreams
            self._check(self.lib.lrpt_restore(self.h, t.ctypes.data), "restore")

    # -- introspection -------------------------------------------------------
    def taps(self):
        n = self.lib.lrpt_get_taps(self.h, None, 0)
        a = np.empty(n, np.float32)
        self.lib.lrpt_get_taps(self.h, a.ctypes.data, n)
        return a

    def launch_count(self):
        return int(self.lib.lrpt_launch_count(self.h))

    def fir_fallbacks(self):
        return int(self.lib.lrpt_fir_fallbacks(self.h))

    def kernel_name(self):
        return self.lib.lrpt_kernel_name(self.h).decode()
