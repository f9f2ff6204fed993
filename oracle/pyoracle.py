"""oracle/pyoracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes bindings for the two CPU checkers:

* ``Oracle``  -- our C restatement (oracle/lrpt_oracle.c, ``_build/liboracle.so``)
* ``Ref``     -- the UNMODIFIED reference sources compiled by oracle/Makefile into
                 ``_ref/libref_strict.so`` (strict IEEE, the bit-exact target) or
                 ``_ref/libref_fma.so`` (the reference's own release flags, FMA-contracted).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module. The product package never does.
"""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_CLI = os.path.join(REF_DIR, "meteor_demod_ref")


def build(verbose=False):
    """Compile the port and, when /root/reference (or a prebuilt _ref) exists, the reference."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", HERE, "port"], stdout=out)
    subprocess.check_call(["make", "-C", HERE, "ref"], stdout=out)


def have_ref(kind="strict"):
    return os.path.exists(os.path.join(REF_DIR, "libref_%s.so" % kind))


class _OracleStruct(C.Structure):
    _fields_ = [
        ("pll_bw", C.c_float), ("sym_bw", C.c_float), ("freq_max", C.c_float),
        ("samplerate", C.c_int), ("symrate", C.c_int), ("interp", C.c_int),
        ("order", C.c_int), ("oqpsk", C.c_int), ("bps", C.c_int),
        ("taps", C.c_int), ("h", C.POINTER(C.c_float)), ("lut_tanh", C.c_float * 32),
        ("t_center", C.c_float), ("t_maxdev", C.c_float), ("t_alpha", C.c_float), ("t_beta", C.c_float),
        ("p_alpha", C.c_float), ("p_beta", C.c_float), ("p_fmax", C.c_float), ("p_bw", C.c_float),
        ("t_phase", C.c_float), ("t_freq", C.c_float), ("t_prev", C.c_float),
        ("t_dual_state", C.c_int), ("oq_inphase", C.c_float),
        ("agc_gain", C.c_float), ("agc_bias_re", C.c_float), ("agc_bias_im", C.c_float),
        ("p_phase", C.c_float), ("p_freq", C.c_float), ("p_err", C.c_float),
        ("p_locked", C.c_int), ("p_locked_once", C.c_int), ("p_updown", C.c_int),
        ("hist", C.POINTER(C.c_float)),
        ("nsamples", C.c_longlong), ("nsymbols", C.c_longlong), ("first_lock_symbol", C.c_longlong),
        ("substep_out", C.c_void_p),
    ]


STATE_FIELDS = ("t_phase", "t_freq", "t_prev", "agc_gain", "agc_bias_re", "agc_bias_im",
                "p_phase", "p_freq", "p_err", "p_locked", "p_locked_once")


def _as_raw(raw, bps):
    dt = {8: np.uint8, 16: np.int16, 32: np.float32}[bps]
    a = np.ascontiguousarray(raw)
    if a.dtype != dt:
        raise TypeError("raw IQ must be %s for bps=%d, got %s" % (dt.__name__, bps, a.dtype))
    if a.size % 2:
        raise ValueError("raw IQ must hold interleaved I,Q pairs")
    return a


class _Result(dict):
    __getattr__ = dict.__getitem__


def _run(fn, raw, bps, cap=None, want_float=True):
    a = _as_raw(raw, bps)
    n = a.size // 2
    if cap is None:
        cap = n + 8          # symbols per sample is < 1 for every sane configuration
    sym = np.empty((cap, 2), np.float32) if want_float else None
    soft = np.empty((cap, 2), np.int8)
    idx = np.empty(cap, np.int64)
    lock = np.empty(cap, np.uint8)
    nsym = fn(a.ctypes.data_as(C.c_void_p), n,
              sym.ctypes.data_as(C.c_void_p) if want_float else None,
              soft.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p),
              lock.ctypes.data_as(C.c_void_p), cap)
    if nsym < 0:
        raise RuntimeError("oracle process failed")
    k = min(nsym, cap)
    return _Result(nsym=int(nsym), sym=None if sym is None else sym[:k], soft=soft[:k],
                   sample_idx=idx[:k], lock_once=lock[:k])


class Oracle:
    """Our C restatement. Re-entrant; keeps state between process() calls."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(PORT_SO):
                build()
            L = C.CDLL(PORT_SO)
            L.lrpt_oracle_init.argtypes = [C.POINTER(_OracleStruct), C.c_float, C.c_float, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.c_int, C.c_float, C.c_int]
            L.lrpt_oracle_init.restype = C.c_int
            L.lrpt_oracle_free.argtypes = [C.POINTER(_OracleStruct)]
            L.lrpt_oracle_process.argtypes = [C.POINTER(_OracleStruct), C.c_void_p, C.c_long, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
            L.lrpt_oracle_process.restype = C.c_long
            L.lrpt_oracle_rrc_coeff.argtypes = [C.c_int, C.c_uint, C.c_float, C.c_float]
            L.lrpt_oracle_rrc_coeff.restype = C.c_float
            L.lrpt_oracle_fir_all.argtypes = [C.POINTER(_OracleStruct), C.c_void_p, C.c_long, C.c_void_p]
            L.lrpt_oracle_fir_all.restype = C.c_int
            for f in ("lrpt_oracle_fast_sin", "lrpt_oracle_fast_cos"):
                getattr(L, f).argtypes = [C.c_float]
                getattr(L, f).restype = C.c_float
            L.lrpt_oracle_cabsf.argtypes = [C.c_float, C.c_float]
            L.lrpt_oracle_cabsf.restype = C.c_float
            L.lrpt_oracle_quantise.argtypes = [C.c_float]
            L.lrpt_oracle_quantise.restype = C.c_int8
            L.lrpt_oracle_freq_delta.argtypes = [C.c_float, C.c_float]
            L.lrpt_oracle_freq_delta.restype = C.c_float
            cls._lib = L
        return cls._lib

    def __init__(self, samplerate=230000, symrate=72000, interp=5, order=32, oqpsk=0, bps=16,
                 pll_bw=1.0, sym_bw=0.00005, freq_max=-1.0):
        self.L = self.lib()
        self.s = _OracleStruct()
        self.bps = bps
        if self.L.lrpt_oracle_init(C.byref(self.s), pll_bw, sym_bw, samplerate, symrate, interp, order,
                                   oqpsk, freq_max, bps):
            raise ValueError("lrpt_oracle_init rejected the configuration")

    def __del__(self):
        try:
            self.L.lrpt_oracle_free(C.byref(self.s))
        except Exception:
            pass

    def process(self, raw, cap=None, want_float=True, want_substep=False):
        """want_substep adds `q` = sample_idx*interp + timing sub-step (the CUDA path's symbol index output)."""
        a = _as_raw(raw, self.bps)
        ncap = (a.size // 2 + 8) if cap is None else cap
        sub = np.zeros(ncap, np.uint8) if want_substep else None
        self.s.substep_out = sub.ctypes.data if want_substep else None
        fn = lambda p, n, sym, soft, idx, lock, cap_: self.L.lrpt_oracle_process(
            C.byref(self.s), p, n, sym, soft, idx, lock, cap_)
        try:
            res = _run(fn, raw, self.bps, ncap, want_float)
        finally:
            self.s.substep_out = None
        if want_substep:
            res["q"] = res.sample_idx * self.s.interp + sub[: res.sample_idx.size].astype(np.int64)
        return res

    def fir_all(self, raw):
        """filter_get at every (sample, sub-step) from a zeroed delay line: float32 [nsamples, interp, 2]."""
        a = _as_raw(raw, self.bps)
        n = a.size // 2
        out = np.empty((n, self.s.interp, 2), np.float32)
        if self.L.lrpt_oracle_fir_all(C.byref(self.s), a.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p)):
            raise MemoryError("lrpt_oracle_fir_all")
        return out

    def taps(self):
        n = self.s.taps * self.s.interp
        return np.ctypeslib.as_array(self.s.h, (n,)).copy()

    def history(self):
        n = max(self.s.taps - 1, 0)
        return np.ctypeslib.as_array(self.s.hist, (2 * n,)).copy().reshape(n, 2)

    def state(self):
        d = {k: getattr(self.s, k) for k in STATE_FIELDS}
        d.update(t_dual_state=self.s.t_dual_state, oq_inphase=self.s.oq_inphase, p_updown=self.s.p_updown,
                 nsamples=self.s.nsamples, nsymbols=self.s.nsymbols,
                 first_lock_symbol=self.s.first_lock_symbol)
        return d

    def set_state(self, **kw):
        for k, v in kw.items():
            setattr(self.s, k, v)

    def set_history(self, hist):
        h = np.ascontiguousarray(hist, np.float32).reshape(-1)
        n = 2 * max(self.s.taps - 1, 0)
        assert h.size == n
        C.memmove(self.s.hist, h.ctypes.data, 4 * n)


class _RefState(C.Structure):
    _fields_ = [(k, C.c_float) for k in
                ("t_prev", "t_phase", "t_freq", "t_center", "t_maxdev", "t_alpha", "t_beta",
                 "agc_gain", "agc_bias_re", "agc_bias_im",
                 "p_freq", "p_phase", "p_alpha", "p_beta", "p_err", "p_fmax")] + \
               [(k, C.c_int) for k in ("p_locked", "p_locked_once", "flt_idx", "flt_size", "flt_interp")]


class Ref:
    """The compiled reference. One private copy of the .so per instance => fresh statics.

    in_place=True loads oracle/_ref/libref_<kind>.so itself (ONE instance per process; the file shows up in the
    process's memory map, which is how the driver tells the reference arm of bench.py from a fallback) and
    enables power_on(): back to the state of a freshly loaded library, for many power-on streams in one process."""

    def __init__(self, samplerate=230000, symrate=72000, interp=5, order=32, oqpsk=0, bps=16,
                 pll_bw=1.0, sym_bw=0.00005, freq_max=-1.0, kind="strict", in_place=False):
        src = os.path.join(REF_DIR, "libref_%s.so" % kind)
        if not os.path.exists(src):
            raise FileNotFoundError(src + " (run `make -C oracle ref` where /root/reference exists)")
        self._init_args = (pll_bw, sym_bw, samplerate, int(symrate), interp, order, oqpsk, freq_max)
        if in_place:
            L = self.L = C.CDLL(src)
            L.ref_save_power_on.restype = C.c_int
            L.ref_power_on.restype = C.c_int
            if L.ref_save_power_on():
                raise RuntimeError("ref_save_power_on failed")
            L.ref_power_on()             # a second in-place instance in the same process starts fresh as well
        else:
            fd, self._tmp = tempfile.mkstemp(prefix="libref_", suffix=".so")
            os.close(fd)
            shutil.copyfile(src, self._tmp)
            L = self.L = C.CDLL(self._tmp)
            os.unlink(self._tmp)             # mapping stays valid; nothing left on disk
        self._in_place = in_place
        L.ref_init.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
        L.ref_process.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_long]
        L.ref_process.restype = C.c_long
        L.ref_get_state.argtypes = [C.POINTER(_RefState)]
        L.ref_get_taps.argtypes = [C.c_void_p, C.c_int]
        L.ref_get_taps.restype = C.c_int
        L.ref_get_history.argtypes = [C.c_void_p, C.c_int]
        L.ref_get_history.restype = C.c_int
        for f in ("ref_fast_sin", "ref_fast_cos", "ref_lut_tanh"):
            getattr(L, f).argtypes = [C.c_float]
            getattr(L, f).restype = C.c_float
        L.ref_rrc_coeff.argtypes = [C.c_int, C.c_uint, C.c_float, C.c_float]
        L.ref_rrc_coeff.restype = C.c_float
        L.ref_cabsf.argtypes = [C.c_float, C.c_float]
        L.ref_cabsf.restype = C.c_float
        for f in ("ref_pll_get_freq", "ref_mm_omega", "ref_agc_get_gain"):
            getattr(L, f).restype = C.c_float
        self.bps = bps
        # main.c:70,187 : the CLI holds symrate as float and demod_init takes int
        L.ref_init(pll_bw, sym_bw, samplerate, int(symrate), interp, order, oqpsk, freq_max)

    def process(self, raw, cap=None, want_float=True):
        fn = lambda p, n, sym, soft, idx, lock, cap_: self.L.ref_process(
            p, n, self.bps, sym, soft, idx, lock, cap_)
        return _run(fn, raw, self.bps, cap, want_float)

    def power_on(self):
        """Fresh process image of every static of the reference, then demod_init again (in_place instances)."""
        if not self._in_place:
            raise RuntimeError("power_on needs in_place=True")
        if self.L.ref_power_on():
            raise RuntimeError("ref_power_on failed")
        self.L.ref_init(*self._init_args)

    def taps(self):
        n = self.L.ref_get_taps(None, 0)
        a = np.empty(n, np.float32)
        self.L.ref_get_taps(a.ctypes.data_as(C.c_void_p), n)
        return a

    def history(self):
        n = self.L.ref_get_history(None, 0)
        a = np.empty((n, 2), np.float32)
        self.L.ref_get_history(a.ctypes.data_as(C.c_void_p), n)
        return a[1:]                     # the reference keeps taps samples; the oldest is never used again

    def state(self):
        s = _RefState()
        self.L.ref_get_state(C.byref(s))
        return {k: getattr(s, k) for k, _ in _RefState._fields_}


def fnv1a32(a):
    """FNV-1a-32 over the bytes of an array (SURVEY.md A.3 tap-bank tripwires)."""
    h = 0x811C9DC5
    for b in np.ascontiguousarray(a).view(np.uint8).tolist():
        h = ((h ^ b) * 0x01000193) & 0xFFFFFFFF
    return h
