"""oracle/pyfrontend.py -- TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/frontend_oracle.c (the CPU restatement of
the LRPT decoder front-end: frame-sync correlation and the CCSDS r=1/2 K=7 Viterbi decoder) plus a reference
TRANSMITTER for test vectors (CADUs with the attached sync marker, convolutional encoder, QPSK soft symbols)."""
import ctypes as C

import numpy as np

from . import pyoracle

CADU, CADU_SYMS, ASM = 1024, 8192, bytes([0x1A, 0xCF, 0xFC, 0x1D])
_lib = None


def lib():
    global _lib
    if _lib is None:
        pyoracle.Oracle.lib()                                  # builds the port when missing
        L = C.CDLL(pyoracle.PORT_SO)
        L.fe_conv_encode.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.POINTER(C.c_uint)]
        L.fe_sync_pattern.argtypes = [C.c_int]
        L.fe_sync_pattern.restype = C.c_uint64
        L.fe_sync_scores.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
        L.fe_window_peaks.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fe_window_peaks.restype = C.c_long
        L.fe_viterbi_cadu.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_int, C.c_void_p]
        L.fe_viterbi_cadu.restype = C.c_int32
        _lib = L
    return _lib


def conv_encode(data, state=0):
    """bytes -> encoded bits (uint8 0/1, I arm then Q arm per input bit), new encoder state."""
    a = np.frombuffer(bytes(data), np.uint8)
    out = np.empty(16 * a.size, np.uint8)
    st = C.c_uint(state)
    lib().fe_conv_encode(a.ctypes.data, 8 * a.size, out.ctypes.data, C.byref(st))
    return out, st.value


def sync_pattern(h):
    return int(lib().fe_sync_pattern(h))


def sync_scores(soft):
    s = np.ascontiguousarray(soft, np.int8).reshape(-1, 2)
    score, hyp = np.empty(s.shape[0], np.uint8), np.empty(s.shape[0], np.uint8)
    lib().fe_sync_scores(s.ctypes.data, s.shape[0], score.ctypes.data, hyp.ctypes.data)
    return score, hyp


def window_peaks(score, hyp, window=CADU_SYMS):
    n = score.size
    nw = (n + window - 1) // window
    off, oh, osc = np.empty(nw, np.uint32), np.empty(nw, np.uint8), np.empty(nw, np.uint8)
    lib().fe_window_peaks(score.ctypes.data, hyp.ctypes.data, n, window, off.ctypes.data, oh.ctypes.data, osc.ctypes.data)
    return off, oh, osc


def viterbi_cadu(soft, start, h):
    s = np.ascontiguousarray(soft, np.int8).reshape(-1, 2)
    out = np.zeros(CADU, np.uint8)
    metric = lib().fe_viterbi_cadu(s.ctypes.data, s.shape[0], int(start), int(h), out.ctypes.data)
    return out, int(metric)


def transmit(frames, amp=60, noise=0.0, turns=0, swap=False, lead=0, seed=1):
    """frames: uint8 [n, 1020] payloads -> int8 soft symbols [nsym, 2] of the continuous encoded stream
    ASM + payload per CADU (encoded 1 -> +amp), `lead` random symbols in front, Gaussian noise, then the channel's
    symmetry: I/Q swap first, then `turns` quarter turns (I,Q) -> (-Q, I)."""
    rng = np.random.default_rng(seed)
    frames = np.ascontiguousarray(frames, np.uint8).reshape(-1, CADU - 4)
    stream = b"".join(ASM + f.tobytes() for f in frames)
    bits, _ = conv_encode(stream)
    sym = (2.0 * bits.astype(np.float64) - 1.0).reshape(-1, 2) * amp
    if lead:
        sym = np.concatenate([rng.choice([-amp, amp], size=(lead, 2)).astype(np.float64), sym])
    sym = sym + noise * rng.standard_normal(sym.shape)
    i, q = sym[:, 0].copy(), sym[:, 1].copy()
    if swap:
        i, q = q, i
    for _ in range(turns % 4):
        i, q = -q, i
    out = np.stack([i, q], axis=1)
    return np.clip(np.rint(out), -127, 127).astype(np.int8)
