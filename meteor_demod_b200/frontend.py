"""Host-side mirror of the decoder front-end entry points (include/lrpt_b200.h, csrc/frontend.cu): frame
synchronisation and Viterbi decoding of the demodulator's int8 soft-symbol stream on the device. Plumbing only
(torch tensors for device memory); all computation is in liblrpt_b200.so."""
import ctypes as C

import torch

from . import _lib

CADU, CADU_SYMS = 1024, 8192


def _check(rc, what):
    if rc:
        raise _lib.LrptError(rc, what)


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def sync_scores(soft):
    """soft: int8 device tensor [nsym, 2] (or flat), 16-byte aligned. Returns (score, hyp) uint8 [nsym]."""
    lib = _lib.load()
    s = soft.reshape(-1)
    nsym = s.numel() // 2
    pad = (nsym + 3) // 4 * 4
    score = torch.empty(pad, dtype=torch.uint8, device=s.device)
    hyp = torch.empty(pad, dtype=torch.uint8, device=s.device)
    words = torch.empty(lib.lrpt_fe_sync_words(nsym), dtype=torch.int32, device=s.device)
    _check(lib.lrpt_fe_sync_device(s.data_ptr(), nsym, score.data_ptr(), hyp.data_ptr(), words.data_ptr(), _stream(s.device)),
           "lrpt_fe_sync_device")
    return score[:nsym], hyp[:nsym]


def window_peaks(score, hyp, window=CADU_SYMS):
    lib = _lib.load()
    nsym = score.numel()
    nw = (nsym + window - 1) // window
    off = torch.empty(nw, dtype=torch.int32, device=score.device)
    oh = torch.empty(nw, dtype=torch.uint8, device=score.device)
    osc = torch.empty(nw, dtype=torch.uint8, device=score.device)
    _check(lib.lrpt_fe_peaks_device(score.data_ptr(), hyp.data_ptr(), nsym, window, off.data_ptr(), oh.data_ptr(),
                                    osc.data_ptr(), _stream(score.device)), "lrpt_fe_peaks_device")
    return off, oh, osc


class Viterbi:
    """Batch decoder: frames (start symbol, symmetry) of one soft stream -> CADUs [n, 1024] uint8 + path metrics."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        self.dev = torch.device("cuda", device)
        self.scratch = torch.empty(self.lib.lrpt_fe_viterbi_scratch_bytes(device), dtype=torch.uint8, device=self.dev)

    def decode(self, soft, frame_off, frame_hyp):
        s = soft.reshape(-1)
        n = int(frame_off.numel())
        cadu = torch.empty((n, CADU), dtype=torch.uint8, device=self.dev)
        metric = torch.empty(n, dtype=torch.int32, device=self.dev)
        off = frame_off.to(torch.int32).contiguous()
        hyp = frame_hyp.to(torch.uint8).contiguous()
        _check(self.lib.lrpt_fe_viterbi_device(s.data_ptr(), s.numel() // 2, off.data_ptr(), hyp.data_ptr(), n, cadu.data_ptr(),
                                               metric.data_ptr(), self.scratch.data_ptr(), self.scratch.numel(),
                                               _stream(self.dev)), "lrpt_fe_viterbi_device")
        return cadu, metric
