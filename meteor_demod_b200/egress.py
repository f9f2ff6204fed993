"""Soft-symbol egress rules of the reference host loop (main.c:303-323), as host logic.

The C ABI returns every symbol, ungated. The reference writes its 512-symbol ring
(main.c:20,34) only when pll_did_lock_once() is true at the moment a ring block
completes (main.c:308-316) and flushes the partial last block unconditionally
(main.c:321). `gate` turns (all symbols, first_lock_symbol) into the bytes the
reference's output file holds.
"""
import numpy as np

RINGSIZE = 512            # symbols, main.c:20


def first_written_symbol(first_lock_symbol):
    """Index of the first symbol that reaches the output file, or None if no full block does."""
    if first_lock_symbol is None or first_lock_symbol < 0:
        return None
    return RINGSIZE * (first_lock_symbol // RINGSIZE)


def gate(soft, first_lock_symbol, ref_compatible_tail=False):
    """soft: int8 [nsym, 2] of ALL symbols of a run. Returns the output file content as bytes.

    Default: full blocks from the first block that completed after lock, then the valid
    bytes of the trailing partial block. ref_compatible_tail=True also reproduces the
    reference's final-flush length bug (main.c:321 writes 2*ring_idx bytes: the valid
    ones followed by stale ring content; bytes past the ring are an out-of-bounds read
    in the reference and are zero-filled here).
    """
    soft = np.ascontiguousarray(soft, np.int8).reshape(-1, 2)
    nsym = soft.shape[0]
    nfull = nsym // RINGSIZE
    tail = soft[nfull * RINGSIZE:].reshape(-1)
    start = first_written_symbol(first_lock_symbol)
    parts = []
    if start is not None and start < nfull * RINGSIZE:
        parts.append(soft[start: nfull * RINGSIZE].reshape(-1))
    parts.append(tail)
    if ref_compatible_tail and tail.size:
        ring = np.zeros(2 * RINGSIZE, np.int8)
        if nfull:
            ring[:] = soft[(nfull - 1) * RINGSIZE: nfull * RINGSIZE].reshape(-1)
        stale = np.zeros(tail.size, np.int8)
        lo, hi = tail.size, min(2 * tail.size, 2 * RINGSIZE)
        if hi > lo:
            stale[: hi - lo] = ring[lo:hi]
        parts.append(stale)
    return np.concatenate(parts).tobytes() if parts else b""


WAV_BLOCK = 32768         # wavfile.c:8


def consumed_samples(nbytes, bps):
    """wav_read only consumes whole 32 KiB blocks (wavfile.c:55): trailing bytes are dropped."""
    return (nbytes // WAV_BLOCK) * WAV_BLOCK // (bps // 4)
