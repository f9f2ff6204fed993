"""Time-sharding of ONE long I/Q stream (BASELINE config 4).

The recurrence of the reference is sequential and decision-chaotic (SURVEY.md finding 3): a
chunk that does not start from the exact float state of its predecessor never re-joins the
sequential trajectory bit for bit. What CAN be done in parallel is *statistical* parity
("Tier-S"), and that is what this module does, on top of the exact batch engine:

  1. plan     the stream is cut at B_c = W + c*C; chunk c is demodulated as an independent
              stream over samples [c*C, c*C + W + C + V): W samples of warm-up from power-on
              state (AGC, timing, Costas loop acquire), C samples it owns, V samples of overlap
              into its successor. All chunks run in ONE batch launch (one recurrence lane each).
  2. quadrant a QPSK Costas loop locks with a k*90 degree ambiguity. On the overlap
              [B_c, B_c+V) chunks c-1 and c demodulate the same samples; the rotation k_c that
              best maps c onto c-1, and the agreement of hard decisions under it, come from the
              paired soft symbols.
  3. scan     K_c = (k_1 + ... + k_c) mod 4 -- a prefix sum over chunk boundaries; across GPUs
              the only exchange is the overlap symbols of a rank's last chunk (to the next
              rank) and one integer per rank (all-gather), i.e. "boundary state handed rank to
              rank and nothing else".
  4. stitch   chunk c contributes the symbols between the cut points (placed mid-way between
              two symbols of the predecessor so that a symbol instant near the boundary is
              neither duplicated nor dropped), rotated by -K_c (exact on int8 pairs).
  5. verify   every boundary reports its agreement (share of overlap symbols whose hard decisions
              match after de-rotation); a low value flags a chunk that had not locked by its
              boundary (warm-up too short for the carrier offset) or a cycle slip. The exact remedy
              is a state hand-off: re-run that chunk from the state its predecessor exports at B_c
              (lrpt_export_state / lrpt_import_state) -- sequential for the flagged chunks only.

Chunk 0 starts from the true power-on state, so its symbols are bit-exact; later chunks are
Tier-S: same symbol count and quadrant, a measured fraction epsilon of symbols off by more
than one LSB -- comparable to the reference's own FMA-vs-strict build difference.

Only host logic lives here (torch tensor ops for the index arithmetic, torch.distributed for the
exchange). Demodulation is `GpuEngine` (liblrpt_b200.so); the tests feed `stitch` with chunks
demodulated by the CPU oracle to check the logic without a GPU, including a 2-rank gloo run.
"""
import ctypes as C
import math
from dataclasses import dataclass

import numpy as np
import torch


NATIVE_ESTIMATOR = True     # seeded warm-ups: the library's kernel (csrc/acquire.cu) instead of the torch.fft form of acquire.py


def _on_device(soft, q, base):
    """Rows as the GPU engine leaves them: int32 row-local sub-step indices + a base per row, on the
    device. Those are joined by the library's stitch kernels (csrc/shard_stitch.cu); anything else (CPU
    tensors in the tests, small int64 tables) by the torch ops below -- same arithmetic."""
    return soft.is_cuda and q.dtype == torch.int32 and base is not None


def _lib_call(name, dev, *args):
    from ._lib import LrptError, load
    with torch.cuda.device(dev):
        rc = getattr(load(), name)(*args, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc:
        raise LrptError(rc, name)


def _p(t):
    return C.c_void_p(t.data_ptr())


def _abs_row(q, base, r):
    """Absolute sub-step indices of row r, int64."""
    row = q[r].to(torch.int64)
    return row if base is None else row + base[r]


@dataclass
class Plan:
    nsamples: int      # samples in the stream
    chunk: int         # C: samples owned per chunk
    warm: int          # W: warm-up before the owned region
    overlap: int       # V: overlap into the successor
    interp: int        # L: timing sub-steps per sample
    nchunks: int = 0
    cut_shift: int = 0 # samples added to every cut target (hand-off scheme: rows start V samples after their boundary)

    def __post_init__(self):
        if self.chunk <= 0 or self.warm < 0 or self.overlap <= 0:
            raise ValueError("chunk > 0, warm >= 0, overlap > 0 required")
        self.nchunks = max(1, math.ceil((self.nsamples - self.warm) / self.chunk))

    def start(self, c):            # first sample chunk c demodulates
        return c * self.chunk

    def boundary(self, c):         # B_c: first sample chunk c owns (chunk 0 owns from 0)
        return 0 if c == 0 else self.warm + c * self.chunk

    def cut_target(self, c):
        """Sub-step index around which chunks c-1 and c are joined: a little INSIDE chunk c's owned region,
        so that both rows have symbols on either side of the cut (in the two-pass scheme row c only starts
        at B_c; a symbol instant falling exactly on B_c is in one row and not the other)."""
        return (self.boundary(c) + self.cut_shift + min(64, self.overlap // 4)) * self.interp

    def cut_targets(self, c0, c1, device=None):
        """cut_target(c) for c in [c0, c1), c0 >= 1, as an int64 tensor (no Python loop: a long stream has 1e5 chunks)."""
        assert c0 >= 1
        c = torch.arange(c0, c1, dtype=torch.int64, device=device)
        return (self.warm + c * self.chunk + self.cut_shift + min(64, self.overlap // 4)) * self.interp

    @property
    def n_main(self):              # samples up to the successor's boundary
        return self.warm + self.chunk

    @property
    def padded(self):              # buffer length that lets every chunk read n_main + 2*overlap samples
        return (self.nchunks - 1) * self.chunk + self.n_main + 2 * self.overlap


def rotate_quarter_turns(soft, k):
    """soft [...,2] int8 (I,Q) times j**k, exact: (I,Q) -> (-Q,I) per quarter turn."""
    i, q = soft[..., 0], soft[..., 1]
    k = k % 4
    if isinstance(k, torch.Tensor):
        k = k.reshape(k.shape + (1,) * (i.dim() - k.dim()))
        ri = torch.where(k == 0, i, torch.where(k == 1, -q, torch.where(k == 2, -i, q)))
        rq = torch.where(k == 0, q, torch.where(k == 1, i, torch.where(k == 2, -q, -i)))
    else:
        ri, rq = [(i, q), (-q, i), (-i, -q), (q, -i)][k]
    return torch.stack((ri, rq), dim=-1)


def _first_at_or_after(q, count, target):
    """Per row: index of the first valid symbol with q >= target (q rows ascending; entries beyond
    count are ignored). q [M,cap] int64, count [M], target [M] -> [M] int64."""
    big = torch.iinfo(torch.int64).max
    cols = torch.arange(q.shape[1], device=q.device)
    qm = torch.where(cols[None, :] < count[:, None], q, torch.full_like(q, big))
    return torch.searchsorted(qm, target[:, None].contiguous()).squeeze(1)


def boundary_quadrants(soft, q, count, Bq, npairs=None, base=None):
    """Rows are consecutive chunks. For every boundary (row c-1 | row c) at sub-step Bq[c-1], pair the
    overlap symbols of the two rows and find the quarter-turn count k mapping row c onto row c-1.

    Returns k [M-1] int64, agreement [M-1] float (share of pairs whose hard decisions match under k
    and whose symbol instants differ by at most 2 sub-steps), cut [M-1] int64 (sub-step index that
    separates the two rows' shares: mid-way between two symbols of row c-1)."""
    M = soft.shape[0]
    dev = soft.device
    if M < 2:
        z = torch.zeros(0, dtype=torch.int64, device=dev)
        return z, torch.zeros(0, device=dev), z
    B = Bq.to(torch.int64)
    if _on_device(soft, q, base):
        assert soft.stride(1) == 2 and soft.stride(2) == 1 and q.stride(1) == 1
        cnt, b64, tgt = count.to(torch.int32).contiguous(), base.to(torch.int64).contiguous(), B.contiguous()
        cut = torch.empty(M - 1, dtype=torch.int64, device=dev)
        ia, ib, nav = (torch.empty(M - 1, dtype=torch.int32, device=dev) for _ in range(3))
        _lib_call("lrpt_shard_find_cuts_device", dev, _p(q), q.stride(0) * 4, _p(cnt), _p(b64), M, _p(tgt), _p(cut),
                  _p(ia), _p(ib), _p(nav))
        nmin = int(nav.min().item())
        npairs = nmin if npairs is None else min(npairs, nmin)
        if npairs < 8:
            return torch.zeros(M - 1, dtype=torch.int64, device=dev), torch.zeros(M - 1, device=dev), cut
        k, same = (torch.empty(M - 1, dtype=torch.int32, device=dev) for _ in range(2))
        _lib_call("lrpt_shard_quadrants_device", dev, _p(soft), soft.stride(0), _p(q), q.stride(0) * 4, _p(b64), M,
                  _p(ia), _p(ib), npairs, _p(k), _p(same))
        return k.to(torch.int64), same.to(torch.float32) / npairs, cut
    if base is not None:
        q = q.to(torch.int64) + base[:, None]
    a_q, a_s, a_n = q[:-1], soft[:-1], count[:-1]
    b_q, b_s, b_n = q[1:], soft[1:], count[1:]
    ia = _first_at_or_after(a_q, a_n, B)                     # first symbol of row c-1 inside the overlap
    prev_q = torch.gather(a_q, 1, (ia - 1).clamp(min=0)[:, None]).squeeze(1)
    next_q = torch.gather(a_q, 1, ia.clamp(max=a_q.shape[1] - 1)[:, None]).squeeze(1)
    cut = torch.where((ia > 0) & (ia < a_n), (prev_q + next_q) // 2, B)
    ib = _first_at_or_after(b_q, b_n, cut + 1)               # first symbol of row c after the cut
    navail = torch.minimum(a_n - ia, b_n - ib).clamp(min=0)
    nmin = int(navail.min().item())
    npairs = nmin if npairs is None else min(npairs, nmin)
    if npairs < 8:
        return torch.zeros(M - 1, dtype=torch.int64, device=dev), torch.zeros(M - 1, device=dev), cut
    j = torch.arange(npairs, device=dev)
    ga, gb = ia[:, None] + j[None, :], ib[:, None] + j[None, :]
    sa = torch.gather(a_s, 1, ga[:, :, None].expand(-1, -1, 2)).to(torch.float32)
    sb = torch.gather(b_s, 1, gb[:, :, None].expand(-1, -1, 2)).to(torch.float32)
    dq = (torch.gather(a_q, 1, ga) - torch.gather(b_q, 1, gb)).abs()
    ai, aq, bi, bq = sa[..., 0], sa[..., 1], sb[..., 0], sb[..., 1]
    # correlation of row c-1 with row c rotated by k quarter turns: (I,Q) -> (-Q,I) per turn
    scores = torch.stack(((ai * bi + aq * bq).sum(1), (-ai * bq + aq * bi).sum(1),
                          (-ai * bi - aq * bq).sum(1), (ai * bq - aq * bi).sum(1)), dim=1)
    k = scores.argmax(dim=1)
    rb = rotate_quarter_turns(sb, k)
    same = (torch.sign(rb[..., 0]) == torch.sign(ai)) & (torch.sign(rb[..., 1]) == torch.sign(aq)) & (dq <= 2)
    return k, same.float().mean(dim=1), cut


def boundary_quadrants_oqpsk(soft, q, count, Bq, half, npairs=None):
    """boundary_quadrants for OQPSK rows (torch ops; q absolute int64). The I arm is sampled half a symbol before
    the Q arm (demod.c:66-83), and the four lock points of the loop are not alike: at an EVEN number of quarter
    turns two rows take their symbols at the same instants and differ by a sign; at an ODD number the later
    row's symbol instants sit half a symbol (`half` sub-steps) off, its Q arm carries the earlier row's I
    stream and its I arm the earlier row's Q stream of the symbol before:
        row b (k = 1):  b.Q_j = -a.I_{m+1},  b.I_j = a.Q_m      with  q_b[j] ~ q_a[m] + half
    (k = 3: both signs flipped). The timing offset between paired symbols tells even from odd, a correlation
    then picks the sign. Returns k, agreement, cut like boundary_quadrants."""
    M = soft.shape[0]
    dev = soft.device
    if M < 2:
        z = torch.zeros(0, dtype=torch.int64, device=dev)
        return z, torch.zeros(0, device=dev), z
    B = Bq.to(torch.int64)
    a_q, a_s, a_n = q[:-1], soft[:-1], count[:-1]
    b_q, b_s, b_n = q[1:], soft[1:], count[1:]
    ia = _first_at_or_after(a_q, a_n, B)
    prev_q = torch.gather(a_q, 1, (ia - 1).clamp(min=0)[:, None]).squeeze(1)
    next_q = torch.gather(a_q, 1, ia.clamp(max=a_q.shape[1] - 1)[:, None]).squeeze(1)
    cut = torch.where((ia > 0) & (ia < a_n), (prev_q + next_q) // 2, B)
    ib = _first_at_or_after(b_q, b_n, cut + 1)
    navail = torch.minimum(a_n - ia - 1, b_n - ib).clamp(min=0)           # one symbol of row a in reserve (m + 1)
    nmin = int(navail.min().item()) if (ia > 0).all() else 0
    npairs = nmin if npairs is None else min(npairs, nmin)
    if npairs < 8:
        return torch.zeros(M - 1, dtype=torch.int64, device=dev), torch.zeros(M - 1, device=dev), cut
    j = torch.arange(npairs, device=dev)
    ga, gb = ia[:, None] + j[None, :], ib[:, None] + j[None, :]
    d = (torch.gather(b_q, 1, gb) - torch.gather(a_q, 1, ga)).to(torch.float32)
    dm = d.median(dim=1).values                                            # timing offset of row b against row a
    odd = dm.abs() > half / 2
    off = torch.where(dm > 0, 0, -1)                                       # m = ia + j + off for odd boundaries
    gm = (ga + off[:, None]).clamp(min=0)

    def col(t, idx, c):
        return torch.gather(t[..., c], 1, idx).to(torch.float32)
    aI, aQ = col(a_s, ga, 0), col(a_s, ga, 1)
    bI, bQ = col(b_s, gb, 0), col(b_s, gb, 1)
    aQm, aIm1 = col(a_s, gm, 1), col(a_s, gm + 1, 0)
    even_score = (aI * bI + aQ * bQ).sum(1)                                # > 0: k = 0, < 0: k = 2
    odd_score = (-aIm1 * bQ + aQm * bI).sum(1)                             # > 0: k = 1, < 0: k = 3
    k = torch.where(odd, torch.where(odd_score >= 0, 1, 3), torch.where(even_score >= 0, 0, 2))
    sg = torch.where((k == 0) | (k == 1), 1.0, -1.0)[:, None]
    same_even = (torch.sign(sg * bI) == torch.sign(aI)) & (torch.sign(sg * bQ) == torch.sign(aQ)) & (d.abs() <= 2)
    same_odd = (torch.sign(-sg * bQ) == torch.sign(aIm1)) & (torch.sign(sg * bI) == torch.sign(aQm)) & \
               ((d - dm[:, None]).abs() <= 2)
    same = torch.where(odd[:, None], same_odd, same_even)
    return k, same.float().mean(dim=1), cut


def turn_oqpsk_state(st, turns):
    """A row's loop state moved from the lock point it acquired to the one `turns` quarter turns back (the
    sequential run's), for OQPSK: the Costas NCO turns as for QPSK (p_phase -= turns*pi/2, pll.c:16); for an
    odd count the arms also change roles -- what was sampled as I (at the pi crossing, timing.c:47-50) is the new
    Q and vice versa -- so the timing NCO moves by pi, the dual-threshold state toggles, and the two remembered
    samples swap with the signs of the turn: new Q' = +-old I, new I' = -+old Q (demod.c:54, timing.c:13).
    st: dict of 1-D float64 tensors (p_phase, t_phase, t_dual_state, t_prev, oq_inphase); returns a new dict."""
    kk = (turns & 3).to(torch.int64)
    out = dict(st)
    out["p_phase"] = (st["p_phase"].float().double() - kk.double() * 1.57079632679489661923).float().double()
    odd = (kk & 1) == 1
    s_q = torch.where(kk == 1, 1.0, -1.0).double()                        # new Q' = s_q * old I ; new I' = -s_q * old Q
    flip = torch.where(kk == 2, -1.0, 1.0).double()                       # half a turn: both arms change sign
    dual = st["t_dual_state"].to(torch.int64)
    pi = 3.14159265358979323846
    ph = st["t_phase"].float().double()
    out["t_phase"] = torch.where(odd, torch.where(dual == 1, ph + pi, ph - pi), ph).float().double()
    out["t_dual_state"] = torch.where(odd, 3 - dual, dual).double()
    out["t_prev"] = torch.where(odd, s_q * st["oq_inphase"], flip * st["t_prev"])
    out["oq_inphase"] = torch.where(odd, -s_q * st["t_prev"], flip * st["oq_inphase"])
    return out


def _neighbour_exchange(dist, to_next, from_prev):
    """One batched point-to-point exchange down the chain of ranks: `to_next` (tensor or None) goes to rank + 1,
    `from_prev` (tensor or None) is filled by rank - 1. With NCCL the pair is one grouped launch
    (batch_isend_irecv = ncclGroupStart/End) instead of separate, serialised sends and receives."""
    rank, world = dist.get_rank(), dist.get_world_size()
    ops = []
    if to_next is not None and rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, to_next, rank + 1))
    if from_prev is not None and rank > 0:
        ops.append(dist.P2POp(dist.irecv, from_prev, rank - 1))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def stitch(soft, q, count, plan, first_chunk=0, dist=None, base=None, oqpsk_half=None):
    """Quadrant scan + concatenation of the owned symbols.

    soft [M,cap,2] int8, q [M,cap] int64 ABSOLUTE sub-step index (sample*interp + sub-step from the
    start of the stream) -- or row-local with `base` [M] int64 to be added, which is how the GPU engine
    leaves it (int32; such rows are joined by the library's stitch kernels, csrc/shard_stitch.cu) --
    count [M] int64: the chunks this process demodulated, chunk indices first_chunk .. first_chunk+M-1. Single process: pass all chunks. Multi-process (`dist` =
    torch.distributed, ranks own consecutive runs of chunks in rank order): the exchange is (a) the
    symbols around the next boundary of a rank's LAST chunk, sent to the next rank, (b) one all-gather
    of an integer per rank -- the prefix sum of quarter turns is then local.

    Returns dict: soft [n,2] int8 (this process's share, in stream order), k / agreement per interior
    boundary, boundary_prev = (k, agreement) of the boundary to the previous rank, K_first, K_last."""
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    M, L, dev = soft.shape[0], plan.interp, soft.device
    big = torch.iinfo(torch.int64).max
    Bq = plan.cut_targets(first_chunk + 1, first_chunk + M, dev)
    if oqpsk_half is not None:
        k, agree, cut = boundary_quadrants_oqpsk(soft, q if base is None else q.to(torch.int64) + base[:, None],
                                                 count, Bq, oqpsk_half)
    else:
        k, agree, cut = boundary_quadrants(soft, q, count, Bq, base=base)

    k_prev, agree_prev, cut_prev = 0, None, None
    if world > 1:
        # (a) hand the neighbourhood of my last chunk's far boundary to the next rank
        width = int(plan.overlap) + 64
        pack = torch.zeros((width + 1, 3), dtype=torch.int64, device=dev)     # row `width` carries the fill count
        if first_chunk + M < plan.nchunks:
            nextB = plan.cut_target(first_chunk + M)
            n = int(count[-1].item())
            q_last = _abs_row(q, base, M - 1)
            sel = (q_last[:n] >= nextB - 64 * L).nonzero().squeeze(1)[:width]
            pack[: sel.numel(), 0] = q_last[sel]
            pack[: sel.numel(), 1:] = soft[-1, sel].to(torch.int64)
            pack[width, 0] = sel.numel()
        recv = torch.zeros_like(pack)
        _neighbour_exchange(dist, pack, recv)
        nrecv = recv[width, 0]
        if rank > 0:
            n = int(nrecv.item())
            cap2 = max(n, soft.shape[1])
            two_s = torch.zeros((2, cap2, 2), dtype=torch.int8, device=dev)
            two_q = torch.zeros((2, cap2), dtype=torch.int64, device=dev)
            two_s[0, :n] = recv[:n, 1:].to(torch.int8)
            two_q[0, :n] = recv[:n, 0]
            two_s[1, : soft.shape[1]] = soft[0]
            two_q[1, : soft.shape[1]] = _abs_row(q, base, 0)
            two_n = torch.stack((torch.tensor(n, dtype=torch.int64, device=dev), count[0]))
            tgt = torch.tensor([plan.cut_target(first_chunk)], device=dev)
            if oqpsk_half is not None:
                kp, ap, cp = boundary_quadrants_oqpsk(two_s, two_q, two_n, tgt, oqpsk_half)
            else:
                kp, ap, cp = boundary_quadrants(two_s, two_q, two_n, tgt)
            k_prev, agree_prev, cut_prev = int(kp[0].item()), float(ap[0].item()), int(cp[0].item())

    # (b) prefix sum of quarter turns over ranks
    K_first = 0
    if world > 1:
        mine = torch.tensor([(k_prev + (int(k.sum().item()) if M > 1 else 0)) % 4], dtype=torch.int64, device=dev)
        lst = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(lst, mine)
        K_first = (sum(int(x.item()) for x in lst[:rank]) + k_prev) % 4

    K = torch.empty(M, dtype=torch.int64, device=dev)
    K[0] = K_first
    if M > 1:
        K[1:] = (K_first + torch.cumsum(k, 0)) % 4
    lo = torch.full((M,), -1, dtype=torch.int64, device=dev)
    hi = torch.full((M,), big, dtype=torch.int64, device=dev)
    if cut_prev is not None:
        lo[0] = cut_prev
    if M > 1:
        lo[1:] = cut
        hi[:-1] = cut
    if first_chunk + M < plan.nchunks:
        # my last chunk ends where the next rank's first chunk begins: same rule, evaluated locally
        nextB = torch.tensor([plan.cut_target(first_chunk + M)], dtype=torch.int64, device=dev)
        q_last = _abs_row(q, base, M - 1)[None, :]
        ia = _first_at_or_after(q_last, count[-1:], nextB)
        pq = torch.gather(q_last, 1, (ia - 1).clamp(min=0)[:, None]).squeeze(1)
        nq = torch.gather(q_last, 1, ia.clamp(max=q.shape[1] - 1)[:, None]).squeeze(1)
        hi[-1] = torch.where((ia > 0) & (ia < count[-1:]), (pq + nq) // 2, nextB)[0]
    else:
        hi[-1] = plan.nsamples * L - 1                       # nothing from the zero padding
    if _on_device(soft, q, base):
        # every row contributes one contiguous run (its sub-step indices ascend): find it, then copy + turn
        cnt, b64 = count.to(torch.int32).contiguous(), base.to(torch.int64).contiguous()
        start, ln = (torch.empty(M, dtype=torch.int32, device=dev) for _ in range(2))
        _lib_call("lrpt_shard_ranges_device", dev, _p(q), q.stride(0) * 4, _p(cnt), _p(b64), M, _p(lo), _p(hi), _p(start), _p(ln))
        ln64 = ln.to(torch.int64)
        off = (torch.cumsum(ln64, 0) - ln64).contiguous()
        total, longest = (int(v) for v in torch.stack((ln64.sum(), ln64.max())).tolist())
        out = torch.empty((total, 2), dtype=torch.int8, device=dev)
        turns = K.to(torch.int32).contiguous()
        _lib_call("lrpt_shard_gather_device", dev, _p(soft), soft.stride(0), M, longest, _p(start), _p(ln), _p(off), _p(turns), _p(out))
    else:
        if base is not None:
            q = q.to(torch.int64) + base[:, None]
        cols = torch.arange(soft.shape[1], device=dev)
        keep = (cols[None, :] < count[:, None]) & (q > lo[:, None]) & (q <= hi[:, None])
        out = rotate_quarter_turns(soft, K)[keep]
    return dict(soft=out, k=k, agreement=agree, boundary_prev=(k_prev, agree_prev),
                K_first=K_first, K_last=int(K[-1].item()))


def split_chunks(nchunks, world, rank):
    """Consecutive run of chunks of `rank`: [c0, c1). Even split: the first nchunks % world ranks take one
    chunk more, so no rank is left without chunks when nchunks >= world."""
    base, extra = divmod(nchunks, world)
    c0 = rank * base + min(rank, extra)
    return c0, c0 + base + (1 if rank < extra else 0)


def check_layout(nchunks, world, handoff=False):
    """The chunk layout of a multi-rank run, validated identically on every rank BEFORE any collective (a
    one-sided exception would leave the other ranks blocked in recv / all_gather): every rank needs a chunk,
    and in the hand-off scheme the rank holding chunk 0 needs two (chunk 1 continues chunk 0 exactly)."""
    if nchunks < world:
        raise ValueError("fewer chunks (%d) than ranks (%d)" % (nchunks, world))
    if handoff and nchunks > 1 and split_chunks(nchunks, world, 0)[1] < 2:
        raise ValueError("the rank holding chunk 0 needs at least two chunks (%d chunks over %d ranks)" % (nchunks, world))


def chunk_turns(res, M):
    """Cumulative quarter turns K_c of the M local chunks, from a stitch() result."""
    K = torch.empty(M, dtype=torch.int64, device=res["k"].device)
    K[0] = res["K_first"]
    if M > 1:
        K[1:] = (res["K_first"] + torch.cumsum(res["k"], 0)) % 4
    return K


class GpuEngine:
    """Chunks of one device-resident raw stream through liblrpt_b200.so, one recurrence lane each."""

    def __init__(self, raw, plan, device=0, first_chunk=0, nchunks=None, raw_first=0, seed_carrier=False, seed_nfft=1 << 17,
                 **cfg):
        """raw: device tensor of the raw dtype whose item 0 is I of stream sample `raw_first` -- the whole
        (padded) stream, or just the span this engine's chunks read: [plan.start(first_chunk),
        plan.start(last chunk) + n_main + 2*overlap). seed_carrier (opt-in, not what the reference does): every
        row but the stream's chunk 0 starts its Costas NCO at a coarse carrier estimate
        (acquire.py), so the warm-up need not cover the reference's slow sweep at large offsets."""
        from .demod import Demod
        self.plan, self.raw, self.first, self.raw_first = plan, raw, first_chunk, raw_first
        self.oqpsk = bool(cfg.get("oqpsk"))
        self.seed, self.seed_nfft = bool(seed_carrier), int(seed_nfft)
        self.M = plan.nchunks - first_chunk if nchunks is None else nchunks
        self.d = Demod(nstreams=self.M, device=device, interp_factor=plan.interp, **cfg)
        n_all = plan.n_main + plan.overlap
        self.cap = (self.d.capacity(n_all) + 7) // 8 * 8
        dev = raw.device
        self.soft = torch.empty((self.M, 2 * self.cap), dtype=torch.int8, device=dev)
        self.q = torch.empty((self.M, self.cap), dtype=torch.int32, device=dev)
        self.nsym = torch.zeros(self.M, dtype=torch.int32, device=dev)
        self.d.set_symbol_index_output(self.q)

    def _view(self, first_sample_of_chunk0, nsamples):
        p = self.plan
        return torch.as_strided(self.raw, (self.M, 2 * nsamples), (2 * p.chunk, 1),
                                storage_offset=2 * (p.start(self.first) + first_sample_of_chunk0 - self.raw_first))

    def _result(self, offset):
        p = self.plan
        base = ((torch.arange(self.first, self.first + self.M, device=self.raw.device, dtype=torch.int64)
                 * p.chunk + offset) * p.interp)
        # rows stay as the kernel wrote them: int32 sub-step indices counted from the row's own start + a base per row
        return self.soft.view(self.M, self.cap, 2), self.q, self.nsym.to(torch.int64), base

    def run(self, stream=None):
        """Single pass: every local chunk (warm-up + owned + overlap) in one batch launch. Returns
        soft [M,cap,2] int8, q [M,cap] int32 row-local sub-step indices, count [M] int64, base [M] int64 (device)."""
        n_all = self.plan.n_main + self.plan.overlap
        self.d.reset(stream=stream, asynchronous=True)
        self.d.process_device(self._view(0, n_all), self.soft, nsym=self.nsym, stream=stream, nsamples=n_all)
        self.d.sync(stream)                                  # the torch ops below run on torch's own stream
        return self._result(0)

    def warm_up(self):
        """Two-pass scheme, pass A: power-on -> boundary over the W warm-up samples; snapshot the states.
        Returns the warm-up symbols of local chunk 0 (they are output only if it is the stream's chunk 0)."""
        W = self.plan.warm
        self.d.reset()
        if self.seed:
            self.seed_carrier(nfft=self.seed_nfft)
        if W:
            self.d.process_device(self._view(0, W), self.soft, nsym=self.nsym, nsamples=W)
        self.d.sync()
        n0 = int(self.nsym[0].item()) if W else 0
        head = self.soft.view(self.M, self.cap, 2)[0, :n0].clone()
        self.d.snapshot()
        return head

    def owned(self, quarter_turns=None):
        """Pass B (quarter_turns None) / pass C: from the snapshot, demodulate owned + overlap samples."""
        p = self.plan
        n = p.chunk + p.overlap
        self.d.restore(None if quarter_turns is None else quarter_turns.cpu().numpy().astype(np.int32))
        self.d.process_device(self._view(p.warm, n), self.soft, nsym=self.nsym, nsamples=n)
        self.d.sync()
        return self._result(p.warm)

    def seed_carrier(self, nfft=1 << 17, group=None):
        """Rows of chunks >= 1: p_freq = the NCO step of a coarse carrier estimate over the row's first nfft samples
        (everything else stays power-on). Chunk 0 is left alone: the head of the stream is the sequential run.
        `group` rows go through the estimator at a time (default: 2^25 samples per batch of torch ops)."""
        from . import acquire
        from ._lib import State
        par = self.d.p
        nfft = min(nfft, self.plan.n_main)
        if group is None:
            group = max(64, (1 << 25) // nfft)
        rows = self.export_rows()
        off = State.p_freq.offset
        pow2 = nfft >= 256 and nfft <= 16384 and (nfft & (nfft - 1)) == 0
        if self.raw.is_cuda and pow2 and NATIVE_ESTIMATOR:
            # one launch of the library's own estimator over all rows (csrc/acquire.cu)
            f = acquire.estimate_cfo_device(self._view(0, nfft), par, nfft)
            pf = acquire.p_freq_for(f, par.symrate, bool(par.oqpsk))
            first = 1 if self.first == 0 else 0
            rows[first:, off: off + 4] = pf[first:].contiguous().view(-1, 1).view(torch.uint8)
        else:
            for r0 in range(0, self.M, group):
                r1 = min(self.M, r0 + group)
                x = acquire.to_complex(self._view(0, nfft)[r0:r1], par.bps)
                f = acquire.estimate_cfo(x, par.samplerate, par.symrate, bool(par.oqpsk))
                pf = acquire.p_freq_for(f, par.symrate, bool(par.oqpsk))
                first = 1 if self.first + r0 == 0 else 0
                rows[r0 + first: r1, off: off + 4] = pf[first:].contiguous().view(-1, 1).view(torch.uint8)
        self.import_rows(rows)

    # -- hand-off scheme (run_handoff) -----------------------------------------------------------------
    def pass_a(self):
        return self.warm_up()

    def pass_b(self):
        """From the live state at B_c: owned + overlap samples; every row ends at B_{c+1} + V."""
        p = self.plan
        n = p.chunk + p.overlap
        self.d.process_device(self._view(p.warm, n), self.soft, nsym=self.nsym, nsamples=n)
        self.d.sync()
        return self._result(p.warm)                          # views of the engine's buffers: valid until pass_c

    def pass_c(self):
        """From the imported states: owned + overlap samples starting V samples after the boundary."""
        p = self.plan
        n = p.chunk + p.overlap
        self.d.process_device(self._view(p.warm + p.overlap, n), self.soft, nsym=self.nsym, nsamples=n)
        self.d.sync()
        return self._result(p.warm + p.overlap)

    def export_rows(self):
        """Complete state of every row as bytes [M, state_size] (lrpt_state_t + delay line per row)."""
        from ._lib import State
        import ctypes as C
        sb = C.sizeof(State)
        buf = torch.empty(self.d.states_size(), dtype=torch.uint8, device=self.raw.device)
        self.d.export_states_device(buf)
        self.d.sync()
        hb = (buf.numel() - self.M * sb) // self.M
        return torch.cat((buf[: self.M * sb].view(self.M, sb), buf[self.M * sb:].view(self.M, hb)), dim=1)

    def import_rows(self, rows):
        from ._lib import State
        import ctypes as C
        sb = C.sizeof(State)
        buf = torch.cat((rows[:, :sb].reshape(-1), rows[:, sb:].reshape(-1))).contiguous()
        self.d.import_states_device(buf, check=True)         # on the handle's own stream, ordered after torch's (recv, cat)
        self.d.sync()

    def rotate_rows(self, rows, turns):
        """Every row's Costas NCO turned back by `turns` quarter turns, exactly as lrpt_restore does:
        p_phase = (float)((double)p_phase - (turns & 3) * M_PI/2)  (pll.c:16). OQPSK rows also move their timing
        NCO and swap the remembered arm samples: turn_oqpsk_state."""
        from ._lib import State
        if self.oqpsk:
            out = rows.clone()
            f32 = ("p_phase", "t_phase", "t_prev", "oq_inphase")
            def field(name, dt):
                o = getattr(State, name).offset
                return out[:, o: o + 4].contiguous().view(dt).reshape(-1)
            st = {k: field(k, torch.float32).double() for k in f32}
            st["t_dual_state"] = field("t_dual_state", torch.int32).double()
            new = turn_oqpsk_state(st, turns.to(out.device))
            for k in f32:
                o = getattr(State, k).offset
                out[:, o: o + 4] = new[k].float().view(-1, 1).view(torch.uint8)
            o = State.t_dual_state.offset
            out[:, o: o + 4] = new["t_dual_state"].to(torch.int32).view(-1, 1).view(torch.uint8)
            return out
        off = State.p_phase.offset
        out = rows.clone()
        ph = out[:, off: off + 4].contiguous().view(torch.float32).reshape(-1)
        new = (ph.double() - (turns.to(ph.device) & 3).double() * 1.57079632679489661923).float()
        out[:, off: off + 4] = new.view(-1, 1).view(torch.uint8)
        return out

    def close(self):
        self.d.set_symbol_index_output(None)
        self.d.close()


class _Phases:
    """Wall time of the phases of one run (every phase ends in a device synchronisation of its own, so the host
    clock brackets the device work). Disabled unless a dict is passed in."""

    def __init__(self, out):
        import time
        self.out, self.clock = out, time.perf_counter
        self.t = self.clock()

    def mark(self, name):
        if self.out is None:
            return
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        now = self.clock()
        self.out[name] = self.out.get(name, 0.0) + (now - self.t) * 1e3
        self.t = now


def run_handoff(eng, plan, first_chunk=0, dist=None, oqpsk_half=None, phase_ms=None):
    """Time-sharding with state hand-off between chunks (Tier-S, same cost as the two-pass scheme, longer
    effective warm-up): pass A = warm-up W from power-on; pass B = owned + overlap, giving the quadrant
    scan AND, at its end, the state of every row V samples past its successor's boundary; pass C = every
    row continues from the state of its PREDECESSOR row (turned to the sequential run's lock point), so
    the symbols it contributes come from a trajectory with W + C + V samples of history instead of W.
    The row after chunk 0 inherits the exact sequential state and stays bit-exact as well. Across ranks
    the predecessor state of a rank's first row arrives from the previous rank (send/recv of one state:
    NCCL between GPUs); the quadrant scan's exchange is stitch()'s.

    eng: pass_a() -> head symbols of row 0, pass_b()/pass_c() -> (soft, q, count), export_rows() ->
    [M, R] tensor, import_rows(rows), rotate_rows(rows, turns); eng.M local rows.
    oqpsk_half: None for QPSK; for OQPSK the number of timing sub-steps in half a symbol (fs*interp/(2*symrate)):
    the quadrant scan is then boundary_quadrants_oqpsk and the engine's rotate_rows must move a row turned by an
    odd count as turn_oqpsk_state does. Checked with the CPU oracle as the engine (tests/test_sharded.py); the
    GPU engine against that emulation byte for byte (test_gpu_oqpsk_handoff_equals_oracle_handoff)."""
    import dataclasses
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    M = eng.M
    if first_chunk == 0 and M < 2 and plan.nchunks > 1:
        raise ValueError("the rank holding chunk 0 needs at least two chunks")
    ph = _Phases(phase_ms)
    head = eng.pass_a()
    ph.mark("pass_a_warm_up")
    soft_b, q_b, n_b, base_b = _rows4(eng.pass_b())
    ph.mark("pass_b_owned")
    scan = stitch(soft_b, q_b, n_b, plan, first_chunk=first_chunk, dist=dist, base=base_b, oqpsk_half=oqpsk_half)
    ph.mark("quadrant_scan")
    K = chunk_turns(scan, M)
    row0 = soft_b[0, : int(n_b[0].item())].clone() if first_chunk == 0 else None   # pass C reuses the buffers
    turned = eng.rotate_rows(eng.export_rows(), K)
    incoming = turned[:1].clone()                                   # placeholder for the row that has no predecessor
    if world > 1:
        # "boundary state handed rank to rank": one state (lrpt_state_t + delay line) down the chain
        _neighbour_exchange(dist, turned[-1:].contiguous() if first_chunk + M < plan.nchunks else None, incoming)
    if first_chunk == 0 and M == 1:                                 # the whole stream is chunk 0: already exact
        res = dict(scan)
        res["soft"] = torch.cat((_as_tensor(head, soft_b.device), row0))
        res["first_pass"] = dict(k=scan["k"], agreement=scan["agreement"], K=K)
        return res
    eng.import_rows(torch.cat((incoming, turned[:-1])))
    ph.mark("state_hand_off")
    soft_c, q_c, n_c, base_c = _rows4(eng.pass_c())
    ph.mark("pass_c_final")
    shifted = dataclasses.replace(plan, cut_shift=plan.overlap)
    if first_chunk == 0:
        # rows 0 and 1 are one exact trajectory: the head, row 0's pass-B symbols, then row 1's pass-C symbols up
        # to its cut with row 2 -- i.e. the final table is rows 1.. with nothing cut off the front of row 1
        res = stitch(soft_c[1:], q_c[1:], n_c[1:], shifted, first_chunk=1, dist=dist,
                     base=None if base_c is None else base_c[1:], oqpsk_half=oqpsk_half)
        res["soft"] = torch.cat((_as_tensor(head, soft_c.device), row0, res["soft"]))
    else:
        res = stitch(soft_c, q_c, n_c, shifted, first_chunk=first_chunk, dist=dist, base=base_c, oqpsk_half=oqpsk_half)
    ph.mark("join")
    res["first_pass"] = dict(k=scan["k"], agreement=scan["agreement"], K=K)
    return res


def _rows4(rows):
    """(soft, q, count[, base]) as an engine returns it -> always four items."""
    return tuple(rows) if len(rows) == 4 else tuple(rows) + (None,)


def _stitch_rows(rows, plan, **kw):
    soft, q, count, base = _rows4(rows)
    return stitch(soft, q, count, plan, base=base, **kw)


def _as_tensor(x, device):
    return x.to(device) if isinstance(x, torch.Tensor) else torch.from_numpy(x).to(device)


class ShardedDemod:
    """Persistent time-sharding engine for one stream buffer (plan, device buffers and the C-ABI handle
    are built once; run() can be called repeatedly, e.g. by bench.py)."""

    def __init__(self, raw, nsamples, chunk=1 << 21, warm=1 << 19, overlap=8192, device=0, dist=None,
                 interp_factor=5, two_pass=True, handoff=False, raw_first=0, seed_carrier=False, seed_nfft=1 << 17, **cfg):
        self.oqpsk_half = None
        if cfg.get("oqpsk"):
            # OQPSK: the join tells even from odd quarter turns by the half-symbol timing offset and the hand-off also
            # moves the timing NCO (boundary_quadrants_oqpsk, turn_oqpsk_state). Only the hand-off scheme knows how;
            # the two-pass scheme turns the Costas NCO alone, which is the QPSK ambiguity.
            if not handoff:
                raise NotImplementedError("OQPSK time shards need the hand-off scheme (handoff=True)")
            self.oqpsk_half = cfg.get("samplerate", 230000) * interp_factor / (2.0 * cfg.get("symrate", 72000))
        self.plan = plan = Plan(nsamples, chunk, warm, overlap, interp_factor)
        self.dist, self.two_pass, self.handoff = dist, two_pass, handoff
        world = dist.get_world_size() if dist is not None else 1
        rank = dist.get_rank() if dist is not None else 0
        check_layout(plan.nchunks, world, handoff)              # same verdict on every rank, before any collective
        self.c0, self.c1 = split_chunks(plan.nchunks, world, rank)
        need = (plan.start(self.c1 - 1) + plan.n_main + 2 * plan.overlap - raw_first) * 2
        if raw_first > plan.start(self.c0) or raw.numel() < need:
            raise ValueError("raw must hold samples [%d, %d) for chunks %d..%d" % (plan.start(self.c0), raw_first + need // 2, self.c0, self.c1 - 1))
        self.eng = GpuEngine(raw, plan, device=device, first_chunk=self.c0, nchunks=self.c1 - self.c0,
                             raw_first=raw_first, seed_carrier=seed_carrier, seed_nfft=seed_nfft, **cfg)

    def run(self, phase_ms=None):
        """phase_ms: optional dict that receives the wall time of each phase in ms (hand-off scheme)."""
        eng, plan, c0, M = self.eng, self.plan, self.c0, self.c1 - self.c0
        l0 = eng.d.launch_count()
        if self.handoff:
            res = run_handoff(eng, plan, first_chunk=c0, dist=self.dist, oqpsk_half=self.oqpsk_half, phase_ms=phase_ms)
        elif not self.two_pass:
            res = _stitch_rows(eng.run(), plan, first_chunk=c0, dist=self.dist)
        else:
            head = eng.warm_up()
            scan = _stitch_rows(eng.owned(), plan, first_chunk=c0, dist=self.dist)
            K = chunk_turns(scan, M)
            res = _stitch_rows(eng.owned(K), plan, first_chunk=c0, dist=self.dist)
            res["first_pass"] = dict(k=scan["k"], agreement=scan["agreement"], K=K)
            if c0 == 0:
                res["soft"] = torch.cat((head, res["soft"]))
        res.update(plan=plan, first_chunk=c0, nchunks_local=M, launches=eng.d.launch_count() - l0)
        return res

    def close(self):
        self.eng.close()


def demod_sharded(raw, nsamples, **kw):
    """One long stream, time-sharded over this process's GPU and, with dist=torch.distributed
    (initialised), over ranks: rank r takes a consecutive run of chunks (split_chunks) and needs only the
    samples those read -- its time slice plus warm-up and overlap, which it over-reads instead of
    communicating samples. raw: 1-D device tensor of the raw dtype starting at stream sample `raw_first`
    (default 0: the whole stream, at least 2*Plan.padded items, zeros after 2*nsamples).

    two_pass=False: one launch; chunks keep whatever lock point they acquired and are de-rotated after
    the fact. Chunks that locked an odd number of quarter turns away see the OTHER bit stream on the Q
    arm, the only input of the timing detector (timing.c:65), and come out with a different timing
    jitter: about 15 % of their symbols differ from the sequential run by more than 1 LSB.
    two_pass=True (default): pass A warm-up -> snapshot; pass B owned+overlap -> quadrant scan; pass C
    again from the snapshot with every Costas NCO turned back by its K_c quarter turns, so all chunks
    run at the sequential run's lock point: eps drops to the 0.2-0.4 % of the reference's own
    FMA-vs-strict builds, at the price of demodulating the owned samples twice.
    Keyword arguments: chunk, warm, overlap, device, dist, interp_factor, two_pass + Demod's (symrate,
    bps, rrc_order, ...). Returns the stitch() dict plus plan, first_chunk, nchunks_local, launches."""
    sd = ShardedDemod(raw, nsamples, **kw)
    try:
        return sd.run()
    finally:
        sd.close()


def process_host(raw, chunk=1 << 18, warm=150000, overlap=8192, out=None, devices=None, seed_nfft=0, **cfg):
    """ONE recording in host memory, time-sharded on one GPU by a single C call (lrpt_sharded_process,
    csrc/shard_run.cu -- the hand-off scheme above without Python in the loop; what host/lrpt_demod --shard
    uses). raw: numpy array of interleaved I,Q in the configured sample format. cfg: make_params' keywords
    (symrate, bps, rrc_order, interp_factor, ...). out: optional int8 array [cap, 2] to receive the symbols
    (page-locked memory makes the final copy 15x faster than a fresh pageable array). devices: list of CUDA
    ordinals -> lrpt_sharded_process_multi (one host thread per GPU, boundary state and overlap symbols over NCCL;
    byte-identical to the one-GPU call). seed_nfft: 0, or the transform length of the coarse carrier estimate chunks
    after the first start from (lrpt_shard_plan_t.seed_nfft). Returns (soft [n,2] int8 -- a view of `out` when given --,
    report dict)."""
    from ._lib import LrptError, ShardPlan, ShardReport, load
    from .demod import make_params, symbol_capacity
    p = make_params(**cfg)
    a = np.ascontiguousarray(raw).reshape(-1)
    n = a.size // 2
    cap = symbol_capacity(n, p.samplerate, p.symrate)
    if out is None:
        soft = np.empty((cap, 2), np.int8)
    else:
        soft = out
        if soft.dtype != np.int8 or not soft.flags["C_CONTIGUOUS"] or soft.ndim != 2 or soft.shape[1] != 2:
            raise ValueError("out must be a C-contiguous int8 array [cap, 2]")
        cap = soft.shape[0]
    nsym, rep, plan = C.c_size_t(0), ShardReport(), ShardPlan(chunk, warm, overlap, seed_nfft)
    if devices:
        devs = (C.c_int * len(devices))(*devices)
        rc = load().lrpt_sharded_process_multi(C.byref(p), C.byref(plan), a.ctypes.data, n, soft.ctypes.data, cap,
                                               C.byref(nsym), C.byref(rep), devs, len(devices))
    else:
        rc = load().lrpt_sharded_process(C.byref(p), C.byref(plan), a.ctypes.data, n, soft.ctypes.data, cap,
                                         C.byref(nsym), C.byref(rep))
    if rc:
        raise LrptError(rc, "lrpt_sharded_process")
    return soft[: nsym.value], {f: getattr(rep, f) for f, _ in ShardReport._fields_ if f != "reserved"}
