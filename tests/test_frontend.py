"""the LRPT decoder front-end behind the demodulator's soft-symbol stream -- frame synchronisation
and the CCSDS r = 1/2, K = 7 Viterbi decoder. CPU: the oracle (oracle/frontend_oracle.c) against the link layer's
known answers and its own transmitter; GPU: csrc/frontend.cu against the oracle, bit for bit, and end to end behind
the demodulator."""
import numpy as np
import pytest

KNOWN_SYNC = {0xFCA2B63DB00D9794, 0x56FBD394DAA4C1C2, 0x035D49C24FF2686B, 0xA9042C6B255B3E3D}


def test_encoded_sync_word_known_answers(oracle_mod):
    """The attached sync marker 0x1ACFFC1D through the rate-1/2 K=7 code (G1 = 171, G2 = 133, from the zero state) is
    0x035D49C24FF2686B; its four quarter-turn images are the constants every LRPT decoder correlates against."""
    from oracle import pyfrontend as fe
    bits, state = fe.conv_encode(fe.ASM)
    v = 0
    for b in bits.tolist():
        v = (v << 1) | b
    assert v == 0x035D49C24FF2686B and state == (0x1D & 0x7F)
    assert fe.sync_pattern(0) == v
    assert {fe.sync_pattern(h) for h in range(4)} == KNOWN_SYNC
    assert len({fe.sync_pattern(h) for h in range(8)}) == 8
    # impulse response of the code = its generator polynomials: 171 octal (1111001) on the I arm, 133 octal (1011011) on Q
    imp, _ = fe.conv_encode(bytes([0x80, 0x00]))
    assert imp[0:14:2].tolist() == [1, 1, 1, 1, 0, 0, 1] and imp[1:14:2].tolist() == [1, 0, 1, 1, 0, 1, 1]


def frames_and_stream(n=4, lead=1234, noise=30.0, turns=0, swap=False, seed=1):
    from oracle import pyfrontend as fe
    rng = np.random.default_rng(seed)
    frames = rng.integers(0, 256, (n, 1020), dtype=np.uint8)
    return frames, fe.transmit(frames, noise=noise, turns=turns, swap=swap, lead=lead, seed=seed + 1)


@pytest.mark.parametrize("turns,swap", [(0, False), (1, False), (2, False), (3, False), (0, True), (1, True), (2, True), (3, True)])
def test_oracle_round_trip_under_every_symmetry(turns, swap, oracle_mod):
    """encode -> 8 symmetries of the constellation + noise -> synchronise -> decode: the payload comes back, the frame
    starts are found at their offsets with the symmetry that was applied."""
    from oracle import pyfrontend as fe
    frames, soft = frames_and_stream(turns=turns, swap=swap)
    score, hyp = fe.sync_scores(soft)
    off, oh, osc = fe.window_peaks(score, hyp)
    want_h = turns + (4 if swap else 0)
    for f in range(frames.shape[0]):
        start = 1234 + f * fe.CADU_SYMS
        assert score[start] >= 50 and hyp[start] == want_h
        cadu, metric = fe.viterbi_cadu(soft, start, hyp[start])
        assert bytes(cadu[:4]) == fe.ASM and np.array_equal(cadu[4:], frames[f]) and metric > 0
    assert off[0] == 1234 and oh[0] == want_h and off[1] == 1234 + fe.CADU_SYMS
    # a wrong symmetry does not decode
    bad, _ = fe.viterbi_cadu(soft, 1234, (want_h + 1) % 8)
    assert bytes(bad[:4]) != fe.ASM


@pytest.mark.gpu
def test_gpu_sync_equals_oracle(lib, oracle_mod):
    import torch
    from meteor_demod_b200 import frontend
    from oracle import pyfrontend as fe
    rng = np.random.default_rng(7)
    for n, turns, swap in ((3, 1, False), (2, 3, True)):
        frames, soft = frames_and_stream(n=n, lead=4321, noise=45.0, turns=turns, swap=swap, seed=10 + n)
        for cut in (soft.shape[0], soft.shape[0] - 5, 8192 + 37, 40, 32):      # ragged ends, tiny streams
            s = np.ascontiguousarray(soft[:cut])
            want_s, want_h = fe.sync_scores(s)
            got_s, got_h = frontend.sync_scores(torch.from_numpy(s).cuda())
            assert np.array_equal(got_s.cpu().numpy(), want_s) and np.array_equal(got_h.cpu().numpy(), want_h), cut
            w_off, w_h, w_sc = fe.window_peaks(want_s, want_h)
            g_off, g_h, g_sc = frontend.window_peaks(got_s, got_h)
            assert np.array_equal(g_off.cpu().numpy().astype(np.uint32), w_off)
            assert np.array_equal(g_h.cpu().numpy(), w_h) and np.array_equal(g_sc.cpu().numpy(), w_sc)
    noise = rng.integers(-128, 128, (100_003, 2), dtype=np.int8)                  # no frames at all
    want_s, want_h = fe.sync_scores(noise)
    got_s, got_h = frontend.sync_scores(torch.from_numpy(noise).cuda())
    assert np.array_equal(got_s.cpu().numpy(), want_s) and np.array_equal(got_h.cpu().numpy(), want_h)


@pytest.mark.gpu
def test_gpu_viterbi_equals_oracle(lib, oracle_mod):
    """Frames at their true offsets, at wrong offsets, at the edges of the stream, under right and wrong symmetries,
    at a noise level where the decoder makes errors: bytes and path metric equal the oracle's everywhere."""
    import torch
    from meteor_demod_b200 import frontend
    from oracle import pyfrontend as fe
    frames, soft = frames_and_stream(n=5, lead=777, noise=52.0, turns=2, swap=True, seed=30)
    nsym = soft.shape[0]
    offs = [777 + f * fe.CADU_SYMS for f in range(5)] + [0, 5, 700, 777 + 3, nsym - 8192, nsym - 4000, nsym - 10]
    hyps = [6] * 5 + [6, 0, 3, 6, 6, 1, 7]
    d_soft = torch.from_numpy(soft).cuda()
    vit = frontend.Viterbi()
    cadu, metric = vit.decode(d_soft, torch.tensor(offs, dtype=torch.int32, device="cuda"), torch.tensor(hyps, dtype=torch.uint8, device="cuda"))
    cadu, metric = cadu.cpu().numpy(), metric.cpu().numpy()
    errors = 0
    for k, (o, h) in enumerate(zip(offs, hyps)):
        want, wm = fe.viterbi_cadu(soft, o, h)
        assert np.array_equal(cadu[k], want), (k, o, h)
        assert int(metric[k]) == wm, (k, o, h)
        if k < 5:
            errors += int(np.unpackbits(want[4:] ^ frames[k]).sum())
    assert 0 < errors < 4000                                   # the channel is bad enough to exercise ties and wrong paths
    # full-scale input (every soft value -128 or 127, no code structure): the widest metric spread and the fastest
    # drift the packed 16-bit metrics of the kernel have to hold
    ext = np.random.default_rng(31).choice(np.array([-128, 127], dtype=np.int8), size=(3 * fe.CADU_SYMS, 2))
    eo, eh = [0, 100, 4096, 8192 + 31, 2 * 8192], [0, 1, 2, 7, 5]
    c2, m2 = vit.decode(torch.from_numpy(ext).cuda(), torch.tensor(eo, dtype=torch.int32, device="cuda"),
                        torch.tensor(eh, dtype=torch.uint8, device="cuda"))
    for k, (o, h) in enumerate(zip(eo, eh)):
        want, wm = fe.viterbi_cadu(ext, o, h)
        assert np.array_equal(c2[k].cpu().numpy(), want) and int(m2[k]) == wm, (k, o, h)
    # many frames: more frames than resident warps, decoded in several rounds
    reps = torch.tensor(offs[:5] * 2400, dtype=torch.int32, device="cuda")
    many, _ = vit.decode(d_soft, reps, torch.full((12000,), 6, dtype=torch.uint8, device="cuda"))
    assert torch.equal(many.view(2400, 5, 1024), many[:5].unsqueeze(0).expand(2400, 5, 1024))


@pytest.mark.gpu
def test_frames_survive_the_whole_chain(lib, oracle_mod):
    """CADUs -> convolutional code -> QPSK at 72 ksym/s -> RRC pulse, carrier offset, noise, 16-bit samples ->
    the demodulator (GPU) -> frame synchronisation + Viterbi (GPU): the payloads come back."""
    import torch
    from meteor_demod_b200 import Demod, frontend, synth
    from oracle import pyfrontend as fe
    rng = np.random.default_rng(3)
    nfr = 9
    frames = rng.integers(0, 256, (nfr, 1020), dtype=np.uint8)
    bits, _ = fe.conv_encode(b"".join(fe.ASM + f.tobytes() for f in frames))
    sym = (2.0 * bits.astype(np.float64) - 1.0).reshape(-1, 2)
    pre = rng.choice([-1.0, 1.0], size=(20000, 2))             # the loops acquire on these
    sym = np.concatenate([pre, sym, rng.choice([-1.0, 1.0], size=(600, 2))])
    raw = synth.modulate(sym[:, 0] + 1j * sym[:, 1], cfo_hz=35.0, esn0_db=11.0, seed=9)
    d = Demod(nstreams=1)
    soft, _ = d.process(raw)
    d_soft = torch.from_numpy(np.ascontiguousarray(soft)).cuda()
    score, hyp = frontend.sync_scores(d_soft)
    off, oh, osc = frontend.window_peaks(score, hyp)
    off, oh, osc = off.cpu().numpy(), oh.cpu().numpy(), osc.cpu().numpy()
    good = osc >= 53                                          # noise alone reaches ~50 of 64 somewhere in 8192 offsets
    assert good.sum() >= nfr - 1
    starts = off[good]
    assert np.all(np.diff(starts) % fe.CADU_SYMS == 0) and len(set(oh[good].tolist())) == 1     # one lock point, frames back to back
    cadu, _ = frontend.Viterbi().decode(d_soft, torch.from_numpy(starts.astype(np.int32)).cuda(), torch.from_numpy(oh[good]).cuda())
    cadu = cadu.cpu().numpy()
    payloads = {f.tobytes() for f in frames}
    hits = sum(1 for c in cadu if bytes(c[:4]) == fe.ASM and c[4:].tobytes() in payloads)
    assert hits >= nfr - 1
