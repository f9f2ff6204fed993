"""Which loop state makes a cold-started chunk differ from the sequential run? (CPU oracle only.)

Sequential run A over a C1 stream with a 700 Hz carrier offset; run B starts from power-on state W samples before
sample s0 and is turned to A's lock point at s0. Printed per W: B's state error at s0 and the share of symbols more
than one LSB from A's in consecutive windows of 20 k symbols after s0 -- cold, and with single fields of B's state
replaced by A's at s0. Result (profiles/r2_warmup_state_convergence.txt): the Costas loop's frequency integrator
p_freq is the slow state (error e-folds every ~52 k samples once locked; getting to the carrier in the first place
takes the reference's 1e-6 rad/symbol^2 sweep, pll.c:126); timing and AGC states are irrelevant after 30 k samples.
That is why time-sharded chunks start from a coarse carrier estimate (acquire.py) and inherit their predecessor's
end state (sharded.run_handoff)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle  # noqa: E402
from meteor_demod_b200 import synth  # noqa: E402

N, s0 = 1_400_000, 800_000
raw = synth.make_raw(N, bps=16, seed=21, cfo_hz=700.0)
A = pyoracle.Oracle()
A.process(raw[: 2 * s0], want_float=False)
stA = A.state()
wA = A.process(raw[2 * s0:], want_float=False)
print("sequential state at s0:", {k: stA[k] for k in ("t_phase", "t_freq", "agc_gain", "p_phase", "p_freq", "p_err", "p_locked")})


def eps(wB):
    n = min(wB.nsym, wA.nsym)
    d = np.abs(wB.soft[:n].astype(np.int16) - wA.soft[:n].astype(np.int16)).max(axis=1)
    return ["%.3f" % float((d[i: i + 20000] > 1).mean()) for i in range(0, n - 20000, 20000)][:8]


def warmed(W):
    B = pyoracle.Oracle()
    B.process(raw[2 * (s0 - W): 2 * s0], want_float=False)
    st = B.state()
    kq = int(np.round((stA["p_phase"] - st["p_phase"]) / (np.pi / 2)))
    B.set_state(p_phase=float(np.float32(st["p_phase"] + kq * np.pi / 2)))
    return B, st


for W in (30000, 60000, 100000, 150000, 300000):
    B, st = warmed(W)
    print("W = %6d  state error at s0:" % W, {k: float("%.3g" % (st[k] - stA[k])) for k in ("t_phase", "t_freq", "agc_gain", "p_freq")})
    print("    cold                  ", eps(B.process(raw[2 * s0:], want_float=False)))
    for name, fields in (("t_freq, t_phase", ("t_freq", "t_phase")), ("agc_gain", ("agc_gain",)), ("p_freq", ("p_freq",))):
        B, _ = warmed(W)
        B.set_state(**{f: stA[f] for f in fields})
        print("    %-22s" % (name + " from A"), eps(B.process(raw[2 * s0:], want_float=False)))
