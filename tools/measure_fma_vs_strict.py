"""The yardstick for Tier-S: how far the reference's OWN two legitimate builds are from each other.

oracle/_ref/libref_fma.so is the reference compiled with its release flags (FMA contraction, what users run),
oracle/_ref/libref_strict.so the strict-IEEE meaning of the same sources (the bit-exact target of the CUDA path).
Both demodulate the same recordings from power-on state; reported per configuration: the share of soft symbols
that differ by more than one int8 LSB (after the quarter turn between the two builds' lock points, which is not
always zero), the share that are identical, and the symbol counts. Decision chaos
(SURVEY.md finding 3) makes this a floor for ANY implementation that is not bit-exact, time-sharded ones included.

  python tools/measure_fma_vs_strict.py [nsamples]   ->  profiles/r2_fma_vs_strict_eps.json   (CPU only)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from meteor_demod_b200 import synth  # noqa: E402
from oracle import pyoracle  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 13_795_328          # 60 s at 230 kS/s, SURVEY.md appendix A.1
CASES = {
    "c1_qpsk72k_s16": dict(symrate=72000, oqpsk=0, bps=16, order=32, interp=5),
    "c2_oqpsk80k_u8": dict(symrate=80000, oqpsk=1, bps=8, order=32, interp=5),
    "c3_qpsk72k_s16_o64_L8": dict(symrate=72000, oqpsk=0, bps=16, order=64, interp=8),
}
out = {"nsamples": N, "signal": "synth.make_raw seed 1 (+700 Hz, Es/N0 12 dB); bench-like stream: tileable period, seed 11",
       "how": "tools/measure_fma_vs_strict.py: libref_fma.so vs libref_strict.so on identical input, all symbols after the first lock"}
for name, cfg in CASES.items():
    rows = {}
    per = synth.baseband(230000, symrate=cfg["symrate"], oqpsk=bool(cfg["oqpsk"]), periodic=True, seed=3).astype(np.complex64)
    signals = {
        "make_raw": synth.make_raw(N, symrate=cfg["symrate"], oqpsk=bool(cfg["oqpsk"]), bps=cfg["bps"], seed=1),
        "bench_stream": synth.device_long_stream(per, N, bps=cfg["bps"], sps=230000 / cfg["symrate"], device="cpu").numpy(),
    }
    for sname, raw in signals.items():
        a = pyoracle.Ref(kind="strict", **cfg).process(raw, want_float=False)
        b = pyoracle.Ref(kind="fma", **cfg).process(raw, want_float=False)
        n = min(a.nsym, b.nsym)
        first = int(np.argmax(a.lock_once[:n])) if a.lock_once[:n].any() else 0
        x, y = a.soft[first:n].astype(np.int16), b.soft[first:n].astype(np.int16)
        # the two builds may even lock at DIFFERENT quadrants (a Costas loop has four lock points, pll.c:143-152):
        # compare under the quarter turn k (fma -> strict) that agrees best, and report k
        turns = [(y[:, 0], y[:, 1]), (-y[:, 1], y[:, 0]), (-y[:, 0], -y[:, 1]), (y[:, 1], -y[:, 0])]
        k = int(np.argmax([float(((np.sign(x[:, 0]) == np.sign(t[0])) & (np.sign(x[:, 1]) == np.sign(t[1]))).mean()) for t in turns]))
        d = np.maximum(np.abs(x[:, 0] - turns[k][0]), np.abs(x[:, 1] - turns[k][1]))
        d0 = np.abs(x - y).max(axis=1)
        rows[sname] = {"symbols_strict": a.nsym, "symbols_fma": b.nsym, "compared_after_lock": int(d.size),
                       "quarter_turns_between_builds": k,
                       "frac_gt_1lsb": float((d > 1).mean()), "frac_identical": float((d == 0).mean()),
                       "max_abs_diff_lsb": int(d.max()), "frac_gt_1lsb_without_turn": float((d0 > 1).mean())}
        print(name, sname, rows[sname], flush=True)
    out[name] = rows
path = os.path.join(ROOT, "profiles", "r2_fma_vs_strict_eps.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path)
