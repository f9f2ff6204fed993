"""Generate the committed golden vectors by RUNNING THE REFERENCE ITSELF.

Run in the CPU container (where /root/reference exists):
    make -C oracle ref && python tests/golden/make_golden.py

For each case: a short synthetic raw I/Q input (committed, so the test does not depend
on numpy/scipy float reproducibility) pushed through oracle/_ref/libref_strict.so -- the
unmodified reference sources compiled with -ffp-contract=off -- recording every float
symbol, int8 soft symbol, producing sample index, lock flag and the final loop state.
tests/test_golden.py checks our oracle port against these files on any machine;
tests/test_gpu_parity.py checks the CUDA path against them on the B200.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from meteor_demod_b200 import synth  # noqa: E402
from oracle.pyoracle import Ref  # noqa: E402

CASES = {
    # name: (config, nsamples, carrier offset Hz)  -- small offset so the PLL locks inside the vector
    "qpsk72k_s16_o32_L5": (dict(symrate=72000, oqpsk=0, bps=16, order=32, interp=5), 65536, 40.0),
    "oqpsk80k_u8_o32_L5": (dict(symrate=80000, oqpsk=1, bps=8, order=32, interp=5), 49152, 60.0),
    "qpsk72k_f32_o16_L3": (dict(symrate=72000, oqpsk=0, bps=32, order=16, interp=3), 24576, -30.0),
    "qpsk72k_s16_o64_L8": (dict(symrate=72000, oqpsk=0, bps=16, order=64, interp=8), 32768, 25.0),
}


def main():
    for name, (cfg, n, cfo) in CASES.items():
        raw = synth.make_raw(n, symrate=cfg["symrate"], oqpsk=bool(cfg["oqpsk"]), bps=cfg["bps"], seed=11,
                             cfo_hz=cfo)
        r = Ref(kind="strict", **cfg)
        out = r.process(raw)
        st = r.state()
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), raw=raw, sym_bits=out.sym.view(np.uint32), soft=out.soft,
            sample_idx=out.sample_idx.astype(np.int32), lock_once=out.lock_once,
            taps_bits=r.taps().view(np.uint32), history_bits=r.history().view(np.uint32),
            state_names=np.array(sorted(st)), state_bits=np.array(
                [np.float32(st[k]).view(np.uint32) if isinstance(st[k], float) else np.uint32(st[k] & 0xffffffff)
                 for k in sorted(st)], np.uint32),
            cfg_names=np.array(sorted(cfg)), cfg_vals=np.array([cfg[k] for k in sorted(cfg)], np.int64))
        print(name, "nsym", out.nsym, "locked_once", int(out.lock_once.max()) if out.nsym else 0,
              "first lock", int(np.argmax(out.lock_once)) if out.lock_once.any() else -1)


if __name__ == "__main__":
    main()
