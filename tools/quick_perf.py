"""Scratch timing of the device-resident batch path (not the bench; see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meteor_demod_b200 import Demod, synth

def run(B, N, kernel="ws", order=32, L=5, oqpsk=0, symrate=72000, bps=16, reps=3, period=None):
    d = Demod(symrate=symrate, oqpsk=oqpsk, bps=bps, rrc_order=order, interp_factor=L, nstreams=B, kernel=kernel)
    raw = synth.device_streams(period, B, N, bps=bps, sps=230000/symrate)
    cap = (d.capacity(N) + 7)//8*8
    soft = torch.empty((B, 2*cap), dtype=torch.int8, device="cuda")
    st = torch.cuda.Stream()
    ts = []
    with torch.cuda.stream(st):
        for r in range(reps):
            d.reset(stream=st)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); d.process_device(raw, soft, stream=st); e1.record(st); st.synchronize()
            ts.append(e0.elapsed_time(e1))
    c = d.counts()
    ms = min(ts)
    fb = d.fir_fallbacks() / max(1, reps)
    print("kernel=%s B=%d N=%d order=%d L=%d oqpsk=%d: %.2f ms -> %.1f MS/s (%.2f MS/s/stream) nsym[0]=%d locked=%d fir_fallbacks/run=%.0f (%.3f%% of symbols)" % (
        kernel, B, N, order, L, oqpsk, ms, B*N/ms/1e3, N/ms/1e3, c[0], d.status(0)["locked"], fb, 100.0*fb/max(1, int(c.sum()))), flush=True)
    d.close()

if __name__ == "__main__":
    t = time.time()
    per = synth.baseband(230000, periodic=True).astype(np.complex64)
    print("period gen %.1fs" % (time.time()-t), flush=True)
    import sys as _s
    kernels = [k for k in ("ws", "spec", "simple", "lane") if ("--" + k) in _s.argv] or ["ws"]
    cfgs = ((1, 1<<18), (2048, 1<<18), (4736, 1<<18))
    for a_ in _s.argv:
        if a_.startswith("--cfg="):
            cfgs = tuple(tuple(int(v) for v in c_.split(":")) for c_ in a_[6:].split(","))
    for kern in kernels:
        for B, N in cfgs:
            run(B, N, kern, period=per)
        if "--all" in _s.argv:
            run(4736, 1<<17, kern, order=64, L=8, period=per)
            per80 = synth.baseband(230000, symrate=80000, oqpsk=True, periodic=True).astype(np.complex64)
            run(4736, 1<<17, kern, oqpsk=1, symrate=80000, bps=8, period=per80)
