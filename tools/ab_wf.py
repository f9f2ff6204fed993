"""A/B of the lane kernel's window format (raw-typed vs float pairs) on the device-resident batch path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meteor_demod_b200 import Demod, synth

def run(raw, B, N, env, order=32, L=5, oqpsk=0, symrate=72000, bps=16, reps=3):
    for k in ("LRPT_LANE_WF", "LRPT_LANE_NT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    d = Demod(symrate=symrate, oqpsk=oqpsk, bps=bps, rrc_order=order, interp_factor=L, nstreams=B, kernel="lane")
    cap = (d.capacity(N) + 7)//8*8
    soft = torch.zeros((B, 2*cap), dtype=torch.int8, device="cuda")
    st = torch.cuda.Stream()
    ts = []
    with torch.cuda.stream(st):
        for r in range(reps):
            d.reset(stream=st)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); d.process_device(raw[:B], soft, stream=st); e1.record(st); st.synchronize()
            ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    c = d.counts()
    print("env=%s B=%d N=%d order=%d L=%d oq=%d bps=%d: %.2f ms -> %.1f MS/s" % (env, B, N, order, L, oqpsk, bps, ms, B*N/ms/1e3), flush=True)
    d.close()
    return soft, c

if __name__ == "__main__":  # noqa
    os.environ["LRPT_LANE_DEBUG"] = "1"
    per = synth.baseband(230000, periodic=True).astype(np.complex64)
    N = 32768
    raw = synth.device_streams(per, 75776, N, bps=16)
    ref, cref = run(raw, 75776, N, {})
    for B, env in ((75776, {"LRPT_LANE_WF": "1"}), (75776, {"LRPT_LANE_WF": "1", "LRPT_LANE_NT": "1"})):
        s, c = run(raw, B, N, env)
        print("   identical to raw-window output:", bool(torch.equal(s, ref[:B])) and bool((c == cref[:B]).all()), flush=True)
    for B, env in ((37888, {}), (37888, {"LRPT_LANE_WF": "1"}), (42624, {"LRPT_LANE_WF": "1"}), (42624, {}),
                   (47360, {"LRPT_LANE_WF": "1", "LRPT_LANE_NT": "1"}), (52096, {"LRPT_LANE_WF": "1", "LRPT_LANE_NT": "1"}),
                   (52096, {}), (18944, {}), (18944, {"LRPT_LANE_WF": "1"}), (4736, {}), (4736, {"LRPT_LANE_WF": "1"})):
        run(raw, B, N, env)
    del raw, ref
    # C3: RRC-64 x8
    raw = synth.device_streams(per, 52096, N, bps=16)
    a, ca = run(raw, 52096, N, {}, order=64, L=8)
    b, cb = run(raw, 52096, N, {"LRPT_LANE_WF": "1"}, order=64, L=8)
    print("   C3 identical:", bool(torch.equal(a, b)), flush=True)
    run(raw, 23680, N, {"LRPT_LANE_WF": "1"}, order=64, L=8)
    run(raw, 28416, N, {"LRPT_LANE_WF": "1", "LRPT_LANE_NT": "1"}, order=64, L=8)
    del raw, a, b
    # C2: OQPSK u8
    per80 = synth.baseband(230000, symrate=80000, oqpsk=True, periodic=True).astype(np.complex64)
    raw = synth.device_streams(per80, 75776, N, bps=8, sps=230000/80000)
    a, ca = run(raw, 75776, N, {}, oqpsk=1, symrate=80000, bps=8)
    b, cb = run(raw, 75776, N, {"LRPT_LANE_WF": "1"}, oqpsk=1, symrate=80000, bps=8)
    print("   C2 identical:", bool(torch.equal(a, b)), flush=True)
    run(raw, 42624, N, {"LRPT_LANE_WF": "1"}, oqpsk=1, symrate=80000, bps=8)
