"""BASELINE config 5: RRC order {16,32,64,128} x oversampling {3,5,8}, QPSK vs OQPSK.
Device-resident throughput (CUDA events) of the exact batch path at 75776 streams (AUTO kernel = lane),
with stream 0 of every configuration checked bit for bit against the CPU oracle. Writes a CSV to stdout.
hbm_gbs = algorithmic bytes (raw in + int8 symbols out) / time; fir_gflop_s counts the reference's own lazy
FIR (4*taps flops per filter_get call), the work the lane kernel actually does."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meteor_demod_b200 import Demod, synth
from oracle import pyoracle

B = int(sys.argv[1]) if len(sys.argv) > 1 else 75776
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 14
KERNEL = sys.argv[3] if len(sys.argv) > 3 else "auto"
periods = {0: synth.baseband(230000, symrate=72000, periodic=True, seed=3).astype(np.complex64),
           1: synth.baseband(230000, symrate=80000, oqpsk=True, periodic=True, seed=3).astype(np.complex64)}
print("mode,symrate,bps,order,taps,interp,streams,samples_per_stream,ms,msps,hbm_gbs,hbm_frac_of_6451.8,fir_gflop_s,oracle_stream0_equal,kernel")
peak = 6451.8
for oq, symrate, bps in ((0, 72000, 16), (1, 80000, 8)):
    raw = synth.device_streams(periods[oq], B, N, bps=bps, sps=230000 / symrate, seed=5)
    for order in (16, 32, 64, 128):
        for L in (3, 5, 8):
            d = Demod(symrate=symrate, oqpsk=oq, bps=bps, rrc_order=order, interp_factor=L, nstreams=B, kernel=KERNEL)
            cap = (d.capacity(N) + 7) // 8 * 8
            soft = torch.empty((B, 2 * cap), dtype=torch.int8, device="cuda")
            st = torch.cuda.Stream()
            best = 1e9
            for _ in range(3):
                d.reset(stream=st)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st); d.process_device(raw, soft, stream=st); e1.record(st); st.synchronize()
                best = min(best, e0.elapsed_time(e1))
            n0 = int(d.counts()[0])
            w = pyoracle.Oracle(symrate=symrate, oqpsk=oq, bps=bps, order=order, interp=L).process(raw[0].cpu().numpy(), want_float=False)
            ok = (w.nsym == n0) and np.array_equal(soft[0, : 2 * n0].cpu().numpy().reshape(-1, 2), w.soft)
            msps = B * N / best / 1e3
            nsym_total = float(d.counts().astype(np.int64).sum())
            gbs = (B * N * (bps // 4) + 2.0 * nsym_total) / (best * 1e-3) / 1e9
            fir = nsym_total * (2 if oq else 1) * 4 * (2 * order + 1) / (best * 1e-3) / 1e9
            print("%s,%d,%d,%d,%d,%d,%d,%d,%.2f,%.0f,%.1f,%.4f,%.0f,%s,%s" % ("oqpsk" if oq else "qpsk", symrate, bps, order, 2 * order + 1, L, B, N,
                  best, msps, gbs, gbs / peak, fir, ok, d.kernel_name()), flush=True)
            d.close()
