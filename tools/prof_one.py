"""One launch of the device-resident batch path, for ncu. args: B N [order L oqpsk kernel bps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meteor_demod_b200 import Demod, synth
B, N = int(sys.argv[1]), int(sys.argv[2])
order = int(sys.argv[3]) if len(sys.argv) > 3 else 32
L = int(sys.argv[4]) if len(sys.argv) > 4 else 5
oq = int(sys.argv[5]) if len(sys.argv) > 5 else 0
kern = sys.argv[6] if len(sys.argv) > 6 else "auto"
bps = int(sys.argv[7]) if len(sys.argv) > 7 else 16
symrate = 80000 if oq else 72000
per = synth.baseband(230000, symrate=symrate, oqpsk=bool(oq), periodic=True).astype(np.complex64)
d = Demod(symrate=symrate, oqpsk=oq, bps=bps, rrc_order=order, interp_factor=L, nstreams=B, kernel=kern)
raw = synth.device_streams(per, B, N, bps=bps, sps=230000/symrate)
cap = (d.capacity(N) + 7)//8*8
soft = torch.empty((B, 2*cap), dtype=torch.int8, device="cuda")
for _ in range(2):
    d.reset(); d.process_device(raw, soft); d.sync()
print("done", d.counts()[:4], d.kernel_name())
