"""Debug: where does one long push differ from two half pushes?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meteor_demod_b200 import Demod, synth

n = (1 << 26) + 300_000
per = synth.baseband(230000, periodic=True, seed=3).astype(np.complex64)
raw = synth.device_long_stream(per, n, cfo_hz=90.0).view(1, -1)

def run(cuts):
    d = Demod(nstreams=1)
    cap = (d.capacity(n) + 7) // 8 * 8
    outs = []
    pos = 0
    for c in cuts + [n]:
        s = torch.zeros((1, 2 * cap), dtype=torch.int8, device="cuda")
        d.process_device(raw[:, 2 * pos: 2 * c].contiguous(), s)
        d.sync()
        k = int(d.counts()[0])
        outs.append(s[0, : 2 * k].clone())
        pos = c
    st = d.export_state()
    d.close()
    return torch.cat(outs), st

a, sa = run([])
b, sb = run([])
print("deterministic:", a.numel() == b.numel() and torch.equal(a, b), sa == sb, a.numel() // 2)
half = (n // 2) // 8 * 8
for cuts in ([half], [1 << 26], [half, 1 << 26], [1000000]):
    c, sc = run(cuts)
    m = min(a.numel(), c.numel())
    ne = (a[:m] != c[:m]).nonzero()
    first = int(ne[0]) // 2 if ne.numel() else None
    print("cuts", cuts, "len", c.numel() // 2, "nmis", ne.numel(), "first mismatch symbol", first, "state eq", sa == sc)
