"""Scratch: which boundaries of a time-sharded run disagree, and what the two rows look like there."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meteor_demod_b200 import sharded, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
C, W, V, L = 1 << 18, 400000, 8192, 5
plan = sharded.Plan(N, C, W, V, L)
per = synth.baseband(230000, periodic=True, seed=3).astype(np.complex64)
raw = synth.device_long_stream(per, N, total=plan.padded)
sd = sharded.ShardedDemod(raw, N, chunk=C, warm=W, overlap=V, symrate=72000, bps=16, rrc_order=32, interp_factor=L)
eng = sd.eng
head = eng.warm_up()
soft, q, count, base = eng.owned()
Bq = torch.tensor([plan.cut_target(c) for c in range(1, plan.nchunks)], dtype=torch.int64, device="cuda")
k, agree, cut = sharded.boundary_quadrants(soft, q, count, Bq, base=base)
q = q.to(torch.int64) + base[:, None]                      # absolute indices for the prints below
bad = (agree < 0.9).nonzero().squeeze(1)
print("pass B: nchunks", plan.nchunks, "bad boundaries", bad.tolist()[:20], "min agree", float(agree.min()))
for b in bad.tolist()[:5]:
    c = b + 1
    st = [eng.d.status(s) for s in (c - 1, c)]
    print(" boundary", c, "k", int(k[b]), "agree", float(agree[b]), "counts", int(count[c-1]), int(count[c]),
          "locked", st[0]["locked"], st[1]["locked"], "first_lock", st[0]["first_lock_symbol"], st[1]["first_lock_symbol"],
          "pllfreq", st[0]["pll_freq"], st[1]["pll_freq"], "omega", st[0]["mm_omega"], st[1]["mm_omega"])
    ia = sharded._first_at_or_after(q[c-1:c], count[c-1:c], Bq[b:b+1])
    ib = sharded._first_at_or_after(q[c:c+1], count[c:c+1], cut[b:b+1] + 1)
    print("   q a:", q[c-1, int(ia)-2:int(ia)+4].tolist(), " q b:", q[c, int(ib):int(ib)+5].tolist(), "cut", int(cut[b]), "B", int(Bq[b]))
