"""A/B of the lane kernel's window formats on 8-bit input (raw u8 pairs, bfloat16 pairs, float pairs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
from ab_wf import run
from meteor_demod_b200 import synth

os.environ["LRPT_LANE_DEBUG"] = "1"
N = 32768
per80 = synth.baseband(230000, symrate=80000, oqpsk=True, periodic=True).astype(np.complex64)
raw = synth.device_streams(per80, 75776, N, bps=8, sps=230000/80000)
ref = None
for B, env in ((75776, {"LRPT_LANE_WF": "0"}), (75776, {"LRPT_LANE_WF": "2"}), (75776, {}), (37888, {"LRPT_LANE_WF": "0"}),
               (37888, {"LRPT_LANE_WF": "2"}), (37888, {"LRPT_LANE_WF": "1"}), (37888, {}), (4736, {"LRPT_LANE_WF": "0"}), (4736, {})):
    s, c = run(raw, B, N, env, oqpsk=1, symrate=80000, bps=8)
    if ref is None:
        ref = (s, c)
    else:
        print("   identical to raw-window output:", bool(torch.equal(s[:B], ref[0][:B])), flush=True)
del raw, ref
per = synth.baseband(230000, periodic=True).astype(np.complex64)
raw = synth.device_streams(per, 75776, N, bps=8)
for env in ({"LRPT_LANE_WF": "0"}, {"LRPT_LANE_WF": "2"}):
    run(raw, 75776, N, env, bps=8)                      # QPSK 72k on 8-bit input
for env in ({"LRPT_LANE_WF": "0"}, {"LRPT_LANE_WF": "2"}, {}):
    run(raw, 52096, N, env, bps=8, order=64, L=8)       # RRC-64 x8 on 8-bit input
