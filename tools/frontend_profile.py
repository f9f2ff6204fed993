"""One pass of the decoder front-end over 8192 synthetic CADUs (for ncu / timing): python tools/frontend_profile.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_demod_b200 import frontend  # noqa: E402

rng = np.random.default_rng(1)
soft = torch.from_numpy(rng.integers(-128, 128, (8192 * 8192, 2), dtype=np.int8)).cuda()
off = torch.arange(8192, dtype=torch.int32, device="cuda") * 8192
hyp = torch.zeros(8192, dtype=torch.uint8, device="cuda")
vit = frontend.Viterbi()
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    vit.decode(soft, off, hyp)
    e1.record()
    torch.cuda.synchronize()
    print("viterbi ms", e0.elapsed_time(e1))
