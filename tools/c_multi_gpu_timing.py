"""The one-call C time-sharded entry point on 1 and on N GPUs of one box (lrpt_sharded_process_multi: a host thread per
GPU, ncclSend/ncclRecv of boundary state and overlap rows): wall time for one recording in page-locked host memory,
H2D + three passes + join + D2H inside, and byte equality of the results. python tools/c_multi_gpu_timing.py [log2 samples]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_demod_b200 import sharded, symbol_capacity, synth  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 31)
ngpu = torch.cuda.device_count()
per = synth.baseband(230000, periodic=True, seed=3).astype(np.complex64)
host = torch.empty(2 * n, dtype=torch.int16, pin_memory=True)
step = 1 << 28
for s0 in range(0, n, step):                                      # generated slice by slice: the device never holds it all
    part = synth.device_long_stream(per, n, total=min(step, n - s0), first=s0, cfo_hz=90.0)
    host[2 * s0: 2 * (s0 + part.numel() // 2)].copy_(part)
    del part
torch.cuda.synchronize()
torch.cuda.empty_cache()
hn = host.numpy()
out = [torch.empty((symbol_capacity(n, 230000, 72000), 2), dtype=torch.int8, pin_memory=True) for _ in range(2)]
res = {}
for k, devs in enumerate([[0]] + ([list(range(ngpu))] if ngpu > 1 else [])):
    kw = dict(chunk=262144, warm=150000, overlap=8192, devices=devs, out=out[k].numpy())
    soft, rep = sharded.process_host(hn, **kw)                    # first call: module load, NCCL communicators
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        soft, rep = sharded.process_host(hn, **kw)
        best = min(best, time.perf_counter() - t0)
    res[len(devs)] = (soft, rep, best)
    print(json.dumps({"gpus": len(devs), "samples": n, "seconds": best, "gsps": n / best / 1e9, "symbols": int(soft.shape[0]),
                      "nchunks": rep["nchunks"], "min_agreement_final": rep["min_agreement_final"], "aligned": rep["aligned"]}), flush=True)
if len(res) == 2:
    a, b = res[1][0], res[ngpu][0]
    print(json.dumps({"byte_identical_1_vs_%d_gpus" % ngpu: bool(a.shape == b.shape and np.array_equal(a, b))}))
