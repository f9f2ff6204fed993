"""Sweep of the time-sharding parameters (chunk size, warm-up) on ONE device-resident stream: throughput and
Tier-S epsilon against the sequential CPU oracle on the head of the stream. Usage:
    python tools/shard_sweep.py [log2_samples=30] [chunk:warm ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meteor_demod_b200 import sharded, synth
from oracle import pyoracle

FS = 230000
N = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
combos = [tuple(int(v) for v in s.split(":")) for s in sys.argv[2:]] or \
    [(262144, 150000), (131072, 150000), (65536, 150000), (32768, 150000), (65536, 100000), (131072, 100000)]
period = synth.baseband(FS, symrate=72000, periodic=True, seed=3).astype(np.complex64)
pad = max(sharded.Plan(N, c, w, 8192, 5).padded for c, w in combos)
t = time.time()
raw = synth.device_long_stream(period, N, total=pad, bps=16, sps=FS / 72000)
torch.cuda.synchronize()
print("stream of %d samples generated in %.1f s" % (N, time.time() - t), flush=True)
ncheck = min(N, 12_000_000)
w = pyoracle.Oracle(symrate=72000, oqpsk=0, bps=16, order=32, interp=5).process(raw[: 2 * ncheck].cpu().numpy(), want_float=False)
for chunk, warm in combos:
    plan = sharded.Plan(N, chunk, warm, 8192, 5)
    try:
        sd = sharded.ShardedDemod(raw, N, chunk=chunk, warm=warm, overlap=8192, device=0, symrate=72000, bps=16,
                                  rrc_order=32, interp_factor=5, two_pass=True, handoff=True)
        res = sd.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            res = sd.run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        got = res["soft"][: w.nsym].cpu().numpy()
        n = min(len(got), w.nsym) - 4096           # the oracle's last symbols sit next to the end of ITS input only
        dlt = np.abs(got[:n].astype(np.int16) - w.soft[:n].astype(np.int16)).max(axis=1)
        print("chunk %7d warm %7d chunks %6d kernel %-5s: %8.2f ms -> %8.1f MS/s  eps(>1LSB) %.4f%% identical %.2f%% min agreement %.4f  peak mem %.1f GB"
              % (chunk, warm, plan.nchunks, sd.eng.d.kernel_name(), ms, N / ms / 1e3, 100 * (dlt > 1).mean(), 100 * (dlt == 0).mean(),
                 float(res["agreement"].min()), torch.cuda.max_memory_allocated() / 1e9), flush=True)
        sd.close()
        del sd, res
    except Exception as e:                           # e.g. out of memory for very many chunks
        print("chunk %d warm %d: %s: %s" % (chunk, warm, type(e).__name__, str(e)[:200]), flush=True)
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
