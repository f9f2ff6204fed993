"""Phase timing of lrpt_sharded_process on a 1-GSample recording in pinned and in pageable host memory."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LRPT_SHARD_TIMING"] = "1"
import numpy as np, torch
from meteor_demod_b200 import sharded, synth
N = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
per = synth.baseband(230000, periodic=True, seed=3).astype(np.complex64)
raw = synth.device_long_stream(per, N)
pinned = torch.empty(2 * N, dtype=torch.int16, pin_memory=True); pinned.copy_(raw); torch.cuda.synchronize()
del raw; torch.cuda.empty_cache()
kw = dict(chunk=1 << 18, warm=150000, overlap=8192, symrate=72000, bps=16)
for name, arr in (("pinned", pinned.numpy()), ("pinned again", pinned.numpy()), ("pageable", pinned.numpy().copy())):
    t = time.perf_counter()
    soft, rep = sharded.process_host(arr, **kw)
    print("%s: %.1f ms total, %d symbols, %s" % (name, (time.perf_counter() - t) * 1e3, soft.shape[0], rep), file=sys.stderr, flush=True)
