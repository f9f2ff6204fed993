"""Next-round check (needs a GPU): OQPSK time shards on the GPU engine, which is still behind
LRPT_EXPERIMENTAL_OQPSK_SHARDS because only the oracle-driven emulation has been run so far
(tests/test_sharded.py::test_oqpsk_handoff_scheme_on_the_oracle). Demodulates the test's stream with the GPU
engine and with the CPU oracle as the engine and compares: the two must be byte-identical (same arithmetic,
bit-exact engines); then reports eps against the sequential run."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["LRPT_EXPERIMENTAL_OQPSK_SHARDS"] = "1"
import numpy as np, torch
import test_sharded as T
from meteor_demod_b200 import sharded
from oracle import pyoracle

raw = T.make_oqpsk_stream()
n = raw.size // 2
plan = sharded.Plan(n, T.CHUNK, T.WARM, T.OVERLAP, T.OQ_CFG["interp"])
want = sharded.run_handoff(T.OracleEngine(raw, plan, cfg=T.OQ_CFG), plan, oqpsk_half=T.OQ_HALF)
dev = torch.zeros(2 * plan.padded, dtype=torch.uint8, device="cuda")              # the padding the oracle engine uses
dev[: raw.size] = torch.from_numpy(raw).cuda()
got = sharded.demod_sharded(dev, n, chunk=T.CHUNK, warm=T.WARM, overlap=T.OVERLAP, symrate=80000, oqpsk=True, bps=8,
                            rrc_order=32, interp_factor=5, handoff=True)
a, b = got["soft"].cpu().numpy(), want["soft"].numpy()
print("first-pass K  gpu", got["first_pass"]["K"].tolist(), " oracle engine", want["first_pass"]["K"].tolist())
print("final k       gpu", got["k"].tolist(), " oracle engine", want["k"].tolist())
print("symbols", a.shape[0], b.shape[0], "identical:", a.shape == b.shape and bool(np.array_equal(a, b)))
seq = pyoracle.Oracle(**T.OQ_CFG).process(raw, want_float=False).soft
print("vs sequential:", T.tier_s_report(a, seq))

# carrier-seeded warm-up (GpuEngine.seed_carrier, also not yet run on a GPU): OQPSK at +1200 Hz, where cold chunks
# cannot lock within a 160 k warm-up; again the GPU engine against the oracle engine, then against the sequential run
from meteor_demod_b200 import synth
n = 2_600_000
raw = synth.make_raw(n, symrate=80000, oqpsk=True, bps=8, cfo_hz=1200.0, seed=4)
plan = sharded.Plan(n, T.CHUNK, T.WARM, T.OVERLAP, 5)
want = sharded.run_handoff(T.OracleEngine(raw, plan, cfg=T.OQ_CFG, seed_carrier=True), plan, oqpsk_half=T.OQ_HALF)
dev = torch.zeros(2 * plan.padded, dtype=torch.uint8, device="cuda")
dev[: raw.size] = torch.from_numpy(raw).cuda()
got = sharded.demod_sharded(dev, n, chunk=T.CHUNK, warm=T.WARM, overlap=T.OVERLAP, symrate=80000, oqpsk=True, bps=8,
                            rrc_order=32, interp_factor=5, handoff=True, seed_carrier=True)
a, b = got["soft"].cpu().numpy(), want["soft"].numpy()
print("seeded: final k gpu", got["k"].tolist(), "symbols", a.shape[0], b.shape[0],
      "identical to the oracle engine:", a.shape == b.shape and bool(np.array_equal(a, b)),
      "(the FFT runs in float32 on both, but on different devices: a last-bit difference in the seed is possible)")
seq = pyoracle.Oracle(**T.OQ_CFG).process(raw, want_float=False).soft
print("seeded vs sequential:", T.tier_s_report(a, seq))
