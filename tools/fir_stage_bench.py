"""g1: the RRC matched-filter stage measured ALONE (lrpt_fir_stage_device, csrc/fir_stage.cu).

Per input sample the stage reads 2*bps/8 bytes, writes 8*L bytes (all L polyphase outputs, float2) and does
4*taps*L flops. Reports, per configuration and mode (exact = multiply and add rounded separately, bit-identical
to filter_get; fma = one rounding per tap): GS/s, algorithmic GB/s and its share of the measured HBM peak,
fp32 Tflop/s and its share of the fp32 pipe peak (148 SMs x 128 lanes x 2 flop x clock; the exact form issues
TWO instructions per multiply-accumulate, so its pipe occupancy is twice its flop share). Writes a JSON line."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from meteor_demod_b200 import _lib, synth  # noqa: E402
from meteor_demod_b200.demod import make_params  # noqa: E402

lib = _lib.load()
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
rows_out = []
for name, symrate, oq, bps, order, L in (("c1", 72000, 0, 16, 32, 5), ("c2", 80000, 1, 8, 32, 5), ("c3", 72000, 0, 16, 64, 8),
                                          ("o16_L3", 72000, 0, 16, 16, 3)):
    p = make_params(symrate=symrate, oqpsk=oq, bps=bps, rrc_order=order, interp_factor=L)
    per = synth.baseband(230000, symrate=symrate, oqpsk=bool(oq), periodic=True, seed=3).astype(np.complex64)
    R, N = 256, 1 << 19                                        # 134 M samples: 0.5 GB in, 5.4 GB out at L = 5
    raw = synth.device_streams(per, R, N, bps=bps, sps=230000 / symrate, seed=5)
    out = torch.empty((R, N * L * 2), dtype=torch.float32, device="cuda")
    taps = 2 * order + 1
    modes = ((0, "exact"), (1, "fma"))
    if "--sweep" in sys.argv:                                  # bits 1-2 of mode: samples per thread (0 = by filter length)
        modes = tuple((m | (ns_ << 1), "%s/ns%d" % (nm, ns_)) for m, nm in modes for ns_ in (1, 2))
    for mode, mname in modes:
        best, times = 1e9, []
        for _ in range(30):                                   # ~0.3 s of launches first: clocks up, code resident
            lib.lrpt_fir_stage_device(C.byref(p), raw.data_ptr(), raw.stride(0) * raw.element_size(), R, N,
                                      out.data_ptr(), out.stride(0) * 4, mode, None)
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            rc = lib.lrpt_fir_stage_device(C.byref(p), raw.data_ptr(), raw.stride(0) * raw.element_size(), R, N,
                                           out.data_ptr(), out.stride(0) * 4, mode, None)
            e1.record()
            torch.cuda.synchronize()
            assert rc == 0, rc
            times.append(e0.elapsed_time(e1))
        best = float(np.median(times))
        ns = R * N
        gbs = ns * (bps // 4 + 8 * L) / (best * 1e-3) / 1e9
        tf = ns * 4.0 * taps * L / (best * 1e-3) / 1e12
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
        row = {"config": name, "mode": mname, "taps": taps, "interp": L, "bps": bps, "samples": ns, "ms": best, "ms_min": min(times), "ms_max": max(times),
               "gsps": ns / best / 1e6, "hbm_gbs": gbs, "hbm_frac_of_measured_peak": gbs / peak, "fp32_tflops": tf,
               "fp32_flop_frac_of_75TF": tf / fp32_peak,
               "fp32_pipe_frac_est": tf / fp32_peak * (1.0 if mode & 1 else 2.0),
               "bytes_per_sample": bps // 4 + 8 * L, "flops_per_sample": 4 * taps * L}
        rows_out.append(row)
        print(json.dumps(row), flush=True)
    del raw, out
    torch.cuda.empty_cache()
name = "r2_fir_stage_sweep.json" if "--sweep" in sys.argv else "r2_fir_stage.json"
json.dump({"hbm_peak_gbs": peak, "rows": rows_out}, open(os.path.join(ROOT, "gpurun_out", name), "w"), indent=1)
