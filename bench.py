#!/usr/bin/env python
"""bench.py -- headline benchmark of the LRPT demodulator hot path (contract: task section 4).

A "step" is one pass of the hot path over one batch of synthetic I/Q: B independent
streams x N samples (QPSK 72 ksym/s at 230 kS/s, 16-bit, RRC order 32, x5 interpolation by
default = the configuration BASELINE.json's metric is quoted on), every stream demodulated
from power-on state to int8 soft symbols, bit-exactly as the strict-IEEE reference does.

  value     whole-job input Msamples/s with the raw I/Q already resident in HBM
            (CUDA events on the launching stream, max over ranks)
  e2e       the same through the C ABI with HOST (pinned) buffers: H2D of the raw I/Q and
            D2H of the soft symbols inside the timed region (lrpt_process_batch)
  roofline  HBM: algorithmic bytes (raw in + soft out) / kernel time vs MEASURED_PEAKS.json
  cpu_baseline / --impl reference
            the reference's own C code (oracle/_ref/libref_fma.so = reference sources with
            the reference's release flags), one process per host core, on a bounded sample
            of the same workload.

Multi-GPU (torchrun): streams are sharded across ranks, no data-path collective, weak scaling
(every rank demodulates B streams); value = all ranks' samples / max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (symrate, oqpsk, bps, order, interp, label)
    "c1": (72000, 0, 16, 32, 5, "QPSK 72ksym/s fs=230kS/s s16 RRC-32 x5 (BASELINE metric config)"),
    "c2": (80000, 1, 8, 32, 5, "OQPSK 80ksym/s fs=230kS/s u8 RRC-32 x5"),
    "c3": (72000, 0, 16, 64, 8, "QPSK 72ksym/s fs=230kS/s s16 RRC-64 x8"),
}
# streams per GPU that fill every SM with as many warps as the lane kernel's shared-memory delay lines allow
# (148 SMs x W warps x 32 lanes; W = 16 at 65 taps, 11 at 129 taps of 16-bit input)
DEFAULT_STREAMS = {"c1": 75776, "c2": 75776, "c3": 52096}
FS = 230000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c1", choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=0, help="independent streams per GPU (default: DEFAULT_STREAMS[workload])")
    ap.add_argument("--samples", type=int, default=1 << 15, help="samples per stream per step")
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--mode", default="batch", choices=["batch", "sharded", "relay"],
                    help="batch: independent streams (exact); sharded: ONE stream of --stream-samples per job, "
                         "time-sharded over chunks and GPUs (Tier-S, reports eps); relay: the batch with its TIME "
                         "axis split over the GPUs, complete state handed rank to rank over NCCL (exact)")
    ap.add_argument("--groups", type=int, default=4, help="relay: stream groups travelling down the ranks")
    ap.add_argument("--stream-samples", type=int, default=1 << 30)
    ap.add_argument("--chunk", type=int, default=1 << 18)
    ap.add_argument("--warm", type=int, default=0, help="sharded: warm-up samples per chunk (default 150000 with the "
                    "hand-off scheme, 400000 otherwise)")
    ap.add_argument("--two-pass", action="store_true", help="sharded: two-pass lock-point alignment instead of hand-off")
    ap.add_argument("--single-pass", action="store_true")
    ap.add_argument("--handoff", action="store_true",
                    help="sharded: chunks continue from their predecessor's state (sharded.run_handoff; the default)")
    ap.add_argument("--seed-carrier", action="store_true", help="sharded: chunks >= 1 start their Costas NCO at a coarse "
                    "carrier estimate (meteor_demod_b200/acquire.py; opt-in, not what the reference does)")
    ap.add_argument("--seed-nfft", type=int, default=4096, help="sharded: samples the coarse carrier estimate looks at")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-locked", action="store_true", help="skip the steady-state (locked, state carried over) pass")
    ap.add_argument("--no-single", action="store_true", help="skip the single exact stream sub-record")
    ap.add_argument("--no-frontend", action="store_true", help="skip the decoder front-end sub-record")
    ap.add_argument("--no-c4", action="store_true", help="skip the time-sharded single-stream sub-record (BASELINE config 4)")
    ap.add_argument("--c4-samples", type=int, default=1 << 33, help="length of the ONE stream of the c4 sub-record")
    ap.add_argument("--c4-unseeded", action="store_true", help="c4 with the reference's own acquisition sweep in every chunk (150 k warm-up)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=6.0, help="target CPU work per core for the baseline")
    return ap.parse_args()


# ------------------------------------------------------------------ clocks sampler --

class Clocks:
    """SM clock, power and throttle reasons sampled DURING the timed region: NVML every 20 ms
    (nvidia-smi every 100 ms if the NVML binding is missing)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
        except Exception:
            self.nvml, self.source = None, "nvidia-smi"

    def sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.dev, n.NVML_CLOCK_SM))
        pw = n.nvmlDeviceGetPowerUsage(self.dev) / 1000.0
        try:
            rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
        except Exception:
            rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
        bits = (0x8, 0x40, 0x20, 0x4)      # HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap
        return [sm, self.max_sm, pw] + ["Active" if rs & b else "Not Active" for b in bits]

    def run(self):
        while not self.stop.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self.sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    f = [x.strip() for x in out.strip().split(",")]
                    if len(f) >= 7:
                        self.rows.append(f)
            except Exception:
                pass
            self.stop.wait(0.02 if self.nvml is not None else 0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = [n for i, n in enumerate(self.NAMES) if any(str(r[3 + i]).lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows), "source": self.source}


# ------------------------------------------------------------------ CPU reference ---

_CPU_PERIOD = None     # tileable baseband period, built once in the parent before fork
CPU_POOL = 8           # distinct streams per CPU process


def _cpu_worker(args):
    """One process = the reference library loaded IN PLACE (oracle/_ref/libref_fma.so shows up in the
    process's memory map), demodulating N-sample streams from POWER-ON state, one after the other -- the
    shape of the GPU arm's step (every stream of the batch starts from power-on state). A pool of
    CPU_POOL distinct streams (same recipe and parameter ranges as synth.device_streams) is cycled
    `reps` times; pyoracle.Ref.power_on() gives every stream a fresh image of the reference's statics."""
    kind, cfg, seed, nsamples, reps = args
    from meteor_demod_b200 import synth
    from oracle import pyoracle
    symrate, oqpsk, bps, order, interp = cfg
    rng = np.random.Generator(np.random.PCG64(seed))
    P = _CPU_PERIOD.size
    pool = []
    for k in range(CPU_POOL):
        z = np.tile(np.roll(_CPU_PERIOD, int(rng.integers(0, P))), nsamples // P + 1)[:nsamples]
        y = synth.impair(z, FS, cfo_hz=float(rng.integers(-1500, 1500)), phase=float(rng.uniform(0, 6.28)),
                         esn0_db=12.0, sps=FS / symrate, seed=seed * 131 + k)
        pool.append(synth.to_raw(y * float(rng.uniform(0.6, 1.2)), bps))
    d = pyoracle.Ref(symrate=symrate, oqpsk=oqpsk, bps=bps, order=order, interp=interp, kind="fma", in_place=True)
    d.process(pool[0][: 2 * 4096], want_float=False)      # touch code + data once
    nsym = 0
    t0 = time.perf_counter()
    for _ in range(reps):
        for raw in pool:
            d.power_on()
            nsym += d.process(raw, want_float=False).nsym
    dt = time.perf_counter() - t0
    return nsamples * reps * CPU_POOL, dt, nsym


def usable_cores():
    """Host threads this container may really use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0))
    for quota_f, period_f in (("/sys/fs/cgroup/cpu.max", None),
                              ("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us")):
        try:
            if period_f is None:
                quota, period = open(quota_f).read().split()[:2]
            else:
                quota, period = open(quota_f).read().strip(), open(period_f).read().strip()
            if quota not in ("max", "-1"):
                n = min(n, max(1, int(np.ceil(int(quota) / int(period)))))
            break
        except Exception:
            continue
    return n


def _cpu_run(kind, cfg, cores, nsamples, reps):
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(kind, cfg, 1000 + i, nsamples, reps) for i in range(cores)], chunksize=1)
    wall = time.perf_counter() - t0
    tot = sum(r[0] for r in res)
    slow = max(r[1] for r in res)
    per_core = float(np.median([r[0] / r[1] for r in res])) / 1e6
    return tot / slow / 1e6, per_core, wall, slow


def cpu_reference(cfg, seconds, nsamples, cores=None):
    """The reference's own C code on the host cores: `cores` concurrent processes (one per usable host
    thread), each demodulating `nsamples`-sample streams from power-on state; throughput = total samples /
    slowest process time. A one-pass calibration sizes the timed run to about `seconds` per process.
    There is no fallback: without oracle/_ref/libref_fma.so (the reference compiled by oracle/Makefile) this
    raises instead of timing our own port under the reference's name."""
    from meteor_demod_b200 import synth
    from oracle import pyoracle
    global _CPU_PERIOD
    pyoracle.build()
    if not pyoracle.have_ref("fma"):
        raise SystemExit("bench.py: oracle/_ref/libref_fma.so is missing (run `make -C oracle ref` where /root/reference "
                         "exists); the reference arm does not fall back to the port")
    kind = "reference"
    cores = cores or usable_cores()
    symrate, oqpsk = cfg[0], cfg[1]
    if _CPU_PERIOD is None:
        _CPU_PERIOD = synth.baseband(FS, symrate=symrate, oqpsk=bool(oqpsk), periodic=True, seed=3).astype(np.complex64)
    value, per_core, wall, slow = _cpu_run(kind, cfg, cores, nsamples, 2)
    reps = int(2 * seconds / max(slow, 1e-3))
    if reps > 2:
        reps = min(reps, 4096)
        value, per_core, wall, slow = _cpu_run(kind, cfg, cores, nsamples, reps)
    else:
        reps = 2
    return {"value": value, "unit": "Msamples/s", "cores": cores, "kind": kind,
            "per_core_msps": per_core, "wall_s": wall,
            "sample": "%d concurrent processes x %d streams of %d samples each, every stream from power-on state "
                      "(oracle/_ref/libref_fma.so loaded in place: reference sources, -O3 -march=x86-64-v3 "
                      "-ftree-vectorize -std=gnu99), timed region = demod only" % (cores, reps * CPU_POOL, nsamples)}


def bind_to_gpu_cpus(index):
    """Restrict this process to the CPU cores NVML reports as local to GPU `index`, so that the pinned
    host buffers of the end-to-end path are first-touched on that GPU's NUMA node (matters with several
    ranks on a two-socket box). Returns the number of cores bound to, or None when NVML gives nothing."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ------------------------------------------------------------------ main -----------

def main():
    a = parse()
    if a.streams <= 0:
        a.streams = DEFAULT_STREAMS[a.workload]
    symrate, oqpsk, bps, order, interp, label = WORKLOADS[a.workload]
    cfg = (symrate, oqpsk, bps, order, interp)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "%s; %d streams/GPU x %d samples, power-on state each step" % (label, a.streams, a.samples),
              "streams_per_gpu": a.streams, "samples_per_stream": a.samples, "parity": "bit-exact vs strict-IEEE reference",
              "l2": "inputs (%.1f GB/GPU) larger than L2" % (a.streams * a.samples * (bps // 4) / 1e9)}

    if a.impl == "reference":
        if rank != 0:
            return
        ncfg = max(1, a.steps)
        vals = []
        secs = max(1.0, min(a.cpu_seconds, 90.0 / ncfg))   # the whole arm stays within a few minutes
        for _ in range(max(0, min(a.warmup, 1))):
            cpu_reference(cfg, min(secs, 2.0), a.samples)
        for _ in range(ncfg):
            vals.append(cpu_reference(cfg, secs, a.samples))
        best = max(vals, key=lambda r: r["value"])
        line = {"impl": "reference", "metric": "IQ Msamples/s", "value": float(np.mean([v["value"] for v in vals])),
                "unit": "Msamples/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1e3 * float(np.mean([v["wall_s"] for v in vals])), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {k: best[k] for k in ("value", "unit", "cores", "kind", "sample", "per_core_msps")},
                "gpu_launches": 0}
        line["cpu_baseline"]["value"] = line["value"]
        line["e2e"] = {"value": line["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        print(json.dumps(line))
        return

    cpu_base = None
    if not a.no_cpu and world == 1:
        cpu_base = cpu_reference(cfg, a.cpu_seconds, a.samples)   # before CUDA is initialised (fork)

    import torch
    import torch.distributed as dist
    from meteor_demod_b200 import Demod, synth

    numa = bind_to_gpu_cpus(local)                          # pinned host buffers land on the GPU's own NUMA node
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the demodulator has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    if a.mode == "sharded":
        return bench_sharded(a, cfg, label, rank, world, local, cpu_base)
    if a.mode == "relay":
        return bench_relay(a, cfg, label, rank, world, local)

    B, N = a.streams, a.samples
    d = Demod(symrate=symrate, oqpsk=oqpsk, bps=bps, rrc_order=order, interp_factor=interp, nstreams=B,
              device=local, kernel=a.kernel)
    period = synth.baseband(FS, symrate=symrate, oqpsk=bool(oqpsk), periodic=True, seed=3).astype(np.complex64)
    raw = synth.device_streams(period, B, N, bps=bps, sps=FS / symrate, seed=7 + rank, device="cuda")
    cap = (d.capacity(N + 64) + 7) // 8 * 8
    soft = torch.empty((B, 2 * cap), dtype=torch.int8, device="cuda")
    nsym = torch.zeros(B, dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, steps, warmup, tag):
        """W untimed + K timed steps on stream `st`: CUDA events per step, total = max over ranks."""
        for _ in range(warmup):
            step()
        barrier()
        l0 = d.launch_count()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        with Clocks(local) as clk:
            barrier()
            torch.cuda.nvtx.range_push(tag)                 # ncu --nvtx --nvtx-include "<tag>/"
            ev[0].record(st)
            for i in range(steps):
                step()
                ev[i + 1].record(st)
            st.synchronize()
            torch.cuda.nvtx.range_pop()
            barrier()
        t = torch.tensor([ev[0].elapsed_time(ev[-1])], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)], d.launch_count() - l0, clk.summary()

    # ---- value: every stream from power-on state, raw I/Q resident in HBM -------------------------------------
    def step():
        d.reset(stream=st)                                  # enqueued on st, ordered with the launch
        d.process_device(raw, soft, nsym=nsym, stream=st)

    total_ms, step_ms, launches, clocks = timed(step, a.steps, a.warmup, "bench_timed")
    counts = d.counts().astype(np.int64)
    kernel_name = d.kernel_name()
    value = world * B * N * a.steps / (total_ms * 1e-3) / 1e6

    # correctness tripwire on the benchmarked data itself: 32 streams spread over warps and CTAs (first and last
    # lane of a warp, first and last CTA) against the CPU oracle, whole step
    check = None
    if rank == 0:
        from oracle import pyoracle
        picks = sorted({0, 31, 32, B - 1, B - 32, B // 2 + 17} | {(k * B) // 26 + (7 * k) % 32 for k in range(26)})
        picks = [s_ for s_ in picks if 0 <= s_ < B]
        rows = raw[picks].cpu().numpy()
        got_all = soft[picks].cpu().numpy()
        check = {"streams": len(picks), "ok": True}
        for i, sidx in enumerate(picks):
            o = pyoracle.Oracle(symrate=symrate, oqpsk=oqpsk, bps=bps, order=order, interp=interp)
            w = o.process(rows[i], want_float=False)
            ok = int(counts[sidx]) == w.nsym and bool(np.array_equal(got_all[i, : 2 * w.nsym].reshape(-1, 2), w.soft))
            check["ok"] = check["ok"] and ok

    # ---- e2e: the same step through the C ABI with HOST (pinned) buffers, H2D + D2H inside the timed region ----
    e2e = None
    if not a.no_e2e:
        h_raw = torch.empty((B, raw.shape[1]), dtype=raw.dtype, pin_memory=True)
        h_raw.copy_(raw)
        h_soft = torch.zeros((B, 2 * cap), dtype=torch.int8, pin_memory=True)
        h_cnt = np.zeros(B, np.uint32)
        lib = d.lib

        def e2e_step():
            d.reset()
            rc = lib.lrpt_process_batch(d.h, h_raw.data_ptr(), h_raw.stride(0) * h_raw.element_size(), N,
                                        h_soft.data_ptr(), h_soft.stride(0), cap, h_cnt.ctypes.data, None, 0)
            assert rc == 0, rc
        for _ in range(max(1, min(a.warmup, 2))):
            e2e_step()
        barrier()
        torch.cuda.nvtx.range_push("bench_e2e")
        t0 = time.perf_counter()
        for _ in range(a.steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        torch.cuda.nvtx.range_pop()
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        # every byte the host received against the device-resident path's output (same input, same power-on state)
        back = h_soft.cuda(non_blocking=True)
        cnt_d = torch.from_numpy(counts).cuda()
        valid = torch.arange(2 * cap, device="cuda")[None, :] < 2 * cnt_d[:, None]
        same_bytes = bool(((back == soft) | ~valid).all().item()) and bool(np.array_equal(h_cnt.astype(np.int64), counts))
        del back, valid
        # what the host link alone does with the same pinned buffers when ALL ranks copy at the same time: one step's
        # input host -> device and, on a second stream, one step's symbols device -> host (the two directions share
        # the host's memory system)
        nsym_max = int(h_cnt.max())
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        barrier()
        ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        ev0.record()
        s_in.wait_event(ev0)
        s_out.wait_event(ev0)
        with torch.cuda.stream(s_in):
            raw.copy_(h_raw, non_blocking=True)
            ev1.record(s_in)
        with torch.cuda.stream(s_out):
            nb_out = B * 2 * nsym_max                           # the bytes a step returns, as ONE contiguous copy
            h_soft.view(-1)[:nb_out].copy_(soft.view(-1)[:nb_out], non_blocking=True)
            ev2.record(s_out)
        torch.cuda.synchronize()
        tl = torch.tensor([max(ev0.elapsed_time(ev1), ev0.elapsed_time(ev2)), ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        both_ms, h2d_ms = float(tl[0].item()), float(tl[1].item())
        h2d_gbs = h_raw.numel() * h_raw.element_size() / (h2d_ms * 1e-3) / 1e9
        e2e = {"value": world * B * N * a.steps / dt / 1e6, "unit": "Msamples/s",
               "h2d_link_gbs_per_rank_concurrent": h2d_gbs,
               "link_bound_msps": world * B * N / (both_ms * 1e-3) / 1e6,
               "link_bound_note": "plain pinned cudaMemcpyAsync of one step's input (host -> device) and of its symbols (device -> "
                                  "host) on two streams, on all %d rank(s) at the same time (max over ranks): the ceiling of any "
                                  "host-buffer path on this host" % world,
               "h2d_bytes_per_step": int(B * N * (bps // 4)), "d2h_bytes_per_step": int(2 * int(h_cnt.max()) * B + 4 * B),
               "ms_per_step": 1e3 * dt / a.steps,
               "matches_device_path": same_bytes, "compared": "every soft-symbol byte and count of all %d streams" % B}
        del h_raw, h_soft

    # ---- value_locked: steady state. Every row is ONE period of a periodic signal (len(period) == samples per
    # row, carrier a multiple of fs/len), so replaying the row with the state carried over is one continuous
    # stream; carrier offset 0..+150 Hz (the reference's acquisition sweep starts upwards at 1e-6 rad/symbol^2,
    # pll.c:126: an offset below zero is found only after the sweep has been to +3.4 kHz and back, ~600 k symbols),
    # three untimed steps of pre-roll: the loops are locked when timing starts.
    locked = None
    if not a.no_locked:
        NL = (N // 115) * 115 + (115 if N % 115 else 0)      # samples per row with NL*symrate/fs integral (72k and 80k)
        del raw
        torch.cuda.empty_cache()
        per_l = synth.baseband(NL, symrate=symrate, oqpsk=bool(oqpsk), periodic=True, seed=5).astype(np.complex64)
        items = (2 * NL * (bps // 8) + 15) // 16 * 16 // (bps // 8)
        raw_l = synth.device_streams(per_l, B, NL, bps=bps, sps=FS / symrate, seed=70 + rank, device="cuda",
                                     cfo_min_hz=0.0, cfo_max_hz=150.0, row_items=items)

        def locked_frac():
            buf = torch.empty(d.states_size(), dtype=torch.uint8, device="cuda")
            d.export_states_device(buf)
            d.sync()
            from meteor_demod_b200._lib import State
            import ctypes as C_
            sb = C_.sizeof(State)
            stt = buf[: B * sb].view(B, sb)
            o1 = State.p_locked.offset
            return float(stt[:, o1: o1 + 4].contiguous().view(torch.int32).ne(0).float().mean().item())

        def step_l():
            d.process_device(raw_l, soft, nsym=nsym, stream=st)

        d.reset()
        pre = 3
        for _ in range(pre):
            step_l()
        st.synchronize()
        lf0 = locked_frac()
        tl_ms, _, launches_l, clocks_l = timed(step_l, a.steps, 0, "bench_locked")
        lf1 = locked_frac()
        counts_l = d.counts().astype(np.int64)
        check_l = None
        if rank == 0:
            from oracle import pyoracle
            reps = pre + a.steps
            check_l = True
            for sidx in (0, 31, B // 2 + 5, B - 1):
                o = pyoracle.Oracle(symrate=symrate, oqpsk=oqpsk, bps=bps, order=order, interp=interp)
                row = raw_l[sidx].cpu().numpy()
                for _ in range(reps - 1):
                    o.process(row, want_float=False)
                w = o.process(row, want_float=False)         # the last timed step's symbols
                got = soft[sidx, : 2 * w.nsym].cpu().numpy().reshape(-1, 2)
                check_l = check_l and int(counts_l[sidx]) == w.nsym and bool(np.array_equal(got, w.soft))
        locked = {"value": world * B * NL * a.steps / (tl_ms * 1e-3) / 1e6, "unit": "Msamples/s",
                  "ms_per_step": tl_ms / a.steps, "samples_per_stream": NL, "pre_roll_steps": pre,
                  "locked_frac_at_start": lf0, "locked_frac_at_end": lf1, "gpu_launches": int(launches_l),
                  "oracle_check_4_streams_continuous": check_l,
                  "workload": "the same %d streams/GPU, state carried from step to step (no reset): each row is one period of a "
                              "periodic signal, carrier offset 0..+150 Hz" % B}
        del raw_l
    d.close()
    del soft, nsym
    torch.cuda.empty_cache()

    # ---- single_stream: BASELINE config 1 as literally stated -- ONE recording, exact ---------------------------
    single = None
    if not a.no_single:
        single = bench_single_stream(cfg, local, rank)

    # ---- frontend: the decoder front-end behind the path ----------------------------------------
    front = None
    if not a.no_frontend:
        front = bench_frontend(local, rank)

    # ---- c4: BASELINE config 4 -- ONE long stream, time-sharded over chunks and ranks with state hand-off ------
    c4 = c4_weak = None
    if not a.no_c4:
        def c4_record(nsamples, scaling):
            a_c4 = argparse.Namespace(**vars(a))
            a_c4.stream_samples, a_c4.steps, a_c4.warmup = nsamples, max(1, min(a.steps, 3)), 1
            a_c4.two_pass = a_c4.single_pass = False
            # chunks >= 1 start their Costas NCO at a coarse carrier estimate instead of sweeping to the carrier
            # at 1e-6 rad/symbol^2 (pll.c:126): 32 Ki samples of warm-up instead of 150 k; chunk 0 stays the sequential run
            a_c4.seed_carrier, a_c4.seed_nfft, a_c4.warm = (not a.c4_unseeded), 4096, (150000 if a.c4_unseeded else 32768)
            rec, sd_, raw_, res_ = sharded_measure(a_c4, cfg, label, rank, world, local)
            sd_.close()
            sd_.eng.raw = None
            del sd_, raw_, res_
            torch.cuda.empty_cache()
            if rec is not None:
                rec = {k: rec[k] for k in ("value", "unit", "ms_per_step", "steps", "scaling", "config", "gpu_launches", "symbols_per_step",
                                           "min_boundary_agreement", "tier_s", "phase_ms", "phase_ms_per_rank", "kernel", "chunks_per_rank")}
                rec["scaling"] = scaling
                rec["stream_samples"] = int(nsamples)
            return rec
        c4 = c4_record(a.c4_samples, "strong")
        if world > 1:
            # the same path with the recording growing with the ranks (every rank keeps the N = 1 share): what the
            # per-lane latency floor of the strong-scaling record hides
            c4_weak = c4_record(a.c4_samples * world, "weak")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = None, "fallback 6650 GB/s (B200_PROFILING.md)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak = 6650.0
    kern_ms = float(np.mean(step_ms))
    alg_bytes = B * N * (bps // 4) + 2.0 * float(counts.sum())          # raw read + soft symbols written, per launch
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None                                          # dram bytes per launch from the committed ncu capture
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))[a.workload]
        traffic = float(tj["bytes_per_stream_sample"]) * B * N
    except Exception:
        pass
    limiter = None                                          # what ncu says binds the kernel (committed capture)
    try:
        lim = json.load(open(os.path.join(ROOT, "profiles", "lane_kernel_limiter.json")))[a.workload]
        if kernel_name == "lane" and B == int(lim.get("streams", B)):
            limiter = lim
    except Exception:
        pass
    fir_flops = float(counts.sum()) * (2 if oqpsk else 1) * 4.0 * (2 * order + 1)   # one filter_get per (half-)symbol
    line = {"metric": "IQ Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "kernel": kernel_name, "gpu_launches": int(launches), "oracle_check": check,
            "symbols_per_step": int(counts.sum()), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "instruction-issue bound, not HBM bound (DESIGN.md section 5); the reference's own lazy "
                                 "FIR (4*taps flops per filter_get) runs at %.2f Tflop/s" % (fir_flops / (kern_ms * 1e-3) / 1e12),
                         "limiter": limiter},
            "e2e": e2e, "value_locked": locked, "single_stream": single, "c4": c4, "c4_weak": c4_weak, "frontend": front,
            "host_cores_bound_to_gpu_numa_node": numa}
    if cpu_base is not None:
        line["cpu_baseline"] = cpu_base
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_frontend(local, rank, nframes_block=64, tiles=128):
    """SURVEY 8(f1), the step behind this path: frame synchronisation + Viterbi decoding of a device-resident soft
    stream (csrc/frontend.cu). 8192 CADUs = 67 M symbols; every decoded payload is compared with what was sent; the
    CPU oracle (oracle/frontend_oracle.c, one core) is timed on a bounded sample."""
    import torch
    from meteor_demod_b200 import frontend
    from oracle import pyfrontend as fe
    rng = np.random.default_rng(11)
    frames = rng.integers(0, 256, (nframes_block, 1020), dtype=np.uint8)
    block = fe.transmit(frames, noise=25.0, turns=1, swap=False, seed=12)            # [64*8192, 2] int8, Eb/N0 ~ 7.6 dB
    soft = torch.from_numpy(block).cuda().repeat(tiles, 1).contiguous()
    nsym = soft.shape[0]
    nfr = nsym // fe.CADU_SYMS
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    vit = frontend.Viterbi(local)
    best_sync, best_vit = 1e9, 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        ev[0].record()
        score, hyp = frontend.sync_scores(soft)
        off, oh, osc = frontend.window_peaks(score, hyp)
        ev[1].record()
        cadu, metric = vit.decode(soft, off, oh)
        ev[2].record()
        torch.cuda.synchronize()
        best_sync, best_vit = min(best_sync, ev[0].elapsed_time(ev[1])), min(best_vit, ev[1].elapsed_time(ev[2]))
    want = torch.from_numpy(frames).cuda().repeat(tiles, 1)
    ok = (cadu[:, 4:] == want).all(dim=1) & (osc >= 53)
    rec = {"symbols": int(nsym), "frames": int(nfr), "sync_ms": best_sync, "viterbi_ms": best_vit,
           "sync_msym_s": nsym / best_sync / 1e3, "viterbi_msym_s": nsym / best_vit / 1e3,
           "frontend_msym_s": nsym / (best_sync + best_vit) / 1e3,
           "sync_hbm_gbs": nsym * 4.25 / (best_sync * 1e-3) / 1e9,
           "frames_recovered_frac": float(ok.float().mean().item()),
           "note": "soft stream resident in HBM; sync = pack + score (8 symmetries) + per-CADU peaks: 2 B read + 2.25 B written per "
                   "symbol; Viterbi = one warp per CADU, 8320 add-compare-select steps + traceback"}
    if rank == 0:
        n_cpu = 1 << 20
        head = soft[:n_cpu].cpu().numpy()
        t0 = time.perf_counter()
        s_c, h_c = fe.sync_scores(head)
        t1 = time.perf_counter()
        for f in range(8):
            fe.viterbi_cadu(block, f * fe.CADU_SYMS, 1)
        t2 = time.perf_counter()
        rec["cpu_oracle_1core"] = {"sync_msym_s": n_cpu / (t1 - t0) / 1e6, "viterbi_msym_s": 8 * fe.CADU_SYMS / (t2 - t1) / 1e6}
        m = n_cpu - 31                                       # the last 31 windows of the head run past its end
        same = bool(np.array_equal(score[:m].cpu().numpy(), s_c[:m]) and np.array_equal(hyp[:m].cpu().numpy(), h_c[:m]))
        c0, _ = fe.viterbi_cadu(head, 3 * fe.CADU_SYMS, 1)
        rec["equals_oracle"] = bool(same and np.array_equal(cadu[3].cpu().numpy(), c0))
    return rec


def bench_single_stream(cfg, local, rank, nsamples=1 << 22):
    """ONE exact stream on one GPU (BASELINE config 1 is one 230 kS/s recording): device-resident, the kernel
    LRPT_KERNEL_AUTO picks for a single stream, every symbol compared with the CPU oracle on rank 0."""
    import torch
    from meteor_demod_b200 import Demod, synth
    symrate, oqpsk, bps, order, interp = cfg
    per = synth.baseband(FS, symrate=symrate, oqpsk=bool(oqpsk), periodic=True, seed=3).astype(np.complex64)
    raw = synth.device_long_stream(per, nsamples, bps=bps, sps=FS / symrate, cfo_hz=90.0).view(1, -1)
    d = Demod(symrate=symrate, oqpsk=oqpsk, bps=bps, rrc_order=order, interp_factor=interp, nstreams=1, device=local)
    cap = (d.capacity(nsamples) + 7) // 8 * 8
    soft = torch.zeros((1, 2 * cap), dtype=torch.int8, device="cuda")
    st = torch.cuda.Stream()
    best = None
    for _ in range(3):
        d.reset(stream=st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        d.process_device(raw, soft, stream=st)
        e1.record(st)
        st.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    n = int(d.counts()[0])
    rec = {"msps": nsamples / best / 1e3, "ms": best, "samples": nsamples, "kernel": d.kernel_name(),
           "x_realtime_at_230kSps": nsamples / (best * 1e-3) / FS}
    if rank == 0:
        from oracle import pyoracle
        w = pyoracle.Oracle(symrate=symrate, oqpsk=oqpsk, bps=bps, order=order, interp=interp).process(
            raw[0].cpu().numpy(), want_float=False)
        rec["exact"] = bool(n == w.nsym and np.array_equal(soft[0, : 2 * n].cpu().numpy().reshape(-1, 2), w.soft))
        rec["symbols_checked"] = int(w.nsym)
    d.close()
    return rec


def _quarter_turn(w, k):
    return [(w[:, 0], w[:, 1]), (-w[:, 1], w[:, 0]), (-w[:, 0], -w[:, 1]), (w[:, 1], -w[:, 0])][k % 4]


def best_align(g, w, search=24):
    """(agreement of hard decisions, offset, quarter turns k) with rot(w[off:], k) ~ g, over +-search symbols."""
    best = None
    n = min(len(g), len(w)) - 2 * search
    gg = g[search: search + n].astype(np.int16)
    for off in range(-search, search + 1):
        ww = w[search + off: search + off + n].astype(np.int16)
        for k in range(4):
            wi, wq = _quarter_turn(ww, k)
            agree = float(((np.sign(gg[:, 0]) == np.sign(wi)) & (np.sign(gg[:, 1]) == np.sign(wq))).mean())
            if best is None or agree > best[0]:
                best = (agree, off, k)
    return best


def tier_s_local(raw_np, got_probe, got, mk, P1, P2, nsamples):
    """Tier-S epsilon deep inside a stream, where the TRUE sequential state is out of reach for a CPU (it would
    have to demodulate everything before). raw_np: raw samples from P1 + P2 before the compared span to its
    end. A CPU-oracle run starts from power-on state, warms up over P1 samples, is compared with the stitched
    symbols `got_probe` (starting P2 before the span) to find its lock point k, has its Costas NCO turned to the
    stitched stream's lock point (p_phase -= k*pi/2, as the hand-off does; an odd k otherwise shows ~15 % of
    symbols off by more than 1 LSB because the timing detector reads only Q, timing.c:65), converges over the
    remaining P2 samples and is then a sequential trajectory at the same lock point: itself within the
    reference's own FMA-vs-strict distance of the true one (SURVEY.md finding 3). Returns (frac > 1 LSB,
    frac identical, symbols) over `got`, the stitched symbols of the span."""
    o = mk()
    o.process(raw_np[: 2 * P1], want_float=False)
    probe = o.process(raw_np[2 * P1: 2 * (P1 + 20000)], want_float=False)
    b = best_align(got_probe[:6000], probe.soft[:6100])
    if b is None or b[0] < 0.9:
        return None
    ph = np.float32(np.float64(o.state()["p_phase"]) - b[2] * np.pi / 2)
    o.set_state(p_phase=float(ph))
    o.process(raw_np[2 * (P1 + 20000): 2 * (P1 + P2)], want_float=False)
    w = o.process(raw_np[2 * (P1 + P2): 2 * (P1 + P2 + nsamples)], want_float=False)
    b2 = best_align(got[:4000], w.soft[:4100])
    if b2 is None or b2[0] < 0.9:
        return None
    off, k = b2[1], b2[2]
    n = min(len(got), len(w.soft)) - 64
    ww = w.soft[24 + off: 24 + off + n - 48].astype(np.int16)
    gg = got[24: 24 + n - 48].astype(np.int16)
    wi, wq = _quarter_turn(ww, k)
    d = np.maximum(np.abs(gg[:, 0] - wi), np.abs(gg[:, 1] - wq))
    return float((d > 1).mean()), float((d == 0).mean()), int(d.size)


def sharded_measure(a, cfg, label, rank, world, local):
    """ONE stream of a.stream_samples, strong scaling over ranks: rank r demodulates a consecutive run of chunks from
    its own time slice of the stream (+ the warm-up and overlap it over-reads); the quadrant scan exchanges boundary
    symbols and the hand-off one state per rank boundary over NCCL. Returns the record (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from meteor_demod_b200 import sharded, synth
    symrate, oqpsk, bps, order, interp = cfg
    N = a.stream_samples
    a.handoff = not (a.two_pass or a.single_pass)
    if a.warm <= 0:
        a.warm = 150000 if a.handoff else 400000
    plan = sharded.Plan(N, a.chunk, a.warm, 8192, interp)
    period = synth.baseband(FS, symrate=symrate, oqpsk=bool(oqpsk), periodic=True, seed=3).astype(np.complex64)
    # every rank holds only its time slice of the stream (+ the warm-up and overlap its chunks over-read)
    c0, c1 = sharded.split_chunks(plan.nchunks, world, rank)
    s0 = plan.start(c0)
    span = plan.start(c1 - 1) + plan.n_main + 2 * plan.overlap - s0
    raw = synth.device_long_stream(period, N, total=span, bps=bps, sps=FS / symrate, first=s0)
    kw = dict(chunk=a.chunk, warm=a.warm, overlap=8192, device=local, dist=dist if world > 1 else None, raw_first=s0,
              symrate=symrate, bps=bps, rrc_order=order, interp_factor=interp, two_pass=not a.single_pass,
              handoff=a.handoff, seed_carrier=a.seed_carrier, seed_nfft=a.seed_nfft, oqpsk=bool(oqpsk))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    sd = sharded.ShardedDemod(raw, N, **kw)
    res = None
    for _ in range(a.warmup):
        res = sd.run()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phases = {}
    with Clocks(local) as clk:
        e0.record()
        launches = 0
        for _ in range(a.steps):
            res = sd.run(phase_ms=phases)
            launches += res["launches"]
        e1.record()
        barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    nsym = torch.tensor([res["soft"].shape[0]], dtype=torch.int64, device="cuda")
    agree = torch.tensor([float(res["agreement"].min()) if res["agreement"].numel() else 1.0], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nsym)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
    total_ms = float(t.item())

    # ---- Tier-S epsilon, on EVERY rank (CPU oracle) -----------------------------------------------------------
    from oracle import pyoracle
    mine = torch.zeros(4, dtype=torch.float64, device="cuda")      # symbols compared, frac > 1 LSB, frac identical, samples
    mk = lambda: pyoracle.Oracle(symrate=symrate, oqpsk=oqpsk, bps=bps, order=order, interp=interp)
    if rank == 0:
        # the head of the stream against the TRUE sequential run (from sample 0, power-on state)
        ncheck = min(N, span, plan.boundary(min(plan.nchunks - 1, 40)))
        w = mk().process(raw[: 2 * ncheck].cpu().numpy(), want_float=False)
        got = res["soft"][: w.nsym].cpu().numpy()
        n = min(len(got), w.nsym) - 64
        dlt = np.abs(got[:n].astype(np.int16) - w.soft[:n].astype(np.int16)).max(axis=1)
        mine[:] = torch.tensor([n, float((dlt > 1).mean()), float((dlt == 0).mean()), ncheck], dtype=torch.float64)
    else:
        # deep in the stream the true sequential state is out of reach for a CPU (it would have to demodulate
        # everything before); two chunks of this rank's share are compared with a sequential run started from
        # power-on state PRE samples earlier -- a converged trajectory, itself within the reference's own
        # FMA-vs-strict distance of the true one (SURVEY.md finding 3)
        M = c1 - c0
        P1, P2 = 700_000, 800_000
        # this rank's share starts at the cut next to its first boundary (hand-off scheme: V samples after it)
        share0 = plan.boundary(c0) + (plan.overlap if a.handoff else 0) + min(64, plan.overlap // 4) - s0
        j = -(-(P1 + P2 - share0) // a.chunk)
        if 1 <= j < M - 3:
            A = share0 + j * a.chunk                                # local sample index where the compared span starts
            sps = FS / symrate
            gp = int(round((j * a.chunk - P2) / sps))               # stitched symbol index by rate (+- the search window)
            g0 = int(round(j * a.chunk / sps))
            nn = int(2 * a.chunk / sps) - 200
            b = tier_s_local(raw[2 * (A - P1 - P2): 2 * (A + 2 * a.chunk)].cpu().numpy(),
                             res["soft"][gp: gp + 6100].cpu().numpy(), res["soft"][g0: g0 + nn].cpu().numpy(),
                             mk, P1, P2, 2 * a.chunk)
            if b is not None:
                mine[:] = torch.tensor([b[2], b[0], b[1], 2 * a.chunk], dtype=torch.float64)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    ph_t = torch.tensor([phases.get(k, 0.0) / max(1, a.steps) for k in PHASES], dtype=torch.float64, device="cuda")
    ph_all = [torch.zeros_like(ph_t) for _ in range(world)]
    if world > 1:
        dist.all_gather(ph_all, ph_t)
        dist.all_reduce(ph_t, op=dist.ReduceOp.MAX)
    else:
        ph_all = [ph_t.clone()]
    line = None
    if rank == 0:
        per_rank = [{"rank": r, "symbols": int(v[0].item()), "frac_gt_1lsb": float(v[1].item()),
                     "frac_identical": float(v[2].item()), "samples": int(v[3].item()),
                     "against": "true sequential run (CPU oracle from sample 0)" if r == 0 else
                                "sequential CPU-oracle run started 1.5 M samples earlier and turned to the stitched stream's lock point (tier_s_local)"}
                    for r, v in enumerate(allr)]
        worst = max(p["frac_gt_1lsb"] for p in per_rank)
        ref_eps = None
        try:
            full = json.load(open(os.path.join(ROOT, "profiles", "r2_fma_vs_strict_eps.json")))
            key = {(72000, 0, 32, 5): "c1_qpsk72k_s16", (80000, 1, 32, 5): "c2_oqpsk80k_u8", (72000, 0, 64, 8): "c3_qpsk72k_s16_o64_L8"}.get((symrate, oqpsk, order, interp))
            if key:
                ref_eps = {"source": "profiles/r2_fma_vs_strict_eps.json (tools/measure_fma_vs_strict.py: the reference's own FMA build "
                                     "against its strict build, same input)",
                           "bench_stream": {k: full[key]["bench_stream"][k] for k in ("frac_gt_1lsb", "frac_identical", "quarter_turns_between_builds")},
                           "make_raw": {k: full[key]["make_raw"][k] for k in ("frac_gt_1lsb", "frac_identical", "quarter_turns_between_builds")}}
        except Exception:
            pass
        line = {"metric": "IQ Msamples/s", "value": N * a.steps / (total_ms * 1e-3) / 1e6, "unit": "Msamples/s",
                "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": total_ms / a.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "%s; ONE stream of %d samples time-sharded: %d chunks of %d, warm-up %d, overlap 8192, %s"
                           % (label, N, plan.nchunks, a.chunk, a.warm, ("state hand-off between chunks and ranks" if a.handoff else "single pass" if a.single_pass else "two-pass lock-point alignment")
                              + ("; chunks >= 1 start at a coarse carrier estimate (x^4 line, %d-point FFT) instead of the reference's sweep" % a.seed_nfft if a.seed_carrier else "")),
                           "parity": "Tier-S (statistical): chunks 0 and 1 bit-exact, later chunks see tier_s",
                           "l2": "stream (%.1f GB, %.1f GB resident per rank) larger than L2" % (N * (bps // 4) / 1e9, span * (bps // 4) / 1e9)},
                "gpu_launches": int(launches), "symbols_per_step": int(nsym.item()),
                "min_boundary_agreement": float(agree.item()),
                "tier_s": {"frac_gt_1lsb": worst, "per_rank": per_rank,
                           "reference_fma_vs_strict": ref_eps},
                "phase_ms": dict(zip(PHASES, [float(v) for v in ph_t.tolist()])),
                "phase_ms_per_rank": [[round(float(v), 2) for v in t_.tolist()] for t_ in ph_all],
                "kernel": sd.eng.d.kernel_name(), "chunks_per_rank": int(c1 - c0), "clocks": clk.summary()}
    return line, sd, raw, res


PHASES = ("pass_a_warm_up", "pass_b_owned", "quadrant_scan", "state_hand_off", "pass_c_final", "join")


def bench_sharded(a, cfg, label, rank, world, local, cpu_base):
    """--mode sharded: the time-sharded single stream as the whole bench line (+ the single-call C path, N = 1)."""
    import torch
    import torch.distributed as dist
    from meteor_demod_b200 import sharded
    symrate, oqpsk, bps, order, interp = cfg
    N = a.stream_samples
    line, sd, raw, res = sharded_measure(a, cfg, label, rank, world, local)
    if rank == 0:
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        if world == 1 and not a.no_e2e and not oqpsk:
            # end to end through the single C call a host makes (lrpt_sharded_process, what host/lrpt_demod --shard
            # runs): the recording in pinned HOST memory, H2D + three passes + join + D2H of the symbols inside
            # the timed region, device buffers allocated and freed by the call. Bounded to 1 GSample of host memory.
            ne = min(N, 1 << 30)
            host = torch.empty(2 * ne, dtype=raw.dtype, pin_memory=True)
            host.copy_(raw[: 2 * ne])
            torch.cuda.synchronize()
            hn = host.numpy()
            from meteor_demod_b200 import symbol_capacity
            out_pinned = torch.empty((symbol_capacity(ne, FS, symrate), 2), dtype=torch.int8, pin_memory=True)
            kwe = dict(chunk=a.chunk, warm=a.warm, overlap=8192, symrate=symrate, bps=bps, rrc_order=order,
                       interp_factor=interp, device=local, out=out_pinned.numpy(),
                       seed_nfft=a.seed_nfft if a.seed_carrier and 256 <= a.seed_nfft <= 16384 else 0)
            want = res["soft"].cpu().numpy() if ne == N else None
            sd.close()
            sd.eng.raw = None
            del raw, res
            torch.cuda.empty_cache()
            soft_e, rep_e = sharded.process_host(hn, **kwe)          # warm-up (module load, first allocations)
            reps, t0 = max(1, min(a.steps, 3)), time.perf_counter()
            for _ in range(reps):
                soft_e, rep_e = sharded.process_host(hn, **kwe)
            dt = (time.perf_counter() - t0) / reps
            line["e2e"] = {"value": ne / dt / 1e6, "unit": "Msamples/s", "ms_per_step": dt * 1e3, "samples": int(ne),
                           "h2d_bytes_per_step": int(ne * (bps // 4)), "d2h_bytes_per_step": int(2 * soft_e.shape[0]),
                           "api": "lrpt_sharded_process (pinned host buffers in and out; device buffers allocated and freed inside the call)",
                           "nchunks": rep_e["nchunks"],
                           "matches_device_path": bool(want is None or np.array_equal(soft_e, want))}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_relay(a, cfg, label, rank, world, local):
    """Exact time-sharding by state relay (meteor_demod_b200/relay.py): every stream is --samples * world
    long; rank r holds and demodulates samples [r*S, (r+1)*S) of ALL streams, group by group, after
    receiving each group's states from rank r-1 (NCCL send/recv, nothing else). Weak scaling: a rank's
    work and memory are fixed, the streams get longer with the ranks."""
    import torch
    import torch.distributed as dist
    from meteor_demod_b200 import relay, synth
    symrate, oqpsk, bps, order, interp = cfg
    G, S = a.groups, a.samples
    per = a.streams // G
    B = per * G
    period = synth.baseband(FS, symrate=symrate, oqpsk=bool(oqpsk), periodic=True, seed=3).astype(np.complex64)
    # one long synthetic batch; every rank keeps only its own time slice resident
    slices = []
    for g in range(G):
        full = synth.device_streams(period, per, S * world, bps=bps, sps=FS / symrate, seed=11 + g, device="cuda")
        slices.append(full[:, 2 * rank * S: 2 * (rank + 1) * S].contiguous())
        if g == 0:
            head = full[0].cpu().numpy() if rank == 0 else None
        del full
    kw = dict(symrate=symrate, oqpsk=oqpsk, bps=bps, rrc_order=order, interp_factor=interp)
    st = torch.cuda.Stream()
    groups = [relay.GpuGroup(per, device=local, stream=st, **kw) for _ in range(G)]
    d = dist if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        with torch.cuda.stream(st):
            if rank == 0:
                for e in groups:
                    e.reset()
            return relay.relay(groups, slices, dist=d)

    res = None
    for _ in range(a.warmup):
        res = step()
    barrier()
    l0 = sum(e.d.launch_count() for e in groups)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with Clocks(local) as clk:
        barrier()
        e0.record(st)
        for _ in range(a.steps):
            res = step()
        e1.record(st)
        barrier()
    launches = sum(e.d.launch_count() for e in groups) - l0
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    # parity on the benchmarked data: stream 0 of group 0, symbols of all ranks in time order, vs the oracle
    soft0, n0 = res[0]
    mine = soft0[0, : int(n0[0].item())].contiguous()
    cnt = torch.tensor([mine.shape[0]], dtype=torch.int64, device="cuda")
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    if world > 1:
        dist.all_gather(cnts, cnt)
    else:
        cnts = [cnt]
    cap = int(max(int(c.item()) for c in cnts))
    pad = torch.zeros((cap, 2), dtype=torch.int8, device="cuda")
    pad[: mine.shape[0]] = mine
    parts = [torch.zeros_like(pad) for _ in range(world)]
    if world > 1:
        dist.all_gather(parts, pad)
    else:
        parts = [pad]
    if rank == 0:
        from oracle import pyoracle
        got = np.concatenate([p[: int(c.item())].cpu().numpy() for p, c in zip(parts, cnts)])
        w = pyoracle.Oracle(symrate=symrate, oqpsk=oqpsk, bps=bps, order=order, interp=interp).process(head, want_float=False)
        ok = bool(got.shape[0] == w.nsym and np.array_equal(got, w.soft))
        line = {"metric": "IQ Msamples/s", "value": world * B * S * a.steps / (total_ms * 1e-3) / 1e6, "unit": "Msamples/s",
                "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": total_ms / a.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "%s; %d streams x %d samples, time axis split over %d rank(s): %d samples per rank, "
                                       "%d groups of %d streams relayed with their complete state (%d bytes per group and boundary)"
                                       % (label, B, S * world, world, S, G, per, groups[0].d.states_size()),
                           "parity": "bit-exact vs strict-IEEE reference (stream 0 checked across all ranks)",
                           "l2": "inputs (%.1f GB/GPU) larger than L2" % (B * S * (bps // 4) / 1e9)},
                "kernel": groups[0].d.kernel_name(), "gpu_launches": int(launches), "oracle_check_stream0_all_ranks": ok,
                "clocks": clk.summary()}
        print(json.dumps(line))
    for e in groups:
        e.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
