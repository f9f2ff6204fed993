"""Exact time-sharding by state relay (meteor_demod_b200/relay.py): a batch of streams, time axis split
over ranks, complete state handed rank to rank. CPU: the protocol with the oracle as the engine, single
process and as a 2-rank gloo job. GPU: the CUDA engine relayed over two time slices in one process
equals the oracle run of the whole streams, bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

CFG = dict(symrate=72000, oqpsk=0, bps=16, order=32, interp=5)
NS, N, GROUPS = 6, 60_000, 3


def make_batch():
    from meteor_demod_b200 import synth
    return np.stack([synth.make_raw(N, cfo_hz=30.0 * (s + 1), seed=500 + s) for s in range(NS)])


class OracleGroup:
    """CPU stand-in for relay.GpuGroup: a list of oracles; the state buffer is a flat float64 tensor."""

    def __init__(self, nstreams):
        from oracle import pyoracle
        self.o = [pyoracle.Oracle(**CFG) for _ in range(nstreams)]
        self.keys = sorted(self.o[0].state())
        self.hlen = self.o[0].history().size
        self.buf = torch.zeros(nstreams * (len(self.keys) + self.hlen), dtype=torch.float64)

    def state_buffer(self):
        return self.buf

    def export_states(self):
        rows = [np.concatenate([[float(o.state()[k]) for k in self.keys], o.history().reshape(-1).astype(np.float64)])
                for o in self.o]
        self.buf.copy_(torch.from_numpy(np.concatenate(rows)))

    def import_states(self):
        a = self.buf.numpy().reshape(len(self.o), -1)
        for o, row in zip(self.o, a):
            cur = o.state()
            o.set_state(**{k: type(cur[k])(v) for k, v in zip(self.keys, row[: len(self.keys)])})
            o.set_history(row[len(self.keys):].astype(np.float32))

    def process(self, raw):
        return [o.process(r, want_float=False).soft for o, r in zip(self.o, raw)]


def sequential():
    from oracle import pyoracle
    return [pyoracle.Oracle(**CFG).process(r, want_float=False).soft for r in make_batch()]


def run_rank(rank, world, dist=None):
    """This rank's slice of every group through relay(); returns per-stream soft symbols of the slice."""
    from meteor_demod_b200 import relay
    raw = make_batch()
    S = N // world
    per = NS // GROUPS
    sl = raw[:, 2 * rank * S: 2 * (N if rank == world - 1 else (rank + 1) * S)]
    groups = [OracleGroup(per) for _ in range(GROUPS)]
    res = relay.relay(groups, [sl[g * per:(g + 1) * per] for g in range(GROUPS)], dist=dist)
    return [s for grp in res for s in grp]


def test_oracle_state_roundtrip_is_exact():
    """The float64 state buffer of the CPU stand-in carries every float32/int field exactly."""
    raw = make_batch()
    a, b = OracleGroup(1), OracleGroup(1)
    first = a.process(raw[:1, : 2 * 20_000])
    a.export_states()
    b.buf.copy_(a.buf)
    b.import_states()
    rest = b.process(raw[:1, 2 * 20_000:])
    assert np.array_equal(np.concatenate([first[0], rest[0]]), sequential()[0])


def _rank_main(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parts = run_rank(rank, world, dist)
    for s, p in enumerate(parts):
        np.save(os.path.join(out_dir, "r%d_s%d.npy" % (rank, s)), p)
    dist.destroy_process_group()


def test_two_rank_gloo_relay_is_bit_exact(tmp_path):
    """2 ranks, gloo: rank 1 demodulates the second half of every stream from the states rank 0 sends;
    rank-0 symbols followed by rank-1 symbols equal the sequential demodulation of the whole stream."""
    import torch.multiprocessing as mp
    port = 31500 + os.getpid() % 2000
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    want = sequential()
    for s in range(NS):
        got = np.concatenate([np.load(tmp_path / ("r%d_s%d.npy" % (r, s))) for r in range(2)])
        assert np.array_equal(got, want[s]), s


@pytest.mark.gpu
def test_gpu_relay_over_two_slices_equals_oracle(lib):
    """CUDA engine: slice 0 on one set of handles, states exported to device buffers, imported into a SECOND
    set of handles (what the next rank does after ncclRecv), slice 1 there. Bit-exact vs the oracle."""
    from meteor_demod_b200 import relay
    raw = make_batch()
    t = torch.from_numpy(raw).cuda()
    S = (N // 2) // 16 * 16
    per = NS // GROUPS
    kw = dict(symrate=72000, bps=16, rrc_order=32, interp_factor=5)
    st = torch.cuda.Stream()
    first = [relay.GpuGroup(per, stream=st, **kw) for _ in range(GROUPS)]
    second = [relay.GpuGroup(per, stream=st, **kw) for _ in range(GROUPS)]
    want = sequential()
    torch.cuda.synchronize()
    for g in range(GROUPS):
        rows = slice(g * per, (g + 1) * per)
        a0, a1 = t[rows, : 2 * S].contiguous(), t[rows, 2 * S:].contiguous()
        torch.cuda.synchronize()
        with torch.cuda.stream(st):
            s0, n0 = first[g].process(a0)
            first[g].export_states()
            second[g].state_buffer().copy_(first[g].state_buffer())
            second[g].import_states()
            s1, n1 = second[g].process(a1)
        st.synchronize()
        s0, n0, s1, n1 = s0.cpu().numpy(), n0.cpu().numpy(), s1.cpu().numpy(), n1.cpu().numpy()
        for i in range(per):
            got = np.concatenate([s0[i, : n0[i]], s1[i, : n1[i]]])
            assert np.array_equal(got, want[g * per + i]), (g, i)
    for e in first + second:
        e.close()
