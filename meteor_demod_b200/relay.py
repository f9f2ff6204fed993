"""Exact (Tier-E) time-sharding of a BATCH of streams across GPUs by state relay.

SURVEY.md section 8e: the recurrence does not shard in time -- each symbol needs the exact float
state its predecessor left -- so the only exact way to spread the TIME axis of a stream over ranks
is a relay: rank r demodulates samples [r*S, (r+1)*S) after it has received the complete state
(lrpt_state_t + FIR delay line, the reference's statics) from rank r-1, and hands its own final
state to rank r+1. One stream alone gains nothing from that (the ranks take turns), but a batch
does: the streams are cut into G groups and the groups travel down the ranks as a pipeline, so
in steady state every rank is busy, every rank holds only ITS time slice of the input and output
(memory scales with the ranks), and the result is bit-identical to one rank demodulating every
stream from start to end. The exchange is exactly one state buffer per group and boundary
(nstreams_g * (88 + 8*(taps-1)) bytes), sent with torch.distributed send/recv: NCCL over NVLink
between GPUs, gloo in the CPU tests. Nothing else is communicated; the delay line inside the state
is the filter-length overlap.

Host plumbing only; demodulation is liblrpt_b200.so (`GpuGroup`) -- the tests plug in the CPU
oracle instead (`tests/test_relay.py`) to check the protocol without a GPU.
"""
import torch


class GpuGroup:
    """One group of streams on this rank's GPU: a Demod handle plus its state buffer."""

    def __init__(self, nstreams, device=0, stream=None, **cfg):
        """stream: the torch.cuda.Stream all of this rank's relay work is ordered on (kernels, state copies
        and -- when the caller makes it current -- the NCCL send/recv). The legacy default stream cannot
        be used: its handle is NULL, which the C ABI reads as "the handle's own stream"."""
        from .demod import Demod
        self.d = Demod(nstreams=nstreams, device=device, **cfg)
        self.dev = torch.device("cuda", device)
        self.stream = stream if stream is not None else torch.cuda.Stream(self.dev)
        self.buf = torch.empty(self.d.states_size(), dtype=torch.uint8, device=self.dev)
        self.cap = None
        self.soft = self.nsym = None

    def state_buffer(self):
        return self.buf

    def import_states(self):
        self.d.import_states_device(self.buf, check=True, stream=self.stream)

    def export_states(self):
        self.d.export_states_device(self.buf, stream=self.stream)

    def process(self, raw):
        """raw: device tensor [nstreams, 2*S] (this rank's time slice). Returns (soft [n,cap,2] int8, counts [n])."""
        n = raw.shape[1] // 2
        cap = (self.d.capacity(n) + 7) // 8 * 8
        if self.cap != cap:
            self.cap = cap
            self.soft = torch.empty((raw.shape[0], 2 * cap), dtype=torch.int8, device=self.dev)
            self.nsym = torch.zeros(raw.shape[0], dtype=torch.int32, device=self.dev)
        self.d.process_device(raw, self.soft, nsym=self.nsym, stream=self.stream)
        return self.soft.view(raw.shape[0], cap, 2), self.nsym

    def reset(self):
        self.d.reset(stream=self.stream, asynchronous=True)

    def close(self):
        self.d.close()


def relay(groups, slices, dist=None):
    """Run this rank's part of the relay.

    groups: G group engines (GpuGroup or anything with state_buffer / import_states / export_states /
    process); slices: G raw arrays, this rank's time slice of each group's streams. Rank 0 starts every
    group from the state its engine is in (power-on after creation, or wherever an earlier call left it);
    every other rank first receives the group's state from its predecessor. Returns the G results of
    `process` in order. The caller keeps them sharded by time: rank r's symbols follow rank r-1's.
    With GpuGroup engines call this under `with torch.cuda.stream(s)` for the stream s the groups were
    created with, so that the NCCL transfers are ordered with the kernels."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    out = []
    for g, (eng, raw) in enumerate(zip(groups, slices)):
        if rank > 0:
            dist.recv(eng.state_buffer(), src=rank - 1, tag=g)
            eng.import_states()
        out.append(eng.process(raw))
        if rank < world - 1:
            eng.export_states()
            # A stream-ordered (not overlapped) send: an NCCL kernel waiting for its peer holds an SM, and
            # the demodulator's one-CTA-per-SM grid would then need a second wave for its last CTA.
            dist.send(eng.state_buffer(), dst=rank + 1, tag=g)
    return out
