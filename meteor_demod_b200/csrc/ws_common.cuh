/*
 * ws_common.cuh -- pieces shared by the warp-specialised kernels (demod_ws.cu, demod_spec.cu):
 * mbarrier wrappers and the exact timing-NCO search of the recurrence warp.
 */
#pragma once
#include "demod_core.cuh"

namespace lrpt {

constexpr int WS_WARPS     = 16;   /* warps per CTA                                                        */
constexpr int WS_PRODUCERS = 12;   /* FIR warps: the ones NOT on SM sub-partition 0 (warp id % 4 != 0)     */
constexpr int WS_THREADS   = 32*WS_WARPS;
constexpr int NCO_CHUNK    = 8;    /* timing sub-steps evaluated per branch in the exhaustive search        */
constexpr int WS_MAX_SAMPLES = 1 << 26;   /* per launch; keeps sub-step indices in int32 */

/* ------------------------------------------------------------- mbarrier ---- */

LRPT_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

LRPT_DEV void mbar_init(uint64_t *bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}

LRPT_DEV void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

LRPT_DEV void mbar_wait(uint64_t *bar, unsigned parity)
{
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra DONE_%=;\n\t"
		"bra WAIT_%=;\n\t"
		"DONE_%=:\n\t}"
		:: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

/* Producers wait here for the recurrence warp most of the time; back off between polls so
 * their polling does not take issue slots away from that warp, which is the critical path. */
LRPT_DEV void mbar_wait_relaxed(uint64_t *bar, unsigned parity)
{
	unsigned done;
	for (;;) {
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
		if (done) break;
		__nanosleep(200);
	}
}

template <int NWARPS = WS_PRODUCERS>
LRPT_DEV void producers_sync()
{
	asm volatile("bar.sync 1, %0;" :: "n"(32*NWARPS) : "memory");
}

/* ------------------------------------------------------------- two-warp recurrence --
 *
 * One warp alone is bound by ISSUING the ~650 instructions of a symbol step, one dependent chain after the
 * other (profiles/r2_ws_single_stream_regions.txt: 1040 cycles per symbol, IPC 0.6). The step is therefore
 * spread over two warps on two SM sub-partitions, and so is the STATE (demod.c:33-43 reordered per variable,
 * never per value):
 *
 *   timing warp "R"   timing NCO search, delay-line pick, DC bias tracker, retime            timing.c:32-95, agc.c:16
 *                     owns t_phase, t_freq, t_prev, t_dual, bias
 *   loop warp   "D"   scaling by the AGC gain, NCO mix, AGC magnitude + gain update, Costas error / phase /
 *                     frequency / lock detector / sweep, next oscillator values, int8 store
 *                     owns gain, p_phase, p_freq, p_err, locked, updown, oq_inphase           agc.c:19-22, pll.c:51-130
 *
 * Per symbol R sends the bias-free filter output x = y - bias, D answers at once with the mixed Q component
 * (all the timing loop ever reads, timing.c:65) and only then runs its long chains -- they overlap R's retime,
 * NCO search and next delay-line pick. The loop-carried cycle is R: retime -> search -> pick -> bias | hand-over |
 * D: scale, mix | hand-over.
 *
 * Mailboxes: one slot per lane and direction in shared memory, payload and round number in ONE vector store /
 * load, so no fence separates data from flag. Rounds are warp-uniform and lock-step: every lane sends a
 * message every round (meta = 0: no timing event for this lane), so both warps count rounds alike. A slot is
 * reused only after its reader has answered (ping-pong), hence single buffering.
 */
struct __align__(16) MsgRD { float xr, xi; unsigned meta; int seq; };   /* meta: see msg_meta */
struct __align__(8)  MsgDR { float oim; int seq; };
/* loop warp -> egress warp "E": one symbol (or none) per lane and round, in a ring of EG_RING rounds so that the
 * loop warp never waits for the stores; meta as in MsgRD plus bit 31 = the PLL had locked once by this symbol */
constexpr int EG_RING = 16;        /* rounds; E acknowledges every 8 */
struct __align__(16) MsgDE { float ore, oim; unsigned meta; int seq; };

/* meta of a timing-warp message: bits 1-0 = 0 no event this round, 1 + half otherwise (half: 0 QPSK symbol,
 * 1 / 2 the OQPSK arms, timing.c:41-57); the event's sub-step index (< 2^30 per launch) above */
constexpr unsigned MSG_STOP = 0xffffffffu;
LRPT_DEV unsigned msg_meta(int half, int Qx) { return ((unsigned)Qx << 2) | (unsigned)(half + 1); }

/* Hand-over between the timing warp and the loop warp: the payload is an ordinary shared-memory store, the signal an
 * mbarrier the 32 writer lanes arrive on (release) and the reader waits for (acquire, try_wait suspends the warp in
 * hardware). Measured round trip between two warps on two sub-partitions (tools/mb/pingpong.cu,
 * profiles/r2_mailbox_round_trip.txt): 168 cycles, against 277-301 for any flavour of polled volatile load. */
LRPT_DEV void chan_signal(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

LRPT_DEV void chan_wait(uint64_t *bar, unsigned parity)
{
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"CW_%=:\n\t"
		"mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n\t"
		"@!p bra CW_%=;\n\t}"
		:: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

LRPT_DEV void mbox_put4(void *slot, float a, float b, unsigned c, int seq)
{
	asm volatile("st.volatile.shared.v4.b32 [%0], {%1, %2, %3, %4};"
	             :: "r"(smem_u32(slot)), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(c), "r"(seq) : "memory");
}

LRPT_DEV void mbox_put2(void *slot, float a, int seq)
{
	asm volatile("st.volatile.shared.v2.b32 [%0], {%1, %2};" :: "r"(smem_u32(slot)), "r"(__float_as_uint(a)), "r"(seq) : "memory");
}

/* spin until every lane's slot carries round `seq` (the writer stores all 32 slots with one instruction, so the
 * lanes see it within a few cycles of each other; a warp-uniform loop needs no reconvergence barrier) */
LRPT_DEV uint4 mbox_get4(const void *slot, int seq)
{
	uint4 v;
	do {
		asm volatile("ld.volatile.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
		             : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(slot)) : "memory");
	} while (__any_sync(0xffffffffu, (int)v.w != seq));
	return v;
}

LRPT_DEV float mbox_get2(const void *slot, int seq)
{
	uint2 v;
	do {
		asm volatile("ld.volatile.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(smem_u32(slot)) : "memory");
	} while (__any_sync(0xffffffffu, (int)v.y != seq));
	return __uint_as_float(v.x);
}

/* ------------------------------------------------------------- timing NCO -- */

/* threshold the NCO phase is compared with next: timing.c:37 (QPSK) / :51 (OQPSK, state*pi) */
LRPT_DEV float nco_threshold(const Loop &r, const lrpt_consts_t &c)
{
	return c.oqpsk ? __fmul_rn((float)r.t_dual, kPiF) : kTwoPiF;
}


/*
 * One chunk of advance_timeslot (timing.c:32-38) / advance_timeslot_dual (:41-57):
 * up to NCO_CHUNK float additions of the NCO step, stopping at the first sum that
 * reaches the threshold. Fast path (step > 0, at least NCO_CHUNK sub-steps left):
 * the sums are non-decreasing, so the number of sums below the threshold IS the
 * index of the first crossing -- no bit scan, no data-dependent branch inside.
 * Returns true when a crossing was found; Qx = its sub-step index.
 */
LRPT_DEV bool nco_chunk(Loop &r, const lrpt_consts_t &c, int &Q, int Qend, int &Qx, int &half)
{
	const float f = r.t_freq;
	const float thr = c.oqpsk ? __fmul_rn((float)r.t_dual, kPiF) : kTwoPiF;
	const int limit = Qend - Q;                                     /* >= 1 */
	bool found;
	if (f > 0.0f && limit >= NCO_CHUNK) {
		float ph[NCO_CHUNK];
		float acc = r.t_phase;
		int below = 0;
#pragma unroll
		for (int j = 0; j < NCO_CHUNK; j++) {
			acc = __fadd_rn(acc, f);
			ph[j] = acc;
			below += (acc >= thr) ? 0 : 1;
		}
		float sel = ph[NCO_CHUNK-1];
#pragma unroll
		for (int j = NCO_CHUNK - 2; j >= 0; j--) sel = (ph[j] >= thr) ? ph[j] : sel;
		r.t_phase = sel;                                            /* first sum >= thr, or the last sum */
		found = below < NCO_CHUNK;
		Qx = Q + below;
		Q += found ? below + 1 : NCO_CHUNK;
	} else {
		/* end of the block, or a non-positive step from an imported state: one sub-step at a time */
		found = false;
		const int n = min(limit, NCO_CHUNK);
		float acc = r.t_phase;
		for (int j = 0; j < n && !found; j++) {
			acc = __fadd_rn(acc, f);
			Qx = Q; Q++;
			found = acc >= thr;
		}
		r.t_phase = acc;
	}
	if (found && c.oqpsk) { half = r.t_dual; r.t_dual = (r.t_dual % 2) + 1; }
	return found;
}

/*
 * Run the timing NCO from sub-step Q to its next crossing. In steady state the number of sub-steps
 * between two crossings is pi-or-2pi / t_center within +-1 (the NCO step is clamped to +-2^-12 of its
 * centre, timing.c:7,84), so the launch fixes n0 = a multiple of 4 safely below that count: n0 plain
 * float adds (warp-uniform loop, no compares), then NCO_WINDOW tested sums, branch-free. The search
 * is accepted only if the crossing provably lies inside the window (no sum before it had crossed,
 * one inside did); otherwise -- acquisition transients, end of block, imported odd states -- the
 * exhaustive chunks above run from the untouched phase. The sums are the reference's adds, in order.
 */
constexpr int NCO_WINDOW = 8;

LRPT_DEV bool nco_to_crossing(Loop &r, const lrpt_consts_t &c, int n0, int &Q, int q1, int Qend,
                              int &Qx, int &half)
{
	const float f = r.t_freq;
	const float thr = c.oqpsk ? __fmul_rn((float)r.t_dual, kPiF) : kTwoPiF;
	float p = r.t_phase;
	for (int b = 0; b < n0; b += 4) {                               /* n0 is warp-uniform and a multiple of 4 */
		p = __fadd_rn(p, f); p = __fadd_rn(p, f); p = __fadd_rn(p, f); p = __fadd_rn(p, f);
	}
	float s[NCO_WINDOW];
	float acc = p;
	int below = 0;
#pragma unroll
	for (int j = 0; j < NCO_WINDOW; j++) {
		acc = __fadd_rn(acc, f);
		s[j] = acc;
		below += (acc >= thr) ? 0 : 1;
	}
	float sel = s[NCO_WINDOW-1];
#pragma unroll
	for (int j = NCO_WINDOW - 2; j >= 0; j--) sel = (s[j] >= thr) ? s[j] : sel;
	if (f > 0.0f && !(p >= thr) && below < NCO_WINDOW && Q + n0 + NCO_WINDOW <= Qend) {
		r.t_phase = sel;
		Qx = Q + n0 + below; Q = Qx + 1;
		if (c.oqpsk) { half = r.t_dual; r.t_dual = (r.t_dual % 2) + 1; }
		return true;
	}
	bool found = false;
	while (!found && Q < q1) found = nco_chunk(r, c, Q, Qend, Qx, half);
	return found;
}

/*
 * nco_to_crossing in two parts, for callers that want the arithmetic of the search (a chain of dependent
 * float adds) in the same straight-line block as other work: nco_try computes the windowed search from the
 * CURRENT timing state without touching it, nco_commit accepts it under the same proof condition as above
 * or runs the exhaustive chunks.
 */
struct NcoTry { float sel; int below; bool ok; };

LRPT_DEV NcoTry nco_try(const Loop &r, const lrpt_consts_t &c, int n0, int Q, int Qend)
{
	const float f = r.t_freq;
	const float thr = c.oqpsk ? __fmul_rn((float)r.t_dual, kPiF) : kTwoPiF;
	float p = r.t_phase;
	for (int b = 0; b < n0; b += 4) {                               /* n0 is warp-uniform and a multiple of 4 */
		p = __fadd_rn(p, f); p = __fadd_rn(p, f); p = __fadd_rn(p, f); p = __fadd_rn(p, f);
	}
	float s[NCO_WINDOW];
	float acc = p;
	int below = 0;
#pragma unroll
	for (int j = 0; j < NCO_WINDOW; j++) {
		acc = __fadd_rn(acc, f);
		s[j] = acc;
		below += (acc >= thr) ? 0 : 1;
	}
	float sel = s[NCO_WINDOW-1];
#pragma unroll
	for (int j = NCO_WINDOW - 2; j >= 0; j--) sel = (s[j] >= thr) ? s[j] : sel;
	NcoTry t;
	t.sel = sel; t.below = below;
	t.ok = f > 0.0f && !(p >= thr) && below < NCO_WINDOW && Q + n0 + NCO_WINDOW <= Qend;
	return t;
}

LRPT_DEV bool nco_commit(Loop &r, const lrpt_consts_t &c, const NcoTry &t, int n0, int &Q, int q1, int Qend,
                         int &Qx, int &half)
{
	if (t.ok) {
		r.t_phase = t.sel;
		Qx = Q + n0 + t.below; Q = Qx + 1;
		if (c.oqpsk) { half = r.t_dual; r.t_dual = (r.t_dual % 2) + 1; }
		return true;
	}
	bool found = false;
	while (!found && Q < q1) found = nco_chunk(r, c, Q, Qend, Qx, half);
	return found;
}

/*
 * The same search with a 4-sum tested window behind n0 plain adds, n0 any warp-uniform count
 * (demod_lane.cu: n0 = nominal - 2, so the window is the four counts a locked loop produces).
 */
LRPT_DEV bool nco_to_crossing4(Loop &r, const lrpt_consts_t &c, int n0, int &Q, int q1, int Qend,
                               int &Qx, int &half)
{
	const float f = r.t_freq;
	const float thr = c.oqpsk ? __fmul_rn((float)r.t_dual, kPiF) : kTwoPiF;
	float p = r.t_phase;
	/* n0 plain adds, n0 warp-uniform: binary decomposition, so that the only control flow is a handful
	 * of uniform branches around straight-line runs of 16, 8, 4, 2 and 1 adds */
#define NCO_RUN(N) do { _Pragma("unroll") for (int j_ = 0; j_ < (N); j_++) p = __fadd_rn(p, f); } while (0)
	for (int b = n0 >> 5; b > 0; b--) NCO_RUN(32);
	if (n0 & 16) NCO_RUN(16);
	if (n0 & 8) NCO_RUN(8);
	if (n0 & 4) NCO_RUN(4);
	if (n0 & 2) NCO_RUN(2);
	if (n0 & 1) NCO_RUN(1);
#undef NCO_RUN
	const float s0 = __fadd_rn(p, f), s1 = __fadd_rn(s0, f), s2 = __fadd_rn(s1, f), s3 = __fadd_rn(s2, f);
	const bool c0 = s0 >= thr, c1 = s1 >= thr, c2 = s2 >= thr, c3 = s3 >= thr;
	if (f > 0.0f && !(p >= thr) && (c0 || c1 || c2 || c3) && Q + n0 + 4 <= Qend) {
		const int first = c0 ? 0 : c1 ? 1 : c2 ? 2 : 3;             /* sums are non-decreasing: first crossing */
		r.t_phase = c0 ? s0 : c1 ? s1 : c2 ? s2 : s3;
		Qx = Q + n0 + first; Q = Qx + 1;
		if (c.oqpsk) { half = r.t_dual; r.t_dual = (r.t_dual % 2) + 1; }
		return true;
	}
	bool found = false;
	while (!found && Q < q1) found = nco_chunk(r, c, Q, Qend, Qx, half);
	return found;
}

/* ===================== loop warp "D" (two-warp recurrence above), one lane per stream =====================
 * Args: the kernel's argument struct (states, nsamples, first_stream). Every lane runs every round in straight-line
 * code; a lane without an event computes on zeros and commits nothing. Symbols leave through the egress warp. */
template <bool OQ, class Args>
LRPT_DEV void loop_warp_run(const lrpt_consts_t &c, const Args &a, const float *lut, MsgRD *r2d, MsgDR *d2r, uint64_t *chan,
                            MsgDE *d2e, volatile int *eack, int lane, bool active, int g0)
{
	const int local = g0 + lane;
	const int sid = a.first_stream + local;
	Loop r;
	loop_load(r, a.states[a.first_stream + (active ? local : g0)]);
	/* oscillator values of the first mix: fast_sin/fast_cos(-p_phase), pll.c:53-54 */
	float os = fast_sin(-r.p_phase), oc = fast_cos(-r.p_phase);
	__syncwarp();
	int round = 1;
	for (;; round++) {
		chan_wait(&chan[0], (unsigned)(round - 1) & 1u);
		uint4 m;
		asm volatile("ld.volatile.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
		             : "=r"(m.x), "=r"(m.y), "=r"(m.z), "=r"(m.w) : "r"(smem_u32(&r2d[lane])) : "memory");
		if (m.z == MSG_STOP) break;                             /* the timing warp has finished the block (all lanes alike) */
		const bool valid = (m.z & 3u) != 0u;
		const int half = (int)(m.z & 3u) - 1;
		Pend pd;
		pd.half = half;
		pd.sr = __fmul_rn(__uint_as_float(m.x), r.gain);        /* agc.c:19: scaled with the OLD gain */
		pd.si = __fmul_rn(__uint_as_float(m.y), r.gain);
		const float mre = __fsub_rn(__fmul_rn(pd.sr, oc), __fmul_rn(pd.si, os));   /* pll.c:60 / demod.c:67,73 */
		const float mim = __fadd_rn(__fmul_rn(pd.sr, os), __fmul_rn(pd.si, oc));
		pd.oim = mim;
		pd.ore = OQ ? r.oq_inphase : mre;                       /* demod.c:76: the remembered I arm */
		asm volatile("st.volatile.shared.f32 [%0], %1;" :: "r"(smem_u32(&d2r[lane].oim)), "f"(mim) : "memory");   /* what retime() reads (timing.c:65) */
		chan_signal(&chan[1]);                                  /* answer first, the long chains after */

		Loop t = r;
		Osc next;
		const bool ok = step_deferred_fast<OQ>(t, c, lut, pd, next);
		if (__any_sync(0xffffffffu, valid && !ok)) {            /* a shortcut was not provably exact: statement by statement */
			if (valid && !ok) { t = r; step_deferred_exact<OQ>(t, c, lut, pd, next); }
			__syncwarp();
		}
		const bool arm_i = OQ && half == 1;
		r.gain = valid ? t.gain : r.gain; r.p_phase = valid ? t.p_phase : r.p_phase;
		r.p_freq = valid ? t.p_freq : r.p_freq; r.p_err = valid ? t.p_err : r.p_err;
		r.locked = valid ? t.locked : r.locked; r.locked_once = valid ? t.locked_once : r.locked_once;
		r.updown = valid ? t.updown : r.updown;
		os = valid ? next.s : os; oc = valid ? next.co : oc;
		r.oq_inphase = (valid && arm_i) ? mre : r.oq_inphase;   /* demod.c:66-71 */
		/* the symbol goes to the egress warp (quantise, stores, counters). Ring slot round % EG_RING is free again once
		 * E has acknowledged round - EG_RING; checked once per ring revolution */
		if ((round & 7) == 0) {                                 /* about to reuse the slots of rounds round-16 .. round-9 */
			while (__any_sync(0xffffffffu, eack[lane] < round - 9)) { }
		}
		mbox_put4(&d2e[(round & (EG_RING - 1))*32 + lane], pd.ore, pd.oim,
		          (valid && !arm_i) ? ((m.z & 0x7ffffffcu) | (r.locked_once ? 0x80000001u : 1u)) : 0u, round);
	}
	if ((round & 7) == 0) {
		while (__any_sync(0xffffffffu, eack[lane] < round - 9)) { }
	}
	mbox_put4(&d2e[(round & (EG_RING - 1))*32 + lane], 0.0f, 0.0f, MSG_STOP, round);
	if (active) {                                               /* this warp's part of the state (lrpt_state_t) */
		lrpt_state_t &s = a.states[sid];
		s.agc_gain = r.gain; s.p_phase = r.p_phase; s.p_freq = r.p_freq; s.p_err = r.p_err;
		s.p_locked = r.locked; s.p_locked_once = r.locked_once; s.p_updown = r.updown;
		s.oq_inphase = r.oq_inphase;
	}
}

/* ===================== egress warp "E": int8 quantiser, stores, symbol counters (main.c:305-312) =====================
 * Receives every round's message from the loop warp through a ring (so D never waits for global stores) and owns
 * nsymbols, first_lock_symbol, the append cursor and the optional float / index side outputs. */
template <class Args>
LRPT_DEV void egress_warp_run(const Args &a, const MsgDE *d2e, volatile int *eack, int lane, bool active, int g0)
{
	const int local = g0 + lane;
	const int sid = a.first_stream + local;
	long long nsymbols = 0, first_lock = -1;
	unsigned off = 0, nsym = 0;
	char2 *out = nullptr; float2 *outf = nullptr; uint32_t *outq = nullptr;
	if (active) {
		nsymbols = a.states[sid].nsymbols;
		first_lock = a.states[sid].first_lock_symbol;
		off = a.out_off ? a.out_off[local] : 0u;
		out = reinterpret_cast<char2 *>(a.soft + (size_t)local*a.soft_stride);
		if (a.symf) outf = reinterpret_cast<float2 *>(reinterpret_cast<char *>(a.symf) + (size_t)local*a.symf_stride);
		if (a.symq) outq = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(a.symq) + (size_t)local*a.symq_stride);
	}
	for (int round = 1; ; round++) {
		const uint4 m = mbox_get4(&d2e[(round & (EG_RING - 1))*32 + lane], round);
		if (m.z == MSG_STOP) break;
		if ((round & 7) == 7) eack[lane] = round;               /* everything up to this round has been read */
		if (m.z & 1u) {
			const float ore = __uint_as_float(m.x), oim = __uint_as_float(m.y);
			if ((m.z & 0x80000000u) && first_lock < 0) first_lock = nsymbols;
			if (off + nsym < a.cap) {
				out[off + nsym] = make_char2((signed char)quantise(ore), (signed char)quantise(oim));
				if (outf) outf[off + nsym] = make_float2(ore, oim);
				if (outq) outq[off + nsym] = a.q_base + ((m.z & 0x7fffffffu) >> 2);
			}
			nsym++; nsymbols++;
		}
		__syncwarp();
	}
	if (active) {
		lrpt_state_t &s = a.states[sid];
		s.nsamples += a.nsamples;
		s.nsymbols = nsymbols;
		s.first_lock_symbol = first_lock;
		if (a.nsym_out) a.nsym_out[local] = nsym;
		if (a.out_off) a.out_off[local] = off + nsym;
	}
}

/* ===================== timing warp "R": the part both kernels share =====================
 * The windowed timing-NCO search (nco_to_crossing) as straight-line code for a launch-constant n0: no loop,
 * the first crossing picked by a select tree over the count of sums below the threshold. */
template <int N0>
LRPT_DEV NcoTry nco_try_fixed(float p, float f, float thr, int Q, int Qend)
{
#pragma unroll
	for (int j = 0; j < N0; j++) p = __fadd_rn(p, f);
	float s[NCO_WINDOW];
	float acc = p;
	int below = 0;
#pragma unroll
	for (int j = 0; j < NCO_WINDOW; j++) {
		acc = __fadd_rn(acc, f);
		s[j] = acc;
		below += (acc >= thr) ? 0 : 1;
	}
	/* sums are non-decreasing for f > 0: the first crossing is s[below] */
	const float a0 = (below & 1) ? s[1] : s[0], a1 = (below & 1) ? s[3] : s[2];
	const float a2 = (below & 1) ? s[5] : s[4], a3 = (below & 1) ? s[7] : s[6];
	const float b0 = (below & 2) ? a1 : a0, b1 = (below & 2) ? a3 : a2;
	NcoTry t;
	t.sel = (below & 4) ? b1 : b0;
	t.below = below;
	t.ok = f > 0.0f && !(p >= thr) && below < NCO_WINDOW && Q + N0 + NCO_WINDOW <= Qend;
	return t;
}

LRPT_DEV NcoTry nco_try_any(float p, float f, float thr, int n0, int Q, int Qend)
{
	switch (n0) {
		case 0:  return nco_try_fixed<0>(p, f, thr, Q, Qend);
		case 4:  return nco_try_fixed<4>(p, f, thr, Q, Qend);
		case 8:  return nco_try_fixed<8>(p, f, thr, Q, Qend);
		case 12: return nco_try_fixed<12>(p, f, thr, Q, Qend);
		case 16: return nco_try_fixed<16>(p, f, thr, Q, Qend);
		case 20: return nco_try_fixed<20>(p, f, thr, Q, Qend);
		case 24: return nco_try_fixed<24>(p, f, thr, Q, Qend);
		default: break;
	}
	for (int b = 0; b < n0; b += 4) {                               /* n0 is warp-uniform and a multiple of 4 */
		p = __fadd_rn(p, f); p = __fadd_rn(p, f); p = __fadd_rn(p, f); p = __fadd_rn(p, f);
	}
	NcoTry t = nco_try_fixed<0>(p, f, thr, Q + n0, Qend);
	return t;
}

/* One round of the timing warp after the delay-line pick: DC tracker, hand-over to the loop warp, retime, search
 * for the next crossing. Straight-line for every lane; `ready` = this lane has a timing event in this round.
 * y: filter_get(flt, i) for the event (anything for a lane without one). Returns with have_x / Qx / half / Q and
 * the timing state advanced exactly as demod.c:33-39 does. */
template <bool OQ>
LRPT_DEV void timing_round(Loop &r, const lrpt_consts_t &c, bool ready, float2 y, int round, MsgRD *r2d, MsgDR *d2r,
                           uint64_t *chan, int lane, int n0, int &Q, int q1, int Qend, int &Qx, int &half, bool &have_x)
{
	/* agc.c:16-19 up to the bias-free sample */
	const float keep = 1.0f - 0.001f;
	const float nb_re = __fadd_rn(__fmul_rn(r.bias_re, keep), __fmul_rn(0.001f, y.x));
	const float nb_im = __fadd_rn(__fmul_rn(r.bias_im, keep), __fmul_rn(0.001f, y.y));
	const float xr = __fsub_rn(y.x, nb_re), xi = __fsub_rn(y.y, nb_im);
	mbox_put4(&r2d[lane], xr, xi, ready ? msg_meta(OQ ? half : 0, Qx) : 0u, round);
	chan_signal(&chan[0]);
	chan_wait(&chan[1], (unsigned)(round - 1) & 1u);
	float oim;
	asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(oim) : "r"(smem_u32(&d2r[lane].oim)) : "memory");
	const bool arm_i = OQ && half == 1;                             /* demod.c:66-71: no retime on the I arm */
	Loop t = r;
	retime(t, c, oim);                                              /* timing.c:60-95 */
	const float ph = arm_i ? r.t_phase : t.t_phase, fq = arm_i ? r.t_freq : t.t_freq;
	const int dual = r.t_dual;
	const float thr = c.oqpsk ? __fmul_rn((float)dual, kPiF) : kTwoPiF;
	const NcoTry tr = nco_try_any(ph, fq, thr, n0, Q, Qend);
	bool slow = false;
	if (ready) {
		r.bias_re = nb_re; r.bias_im = nb_im;
		r.t_prev = arm_i ? r.t_prev : t.t_prev;
		r.t_freq = fq;
		have_x = false;
		if (tr.ok) {
			r.t_phase = tr.sel;
			Qx = Q + n0 + tr.below; Q = Qx + 1;
			if (OQ) { half = dual; r.t_dual = (dual % 2) + 1; }
			have_x = true;
		} else {
			r.t_phase = ph;
			slow = Q < q1;
		}
	}
	if (slow) {                                                     /* acquisition transients, end of block, odd states */
		bool found = false;
		while (!found && Q < q1) found = nco_chunk(r, c, Q, Qend, Qx, half);
		have_x = found;
	}
}

} // namespace lrpt
