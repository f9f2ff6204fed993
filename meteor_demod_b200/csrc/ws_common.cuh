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

LRPT_DEV void producers_sync()
{
	asm volatile("bar.sync 1, %0;" :: "n"(32*WS_PRODUCERS) : "memory");
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

} // namespace lrpt
