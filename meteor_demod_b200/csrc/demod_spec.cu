/*
 * demod_spec.cu -- warp-specialised exact demodulator with a SPECULATIVE polyphase FIR.
 *
 * demod_ws.cu computes all L polyphase outputs of every sample although the timing loop
 * picks only one output per ~fs*L/symrate sub-steps (demod.c:33-35): 16x the reference's
 * FIR work at the default configuration. Here the FIR warps compute only the outputs the
 * recurrence is LIKELY to pick: the recurrence lane publishes its timing state after every
 * event (last crossing sub-step, NCO phase after retime, NCO step, next threshold); from
 * that the FIR lanes extrapolate the next crossings (step and phase move by < 2^-12 and
 * ~1e-2 rad per symbol, timing.c:7,80) and evaluate SP_NC = 3 consecutive sub-steps around
 * each -- one lane per predicted event, each output still one in-order mul-then-add chain
 * over the taps, i.e. bit-identical to filter_get (filter.c:46-65).
 *
 * Exactness never depends on the prediction: the recurrence lane looks its ACTUAL crossing
 * sub-step up in the candidate list; on a miss (acquisition transients, start of a stream)
 * it evaluates filter_get itself from the shared delay line. Prediction quality only moves
 * time between the two paths.
 *
 * Everything else (delay-line windows with epochs, tile hand-over by mbarriers, recurrence
 * warp alone on SM sub-partition 0, warp-uniform rounds, exact symbol step) is demod_ws.cu's.
 */
#include <algorithm>
#include "ws_common.cuh"
#include "kernels.h"

namespace lrpt {

#ifndef LRPT_SPEC_PIPE
#define LRPT_SPEC_PIPE 1          /* 1: symbol step split in two (demod_core.cuh), as in demod_ws.cu */
#endif
#ifndef LRPT_SPEC_SPLIT
#define LRPT_SPEC_SPLIT 0         /* 1: timing / loop / egress warps as in demod_ws.cu (ws_common.cuh). Measured on B200
                                     (profiles/r2_recurrence_split_ab.log): it wins below ~16 streams per SM, where demod_ws.cu
                                     is used anyway, and loses at 32 streams per SM, where the five FIR warps it costs are
                                     missed (19.7 vs 22.2 GS/s at 4736 streams) -- so this kernel keeps one recurrence warp */
#endif
constexpr int SP_P = LRPT_SPEC_SPLIT ? 7 : WS_PRODUCERS;   /* FIR warps */
constexpr int SP_SLOTS  = 2;      /* candidate tile ring depth                       */
constexpr int SP_NC     = 3;      /* candidate sub-steps per predicted event         */
constexpr int SP_KMAX   = 12;     /* predicted events per stream and tile            */
constexpr int SP_MAX_G  = 32;     /* streams per CTA                                 */
constexpr int SP_MAX_TAPS = 257;
constexpr int SP_MAX_L  = 8;
constexpr int SP_CSTR   = SP_SLOTS*SP_KMAX*SP_NC + 1;   /* float2 per stream of candidates: 146 words, so that
                                                           the 32 recurrence lanes hit different banks */
constexpr int SP_KSTR   = SP_SLOTS*SP_KMAX + 1;         /* ints per stream of centre sub-steps (odd)     */
constexpr int SP_PAD    = 4;      /* window guard entries (reads up to 2 samples past a tile, 1 before) */

struct SpArgs {
	const float  *taps;
	lrpt_state_t *states;
	float2       *hist;
	const uint8_t *raw; size_t raw_stride;
	int           nsamples;
	int8_t       *soft; size_t soft_stride;
	float        *symf; size_t symf_stride;
	uint32_t     *symq; size_t symq_stride; uint32_t q_base;
	unsigned      cap;
	uint32_t     *nsym_out, *out_off;
	unsigned long long *miss_count;   /* optional: FIR evaluations the recurrence had to do itself */
	int           first_stream, nstreams;
	int           G;           /* streams per CTA */
	int           T;           /* samples per tile (<= 32): about SP_KMAX-2 timing events */
	int           win;         /* delay-line window entries per stream: H + NT*T */
	int           NT;          /* tiles per window epoch: 4 + ceil(H/T) */
	int           nco_n0;      /* plain NCO adds before the tested window (multiple of 4) */
	float         ev_phase;    /* NCO phase between two timing events: 2*pi (QPSK) or pi (OQPSK) */
};

template <int L> struct SpTapPad { static constexpr int value = (L <= 4) ? 4 : 8; };

/* filter_get(flt, i) for one (sample, sub-step), by one thread: the recurrence lane's own FIR on a miss.
 * w points at the oldest sample of the window; bank = L-1-i (filter.c:52). */
template <int LP>
LRPT_DEV float2 fir_single(const float2 *__restrict__ w, const float *__restrict__ hT, int taps, int bank)
{
	float ar = 0.0f, ai = 0.0f;
	for (int k = 0; k < taps; k++) {
		const float2 x = w[k];
		const float h = hT[k*LP + bank];
		ar = __fadd_rn(ar, __fmul_rn(x.x, h));
		ai = __fadd_rn(ai, __fmul_rn(x.y, h));
	}
	return make_float2(ar, ai);
}

template <int L, bool OQ>
__global__ void __launch_bounds__(WS_THREADS, 1)
demod_spec_kernel(const lrpt_consts_t c, const SpArgs a)
{
	constexpr int LP = SpTapPad<L>::value;
	constexpr int S = SP_SLOTS, P = SP_P, KM = SP_KMAX, NC = SP_NC;
	extern __shared__ __align__(16) unsigned char smem_raw[];

	const int taps = c.taps, H = taps - 1;
	const int T = a.T, win = a.win, NT = a.NT;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int g0 = blockIdx.x*a.G;
	const int Gc = min(a.G, a.nstreams - g0);
	const int ntiles = (a.nsamples + T - 1)/T;
	const int wstride = win + SP_PAD;                               /* entries per stream window incl. guards */

	/* shared memory carve-up */
	uint64_t *full  = reinterpret_cast<uint64_t *>(smem_raw);       /* [S] */
	uint64_t *empty = full + S;                                     /* [S] */
	uint64_t *chan  = empty + S;                                    /* [2] timing warp <-> loop warp hand-over signals */
	float *lut = reinterpret_cast<float *>(chan + 2);               /* [32] */
	MsgRD *r2d = reinterpret_cast<MsgRD *>(lut + 32);               /* [32] timing warp -> loop warp */
	MsgDR *d2r = reinterpret_cast<MsgDR *>(r2d + 32);               /* [32] loop warp -> timing warp */
	MsgDE *d2e = reinterpret_cast<MsgDE *>(d2r + 32);               /* [EG_RING][32] loop warp -> egress warp */
	volatile int *eack = reinterpret_cast<volatile int *>(d2e + EG_RING*32);   /* [32] rounds the egress warp has consumed */
	float *hT  = reinterpret_cast<float *>(const_cast<int *>(eack) + 32);      /* [taps][LP] */
	float4 *pubs = reinterpret_cast<float4 *>(hT + ((taps*LP + 3) & ~3));     /* [G] published timing state */
	float2 *wins = reinterpret_cast<float2 *>(pubs + a.G);          /* [G][wstride] delay-line windows */
	float2 *cand = wins + (size_t)a.G*wstride;                      /* [G][S][KM][NC] candidate FIR outputs */
	int    *cks  = reinterpret_cast<int *>(cand + (size_t)a.G*SP_CSTR);       /* [G][S][KM] their centre sub-steps */

	if (threadIdx.x == 0) {
		for (int s = 0; s < S; s++) { mbar_init(&full[s], P); mbar_init(&empty[s], 1); }
		mbar_init(&chan[0], 32); mbar_init(&chan[1], 32);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (threadIdx.x < 32) {
		lut[threadIdx.x] = c.lut_tanh[threadIdx.x];
		r2d[threadIdx.x].seq = -1; d2r[threadIdx.x].seq = -1;         /* no round yet */
		eack[threadIdx.x] = -1;
	}
	for (int i = threadIdx.x; i < EG_RING*32; i += WS_THREADS) d2e[i].seq = -1;
	for (int i = threadIdx.x; i < taps*LP; i += WS_THREADS) {
		const int k = i/LP, p = i - k*LP;
		hT[i] = (p < L) ? a.taps[p*taps + k] : 0.0f;
	}
	for (int i = threadIdx.x; i < a.G; i += WS_THREADS) pubs[i] = make_float4(__int_as_float(-1), 0.f, 1.f, 0.f);
	__syncthreads();

#if LRPT_SPEC_SPLIT
	if (warp >= 4 && (warp & 3) < 2) return;                        /* sub-partitions 0 and 1: timing warp, loop warp */
	if (warp == 1) {
		loop_warp_run<OQ>(c, a, lut, r2d, d2r, chan, d2e, eack, lane, lane < Gc, g0);
	} else if (warp == 2) {
		egress_warp_run(a, d2e, eack, lane, lane < Gc, g0);
	} else if (warp == 0) {
		/* ===================== timing warp "R" (ws_common.cuh, two-warp recurrence) ===================== */
		const bool active = lane < Gc;
		const int local = g0 + lane;
		const int sid = a.first_stream + local;
		Loop r;
		unsigned misses = 0;
		loop_load(r, a.states[a.first_stream + (active ? local : g0)]);
		const int Qend = a.nsamples*L;
		int Q = 0;
		bool have_x = false; int Qx = 0, half = 0;
		const float2 *my_win = wins + (size_t)lane*wstride + 1;     /* +1: guard entry in front */
		const float2 *my_cand = cand + (size_t)lane*SP_CSTR;
		const int *my_ck = cks + (size_t)lane*SP_KSTR;
		int round = 0;
		for (int t = 0; t < ntiles; t++) {
			const int slot = t % S;
			mbar_wait(&full[slot], (unsigned)(t/S) & 1u);
			const int q1 = min((t + 1)*T, a.nsamples)*L;
			const float2 *tc = my_cand + slot*KM*NC;
			const int *tk = my_ck + slot*KM;
			int ks = 0;
			int ckn = -0x40000000;
			float2 c0 = make_float2(0.f, 0.f), c1 = c0, c2 = c0;
			if (active) { ckn = tk[0]; c0 = tc[0]; c1 = tc[1]; c2 = tc[2]; }
			if (active && !have_x && Q < q1)
				have_x = nco_to_crossing(r, c, a.nco_n0, Q, q1, Qend, Qx, half);
			while (true) {
				__syncwarp();
				const bool ready = active && have_x && Qx < q1;
				if (!__any_sync(0xffffffffu, ready)) break;
				round++;
				float2 y = make_float2(0.f, 0.f);
				bool hit = false;
				if (ready) {
					/* filter_get(flt, i) for sub-step Qx: from the candidates, or evaluated here */
					while (ks < KM - 1 && ckn + 1 < Qx) {
						ks++;
						ckn = tk[ks]; c0 = tc[ks*NC]; c1 = tc[ks*NC + 1]; c2 = tc[ks*NC + 2];
					}
					const int d = Qx - ckn + 1;
					hit = d >= 0 && d < NC;
					if (hit) y = (d == 0) ? c0 : (d == 1) ? c1 : c2;
					else {
						const int n = Qx/L, i = Qx - n*L;
						y = fir_single<LP>(my_win + (t % NT)*T + (n - t*T), hT, taps, L - 1 - i);
						misses++;
					}
				}
				__syncwarp();
				const int Qsym = Qx;
				timing_round<OQ>(r, c, ready, y, round, r2d, d2r, chan, lane, a.nco_n0, Q, q1, Qend, Qx, half, have_x);
				if (ready) {
					/* publish the timing state the FIR warps extrapolate from */
					pubs[lane] = make_float4(__int_as_float(Qsym), r.t_phase, r.t_freq, nco_threshold(r, c));
					if (hit) {                                       /* the next event will look at the next entry */
						if (ks < KM - 1) {
							ks++;
							ckn = tk[ks]; c0 = tc[ks*NC]; c1 = tc[ks*NC + 1]; c2 = tc[ks*NC + 2];
						} else ckn = -0x40000000;
					}
				}
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&empty[slot]);
		}
		round++;
		mbox_put4(&r2d[lane], 0.0f, 0.0f, MSG_STOP, round);
		chan_signal(&chan[0]);
		if (active) {                                               /* this warp's half of the state */
			lrpt_state_t &s = a.states[sid];
			s.t_phase = r.t_phase; s.t_freq = r.t_freq; s.t_prev = r.t_prev; s.t_dual_state = r.t_dual;
			s.agc_bias_re = r.bias_re; s.agc_bias_im = r.bias_im;
			if (a.miss_count && misses) atomicAdd(a.miss_count, (unsigned long long)misses);
		}
	} else {
#else
	if (warp != 0 && (warp & 3) == 0) return;                       /* sub-partition 0 belongs to the recurrence warp */
	if (warp == 0) {
		/* ===================== recurrence warp: one lane per stream ===================== */
		const bool active = lane < Gc;
		const int local = g0 + lane;
		const int sid = a.first_stream + local;
		Loop r;
		long long nsymbols = 0, first_lock = -1;
		unsigned off = 0, nsym = 0, misses = 0;
		char2 *out = nullptr; float2 *outf = nullptr; uint32_t *outq = nullptr;
		if (active) {
			loop_load(r, a.states[sid]);
			nsymbols = a.states[sid].nsymbols;
			first_lock = a.states[sid].first_lock_symbol;
			off = a.out_off ? a.out_off[local] : 0u;
			out = reinterpret_cast<char2 *>(a.soft + (size_t)local*a.soft_stride);
			if (a.symf) outf = reinterpret_cast<float2 *>(reinterpret_cast<char *>(a.symf) + (size_t)local*a.symf_stride);
			if (a.symq) outq = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(a.symq) + (size_t)local*a.symq_stride);
		}
		const int Qend = a.nsamples*L;
		int Q = 0;
		bool have_x = false; int Qx = 0, half = 0;
		const float2 *my_win = wins + (size_t)lane*wstride + 1;     /* +1: guard entry in front */
		const float2 *my_cand = cand + (size_t)lane*SP_CSTR;
		const int *my_ck = cks + (size_t)lane*SP_KSTR;
#if LRPT_SPEC_PIPE
		Osc osc; osc.s = fast_sin(-r.p_phase); osc.co = fast_cos(-r.p_phase); osc.bad = false;   /* pll.c:53-54 */
#endif

		for (int t = 0; t < ntiles; t++) {
			const int slot = t % S;
			mbar_wait(&full[slot], (unsigned)(t/S) & 1u);
			{
				const int q1 = min((t + 1)*T, a.nsamples)*L;
				const float2 *tc = my_cand + slot*KM*NC;
				const int *tk = my_ck + slot*KM;
				/* candidate events are stored in time order; the entry the next event is expected to hit is
				 * fetched ahead (centre sub-step + its three FIR outputs), so that the look-up costs the
				 * recurrence no shared-memory round trip once the crossing is known */
				int ks = 0;
				int ckn = -0x40000000;
				float2 c0 = make_float2(0.f, 0.f), c1 = c0, c2 = c0;
				if (active) { ckn = tk[0]; c0 = tc[0]; c1 = tc[1]; c2 = tc[2]; }
#if LRPT_SPEC_PIPE
				/* symbol step split in two, the deferred half side by side with the next NCO search: see demod_ws.cu */
				if (active && !have_x && Q < q1)
					have_x = nco_to_crossing(r, c, a.nco_n0, Q, q1, Qend, Qx, half);
				while (true) {
					__syncwarp();
					const bool ready = active && have_x && Qx < q1;
					if (!__any_sync(0xffffffffu, ready)) break;
					if (ready) {
						/* filter_get(flt, i) for sub-step Qx: from the candidates, or evaluated here */
						while (ks < KM - 1 && ckn + 1 < Qx) {
							ks++;
							ckn = tk[ks]; c0 = tc[ks*NC]; c1 = tc[ks*NC + 1]; c2 = tc[ks*NC + 2];
						}
						const int d = Qx - ckn + 1;
						float2 y;
						const bool hit = d >= 0 && d < NC;
						if (hit) y = (d == 0) ? c0 : (d == 1) ? c1 : c2;
						else {
							const int n = Qx/L, i = Qx - n*L;
							y = fir_single<LP>(my_win + (t % NT)*T + (n - t*T), hT, taps, L - 1 - i);
							misses++;
						}
						const int Qsym = Qx;
						Pend pd;
						step_critical<OQ>(r, c, half, y.x, y.y, osc.s, osc.co, pd);
						/* publish the timing state the FIR warps extrapolate from */
						pubs[lane] = make_float4(__int_as_float(Qsym), r.t_phase, r.t_freq, nco_threshold(r, c));
						NcoTry tr = nco_try(r, c, a.nco_n0, Q, Qend);
						if (hit) {                                       /* the next event will look at the next entry */
							if (ks < KM - 1) {
								ks++;
								ckn = tk[ks]; c0 = tc[ks*NC]; c1 = tc[ks*NC + 1]; c2 = tc[ks*NC + 2];
							} else ckn = -0x40000000;
						}
						const float s_gain = r.gain, s_pp = r.p_phase, s_pf = r.p_freq, s_pe = r.p_err;
						const int s_lk = r.locked, s_lo = r.locked_once, s_ud = r.updown;
						Osc next;
						if (!step_deferred_fast<OQ>(r, c, lut, pd, next)) {   /* a shortcut was not provably exact */
							r.gain = s_gain; r.p_phase = s_pp; r.p_freq = s_pf; r.p_err = s_pe;
							r.locked = s_lk; r.locked_once = s_lo; r.updown = s_ud;
							step_deferred_exact<OQ>(r, c, lut, pd, next);
						}
						osc = next;
						if (!(OQ && pd.half == 1)) {
							if (r.locked_once && first_lock < 0) first_lock = nsymbols;
							if (off + nsym < a.cap) {
								out[off + nsym] = make_char2((signed char)quantise(pd.ore), (signed char)quantise(pd.oim));
								if (outf) outf[off + nsym] = make_float2(pd.ore, pd.oim);
								if (outq) outq[off + nsym] = a.q_base + (uint32_t)Qsym;
							}
							nsym++; nsymbols++;
						}
						have_x = false;
						if (Q < q1) have_x = nco_commit(r, c, tr, a.nco_n0, Q, q1, Qend, Qx, half);
					}
				}
			}
#else
				while (true) {
					if (active && !have_x && Q < q1)
						have_x = nco_to_crossing(r, c, a.nco_n0, Q, q1, Qend, Qx, half);
					__syncwarp();
					const bool ready = active && have_x && Qx < q1;
					if (!__any_sync(0xffffffffu, ready)) break;
					if (ready) {
						/* filter_get(flt, i) for sub-step Qx: from the candidates, or evaluated here */
						while (ks < KM - 1 && ckn + 1 < Qx) {
							ks++;
							ckn = tk[ks]; c0 = tc[ks*NC]; c1 = tc[ks*NC + 1]; c2 = tc[ks*NC + 2];
						}
						const int d = Qx - ckn + 1;
						float2 y;
						if (d >= 0 && d < NC) {
							y = (d == 0) ? c0 : (d == 1) ? c1 : c2;
							if (ks < KM - 1) {                               /* the next event will look at the next entry */
								ks++;
								ckn = tk[ks]; c0 = tc[ks*NC]; c1 = tc[ks*NC + 1]; c2 = tc[ks*NC + 2];
							} else ckn = -0x40000000;
						} else {
							const int n = Qx/L, i = Qx - n*L;
							y = fir_single<LP>(my_win + (t % NT)*T + (n - t*T), hT, taps, L - 1 - i);
							misses++;
						}
						const int Qsym = Qx;
						const Loop saved = r;
						float ore, oim; bool emitted;
						if (!symbol_fast<OQ>(r, c, lut, half, y.x, y.y, ore, oim, emitted)) {
							r = saved;
							emitted = symbol_event(r, c, lut, half, y.x, y.y, ore, oim);
						}
						/* publish the timing state the FIR warps extrapolate from */
						pubs[lane] = make_float4(__int_as_float(Qsym), r.t_phase, r.t_freq, nco_threshold(r, c));
						if (emitted) {
							if (r.locked_once && first_lock < 0) first_lock = nsymbols;
							if (off + nsym < a.cap) {
								out[off + nsym] = make_char2((signed char)quantise(ore), (signed char)quantise(oim));
								if (outf) outf[off + nsym] = make_float2(ore, oim);
								if (outq) outq[off + nsym] = a.q_base + (uint32_t)Qsym;
							}
							nsym++; nsymbols++;
						}
						have_x = false;
					}
					__syncwarp();
				}
			}
#endif
			__syncwarp();
			if (lane == 0) mbar_arrive(&empty[slot]);
		}
		if (active) {
			loop_store(r, a.states[sid]);
			a.states[sid].nsamples += a.nsamples;
			a.states[sid].nsymbols = nsymbols;
			a.states[sid].first_lock_symbol = first_lock;
			if (a.nsym_out) a.nsym_out[local] = nsym;
			if (a.out_off) a.out_off[local] = off + nsym;
			if (a.miss_count && misses) atomicAdd(a.miss_count, (unsigned long long)misses);
		}
	} else {
#endif
		/* ===================== FIR warps: ingest + speculative FIR ===================== */
#if LRPT_SPEC_SPLIT
		const int pw = (warp >> 2)*2 + (warp & 3) - 3;              /* warps 3, 6,7, 10,11, 14,15 -> 0..P-1 */
#else
		const int pw = (warp >> 2)*3 + (warp & 3) - 1;              /* 0..P-1 */
#endif
		const int ptid = pw*32 + lane;
		constexpr int MAXU = (SP_MAX_G + P - 1)/P;                   /* ingest units (streams) per warp */
		const int npairs = Gc*KM;                                   /* (stream, predicted event) pairs per tile */
		const float E = a.ev_phase;

		/* delay-line windows as in demod_ws.cu; one guard entry in front, NT two tiles larger because the
		 * recurrence warp may still read the window of tile t (miss path) while tile t+2 is appended */
		for (int i = ptid; i < Gc*(H + T); i += 32*P) {
			const int g = i/(H + T), j = i - g*(H + T);
			const int sid = a.first_stream + g0 + g;
			float2 v;
			if (j < H) v = a.hist[(size_t)sid*H + j];
			else {
				const int m = j - H;
				v = (m < a.nsamples) ? ingest(a.raw + (size_t)(g0 + g)*a.raw_stride, c.bps, m) : make_float2(0.f, 0.f);
			}
			wins[(size_t)g*wstride + 1 + j] = v;
		}
		producers_sync<P>();

		for (int t = 0; t < ntiles; t++) {
			const int slot = t % S;
			const int te = t % NT;
			const bool more = t + 1 < ntiles;
			const int q0 = t*T*L;
			const int q1 = min((t + 1)*T, a.nsamples)*L;
			/* 1. prefetch this warp's share of tile t+1 (lane = sample) */
			float2 nxt[MAXU];
#pragma unroll
			for (int m = 0; m < MAXU; m++) {
				const int g = pw + P*m;
				nxt[m] = make_float2(0.f, 0.f);
				const int n = (t + 1)*T + lane;
				if (more && g < Gc && lane < T && n < a.nsamples)
					nxt[m] = ingest(a.raw + (size_t)(g0 + g)*a.raw_stride, c.bps, n);
			}
			/* 2. wait until the recurrence warp has released this slot, then fill it */
			if (t >= S) mbar_wait_relaxed(&empty[slot], (unsigned)(t/S - 1) & 1u);
#pragma unroll 1
			for (int j = ptid; j < npairs; j += 32*P) {
				const int g = j/KM, k = j - g*KM;
				/* extrapolate the k-th timing event at or after sub-step q0-1 from the published state:
				 * event m >= 1 is expected at Qp + ceil((thr + E*(m-1) - ph)/f) */
				const float4 pb = pubs[g];
				const int Qp = __float_as_int(pb.x);
				const float ph = pb.y, f = pb.z, thr = pb.w;
				int ck = -0x40000000;                                /* "no candidate": never within 1 of a sub-step */
				if (Qp >= 0 && f > 0.0f) {
					const float inv = __frcp_rn(f);
					float m0 = ceilf((((float)(q0 - 1 - Qp))*f + ph - thr)/E) + 1.0f;
					m0 = fmaxf(m0, 1.0f);
					int cm = Qp + (int)ceilf((thr + E*(m0 - 1.0f) - ph)*inv);
					/* float slack: step back / forward so that cm is the FIRST event >= q0-1 */
					for (int it = 0; it < 8 && m0 > 1.0f; it++) {
						const int cb = Qp + (int)ceilf((thr + E*(m0 - 2.0f) - ph)*inv);
						if (cb < q0 - 1) break;
						m0 -= 1.0f; cm = cb;
					}
					for (int it = 0; it < 8 && cm < q0 - 1; it++) {
						m0 += 1.0f; cm = Qp + (int)ceilf((thr + E*(m0 - 1.0f) - ph)*inv);
					}
					const int cc = Qp + (int)ceilf((thr + E*(m0 - 1.0f + (float)k) - ph)*inv);
					if (cm >= q0 - 1 && cc >= q0 - 1 && cc <= q1) ck = cc;
				}
				cks[(size_t)g*SP_KSTR + slot*KM + k] = ck;
				if (ck < 0) continue;
				/* candidates: sub-steps ck-1, ck, ck+1 (those inside [q0,q1)); sample of the first one = n0 */
				const int qa = ck - 1;
				const int n0 = (qa >= 0 ? qa : 0)/L;
				int bank[NC], sh[NC]; bool use[NC];
#pragma unroll
				for (int d = 0; d < NC; d++) {
					const int q = qa + d;
					use[d] = q >= q0 && q < q1;
					const int qq = use[d] ? q : n0*L;                /* harmless stand-in */
					const int n = qq/L;
					sh[d] = n - n0;                                  /* 0, 1 (2 only when L == 1) */
					bank[d] = L - 1 - (qq - n*L);
				}
				/* one pass over the window feeds all three chains: the chain of a candidate on sample n0+s
				 * sees window entry kk as its tap kk-s, oldest first, as filter_get does */
				const float2 *w = wins + (size_t)g*wstride + 1 + te*T + (n0 - t*T);
				float ar[NC], ai[NC];
#pragma unroll
				for (int d = 0; d < NC; d++) { ar[d] = 0.0f; ai[d] = 0.0f; }
				constexpr int SMAX = 2;
				for (int kk = 0; kk < SMAX; kk++) {                 /* head: shifted chains have not started */
					const float2 x = w[kk];
#pragma unroll
					for (int d = 0; d < NC; d++) {
						const int tp = kk - sh[d];
						if (tp >= 0 && tp < taps) {
							const float h = hT[tp*LP + bank[d]];
							ar[d] = __fadd_rn(ar[d], __fmul_rn(x.x, h));
							ai[d] = __fadd_rn(ai[d], __fmul_rn(x.y, h));
						}
					}
				}
				const float *hp[NC];
#pragma unroll
				for (int d = 0; d < NC; d++) hp[d] = hT + bank[d] - sh[d]*LP;
#pragma unroll 4
				for (int kk = SMAX; kk < taps; kk++) {              /* body: every chain is active */
					const float2 x = w[kk];
#pragma unroll
					for (int d = 0; d < NC; d++) {
						const float h = hp[d][kk*LP];
						ar[d] = __fadd_rn(ar[d], __fmul_rn(x.x, h));
						ai[d] = __fadd_rn(ai[d], __fmul_rn(x.y, h));
					}
				}
				for (int kk = taps; kk < taps + SMAX; kk++) {       /* tail: shifted chains finish */
					const float2 x = w[kk];
#pragma unroll
					for (int d = 0; d < NC; d++) {
						const int tp = kk - sh[d];
						if (tp < taps && use[d]) {
							const float h = hT[tp*LP + bank[d]];
							ar[d] = __fadd_rn(ar[d], __fmul_rn(x.x, h));
							ai[d] = __fadd_rn(ai[d], __fmul_rn(x.y, h));
						}
					}
				}
				float2 *o = cand + (size_t)g*SP_CSTR + (slot*KM + k)*NC;
#pragma unroll
				for (int d = 0; d < NC; d++) if (use[d]) o[d] = make_float2(ar[d], ai[d]);
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&full[slot]);
			/* 3. append tile t+1 to the delay lines (new epoch: move the last H samples first) */
			if (more) {
				const bool wrap = (te + 1 == NT);
#pragma unroll
				for (int m = 0; m < MAXU; m++) {
					const int g = pw + P*m;
					if (g < Gc) {
						float2 *w = wins + (size_t)g*wstride + 1;
						if (wrap)
							for (int j = lane; j < H; j += 32) w[j] = w[NT*T + j];
						if (lane < T) w[(wrap ? H : (te + 1)*T + H) + lane] = nxt[m];
					}
				}
			}
			producers_sync<P>();
		}

		const int e_last = (ntiles - 1)/NT;
		for (int i = ptid; i < Gc*H; i += 32*P) {
			const int g = i/H, j = i - g*H;
			const int sid = a.first_stream + g0 + g;
			const int m = a.nsamples - H + j;
			a.hist[(size_t)sid*H + j] = wins[(size_t)g*wstride + 1 + (m - e_last*NT*T + H)];
		}
	}
}

/* ------------------------------------------------------------- host side --- */

static int sp_num_sms = 0, sp_max_smem = 0;

static int sp_tile(const lrpt_consts_t &c)
{
	/* about SP_KMAX-2 timing events per tile: T*L sub-steps / (sub-steps per event) */
	const double spacing = (c.oqpsk ? 3.14159265358979 : 6.28318530717959)/(double)c.t_center;
	int T = (int)((SP_KMAX - 2)*spacing/(double)c.interp);
	return std::max(4, std::min(T, 32));
}

static int sp_nt(int taps, int T) { return 4 + (taps - 1 + T - 1)/T; }

static size_t sp_fixed_smem(int taps, int L)
{
	const int LP = (L <= 4) ? 4 : 8;
	return (2*SP_SLOTS + 2)*sizeof(uint64_t) + 32*sizeof(float) + 32*(sizeof(MsgRD) + sizeof(MsgDR) + sizeof(int)) + EG_RING*32*sizeof(MsgDE) + (size_t)((taps*LP + 3) & ~3)*sizeof(float);
}

static size_t sp_stream_smem(int taps, int T)
{
	const int win = (taps - 1) + sp_nt(taps, T)*T;
	return sizeof(float4) + (size_t)(win + SP_PAD)*sizeof(float2) +
	       (size_t)SP_CSTR*sizeof(float2) + (size_t)SP_KSTR*sizeof(int);
}

bool spec_supported(const lrpt_consts_t &c)
{
	return c.interp >= 2 && c.interp <= SP_MAX_L && c.taps >= 3 && c.taps <= SP_MAX_TAPS;
}

template <int L, bool OQ> static cudaError_t sp_attr1()
{
	cudaError_t e = cudaFuncSetAttribute(demod_spec_kernel<L, OQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, sp_max_smem);
	if (e) return e;
	return cudaFuncSetAttribute(demod_spec_kernel<L, OQ>, cudaFuncAttributePreferredSharedMemoryCarveout,
	                            cudaSharedmemCarveoutMaxShared);
}
template <int L> static cudaError_t sp_attr() { cudaError_t e = sp_attr1<L, false>(); return e ? e : sp_attr1<L, true>(); }

cudaError_t spec_prepare(int device)
{
	cudaError_t e;
	if ((e = cudaDeviceGetAttribute(&sp_num_sms, cudaDevAttrMultiProcessorCount, device))) return e;
	if ((e = cudaDeviceGetAttribute(&sp_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device))) return e;
	if ((e = sp_attr<2>()) || (e = sp_attr<3>()) || (e = sp_attr<4>()) || (e = sp_attr<5>()) ||
	    (e = sp_attr<6>()) || (e = sp_attr<7>()) || (e = sp_attr<8>())) return e;
	return cudaSuccess;
}

template <int L> static void sp_launch_one(const lrpt_consts_t &c, const SpArgs &w, int blocks, size_t smem, cudaStream_t st)
{
	if (c.oqpsk) demod_spec_kernel<L, true><<<blocks, WS_THREADS, smem, st>>>(c, w);
	else         demod_spec_kernel<L, false><<<blocks, WS_THREADS, smem, st>>>(c, w);
}

cudaError_t launch_spec(const LaunchArgs &a, cudaStream_t st, int *launches, unsigned long long *d_miss)
{
	const lrpt_consts_t &c = *a.c;
	const int L = c.interp, taps = c.taps, T = sp_tile(c);
	const size_t fixed = sp_fixed_smem(taps, L), per = sp_stream_smem(taps, T);
	int gfit = (int)(((size_t)sp_max_smem - fixed)/per);
	if (gfit < 1) return cudaErrorInvalidConfiguration;
	gfit = std::min(gfit, SP_MAX_G);
	int G = std::max(1, std::min((a.nstreams + sp_num_sms - 1)/sp_num_sms, gfit));
	const int blocks0 = (a.nstreams + G - 1)/G, waves = (blocks0 + sp_num_sms - 1)/sp_num_sms;
	G = std::max(1, (a.nstreams + waves*sp_num_sms - 1)/(waves*sp_num_sms));
	int n = 0;
	size_t done = 0;
	while (done < a.nsamples) {
		const size_t ns = std::min(a.nsamples - done, (size_t)WS_MAX_SAMPLES);
		if (done && !a.d_out_off) return cudaErrorInvalidValue;
		const int blocks = (a.nstreams + G - 1)/G;
		SpArgs w;
		w.taps = a.d_taps; w.states = a.d_states; w.hist = a.d_hist;
		w.raw = reinterpret_cast<const uint8_t *>(a.d_raw) + done*(size_t)(c.bps/4); w.raw_stride = a.raw_stride;
		w.nsamples = (int)ns;
		w.soft = a.d_soft; w.soft_stride = a.soft_stride; w.symf = a.d_symf; w.symf_stride = a.symf_stride;
		w.symq = a.d_symq; w.symq_stride = a.symq_stride; w.q_base = (uint32_t)(done*(size_t)L);
		w.cap = a.cap; w.nsym_out = a.d_nsym; w.out_off = a.d_out_off; w.miss_count = d_miss;
		w.first_stream = a.first_stream; w.nstreams = a.nstreams; w.G = G;
		w.T = T; w.NT = sp_nt(taps, T); w.win = (taps - 1) + w.NT*T;
		w.ev_phase = c.oqpsk ? 3.14159274101257324f : 6.28318548202514648f;
		{
			const double nominal = (c.oqpsk ? 3.14159265358979 : 6.28318530717959)/(double)c.t_center;
			const int cmin = (int)nominal - 1;
			w.nco_n0 = cmin > 1 ? 4*((cmin - 1)/4) : 0;
		}
		const size_t smem = fixed + per*(size_t)G;
		switch (L) {
			case 2: sp_launch_one<2>(c, w, blocks, smem, st); break;
			case 3: sp_launch_one<3>(c, w, blocks, smem, st); break;
			case 4: sp_launch_one<4>(c, w, blocks, smem, st); break;
			case 5: sp_launch_one<5>(c, w, blocks, smem, st); break;
			case 6: sp_launch_one<6>(c, w, blocks, smem, st); break;
			case 7: sp_launch_one<7>(c, w, blocks, smem, st); break;
			case 8: sp_launch_one<8>(c, w, blocks, smem, st); break;
			default: return cudaErrorInvalidConfiguration;
		}
		n++;
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) { if (launches) *launches = n; return e; }
		done += ns;
	}
	if (launches) *launches = n;
	return cudaSuccess;
}

} // namespace lrpt
