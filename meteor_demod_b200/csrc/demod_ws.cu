/*
 * demod_ws.cu -- warp-specialised exact demodulator (the fast path).
 *
 * Observation: only the symbol-rate recurrence (timing NCO,
 * AGC, Costas PLL, retime; demod.c:33-43) is sequential. The polyphase FIR
 * (filter.c:46-65) is feed-forward: its value at (sample n, sub-step i) does not
 * depend on loop state, only WHICH (n, i) is read does. So one CTA runs
 *
 *   producer warps : ingest raw I/Q (wavfile.c:58-69), keep a float2 delay-line ring
 *                    in shared memory, and compute ALL L polyphase outputs of every
 *                    sample of a tile -- each output as one thread's in-order
 *                    mul-then-add chain over the taps, i.e. bit-identical to
 *                    filter_get -- into a shared-memory tile ring;
 *   consumer warp  : one LANE per stream runs the reference recurrence exactly
 *                    (demod_core.cuh), picking its FIR outputs from the tile ring;
 *                    the L timing sub-steps per sample are evaluated in unrolled
 *                    branch-free chunks of NCO_CHUNK float adds.
 *
 * Tiles are handed over with mbarriers (full/empty per ring slot). Nothing but raw
 * samples is read from HBM and nothing but soft symbols (+ state) is written.
 */
#include <algorithm>
#include "ws_common.cuh"
#include "kernels.h"

namespace lrpt {

#ifndef LRPT_WS_PIPE
#define LRPT_WS_PIPE 1             /* 1: symbol step split in two, deferred half side by side with the next NCO search */
#endif
#ifndef LRPT_WS_SPLIT
#define LRPT_WS_SPLIT 1            /* 1: the two halves of the symbol step on two warps (ws_common.cuh, "two-warp recurrence") */
#endif
/* FIR warps: with the two-warp recurrence SM sub-partitions 0 and 1 belong to the timing warp and the loop warp */
constexpr int WS_P = LRPT_WS_SPLIT ? 7 : WS_PRODUCERS;
constexpr int WS_T        = 32;    /* samples per tile = one FIR unit per stream  */
constexpr int WS_SLOTS    = 2;     /* FIR tile ring depth                        */
constexpr int WS_MAX_G    = 32;    /* streams per CTA = consumer lanes           */
constexpr int WS_CTAS_PER_SM = 1;  /* one big CTA per SM: the recurrence warp's time per symbol does not
                                      depend on how many of its 32 lanes carry a stream, so lanes are filled first */
constexpr int WS_MAX_TAPS = 257;
constexpr int WS_MAX_L    = 8;

struct WsArgs {
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
	int           first_stream, nstreams;
	int           G;           /* streams per CTA */
	int           win;         /* delay-line window entries per stream: H + NT*T */
	int           NT;          /* tiles per window epoch: 2 + ceil(H/T) */
	int           nco_n0;      /* plain NCO adds before the tested window (multiple of 4) */
};

/* ------------------------------------------------------------- producer ---- */

template <int L> struct TapPad { static constexpr int value = (L <= 4) ? 4 : 8; };

/*
 * All L polyphase outputs for one sample: lane's window starts at ring position
 * `pos` (oldest sample). acc[p] is bank p (filter.c:18-22); sub-step i reads bank
 * L-1-i (filter.c:52), so out[i] = acc[L-1-i]. Each accumulator is the reference's
 * chain: acc = acc + x*h, oldest tap first, multiply and add rounded separately.
 */
template <int L>
LRPT_DEV void fir_all_phases(const float2 *__restrict__ w, const float *__restrict__ hT, int taps,
                             float2 *__restrict__ out)
{
	constexpr int LP = TapPad<L>::value;
	float ar[L], ai[L];
#pragma unroll
	for (int p = 0; p < L; p++) { ar[p] = 0.0f; ai[p] = 0.0f; }
#pragma unroll 8
	for (int k = 0; k < taps; k++) {
		const float2 x = w[k];
		float hv[LP];
		*reinterpret_cast<float4 *>(hv) = *reinterpret_cast<const float4 *>(hT + k*LP);
		if (LP == 8) *reinterpret_cast<float4 *>(hv + 4) = *reinterpret_cast<const float4 *>(hT + k*LP + 4);
#pragma unroll
		for (int p = 0; p < L; p++) {
			ar[p] = __fadd_rn(ar[p], __fmul_rn(x.x, hv[p]));
			ai[p] = __fadd_rn(ai[p], __fmul_rn(x.y, hv[p]));
		}
	}
#pragma unroll
	for (int i = 0; i < L; i++) out[i] = make_float2(ar[L-1-i], ai[L-1-i]);
}

/* ------------------------------------------------------------- kernel ------ */

template <int L, bool OQ>
__global__ void __launch_bounds__(WS_THREADS, WS_CTAS_PER_SM)
demod_ws_kernel(const lrpt_consts_t c, const WsArgs a)
{
	constexpr int LP = TapPad<L>::value;
	constexpr int T = WS_T, S = WS_SLOTS, P = WS_P;
	extern __shared__ __align__(16) unsigned char smem_raw[];

	const int taps = c.taps, H = taps - 1;
	const int win = a.win, NT = a.NT;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int g0 = blockIdx.x*a.G;                                  /* first stream (launch-local) of this CTA */
	const int Gc = min(a.G, a.nstreams - g0);                       /* streams this CTA serves */
	const int ntiles = (a.nsamples + T - 1)/T;

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
	float2 *wins  = reinterpret_cast<float2 *>(hT + ((taps*LP + 3) & ~3));   /* [G][win]  delay-line windows */
	float2 *tiles = wins + (size_t)a.G*win;                         /* [G][S][T*L] FIR outputs */

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
	__syncthreads();

	/*
	 * Warp roles. Warps map to the four SM sub-partitions (schedulers) by warp id % 4. The recurrence
	 * warp (warp 0) is the critical path and latency-bound, so it gets sub-partition 0 to itself:
	 * warps 4, 8, 12 retire at once and the twelve FIR warps fill sub-partitions 1-3, where they
	 * are issue-bound without slowing the recurrence down.
	 */
#if LRPT_WS_SPLIT
	if (warp >= 4 && (warp & 3) < 2) return;                        /* sub-partitions 0 and 1: timing warp, loop warp */
	if (warp == 1) {
		loop_warp_run<OQ>(c, a, lut, r2d, d2r, chan, d2e, eack, lane, lane < Gc, g0);
	} else if (warp == 2) {
		egress_warp_run(a, d2e, eack, lane, lane < Gc, g0);
	} else if (warp == 0) {
		/* ===================== timing warp "R": timing NCO, delay-line pick, bias, scale, mix, retime ===================== */
		const bool active = lane < Gc;
		const int local = g0 + lane;
		const int sid = a.first_stream + local;
		Loop r;
		loop_load(r, a.states[a.first_stream + (active ? local : g0)]);
		const int Qend = a.nsamples*L;
		int Q = 0;
		bool have_x = false; int Qx = 0, half = 0;
		const float2 *my_tiles = tiles + (size_t)lane*S*T*L;
		int round = 0;
		for (int t = 0; t < ntiles; t++) {
			const int slot = t % S;
			mbar_wait(&full[slot], (unsigned)(t/S) & 1u);
			const int q0 = t*T*L;
			const int q1 = min((t + 1)*T, a.nsamples)*L;
			const float2 *tile = my_tiles + slot*T*L;
			if (active && !have_x && Q < q1)
				have_x = nco_to_crossing(r, c, a.nco_n0, Q, q1, Qend, Qx, half);
			while (true) {
				__syncwarp();
				const bool ready = active && have_x && Qx < q1;
				if (!__any_sync(0xffffffffu, ready)) break;
				round++;
				float2 y = make_float2(0.f, 0.f);
				if (ready) y = tile[Qx - q0];                        /* filter_get(flt, i); lanes without a stream own no tile */
				timing_round<OQ>(r, c, ready, y, round, r2d, d2r, chan, lane, a.nco_n0, Q, q1, Qend, Qx, half, have_x);
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
		}
	} else {
#else
	if (warp != 0 && (warp & 3) == 0) return;
	if (warp == 0) {
		/* ===================== consumer: one lane per stream ===================== */
		const bool active = lane < Gc;
		const int local = g0 + lane;                                /* launch-local stream index */
		const int sid = a.first_stream + local;
		Loop r;
		long long nsymbols = 0, first_lock = -1;
		unsigned off = 0, nsym = 0;
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
		const int Qend = a.nsamples*L;          /* total timing sub-steps of this launch */
		int Q = 0;                              /* sub-steps already taken               */
		bool have_x = false; int Qx = 0, half = 0;
		const float2 *my_tiles = tiles + (size_t)lane*S*T*L;
#if LRPT_WS_PIPE
		/* oscillator values the next mix will use (fast_sin/fast_cos(-p_phase), pll.c:53-54): known as soon
		 * as the previous symbol step has written p_phase, so they are produced by THAT step's deferred half */
		Osc osc; osc.s = fast_sin(-r.p_phase); osc.co = fast_cos(-r.p_phase); osc.bad = false;
#endif

		for (int t = 0; t < ntiles; t++) {
			const int slot = t % S;
			mbar_wait(&full[slot], (unsigned)(t/S) & 1u);
			{
				const int q0 = t*T*L;
				const int q1 = min((t + 1)*T, a.nsamples)*L;
				const float2 *tile = my_tiles + slot*T*L;
#if LRPT_WS_PIPE
				/* Symbol step split in two (demod_core.cuh, "split in two"). What the NEXT timing decision waits
				 * for -- filter output -> bias/scale -> mix -> retime -> NCO search -> filter output -- is kept
				 * short; the rest of the step (AGC magnitude and gain, Costas error, loop, lock detector, next
				 * oscillator values, int8 store) sits in the same straight-line block as that search, so the
				 * two dependency chains run side by side instead of one after the other. Same values, same
				 * order per variable as demod.c:35-43 / :66-83. Rounds stay warp-uniform (lanes converged). */
				if (active && !have_x && Q < q1)
					have_x = nco_to_crossing(r, c, a.nco_n0, Q, q1, Qend, Qx, half);
				while (true) {
					__syncwarp();
					const bool ready = active && have_x && Qx < q1;
					if (!__any_sync(0xffffffffu, ready)) break;
					if (ready) {
						const float2 y = tile[Qx - q0];              /* filter_get(flt, i) */
						const int Qsym = Qx;
						Pend pd;
						step_critical<OQ>(r, c, half, y.x, y.y, osc.s, osc.co, pd);
						/* timing-NCO search for the next symbol, tried ahead of the deferred half it overlaps with */
						NcoTry tr = nco_try(r, c, a.nco_n0, Q, Qend);
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
#else
				/* Warp-uniform rounds, so that the lanes (streams) stay converged: in each round
				 * every lane that still owes this tile a crossing runs its NCO search, then every
				 * lane holding a crossing inside the tile takes its symbol step, all together. */
				while (true) {
					if (active && !have_x && Q < q1)
						have_x = nco_to_crossing(r, c, a.nco_n0, Q, q1, Qend, Qx, half);
					__syncwarp();
					const bool ready = active && have_x && Qx < q1;
					if (!__any_sync(0xffffffffu, ready)) break;
					if (ready) {
						/* the symbol step (demod.c:35-43 / :66-83) */
						const float2 y = tile[Qx - q0];              /* filter_get(flt, i) */
						const int Qsym = Qx;
						const Loop saved = r;
						float ore, oim; bool emitted;
						if (!symbol_fast<OQ>(r, c, lut, half, y.x, y.y, ore, oim, emitted)) {
							r = saved;                               /* a shortcut was not provably exact */
							emitted = symbol_event(r, c, lut, half, y.x, y.y, ore, oim);
						}
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
#endif
			}
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
		}
	} else {
#endif
		/* ===================== producers: ingest + all-phase FIR ===================== */
#if LRPT_WS_SPLIT
		const int pw = (warp >> 2)*2 + (warp & 3) - 3;              /* warps 3, 6,7, 10,11, 14,15 -> 0..P-1 */
#else
		const int pw = (warp >> 2)*3 + (warp & 3) - 1;              /* 0..P-1 */
#endif
		const int ptid = pw*32 + lane;
		constexpr int SLABS = T/32;
		const int units = Gc*SLABS;                                 /* (stream, 32-sample slab) per tile */
		constexpr int MAXU = (WS_MAX_G*SLABS + P - 1)/P;

		/*
		 * Delay-line window of stream g: `win` = H + NT*T float2, linear (no wrap inside a
		 * FIR). During epoch e = t/NT sample m sits at position m - e*NT*T + H, so the FIR
		 * of local sample nl of tile t reads positions [(t%NT)*T + nl, ... + taps). When a
		 * new epoch starts the last H samples are moved to the front; NT >= 2 + H/T keeps
		 * that move clear of the FIR reads of the tile still in flight.
		 * Prologue: history (taps-1 samples) at [0,H), tile 0 at [H, H+T).
		 */
		for (int i = ptid; i < Gc*(H + T); i += 32*P) {
			const int g = i/(H + T), j = i - g*(H + T);
			const int sid = a.first_stream + g0 + g;
			float2 v;
			if (j < H) v = a.hist[(size_t)sid*H + j];
			else {
				const int m = j - H;
				v = (m < a.nsamples) ? ingest(a.raw + (size_t)(g0 + g)*a.raw_stride, c.bps, m) : make_float2(0.f, 0.f);
			}
			wins[(size_t)g*win + j] = v;
		}
		producers_sync<P>();

		for (int t = 0; t < ntiles; t++) {
			const int slot = t % S;
			const int te = t % NT;                                      /* tile within the epoch */
			const bool more = t + 1 < ntiles;
			/* 1. prefetch this warp's share of tile t+1 (registers) */
			float2 nxt[MAXU];
#pragma unroll
			for (int m = 0; m < MAXU; m++) {
				const int u = pw + P*m;
				nxt[m] = make_float2(0.f, 0.f);
				if (more && u < units) {
					const int g = u/SLABS, sl = u - g*SLABS;
					const int n = (t + 1)*T + sl*32 + lane;
					if (n < a.nsamples) nxt[m] = ingest(a.raw + (size_t)(g0 + g)*a.raw_stride, c.bps, n);
				}
			}
			/* 2. wait until the consumer has released this slot, then fill it */
			if (t >= S) mbar_wait_relaxed(&empty[slot], (unsigned)(t/S - 1) & 1u);
#pragma unroll 1
			for (int u = pw; u < units; u += P) {
				const int g = u/SLABS, sl = u - g*SLABS;
				const int nl = sl*32 + lane;                            /* sample within tile */
				fir_all_phases<L>(wins + (size_t)g*win + te*T + nl, hT, taps,
				                  tiles + ((size_t)g*S + slot)*T*L + nl*L);
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&full[slot]);
			/* 3. append tile t+1 to the delay lines (new epoch: move the last H samples first) */
			if (more) {
				const bool wrap = (te + 1 == NT);
#pragma unroll
				for (int m = 0; m < MAXU; m++) {
					const int u = pw + P*m;
					if (u < units) {
						const int g = u/SLABS, sl = u - g*SLABS;
						float2 *w = wins + (size_t)g*win;
						if (wrap && sl == 0)
							for (int j = lane; j < H; j += 32) w[j] = w[NT*T + j];
						const int base = wrap ? H : (te + 1)*T + H;
						w[base + sl*32 + lane] = nxt[m];
					}
				}
			}
			producers_sync<P>();
		}

		/* epilogue: the last taps-1 samples become the next call's delay line. The window still
		 * has the layout of the last tile's epoch (step 3 is skipped after the last tile). */
		const int e_last = (ntiles - 1)/NT;
		for (int i = ptid; i < Gc*H; i += 32*P) {
			const int g = i/H, j = i - g*H;
			const int sid = a.first_stream + g0 + g;
			const int m = a.nsamples - H + j;                           /* may be negative: old history */
			a.hist[(size_t)sid*H + j] = wins[(size_t)g*win + (m - e_last*NT*T + H)];
		}
	}
}

/* ------------------------------------------------------------- host side --- */

static int g_num_sms = 0;
static int g_max_smem = 0;       /* opt-in per block */
static int g_sm_smem = 0;        /* per SM */

static size_t ws_fixed_smem(int taps, int L)
{
	const int LP = (L <= 4) ? 4 : 8;
	return (2*WS_SLOTS + 2)*sizeof(uint64_t) + 32*sizeof(float) + 32*(sizeof(MsgRD) + sizeof(MsgDR) + sizeof(int)) + EG_RING*32*sizeof(MsgDE) + (size_t)((taps*LP + 3) & ~3)*sizeof(float);
}

static int ws_nt(int taps) { return 2 + (taps - 1 + WS_T - 1)/WS_T; }
static int ws_win(int taps) { return (taps - 1) + ws_nt(taps)*WS_T; }

static size_t ws_stream_smem(int taps, int L)
{
	return (size_t)ws_win(taps)*sizeof(float2) + (size_t)WS_SLOTS*WS_T*L*sizeof(float2);
}

bool ws_supported(const lrpt_consts_t &c)
{
	return c.interp >= 1 && c.interp <= WS_MAX_L && c.taps >= 1 && c.taps <= WS_MAX_TAPS;
}

template <int L, bool OQ> static cudaError_t ws_set_attr1()
{
	cudaError_t e = cudaFuncSetAttribute(demod_ws_kernel<L, OQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem);
	if (e) return e;
	return cudaFuncSetAttribute(demod_ws_kernel<L, OQ>, cudaFuncAttributePreferredSharedMemoryCarveout,
	                            cudaSharedmemCarveoutMaxShared);
}

template <int L> static cudaError_t ws_set_attr()
{
	cudaError_t e = ws_set_attr1<L, false>();
	return e ? e : ws_set_attr1<L, true>();
}

cudaError_t ws_prepare(int device)
{
	cudaError_t e;
	if ((e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, device))) return e;
	if ((e = cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device))) return e;
	if ((e = cudaDeviceGetAttribute(&g_sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device))) return e;
	if ((e = ws_set_attr<1>()) || (e = ws_set_attr<2>()) || (e = ws_set_attr<3>()) || (e = ws_set_attr<4>()) ||
	    (e = ws_set_attr<5>()) || (e = ws_set_attr<6>()) || (e = ws_set_attr<7>()) || (e = ws_set_attr<8>())) return e;
	return cudaSuccess;
}

template <int L> static void ws_launch_one(const lrpt_consts_t &c, const WsArgs &w, int blocks, size_t smem, cudaStream_t st)
{
	if (c.oqpsk) demod_ws_kernel<L, true><<<blocks, WS_THREADS, smem, st>>>(c, w);
	else         demod_ws_kernel<L, false><<<blocks, WS_THREADS, smem, st>>>(c, w);
}

/* Streams per CTA: spread the batch over every SM with up to WS_CTAS_PER_SM co-resident
 * CTAs (each brings one recurrence warp and WS_PRODUCERS FIR warps), within shared memory. */
static int ws_pick_g(int nstreams, size_t fixed, size_t per)
{
	int gfit = (int)(((size_t)g_max_smem - fixed)/per);
	if (gfit < 1) return 0;
	gfit = std::min(gfit, WS_MAX_G);
	const int slots = g_num_sms*WS_CTAS_PER_SM;
	int G = std::max(1, std::min((nstreams + slots - 1)/slots, gfit));
	/* when shared memory caps G below that, the grid needs several waves: level them instead of
	 * running one full wave and a nearly empty one */
	const int blocks = (nstreams + G - 1)/G;
	const int waves = (blocks + slots - 1)/slots;
	return std::max(1, (nstreams + waves*slots - 1)/(waves*slots));
}

cudaError_t launch_ws(const LaunchArgs &a, cudaStream_t st, int *launches)
{
	const lrpt_consts_t &c = *a.c;
	const int L = c.interp, taps = c.taps;
	const size_t fixed = ws_fixed_smem(taps, L), per = ws_stream_smem(taps, L);
	const int G = ws_pick_g(a.nstreams, fixed, per);
	if (G < 1) return cudaErrorInvalidConfiguration;
	int n = 0;
	/* long blocks are cut so that sub-step indices stay in int32; state and cursors carry over */
	size_t done = 0;
	while (done < a.nsamples) {
		const size_t ns = std::min(a.nsamples - done, (size_t)WS_MAX_SAMPLES);
		if (done && !a.d_out_off) return cudaErrorInvalidValue;     /* appending needs a cursor */
		const int blocks = (a.nstreams + G - 1)/G;
		WsArgs w;
		w.taps = a.d_taps; w.states = a.d_states; w.hist = a.d_hist;
		w.raw = reinterpret_cast<const uint8_t *>(a.d_raw) + done*(size_t)(c.bps/4); w.raw_stride = a.raw_stride;
		w.nsamples = (int)ns;
		w.soft = a.d_soft; w.soft_stride = a.soft_stride; w.symf = a.d_symf; w.symf_stride = a.symf_stride;
		w.symq = a.d_symq; w.symq_stride = a.symq_stride; w.q_base = (uint32_t)(done*(size_t)L);
		w.cap = a.cap; w.nsym_out = a.d_nsym; w.out_off = a.d_out_off;
		w.first_stream = a.first_stream; w.nstreams = a.nstreams; w.G = G;
		w.win = ws_win(taps); w.NT = ws_nt(taps);
		{
			/* sub-steps between timing events: 2*pi/step (QPSK) or pi/step (OQPSK halves) */
			const double nominal = (c.oqpsk ? 3.14159265358979 : 6.28318530717959)/(double)c.t_center;
			const int cmin = (int)nominal - 1;
			w.nco_n0 = cmin > 1 ? 4*((cmin - 1)/4) : 0;
		}
		const size_t smem = fixed + per*(size_t)G;
		switch (L) {
			case 1: ws_launch_one<1>(c, w, blocks, smem, st); break;
			case 2: ws_launch_one<2>(c, w, blocks, smem, st); break;
			case 3: ws_launch_one<3>(c, w, blocks, smem, st); break;
			case 4: ws_launch_one<4>(c, w, blocks, smem, st); break;
			case 5: ws_launch_one<5>(c, w, blocks, smem, st); break;
			case 6: ws_launch_one<6>(c, w, blocks, smem, st); break;
			case 7: ws_launch_one<7>(c, w, blocks, smem, st); break;
			case 8: ws_launch_one<8>(c, w, blocks, smem, st); break;
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
