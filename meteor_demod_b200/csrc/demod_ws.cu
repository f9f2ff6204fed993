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
#include "demod_core.cuh"
#include "kernels.h"

namespace lrpt {

constexpr int WS_T        = 64;    /* samples per tile                           */
constexpr int WS_SLOTS    = 3;     /* FIR tile ring depth                        */
constexpr int WS_PRODUCERS = 4;    /* producer warps                             */
constexpr int WS_THREADS  = 32*(1 + WS_PRODUCERS);
constexpr int WS_MAX_G    = 32;    /* streams per CTA = consumer lanes           */
constexpr int NCO_CHUNK   = 8;     /* timing sub-steps evaluated per branch      */
constexpr int WS_MAX_TAPS = 257;
constexpr int WS_MAX_L    = 8;
constexpr int WS_MAX_SAMPLES = 1 << 26;   /* per launch; keeps sub-step indices in int32 */

struct WsArgs {
	const float  *taps;
	lrpt_state_t *states;
	float2       *hist;
	const uint8_t *raw; size_t raw_stride;
	int           nsamples;
	int8_t       *soft; size_t soft_stride;
	float        *symf; size_t symf_stride;
	unsigned      cap;
	uint32_t     *nsym_out, *out_off;
	int           first_stream, nstreams;
	int           G;           /* streams per CTA */
	int           ring;        /* delay-line ring entries per stream (power of two) */
};

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

LRPT_DEV void producers_sync()
{
	asm volatile("bar.sync 1, %0;" :: "n"(32*WS_PRODUCERS) : "memory");
}

/* ------------------------------------------------------------- producer ---- */

template <int L> struct TapPad { static constexpr int value = (L <= 4) ? 4 : 8; };

/*
 * All L polyphase outputs for one sample: lane's window starts at ring position
 * `pos` (oldest sample). acc[p] is bank p (filter.c:18-22); sub-step i reads bank
 * L-1-i (filter.c:52), so out[i] = acc[L-1-i]. Each accumulator is the reference's
 * chain: acc = acc + x*h, oldest tap first, multiply and add rounded separately.
 */
template <int L>
LRPT_DEV void fir_all_phases(const float2 *__restrict__ ring, int mask, int pos,
                             const float *__restrict__ hT, int taps, float2 *__restrict__ out)
{
	constexpr int LP = TapPad<L>::value;
	float ar[L], ai[L];
#pragma unroll
	for (int p = 0; p < L; p++) { ar[p] = 0.0f; ai[p] = 0.0f; }
#pragma unroll 4
	for (int k = 0; k < taps; k++) {
		const float2 x = ring[(pos + k) & mask];
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

template <int L>
__global__ void __launch_bounds__(WS_THREADS, 1)
demod_ws_kernel(const lrpt_consts_t c, const WsArgs a)
{
	constexpr int LP = TapPad<L>::value;
	constexpr int T = WS_T, S = WS_SLOTS, P = WS_PRODUCERS;
	extern __shared__ __align__(16) unsigned char smem_raw[];

	const int taps = c.taps, H = taps - 1;
	const int ring = a.ring, mask = ring - 1;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int g0 = blockIdx.x*a.G;                                  /* first stream (launch-local) of this CTA */
	const int Gc = min(a.G, a.nstreams - g0);                       /* streams this CTA serves */
	const int ntiles = (a.nsamples + T - 1)/T;

	/* shared memory carve-up */
	uint64_t *full  = reinterpret_cast<uint64_t *>(smem_raw);       /* [S] */
	uint64_t *empty = full + S;                                     /* [S] */
	float *lut = reinterpret_cast<float *>(empty + S);              /* [32] */
	float *hT  = lut + 32;                                          /* [taps][LP] */
	float2 *rings = reinterpret_cast<float2 *>(hT + ((taps*LP + 3) & ~3));   /* [G][ring] */
	float2 *tiles = rings + (size_t)a.G*ring;                       /* [G][S][T*L] */

	if (threadIdx.x == 0) {
		for (int s = 0; s < S; s++) { mbar_init(&full[s], P); mbar_init(&empty[s], 1); }
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (threadIdx.x < 32) lut[threadIdx.x] = c.lut_tanh[threadIdx.x];
	for (int i = threadIdx.x; i < taps*LP; i += WS_THREADS) {
		const int k = i/LP, p = i - k*LP;
		hT[i] = (p < L) ? a.taps[p*taps + k] : 0.0f;
	}
	__syncthreads();

	if (warp == 0) {
		/* ===================== consumer: one lane per stream ===================== */
		const bool active = lane < Gc;
		const int local = g0 + lane;                                /* launch-local stream index */
		const int sid = a.first_stream + local;
		Loop r;
		long long nsymbols = 0, first_lock = -1;
		unsigned off = 0, nsym = 0;
		char2 *out = nullptr; float2 *outf = nullptr;
		if (active) {
			loop_load(r, a.states[sid]);
			nsymbols = a.states[sid].nsymbols;
			first_lock = a.states[sid].first_lock_symbol;
			off = a.out_off ? a.out_off[local] : 0u;
			out = reinterpret_cast<char2 *>(a.soft + (size_t)local*a.soft_stride);
			if (a.symf) outf = reinterpret_cast<float2 *>(reinterpret_cast<char *>(a.symf) + (size_t)local*a.symf_stride);
		}
		const int Qend = a.nsamples*L;          /* total timing sub-steps of this launch */
		int Q = 0;                              /* sub-steps already taken               */
		bool have_x = false; int Qx = 0, half = 0;
		const float2 *my_tiles = tiles + (size_t)lane*S*T*L;

		for (int t = 0; t < ntiles; t++) {
			const int slot = t % S;
			mbar_wait(&full[slot], (unsigned)(t/S) & 1u);
			if (active) {
				const int q0 = t*T*L;
				const int q1 = min((t + 1)*T, a.nsamples)*L;
				const float2 *tile = my_tiles + slot*T*L;
				while (true) {
					if (!have_x) {
						if (Q >= q1) break;
						/* advance_timeslot x NCO_CHUNK (timing.c:32-38 / :41-57), branch-free */
						const float f = r.t_freq;
						const float thr = c.oqpsk ? __fmul_rn((float)r.t_dual, kPiF) : kTwoPiF;
						float ph[NCO_CHUNK];
						float acc = r.t_phase;
						unsigned hit = 0;
#pragma unroll
						for (int j = 0; j < NCO_CHUNK; j++) {
							acc = __fadd_rn(acc, f);
							ph[j] = acc;
							hit |= (acc >= thr) ? (1u << j) : 0u;
						}
						const int limit = Qend - Q;                 /* >= 1 */
						const int first = __ffs(hit);               /* 1-based, 0 = none */
						if (first != 0 && first <= limit) {
							float sel = ph[0];
#pragma unroll
							for (int j = 1; j < NCO_CHUNK; j++) sel = (first == j + 1) ? ph[j] : sel;
							r.t_phase = sel;
							Qx = Q + first - 1; Q += first; have_x = true;
							if (c.oqpsk) { half = r.t_dual; r.t_dual = (r.t_dual % 2) + 1; }
						} else if (limit >= NCO_CHUNK) {
							r.t_phase = ph[NCO_CHUNK-1]; Q += NCO_CHUNK;
						} else {                                     /* end of the block */
							float sel = ph[0];
#pragma unroll
							for (int j = 1; j < NCO_CHUNK; j++) sel = (limit == j + 1) ? ph[j] : sel;
							r.t_phase = sel; Q += limit;
						}
					}
					if (have_x) {
						if (Qx >= q1) break;                         /* belongs to a later tile */
						const float2 y = tile[Qx - q0];              /* filter_get(flt, i), demod.c:35 */
						float ore, oim;
						if (symbol_event(r, c, lut, half, y.x, y.y, ore, oim)) {
							if (r.locked_once && first_lock < 0) first_lock = nsymbols;
							if (off + nsym < a.cap) {
								out[off + nsym] = make_char2((signed char)quantise(ore), (signed char)quantise(oim));
								if (outf) outf[off + nsym] = make_float2(ore, oim);
							}
							nsym++; nsymbols++;
						}
						have_x = false;
					}
				}
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
		/* ===================== producers: ingest + all-phase FIR ===================== */
		const int pw = warp - 1;                                    /* 0..P-1 */
		const int ptid = threadIdx.x - 32;
		constexpr int SLABS = T/32;
		const int units = Gc*SLABS;                                 /* (stream, 32-sample slab) per tile */
		constexpr int MAXU = (WS_MAX_G*SLABS + P - 1)/P;

		/* prologue: delay line (taps-1 samples of history) + tile 0 into the rings.
		 * ring position of sample m is (m + H) & mask. */
		for (int i = ptid; i < Gc*(H + T); i += 32*P) {
			const int g = i/(H + T), j = i - g*(H + T);
			const int sid = a.first_stream + g0 + g;
			float2 v;
			if (j < H) v = a.hist[(size_t)sid*H + j];
			else {
				const int m = j - H;
				v = (m < a.nsamples) ? ingest(a.raw + (size_t)(g0 + g)*a.raw_stride, c.bps, m) : make_float2(0.f, 0.f);
			}
			rings[(size_t)g*ring + (j & mask)] = v;
		}
		producers_sync();

		for (int t = 0; t < ntiles; t++) {
			const int slot = t % S;
			/* 1. prefetch this warp's share of tile t+1 (registers) */
			float2 nxt[MAXU];
#pragma unroll
			for (int m = 0; m < MAXU; m++) {
				const int u = pw + P*m;
				nxt[m] = make_float2(0.f, 0.f);
				if (u < units) {
					const int g = u/SLABS, sl = u - g*SLABS;
					const int n = (t + 1)*T + sl*32 + lane;
					if (n < a.nsamples) nxt[m] = ingest(a.raw + (size_t)(g0 + g)*a.raw_stride, c.bps, n);
				}
			}
			/* 2. wait until the consumer has released this slot, then fill it */
			if (t >= S) mbar_wait(&empty[slot], (unsigned)(t/S - 1) & 1u);
#pragma unroll 1
			for (int u = pw; u < units; u += P) {
				const int g = u/SLABS, sl = u - g*SLABS;
				const int nl = sl*32 + lane;                        /* sample within tile */
				fir_all_phases<L>(rings + (size_t)g*ring, mask, t*T + nl, hT, taps,
				                  tiles + ((size_t)g*S + slot)*T*L + nl*L);
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&full[slot]);
			/* 3. append tile t+1 to the delay lines */
#pragma unroll
			for (int m = 0; m < MAXU; m++) {
				const int u = pw + P*m;
				if (u < units) {
					const int g = u/SLABS, sl = u - g*SLABS;
					const int n = (t + 1)*T + sl*32 + lane;
					rings[(size_t)g*ring + ((n + H) & mask)] = nxt[m];
				}
			}
			producers_sync();
		}

		/* epilogue: the last taps-1 samples become the next call's delay line */
		for (int i = ptid; i < Gc*H; i += 32*P) {
			const int g = i/H, j = i - g*H;
			const int sid = a.first_stream + g0 + g;
			const int m = a.nsamples - H + j;                       /* may be negative: old history */
			a.hist[(size_t)sid*H + j] = rings[(size_t)g*ring + ((m + H) & mask)];
		}
	}
}

/* ------------------------------------------------------------- host side --- */

static int g_num_sms = 0;
static int g_max_smem = 0;

static size_t ws_fixed_smem(int taps, int L)
{
	const int LP = (L <= 4) ? 4 : 8;
	return 2*WS_SLOTS*sizeof(uint64_t) + 32*sizeof(float) + (size_t)((taps*LP + 3) & ~3)*sizeof(float);
}

static int ws_ring(int taps)
{
	int need = (taps - 1) + 2*WS_T, r = 64;
	while (r < need) r <<= 1;
	return r;
}

static size_t ws_stream_smem(int taps, int L)
{
	return (size_t)ws_ring(taps)*sizeof(float2) + (size_t)WS_SLOTS*WS_T*L*sizeof(float2);
}

bool ws_supported(const lrpt_consts_t &c)
{
	return c.interp >= 1 && c.interp <= WS_MAX_L && c.taps >= 1 && c.taps <= WS_MAX_TAPS;
}

template <int L> static cudaError_t ws_set_attr()
{
	return cudaFuncSetAttribute(demod_ws_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem);
}

cudaError_t ws_prepare(int device)
{
	cudaError_t e;
	if ((e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, device))) return e;
	if ((e = cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device))) return e;
	if ((e = ws_set_attr<1>()) || (e = ws_set_attr<2>()) || (e = ws_set_attr<3>()) || (e = ws_set_attr<4>()) ||
	    (e = ws_set_attr<5>()) || (e = ws_set_attr<6>()) || (e = ws_set_attr<7>()) || (e = ws_set_attr<8>())) return e;
	return cudaSuccess;
}

template <int L> static void ws_launch_one(const lrpt_consts_t &c, const WsArgs &w, int blocks, size_t smem, cudaStream_t st)
{
	demod_ws_kernel<L><<<blocks, WS_THREADS, smem, st>>>(c, w);
}

cudaError_t launch_ws(const LaunchArgs &a, cudaStream_t st, int *launches)
{
	const lrpt_consts_t &c = *a.c;
	const int L = c.interp, taps = c.taps;
	const size_t fixed = ws_fixed_smem(taps, L), per = ws_stream_smem(taps, L);
	int gfit = (int)(((size_t)g_max_smem - fixed)/per);
	if (gfit < 1) return cudaErrorInvalidConfiguration;
	gfit = std::min(gfit, WS_MAX_G);
	int n = 0;
	/* long blocks are cut so that sub-step indices stay in int32; state carries over */
	size_t done = 0;
	uint32_t *cursor = a.d_out_off;
	do {
		const size_t ns = std::min(a.nsamples - done, (size_t)WS_MAX_SAMPLES);
		int G = (a.nstreams + g_num_sms - 1)/g_num_sms;             /* spread streams over all SMs */
		G = std::max(1, std::min(G, gfit));
		const int blocks = (a.nstreams + G - 1)/G;
		WsArgs w;
		w.taps = a.d_taps; w.states = a.d_states; w.hist = a.d_hist;
		w.raw = reinterpret_cast<const uint8_t *>(a.d_raw) + done*(size_t)(c.bps/4); w.raw_stride = a.raw_stride;
		w.nsamples = (int)ns;
		w.soft = a.d_soft; w.soft_stride = a.soft_stride; w.symf = a.d_symf; w.symf_stride = a.symf_stride;
		w.cap = a.cap; w.nsym_out = a.d_nsym; w.out_off = cursor;
		w.first_stream = a.first_stream; w.nstreams = a.nstreams; w.G = G; w.ring = ws_ring(taps);
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
	} while (done < a.nsamples);
	if (launches) *launches = n;
	return cudaSuccess;
}

} // namespace lrpt
