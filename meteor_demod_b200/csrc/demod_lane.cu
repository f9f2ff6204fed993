/*
 * demod_lane.cu -- "one lane, one stream, lazily": the throughput kernel for large batches.
 *
 * Every warp is the same: lane = stream. A lane keeps the raw samples its delay line needs
 * (the last taps-1 samples plus the tile in flight, filter.c:39-43) in shared memory IN THE
 * INPUT'S OWN TYPE (4 bytes per 16-bit I/Q sample, 2 per 8-bit, 8 per float), runs the exact
 * symbol-rate recurrence (demod_core.cuh) and, at every timing crossing, evaluates the ONE
 * polyphase output the reference evaluates (filter_get, filter.c:46-65) -- as the reference's
 * own in-order mul-then-add chain over the taps. No FIR output is computed that the timing
 * loop does not pick, so the arithmetic per symbol is the reference's: ~4*taps flops of FIR
 * plus the loop, against 16x (demod_ws.cu, all phases) or 3x + prediction (demod_spec.cu).
 *
 * What makes it fast is not a lane's latency (the FIR chain now heads each symbol's dependency
 * chain) but how many lanes fit: the raw-typed delay line is 0.5-0.8 KB per stream, so one SM
 * holds 8-16 warps = 256-512 streams whose chains hide each other; there are no producer
 * warps, no mbarriers and no prediction. demod_spec.cu stays the better kernel below roughly
 * 40 streams per SM, where one lane's latency is what is measured.
 *
 * Shared-memory layout: per warp a window of NE entries x 32 lanes, entry-major
 * (element of lane l, entry e at [e*32 + l]), so every lane always hits its own bank whatever
 * its timing phase is. The window is linear: sample m of epoch E sits at entry m - E.start + H;
 * when it is full the last H entries move to the front (once every NT tiles), which keeps the
 * tap loop free of index arithmetic (immediate offsets only).
 *
 * Global traffic: each lane reads its own row 16 bytes at a time, one tile (LN_T samples)
 * ahead of use, and writes its int8 symbols; whole 32-byte sectors are consumed, nothing is
 * read twice.
 */
#include <algorithm>
#include <stdio.h>
#include <stdlib.h>
#include "ws_common.cuh"
#include "kernels.h"

namespace lrpt {

#ifndef LRPT_LANE_T
#define LRPT_LANE_T 16
#endif
constexpr int LN_T         = LRPT_LANE_T;   /* samples per tile (per lane) */
constexpr int LN_MAX_WARPS = 16;   /* warps per CTA = 512 streams                          */
constexpr int LN_MAX_TAPS  = 1025;
constexpr int LN_MAX_L     = 8;

#ifndef LRPT_LANE_CVT
#define LRPT_LANE_CVT 1            /* 0: integer->float conversion instructions, 1: exponent-splice + subtract */
#endif
#ifndef LRPT_LANE_PIPE
#define LRPT_LANE_PIPE 1           /* 1: deferred half of symbol k woven into the tap loop of symbol k+1 */
#endif
#ifndef LRPT_LANE_PACKED
#define LRPT_LANE_PACKED 1         /* 1: the I and Q chains as one packed f32x2 chain (sm_100 FMUL2/FFMA2/FADD2) */
#endif

/* ------------------------------------------------------- packed f32x2 helpers -- *
 * sm_100 has two-wide fp32 instructions on 64-bit register pairs; each half is an ordinary IEEE
 * operation, so (re, im) of one tap can share instructions without changing a bit. One trap:
 * ptxas CONTRACTS mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false (measured:
 * tools/mb/f32x2.cu). The accumulate step is therefore written p*one + acc with `one` = 1.0f
 * arriving as a kernel argument, which ptxas cannot fold: fma(p, 1, acc) = RN(p + acc) exactly. */
typedef unsigned long long f32x2_t;
LRPT_DEV f32x2_t pk2(float a, float b) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
LRPT_DEV float2 upk2(f32x2_t v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
LRPT_DEV f32x2_t mul2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
LRPT_DEV f32x2_t add2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
LRPT_DEV f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

struct LaneArgs {
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
	int           W;           /* warps per CTA */
	int           NT;          /* tiles per window epoch */
	int           nco_n0;
	int           TS;          /* floats per bank of the tap table: >= taps, a multiple of 4, TS/4 odd */
	int           div_magic;   /* (x*div_magic) >> 16 == x / interp for 0 <= x < LN_T*interp */
	float         one;         /* 1.0f, opaque to the compiler (see the packed f32x2 helpers) */
};

/* ------------------------------------------------------- raw sample formats -- */

/* wavfile.c:58-69 per input type. `elem` is what the window holds; cvt() yields exactly the
 * float pair wav_read produces. */
template <int BPS> struct RawT;

template <> struct RawT<16> {
	typedef uint32_t elem;                       /* I in the low half, Q in the high half */
	static constexpr int NV = LN_T*4/16;         /* 16-byte vectors per tile and lane     */
#if LRPT_LANE_CVT
	/* halves are stored biased (s + 32768, an unsigned 16-bit number); spliced under the exponent
	 * of 2^23 they read 8388608 + s + 32768 exactly, and the subtraction is exact too */
	LRPT_DEV static elem prep(uint32_t w) { return w ^ 0x80008000u; }
	LRPT_DEV static float2 cvt(elem e)
	{
		const float i = __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7610));
		const float q = __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7632));
		return make_float2(__fsub_rn(i, 8421376.0f), __fsub_rn(q, 8421376.0f));
	}
	LRPT_DEV static f32x2_t cvt2(elem e)
	{
		const float i = __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7610));
		const float q = __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7632));
		return add2(pk2(i, q), pk2(-8421376.0f, -8421376.0f));
	}
#else
	LRPT_DEV static elem prep(uint32_t w) { return w; }
	LRPT_DEV static float2 cvt(elem e)
	{
		return make_float2((float)(short)(e & 0xffffu), (float)(short)(e >> 16));
	}
	LRPT_DEV static f32x2_t cvt2(elem e) { const float2 v = cvt(e); return pk2(v.x, v.y); }
#endif
	LRPT_DEV static elem from_float(float2 v)
	{
		const uint32_t i = (uint32_t)(int)v.x & 0xffffu, q = (uint32_t)(int)v.y & 0xffffu;
		return prep(i | (q << 16));
	}
	/* vector j of a tile -> its 4 elements */
	LRPT_DEV static void unpack(const uint4 &v, elem (&e)[16/sizeof(elem)])
	{
		e[0] = prep(v.x); e[1] = prep(v.y); e[2] = prep(v.z); e[3] = prep(v.w);
	}
	LRPT_DEV static elem load1(const uint8_t *row, int idx) { return prep(reinterpret_cast<const uint32_t *>(row)[idx]); }
	LRPT_DEV static elem zero() { return prep(0u); }
};

template <> struct RawT<8> {
	typedef uint16_t elem;                       /* I low byte, Q high byte, both offset-128 (wavfile.c:59-61) */
	static constexpr int NV = LN_T*2/16;
	LRPT_DEV static float2 cvt(elem e)
	{
		const uint32_t w = e;
		const float i = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440));
		const float q = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7441));
		return make_float2(__fsub_rn(i, 8388736.0f), __fsub_rn(q, 8388736.0f));
	}
	LRPT_DEV static f32x2_t cvt2(elem e)
	{
		const uint32_t w = e;
		const float i = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440));
		const float q = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7441));
		return add2(pk2(i, q), pk2(-8388736.0f, -8388736.0f));
	}
	LRPT_DEV static elem from_float(float2 v)
	{
		return (elem)((((int)v.x + 128) & 0xff) | ((((int)v.y + 128) & 0xff) << 8));
	}
	LRPT_DEV static void unpack(const uint4 &v, elem (&e)[16/sizeof(elem)])
	{
		e[0] = (elem)(v.x & 0xffffu); e[1] = (elem)(v.x >> 16); e[2] = (elem)(v.y & 0xffffu); e[3] = (elem)(v.y >> 16);
		e[4] = (elem)(v.z & 0xffffu); e[5] = (elem)(v.z >> 16); e[6] = (elem)(v.w & 0xffffu); e[7] = (elem)(v.w >> 16);
	}
	LRPT_DEV static elem load1(const uint8_t *row, int idx) { return reinterpret_cast<const uint16_t *>(row)[idx]; }
	LRPT_DEV static elem zero() { return (elem)0x8080u; }
};

template <> struct RawT<32> {
	typedef float2 elem;
	static constexpr int NV = LN_T*8/16;
	LRPT_DEV static float2 cvt(elem e) { return e; }
	LRPT_DEV static f32x2_t cvt2(elem e) { return pk2(e.x, e.y); }
	LRPT_DEV static elem from_float(float2 v) { return v; }
	LRPT_DEV static void unpack(const uint4 &v, elem (&e)[16/sizeof(elem)])
	{
		e[0] = make_float2(__uint_as_float(v.x), __uint_as_float(v.y));
		e[1] = make_float2(__uint_as_float(v.z), __uint_as_float(v.w));
	}
	LRPT_DEV static elem load1(const uint8_t *row, int idx) { return reinterpret_cast<const float2 *>(row)[idx]; }
	LRPT_DEV static elem zero() { return make_float2(0.f, 0.f); }
};

/* 8-bit samples as a bfloat16 pair: an integer in [-128, 127] has at most 8 significant bits, so it IS a
 * bfloat16, and a bfloat16 becomes the float of the same value by a 16-bit shift -- one integer operation
 * per component and no subtraction (a tap costs 5.75 instructions instead of 6.75), for 4 bytes per entry. */
struct FmtBF16 {
	typedef uint32_t elem;                       /* I in the low half, Q in the high half */
	LRPT_DEV static float2 cvt(elem e) { return make_float2(__uint_as_float(e << 16), __uint_as_float(e & 0xffff0000u)); }
	LRPT_DEV static f32x2_t cvt2(elem e) { return pk2(__uint_as_float(e << 16), __uint_as_float(e & 0xffff0000u)); }
	LRPT_DEV static elem from_float(float2 v) { return __byte_perm(__float_as_uint(v.x), __float_as_uint(v.y), 0x7632); }
};

/* What the window holds (WF): 0 the raw element (smallest footprint, converted at every tap), 1 the float
 * pair it converts to (converted once per sample when its tile is appended; 8 bytes per entry, so fewer
 * warps fit, but a tap costs three instructions less), 2 (8-bit input only) a bfloat16 pair. */
template <int BPS, int WF> struct WinT {
	typedef RawT<BPS> W;                          /* format the FIR reads */
	typedef typename RawT<BPS>::elem elem;
	static constexpr int WBPS = BPS;              /* bytes per entry * 4 */
	LRPT_DEV static elem from_raw(typename RawT<BPS>::elem e) { return e; }
};
template <int BPS> struct WinT<BPS, 1> {
	typedef RawT<32> W;
	typedef float2 elem;
	static constexpr int WBPS = 32;
	LRPT_DEV static elem from_raw(typename RawT<BPS>::elem e) { return RawT<BPS>::cvt(e); }
};
template <> struct WinT<8, 2> {
	typedef FmtBF16 W;
	typedef uint32_t elem;
	static constexpr int WBPS = 16;
	LRPT_DEV static elem from_raw(RawT<8>::elem e) { return FmtBF16::from_float(RawT<8>::cvt(e)); }
};

/* One FULL tile (LN_T samples starting at s0) of a lane's row into registers. */
template <int BPS>
LRPT_DEV void tile_load(const uint8_t *row, int s0, uint4 (&pf)[RawT<BPS>::NV])
{
	typedef RawT<BPS> R;
	const uint4 *src = reinterpret_cast<const uint4 *>(row + (size_t)s0*sizeof(typename R::elem));
#pragma unroll
	for (int j = 0; j < R::NV; j++) pf[j] = __ldg(src + j);
}

/* The last, partial tile goes from global memory to the window element by element. */
template <int BPS, int WF>
LRPT_DEV void tile_copy_partial(typename WinT<BPS, WF>::elem *col, int e0, const uint8_t *row, int s0, int nsamples)
{
	typedef RawT<BPS> R;
#pragma unroll 1
	for (int i = 0; i < LN_T; i++)
		col[(e0 + i)*32] = WinT<BPS, WF>::from_raw((s0 + i < nsamples) ? R::load1(row, s0 + i) : R::zero());
}

/* Registers -> window entries [e0, e0+LN_T) of this lane's column. */
template <int BPS, int WF>
LRPT_DEV void tile_store(typename WinT<BPS, WF>::elem *col, int e0, const uint4 (&pf)[RawT<BPS>::NV])
{
	typedef RawT<BPS> R;
	constexpr int PER = 16/sizeof(typename R::elem);
#pragma unroll
	for (int j = 0; j < R::NV; j++) {
		typename R::elem e[PER];
		R::unpack(pf[j], e);
#pragma unroll
		for (int i = 0; i < PER; i++) col[(e0 + j*PER + i)*32] = WinT<BPS, WF>::from_raw(e[i]);
	}
}

/* filter_get(flt, i), filter.c:46-65: w = this lane's column at the entry of the oldest sample,
 * hb = this lane's bank of the tap table (16-byte aligned, taps contiguous). acc = acc + x*h,
 * oldest first, multiply and add rounded separately; (re, im) ride one packed f32x2 chain.
 * Coefficients arrive four at a time (lanes on different banks hit different bank groups). */
#if LRPT_LANE_PACKED
#define LN_TAP(X, H) acc = fma2(mul2(R::cvt2(X), pk2((H), (H))), one2, acc)
#else
#define LN_TAP(X, H) do { const float2 x_ = R::cvt(X); ar = __fadd_rn(ar, __fmul_rn(x_.x, (H))); ai = __fadd_rn(ai, __fmul_rn(x_.y, (H))); } while (0)
#endif
template <class R>
LRPT_DEV float2 fir_lazy(const typename R::elem *__restrict__ w, const float *__restrict__ hb, int taps, float one)
{
#if LRPT_LANE_PACKED
	f32x2_t acc = pk2(0.0f, 0.0f);
	const f32x2_t one2 = pk2(one, one);
#else
	(void)one;
	float ar = 0.0f, ai = 0.0f;
#endif
	int left = taps;
#pragma unroll 1
	for (; left >= 16; left -= 16, w += 16*32, hb += 16) {
#pragma unroll
		for (int q = 0; q < 4; q++) {
			const float4 h4 = *reinterpret_cast<const float4 *>(hb + 4*q);
			LN_TAP(w[(4*q + 0)*32], h4.x); LN_TAP(w[(4*q + 1)*32], h4.y);
			LN_TAP(w[(4*q + 2)*32], h4.z); LN_TAP(w[(4*q + 3)*32], h4.w);
		}
	}
#pragma unroll 1
	for (; left >= 4; left -= 4, w += 4*32, hb += 4) {
		const float4 h4 = *reinterpret_cast<const float4 *>(hb);
		LN_TAP(w[0*32], h4.x); LN_TAP(w[1*32], h4.y); LN_TAP(w[2*32], h4.z); LN_TAP(w[3*32], h4.w);
	}
#pragma unroll 1
	for (; left > 0; left--, w += 32, hb++) LN_TAP(w[0], hb[0]);
#if LRPT_LANE_PACKED
	return upk2(acc);
#else
	return make_float2(ar, ai);
#endif
}

/* The same filter with the DEFERRED half of the previous symbol step (demod_core.cuh, "split in two")
 * placed in the basic block of the first 64 taps: ptxas interleaves its long double-precision
 * chains with the issue-bound tap arithmetic, so they no longer cost time of their own. */
template <class R, bool OQ>
LRPT_DEV float2 fir_lazy_with_deferred(const typename R::elem *__restrict__ w, const float *__restrict__ hb,
                                       int taps, float one, Loop &r, const lrpt_consts_t &c, const float *lut,
                                       const Pend &pd, Osc &next, bool &ok)
{
#if LRPT_LANE_PACKED
	f32x2_t acc = pk2(0.0f, 0.0f);
	const f32x2_t one2 = pk2(one, one);
#else
	(void)one;
	float ar = 0.0f, ai = 0.0f;
#endif
	int left = taps;
	if (left >= 64) {
		ok = step_deferred_fast<OQ>(r, c, lut, pd, next);
#pragma unroll
		for (int q = 0; q < 16; q++) {
			const float4 h4 = *reinterpret_cast<const float4 *>(hb + 4*q);
			LN_TAP(w[(4*q + 0)*32], h4.x); LN_TAP(w[(4*q + 1)*32], h4.y);
			LN_TAP(w[(4*q + 2)*32], h4.z); LN_TAP(w[(4*q + 3)*32], h4.w);
		}
		left -= 64; w += 64*32; hb += 64;
	} else {
		ok = step_deferred_fast<OQ>(r, c, lut, pd, next);
	}
#pragma unroll 1
	for (; left >= 16; left -= 16, w += 16*32, hb += 16) {
#pragma unroll
		for (int q = 0; q < 4; q++) {
			const float4 h4 = *reinterpret_cast<const float4 *>(hb + 4*q);
			LN_TAP(w[(4*q + 0)*32], h4.x); LN_TAP(w[(4*q + 1)*32], h4.y);
			LN_TAP(w[(4*q + 2)*32], h4.z); LN_TAP(w[(4*q + 3)*32], h4.w);
		}
	}
#pragma unroll 1
	for (; left >= 4; left -= 4, w += 4*32, hb += 4) {
		const float4 h4 = *reinterpret_cast<const float4 *>(hb);
		LN_TAP(w[0*32], h4.x); LN_TAP(w[1*32], h4.y); LN_TAP(w[2*32], h4.z); LN_TAP(w[3*32], h4.w);
	}
#pragma unroll 1
	for (; left > 0; left--, w += 32, hb++) LN_TAP(w[0], hb[0]);
#if LRPT_LANE_PACKED
	return upk2(acc);
#else
	return make_float2(ar, ai);
#endif
}
#undef LN_TAP

/* ------------------------------------------------------------- kernel ------ */

template <bool OQ, int BPS, bool AUX, int WF>
__global__ void __launch_bounds__(32*LN_MAX_WARPS, 1)
demod_lane_kernel(const lrpt_consts_t c, const LaneArgs a)
{
	typedef RawT<BPS> R;
	typedef WinT<BPS, WF> WT;
	typedef typename WT::W WR;                                      /* the window's format */
	typedef typename WT::elem elem;
	constexpr int T = LN_T;
	extern __shared__ __align__(16) unsigned char smem_raw[];

	const int taps = c.taps, H = taps - 1, L = c.interp;
	const int NT = a.NT;
	const int NE = H + NT*T;                                        /* window entries */
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	float *lut = reinterpret_cast<float *>(smem_raw);               /* [32] */
	float *hT  = lut + 32;                                          /* [L][TS]: bank p, tap k at hT[p*TS + k] */
	elem *wins = reinterpret_cast<elem *>(hT + L*a.TS);            /* [W][NE][32] */

	if (threadIdx.x < 32) lut[threadIdx.x] = c.lut_tanh[threadIdx.x];
	for (int i = threadIdx.x; i < L*a.TS; i += blockDim.x) {
		const int p = i/a.TS, k = i - p*a.TS;
		hT[i] = (k < taps) ? a.taps[p*taps + k] : 0.0f;
	}
	__syncthreads();

	const int local = (blockIdx.x*a.W + warp)*32 + lane;            /* launch-local stream index */
	if ((blockIdx.x*a.W + warp)*32 >= a.nstreams) return;           /* whole warp without streams */
	const bool active = local < a.nstreams;
	const int lrow = active ? local : a.nstreams - 1;               /* idle lanes shadow a valid row (loads only) */
	const int sid = a.first_stream + local;
	const uint8_t *row = a.raw + (size_t)lrow*a.raw_stride;
	elem *col = wins + (size_t)warp*NE*32 + lane;

	Loop r;
	long long nsymbols = 0, first_lock = -1;
	unsigned off = 0, nsym = 0;
	int lock_at = -1;                                               /* symbol of this launch at which the PLL had locked once */
	char2 *out = nullptr; float2 *outf = nullptr; uint32_t *outq = nullptr;
	loop_load(r, a.states[a.first_stream + lrow]);
	if (active) {
		nsymbols = a.states[sid].nsymbols;
		first_lock = a.states[sid].first_lock_symbol;
		off = a.out_off ? a.out_off[local] : 0u;
		out = reinterpret_cast<char2 *>(a.soft + (size_t)local*a.soft_stride);
		if (AUX && a.symf) outf = reinterpret_cast<float2 *>(reinterpret_cast<char *>(a.symf) + (size_t)local*a.symf_stride);
		if (AUX && a.symq) outq = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(a.symq) + (size_t)local*a.symq_stride);
	}

	/* prologue: delay line (taps-1 samples, oldest first) at entries [0,H), tile 0 behind it */
	{
		const float2 *hs = a.hist + (size_t)(a.first_stream + lrow)*H;
		for (int j = 0; j < H; j++) col[j*32] = WR::from_float(hs[j]);
	}
	uint4 pf[R::NV];
	if (T <= a.nsamples) tile_load<BPS>(row, 0, pf);

	const int ntiles = (a.nsamples + T - 1)/T;
	const int Qend = a.nsamples*L;
	int Q = 0;
	bool have_x = false; int Qx = 0, half = 0;
#if LRPT_LANE_PIPE
	Osc osc; osc.s = fast_sin(-r.p_phase); osc.co = fast_cos(-r.p_phase); osc.bad = false;   /* pll.c:53-54 */
	Pend pd; pd.sr = pd.si = pd.ore = pd.oim = 1.0f; pd.half = 0;
	bool pend = false;                                              /* a deferred half is outstanding */
	int pd_q = 0;
#else
	Osc osc = osc_for(r.p_phase);                                   /* fast_sin/fast_cos(-p_phase), pll.c:53-54 */
#endif
	int ep0 = 0;                                                    /* first sample of the window epoch */

	for (int t = 0; t < ntiles; t++) {
		/* append tile t; start the loads of tile t+1 */
		const int te = t - (ep0/T);
		if ((t + 1)*T <= a.nsamples) tile_store<BPS, WF>(col, H + te*T, pf);
		else tile_copy_partial<BPS, WF>(col, H + te*T, row, t*T, a.nsamples);
		if ((t + 2)*T <= a.nsamples) tile_load<BPS>(row, (t + 1)*T, pf);
		__syncwarp();

		const int q0 = t*T*L;
		const int q1 = min((t + 1)*T, a.nsamples)*L;
		/* Warp-uniform rounds, so that the lanes (streams) stay converged: in each round every
		 * lane that still owes this tile a crossing runs its NCO search, then every lane holding
		 * a crossing inside the tile takes its FIR + symbol step, all together. */
		while (true) {
			if (active && !have_x && Q < q1)
				have_x = nco_to_crossing4(r, c, a.nco_n0, Q, q1, Qend, Qx, half);
			__syncwarp();
			const bool ready = active && have_x && Qx < q1;
			if (!__any_sync(0xffffffffu, ready)) break;
			if (ready) {
				/* filter_get(flt, i) at sub-step Qx = n*L + i (demod.c:33-35) */
				const int dq = Qx - q0;
				const int nr = (dq*a.div_magic) >> 16, i = dq - nr*L;       /* sample within the tile, sub-step */
#if LRPT_LANE_PIPE
				/* deferred half of the previous symbol, woven into this symbol's taps */
				const float s_gain = r.gain, s_pp = r.p_phase, s_pf = r.p_freq, s_pe = r.p_err;
				const int s_lk = r.locked, s_lo = r.locked_once, s_ud = r.updown;
				Osc next; bool ok;
				const float2 y = fir_lazy_with_deferred<WR, OQ>(col + (te*T + nr)*32, hT + (L - 1 - i)*a.TS, taps, a.one,
				                                                 r, c, lut, pd, next, ok);
				if (!(ok && pend)) {                                 /* rare: nothing was outstanding (first event of the
				                                                        launch), or a shortcut was not provably exact */
					r.gain = s_gain; r.p_phase = s_pp; r.p_freq = s_pf; r.p_err = s_pe;
					r.locked = s_lk; r.locked_once = s_lo; r.updown = s_ud;
					if (pend) step_deferred_exact<OQ>(r, c, lut, pd, next);
					else next = osc;
				}
				if (pend && !(OQ && pd.half == 1)) {
					lock_at = (r.locked_once && lock_at < 0) ? (int)nsym : lock_at;
					if (off + nsym < a.cap) {
						out[off + nsym] = make_char2((signed char)quantise(pd.ore), (signed char)quantise(pd.oim));
						if (AUX && outf) outf[off + nsym] = make_float2(pd.ore, pd.oim);
						if (AUX && outq) outq[off + nsym] = a.q_base + (uint32_t)pd_q;
					}
					nsym++;
				}
				osc = next;
				/* the half of THIS symbol the next timing decision waits for */
				step_critical<OQ>(r, c, half, y.x, y.y, osc.s, osc.co, pd);
				pend = true; pd_q = Qx;
#else
				const float2 y = fir_lazy<WR>(col + (te*T + nr)*32, hT + (L - 1 - i)*a.TS, taps, a.one);
				const Loop saved = r;
				float ore, oim; bool emitted; Osc next;
				if (!symbol_fast_osc<OQ>(r, c, lut, half, y.x, y.y, osc, ore, oim, emitted, next)) {
					r = saved;                                       /* a shortcut was not provably exact */
					emitted = symbol_event(r, c, lut, half, y.x, y.y, ore, oim);
					next = osc_for(r.p_phase);
				}
				osc = next;
				if (emitted) {
					lock_at = (r.locked_once && lock_at < 0) ? (int)nsym : lock_at;
					if (off + nsym < a.cap) {
						out[off + nsym] = make_char2((signed char)quantise(ore), (signed char)quantise(oim));
						if (AUX && outf) outf[off + nsym] = make_float2(ore, oim);
						if (AUX && outq) outq[off + nsym] = a.q_base + (uint32_t)Qx;
					}
					nsym++;
				}
#endif
				have_x = false;
			}
			__syncwarp();
		}

		/* window full: the last H samples become the head of the next epoch */
		if (te + 1 == NT && t + 1 < ntiles) {
			/* ascending in-place copy towards lower entries, 16 loads then 16 stores at a time (the loads of a
			 * batch overlap each other): safe when source and destination overlap, because a batch's
			 * destination lies below every source still to be read (NT*T >= 32 > 15) */
			const int shift = NT*T;
			int j = 0;
#pragma unroll 1
			for (; j + 16 <= H; j += 16) {
				elem v[16];
#pragma unroll
				for (int u = 0; u < 16; u++) v[u] = col[(shift + j + u)*32];
#pragma unroll
				for (int u = 0; u < 16; u++) col[(j + u)*32] = v[u];
			}
			for (; j < H; j++) col[j*32] = col[(shift + j)*32];
			ep0 += NT*T;
		}
	}

#if LRPT_LANE_PIPE
	if (active && pend) {                                           /* the last symbol's deferred half */
		Osc next;
		step_deferred_exact<OQ>(r, c, lut, pd, next);
		if (!(OQ && pd.half == 1)) {
			lock_at = (r.locked_once && lock_at < 0) ? (int)nsym : lock_at;
			if (off + nsym < a.cap) {
				out[off + nsym] = make_char2((signed char)quantise(pd.ore), (signed char)quantise(pd.oim));
				if (AUX && outf) outf[off + nsym] = make_float2(pd.ore, pd.oim);
				if (AUX && outq) outq[off + nsym] = a.q_base + (uint32_t)pd_q;
			}
			nsym++;
		}
	}
#endif

	/* epilogue: the last taps-1 samples are the next call's delay line */
	if (active) {
		float2 *hs = a.hist + (size_t)sid*H;
		for (int j = 0; j < H; j++) {
			const int m = a.nsamples - H + j;                       /* may be negative: still in the old history */
			hs[j] = WR::cvt(col[(m - ep0 + H)*32]);
		}
		loop_store(r, a.states[sid]);
		a.states[sid].nsamples += a.nsamples;
		a.states[sid].nsymbols = nsymbols + nsym;
		a.states[sid].first_lock_symbol = (first_lock < 0 && lock_at >= 0) ? nsymbols + lock_at : first_lock;   /* main.c:312 */
		if (a.nsym_out) a.nsym_out[local] = nsym;
		if (a.out_off) a.out_off[local] = off + nsym;
	}
}

/* ------------------------------------------------------------- host side --- */

static int ln_num_sms = 0, ln_max_smem = 0;

static int ln_ts(int taps)
{
	int ts = 4*((taps + 3)/4);
	if (((ts/4) & 1) == 0) ts += 4;                /* TS/4 odd: banks of the table start in different bank groups */
	return ts;
}

static size_t ln_fixed_smem(int taps, int L) { return 32*sizeof(float) + (size_t)L*ln_ts(taps)*sizeof(float); }

static size_t ln_warp_smem(int taps, int NT, int bps) { return (size_t)((taps - 1) + NT*LN_T)*32*(size_t)(bps/4); }

bool lane_supported(const lrpt_consts_t &c)
{
	if (!(c.interp >= 1 && c.interp <= LN_MAX_L && c.taps >= 1 && c.taps <= LN_MAX_TAPS &&
	      (c.bps == 8 || c.bps == 16 || c.bps == 32))) return false;
	/* one warp's delay lines (shortest epoch) + the tap table must fit the 227 KB a CTA can opt into on sm_100 */
	return ln_fixed_smem(c.taps, c.interp) + ln_warp_smem(c.taps, 1, c.bps) <= (size_t)232448;
}

template <bool OQ, int BPS, int WF> static cudaError_t ln_attr2()
{
	cudaError_t e = cudaFuncSetAttribute(demod_lane_kernel<OQ, BPS, false, WF>, cudaFuncAttributeMaxDynamicSharedMemorySize, ln_max_smem);
	if (e) return e;
	return cudaFuncSetAttribute(demod_lane_kernel<OQ, BPS, true, WF>, cudaFuncAttributeMaxDynamicSharedMemorySize, ln_max_smem);
}

template <bool OQ> static cudaError_t ln_attr1()
{
	cudaError_t e;
	if ((e = ln_attr2<OQ, 32, 0>()) || (e = ln_attr2<OQ, 16, 0>()) || (e = ln_attr2<OQ, 16, 1>()) ||
	    (e = ln_attr2<OQ, 8, 0>()) || (e = ln_attr2<OQ, 8, 1>()) || (e = ln_attr2<OQ, 8, 2>())) return e;
	return cudaSuccess;
}

cudaError_t lane_prepare(int device)
{
	cudaError_t e;
	if ((e = cudaDeviceGetAttribute(&ln_num_sms, cudaDevAttrMultiProcessorCount, device))) return e;
	if ((e = cudaDeviceGetAttribute(&ln_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device))) return e;
	if ((e = ln_attr1<false>()) || (e = ln_attr1<true>())) return e;
	return cudaSuccess;
}

template <bool OQ> static void ln_launch1(const lrpt_consts_t &c, const LaneArgs &w, int blocks, size_t smem, cudaStream_t st, int wf)
{
	const int threads = 32*w.W;
	const bool aux = w.symf || w.symq;                              /* optional float / index side outputs */
#define LN_GO(B, F) do { if (aux) demod_lane_kernel<OQ, B, true, F><<<blocks, threads, smem, st>>>(c, w); \
                         else     demod_lane_kernel<OQ, B, false, F><<<blocks, threads, smem, st>>>(c, w); } while (0)
	if (c.bps == 16)     { if (wf == 1) LN_GO(16, 1); else LN_GO(16, 0); }
	else if (c.bps == 8) { if (wf == 1) LN_GO(8, 1); else if (wf == 2) LN_GO(8, 2); else LN_GO(8, 0); }
	else                 LN_GO(32, 0);
#undef LN_GO
}

/* Warps per CTA and tiles per window epoch for `nwarps` warps of streams and a window entry of wbps/4
 * bytes: spread the batch over every SM first, then stack warps (one CTA per SM). The epoch is at
 * least 2 tiles (the head move is an in-place ascending copy, so it may overlap its source) and as
 * long as shared memory allows, which amortises that move. false: one warp does not fit. */
static bool ln_plan(int nwarps, int taps, int L, int wbps, int &W, int &NT)
{
	const size_t fixed = ln_fixed_smem(taps, L), lim = (size_t)ln_max_smem;
	W = std::max(1, std::min((nwarps + ln_num_sms - 1)/ln_num_sms, LN_MAX_WARPS));
	NT = 2;
	while (W > 1 && fixed + W*ln_warp_smem(taps, NT, wbps) > lim) W--;
	if (fixed + W*ln_warp_smem(taps, NT, wbps) > lim) NT = 1;
	if (fixed + W*ln_warp_smem(taps, NT, wbps) > lim) return false;
	/* level the waves when shared memory caps W */
	const int per_wave = ln_num_sms*W;
	const int waves = (nwarps + per_wave - 1)/per_wave;
	W = std::max(1, std::min(W, (nwarps + waves*ln_num_sms - 1)/(waves*ln_num_sms)));
	while (NT < 16 && fixed + W*ln_warp_smem(taps, NT + 1, wbps) <= lim) NT++;
	return true;
}

/* Window format for a launch (WinT). A wider entry saves instructions per tap (float pairs three, bfloat16
 * pairs one) but costs shared memory; measured (tools/ab_wf.py, B200): float pairs +2..10 % when the same
 * number of warps per SM fits either way, -18 % when they halve it. So: the widest format that does not
 * cost a warp. LRPT_LANE_WF=0/1/2 overrides (A/B runs, tests). */
static int ln_window_format(const lrpt_consts_t &c, int nwarps)
{
	if (c.bps == 32) return 0;
	if (const char *e = getenv("LRPT_LANE_WF")) {
		const int v = atoi(e);
		return (v == 1 || (v == 2 && c.bps == 8)) ? v : 0;
	}
	int Wr, NTr, Wf, NTf;
	if (!ln_plan(nwarps, c.taps, c.interp, c.bps, Wr, NTr)) return 0;
	if (ln_plan(nwarps, c.taps, c.interp, 32, Wf, NTf) && Wf == Wr && NTf >= 2) return 1;
	if (c.bps == 8 && ln_plan(nwarps, c.taps, c.interp, 16, Wf, NTf) && Wf == Wr && NTf >= 2) return 2;
	return 0;
}

cudaError_t launch_lane(const LaunchArgs &a, cudaStream_t st, int *launches)
{
	const lrpt_consts_t &c = *a.c;
	const int L = c.interp, taps = c.taps;
	const int nwarps = (a.nstreams + 31)/32;
	const int wf = ln_window_format(c, nwarps);
	const int wbps = wf == 1 ? 32 : wf == 2 ? 16 : c.bps;           /* window entry: wbps/4 bytes */
	const size_t fixed = ln_fixed_smem(taps, L);
	int W, NT;
	if (!ln_plan(nwarps, taps, L, wbps, W, NT)) return cudaErrorInvalidConfiguration;
	const int blocks = (nwarps + W - 1)/W;
	const size_t smem = fixed + W*ln_warp_smem(taps, NT, wbps);
	if (getenv("LRPT_LANE_DEBUG")) fprintf(stderr, "lane: streams %d W %d NT %d blocks %d smem %zu wf %d\n", a.nstreams, W, NT, blocks, smem, wf);

	int n = 0;
	size_t done = 0;
	while (done < a.nsamples) {
		const size_t ns = std::min(a.nsamples - done, (size_t)WS_MAX_SAMPLES);
		if (done && !a.d_out_off) return cudaErrorInvalidValue;
		LaneArgs w;
		w.taps = a.d_taps; w.states = a.d_states; w.hist = a.d_hist;
		w.raw = reinterpret_cast<const uint8_t *>(a.d_raw) + done*(size_t)(c.bps/4); w.raw_stride = a.raw_stride;
		w.nsamples = (int)ns;
		w.soft = a.d_soft; w.soft_stride = a.soft_stride; w.symf = a.d_symf; w.symf_stride = a.symf_stride;
		w.symq = a.d_symq; w.symq_stride = a.symq_stride; w.q_base = (uint32_t)(done*(size_t)L);
		w.cap = a.cap; w.nsym_out = a.d_nsym; w.out_off = a.d_out_off;
		w.first_stream = a.first_stream; w.nstreams = a.nstreams; w.W = W; w.NT = NT;
		{
			const double nominal = (c.oqpsk ? 3.14159265358979 : 6.28318530717959)/(double)c.t_center;
			const int cmin = (int)nominal - 1;
			w.nco_n0 = cmin > 1 ? cmin - 1 : 0;                      /* tested sums: cmin .. cmin+3 */
		}
		w.TS = ln_ts(taps);
		w.one = 1.0f;
		w.div_magic = (65536 + L - 1)/L;
		for (int x = 0; x < LN_T*L; x++)
			if (((x*w.div_magic) >> 16) != x/L) return cudaErrorInvalidConfiguration;
		if (c.oqpsk) ln_launch1<true>(c, w, blocks, smem, st, wf);
		else         ln_launch1<false>(c, w, blocks, smem, st, wf);
		n++;
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) { if (launches) *launches = n; return e; }
		done += ns;
	}
	if (launches) *launches = n;
	return cudaSuccess;
}

} // namespace lrpt
