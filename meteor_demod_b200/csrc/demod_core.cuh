/*
 * demod_core.cuh -- the per-symbol recurrence as exact device arithmetic.
 *
 * Every function reproduces one reference function with the reference's operand
 * widths, evaluation order and IEEE round-to-nearest, and with NO fused
 * multiply-add (all float/double arithmetic goes through __f*_rn / __d*_rn
 * intrinsics, which the compiler never contracts). Bit-exactness against the
 * strict-IEEE build of the reference is the contract.
 *
 * Shared by the simple (one thread per stream) and the warp-specialised kernels.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "lrpt_internal.h"

namespace lrpt {

#define LRPT_DEV __device__ __forceinline__

constexpr float  kTwoPiF  = 6.28318548202514648f;      /* 2*(float)M_PI, timing.c:37 */
constexpr float  kPiF     = 3.14159274101257324f;      /* (float)M_PI,   timing.c:51 */
constexpr double kTwoPiD  = 6.28318530717958647692;    /* 2*M_PI                     */
constexpr double kHalfPiD = 1.57079632679489661923;    /* M_PI/2, sincos.c:39        */
constexpr double kInvTwoPiD = 0.15915494309189533577;  /* only used by the exact fast path below */

/* Running state of one demodulator, held in registers while a kernel runs. */
struct Loop {
	float t_phase, t_freq, t_prev;
	int   t_dual;
	float oq_inphase;
	float gain, bias_re, bias_im;
	float p_phase, p_freq, p_err;
	int   locked, locked_once, updown;
};

LRPT_DEV void loop_load(Loop &r, const lrpt_state_t &s)
{
	r.t_phase = s.t_phase; r.t_freq = s.t_freq; r.t_prev = s.t_prev; r.t_dual = s.t_dual_state;
	r.oq_inphase = s.oq_inphase;
	r.gain = s.agc_gain; r.bias_re = s.agc_bias_re; r.bias_im = s.agc_bias_im;
	r.p_phase = s.p_phase; r.p_freq = s.p_freq; r.p_err = s.p_err;
	r.locked = s.p_locked; r.locked_once = s.p_locked_once; r.updown = s.p_updown;
}

LRPT_DEV void loop_store(const Loop &r, lrpt_state_t &s)
{
	s.t_phase = r.t_phase; s.t_freq = r.t_freq; s.t_prev = r.t_prev; s.t_dual_state = r.t_dual;
	s.oq_inphase = r.oq_inphase;
	s.agc_gain = r.gain; s.agc_bias_re = r.bias_re; s.agc_bias_im = r.bias_im;
	s.p_phase = r.p_phase; s.p_freq = r.p_freq; s.p_err = r.p_err;
	s.p_locked = r.locked; s.p_locked_once = r.locked_once; s.p_updown = r.updown;
}

/* ------------------------------------------------------------------ sincos -- */

/*
 * fast_sin, sincos.c:13-34: phase -> Q16 turn fraction by a DOUBLE division and a
 * truncating conversion that wraps to int16, then a 4th-order Q14 polynomial.
 *
 * trunc(RN(x / 2pi)) is evaluated without a division: q1 = x * RN(1/2pi) is within
 * 3e-11 of both the real quotient and its correctly rounded double for |x| < 2^20,
 * so truncation agrees unless q1 lies within 1e-9 of an integer; only then the real
 * __ddiv_rn runs. The result is identical to the division in every case.
 */
static __device__ __noinline__ double turn_fraction_slow(double x) { return __ddiv_rn(x, kTwoPiD); }

LRPT_DEV float fast_sin(float fx)
{
	const double x = (double)__fmul_rn(fx, 65536.0f);
	double q = __dmul_rn(x, kInvTwoPiD);
	/* a call keeps the division out of the straight-line path (it is needed about once in 1e9) */
	if (fabs(q - rint(q)) < 1e-9 || !(fabs(q) < 1048576.0)) q = turn_fraction_slow(x);
	const int wide = __double2int_rz(q);
	int v = (int)(short)(wide & 0xffff);       /* int16 wrap, as cvttsd2si + 16-bit store does */
	const int sign = v;
	v = (v & 0x7fff) - 16384;
	const int v2 = (v*v) >> 14;
	int y = 19900 - ((v2*3516) >> 14);
	y = 16384 - ((v2*y) >> 14);
	return __fmul_rn((float)(sign < 0 ? -y : y), 6.103515625e-05f);   /* /16384, exact */
}

/* fast_cos, sincos.c:37-40: double add narrowed to the float parameter */
LRPT_DEV float fast_cos(float fx)
{
	return fast_sin(__double2float_rn(__dadd_rn((double)fx, kHalfPiD)));
}

/* ------------------------------------------------------------------- agc ---- */

/* libm cabsf for finite input: (float)sqrt((double)x*x + (double)y*y); both squares
 * are exact in double, so the single rounding of the sum is all there is. */
LRPT_DEV float cabsf_exact(float re, float im)
{
	const double a = (double)re, b = (double)im;
	return __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b))));
}

/* agc_apply, agc.c:13-25. Returns the scaled sample (old gain), updates bias + gain. */
LRPT_DEV void agc_apply(Loop &r, float &re, float &im)
{
	const float keep = 1.0f - 0.001f;
	r.bias_re = __fadd_rn(__fmul_rn(r.bias_re, keep), __fmul_rn(0.001f, re));
	r.bias_im = __fadd_rn(__fmul_rn(r.bias_im, keep), __fmul_rn(0.001f, im));
	const float sr = __fmul_rn(__fsub_rn(re, r.bias_re), r.gain);
	const float si = __fmul_rn(__fsub_rn(im, r.bias_im), r.gain);
	const float g = __fadd_rn(r.gain, __fmul_rn(0.0001f, __fsub_rn(190.0f, cabsf_exact(sr, si))));
	r.gain = (0.0f > g) ? 0.0f : g;
	re = sr; im = si;
}

/* ------------------------------------------------------------------- pll ---- */

/* NCO advance common to pll_mix / pll_mix_i / pll_mix_q, pll.c:61-62 */
LRPT_DEV void pll_advance(Loop &r)
{
	r.p_phase = __fadd_rn(r.p_phase, r.p_freq);
	/* (double)p >= 2*M_PI  <=>  p >= kTwoPiF for float p: kTwoPiF is the float just above 2*M_PI */
	if (r.p_phase >= kTwoPiF)
		r.p_phase = __double2float_rn(__dsub_rn((double)r.p_phase, kTwoPiD));
}

/* lut_tanh, pll.c:154-159 */
LRPT_DEV float lut_tanh(const float *lut, float v)
{
	/* v > 15 -> 1 and v < -16 -> -1 coincide with the clamped lookups because
	 * lut[31] = (float)tanh(15) = 1 and lut[0] = (float)tanh(-16) = -1 exactly. */
	const int i = min(max(__float2int_rz(v), -16), 15);
	return lut[i + 16];
}

static __device__ __noinline__ float fmod_two_pi_slow(float x) { return __double2float_rn(fmod((double)x, kTwoPiD)); }

/* pll_update_estimate -> compute_error + update_estimate, pll.c:100-130,143-152 */
LRPT_DEV void pll_update(Loop &r, const lrpt_consts_t &c, const float *lut, float i, float q)
{
	const float error = __fsub_rn(__fmul_rn(lut_tanh(lut, i), q), __fmul_rn(lut_tanh(lut, q), i));

	/* _phase = fmod(_phase + _alpha*error, 2*M_PI): float sum, double fmod (exact), narrowed.
	 * |x| < 2*M_PI <=> |x| < kTwoPiF for float x, and then fmod returns x itself. */
	const float ph = __fadd_rn(r.p_phase, __fmul_rn(c.p_alpha, error));
	r.p_phase = (fabsf(ph) < kTwoPiF) ? ph : fmod_two_pi_slow(ph);
	r.p_freq = __fadd_rn(r.p_freq, __fmul_rn(c.p_beta, error));

	/* lock detector: float product, double |e|*pole, double sum, narrowed */
	r.p_err = __double2float_rn(__dadd_rn((double)__fmul_rn(r.p_err, 1.0f - 0.001f),
	                                      __dmul_rn(fabs((double)error), (double)0.001f)));
	if (r.p_err < 85.0f && !r.locked) { r.locked = 1; r.locked_once = 1; }
	else if (r.p_err > 105.0f && r.locked) r.locked = 0;

	/* frequency sweep while unlocked (double add), direction flip, clamp */
	if (!r.locked)
		r.p_freq = __double2float_rn(__dadd_rn((double)r.p_freq, r.updown > 0 ? 0.000001 : -0.000001));
	r.updown = (r.p_freq >= c.p_fmax) ? -1 : (r.p_freq <= -c.p_fmax) ? 1 : r.updown;
	const float f = (c.p_fmax < r.p_freq) ? c.p_fmax : r.p_freq;
	r.p_freq = (-c.p_fmax > f) ? -c.p_fmax : f;
}

/* ----------------------------------------------------------------- timing --- */

/* retime -> mm_err -> update_estimate, timing.c:60-95. `cur` = imaginary part. */
LRPT_DEV void retime(Loop &r, const lrpt_consts_t &c, float cur)
{
	const float prev = r.t_prev;
	const float err = __fsub_rn(prev < 0.0f ? -cur : cur, cur < 0.0f ? -prev : prev);
	r.t_prev = cur;
	float fd = __fsub_rn(r.t_freq, c.t_center);
	r.t_phase = __double2float_rn(__dsub_rn((double)r.t_phase,
	                              __dadd_rn(kTwoPiD, (double)__fmul_rn(c.t_alpha, err))));
	fd = __fsub_rn(fd, __fmul_rn(c.t_beta, err));
	const float m = (c.t_maxdev < fd) ? c.t_maxdev : fd;
	fd = (-c.t_maxdev > m) ? -c.t_maxdev : m;
	r.t_freq = __fadd_rn(c.t_center, fd);
}

/* ----------------------------------------------------------------- egress --- */

/* main.c:305-306: MAX(-127, MIN(127, v/2)) in float, then truncation to int8 */
LRPT_DEV int quantise(float v)
{
	const float hlf = __fmul_rn(v, 0.5f);
	const float m = (127.0f < hlf) ? 127.0f : hlf;
	const float q = (-127.0f > m) ? -127.0f : m;
	return __float2int_rz(q);
}

/* ------------------------------------------------------------------ ingest -- */

/* wav_read's conversions, wavfile.c:58-69 (no scaling) */
LRPT_DEV float2 ingest(const void *raw, int bps, long long idx)
{
	if (bps == 16) {
		const short2 v = reinterpret_cast<const short2 *>(raw)[idx];
		return make_float2((float)v.x, (float)v.y);
	} else if (bps == 8) {
		const uchar2 v = reinterpret_cast<const uchar2 *>(raw)[idx];
		return make_float2((float)((int)v.x - 128), (float)((int)v.y - 128));
	}
	return reinterpret_cast<const float2 *>(raw)[idx];
}

/*
 * One symbol event of the recurrence, after the FIR output (re, im) for the chosen
 * sub-step is known. QPSK: demod.c:35-43. OQPSK: demod.c:66-83 (`half` = value
 * advance_timeslot_dual returned, 1 = I arm, 2 = Q arm + symbol).
 * Returns true when a symbol (out_re, out_im) was produced.
 */
LRPT_DEV bool symbol_event(Loop &r, const lrpt_consts_t &c, const float *lut, int half,
                           float re, float im, float &out_re, float &out_im)
{
	agc_apply(r, re, im);
	const float s = fast_sin(-r.p_phase);
	const float co = fast_cos(-r.p_phase);
	if (!c.oqpsk) {
		out_re = __fsub_rn(__fmul_rn(re, co), __fmul_rn(im, s));
		out_im = __fadd_rn(__fmul_rn(re, s), __fmul_rn(im, co));
		pll_advance(r);
		retime(r, c, out_im);
		pll_update(r, c, lut, out_re, out_im);
		return true;
	}
	if (half == 1) {
		r.oq_inphase = __fsub_rn(__fmul_rn(re, co), __fmul_rn(im, s));
		pll_advance(r);
		return false;
	}
	out_im = __fadd_rn(__fmul_rn(re, s), __fmul_rn(im, co));
	pll_advance(r);
	out_re = r.oq_inphase;
	retime(r, c, out_im);
	pll_update(r, c, lut, out_re, out_im);
	return true;
}

/* ================================================================== fast path ==
 *
 * symbol_event above is the reference, statement by statement. symbol_fast computes
 * the SAME values with the common case laid out as one branch-free block (so the
 * independent chains -- two sin/cos evaluations, the AGC magnitude, the lock
 * detector, the timing update -- overlap in the pipeline) and reports `false`
 * whenever one of its shortcuts is not provably exact for the operands at hand;
 * the caller then restores the loop state and runs symbol_event. Shortcuts:
 *
 *  - turn fraction: trunc(x * RN(1/2pi)) instead of trunc(RN(x / 2pi)); identical
 *    unless the product lies within 1e-9 of an integer (see fast_sin) -> fallback;
 *  - cabsf: sqrt by one Newton step on rsqrt.approx in double (relative error
 *    < 2^-43); accepted only if the estimate scaled by (1 -+ 2^-40) rounds to the
 *    same float, which then equals (float)sqrt(double) because rounding is monotone;
 *  - fmod(x, 2pi) = x for |x| < 2pi; anything else -> fallback.
 * Every other operation is the reference's own, with selects instead of branches.
 */

LRPT_DEV float sincos_poly(int wide)
{
	int v = (int)(short)(wide & 0xffff);
	const int sign = v;
	v = (v & 0x7fff) - 16384;
	const int v2 = (v*v) >> 14;
	int y = 19900 - ((v2*3516) >> 14);
	y = 16384 - ((v2*y) >> 14);
	return __fmul_rn((float)(sign < 0 ? -y : y), 6.103515625e-05f);
}

/* turn fraction of fast_sin's argument; sets `bad` when the division-free form is not provably exact */
LRPT_DEV int turn_fraction_fast(float fx, bool &bad)
{
	const double q = __dmul_rn((double)__fmul_rn(fx, 65536.0f), kInvTwoPiD);
	bad |= !(fabs(q - rint(q)) >= 1e-9) || !(fabs(q) < 1048576.0);
	return __double2int_rz(q);
}

/* (float)sqrt(s) for a double s = x*x + y*y, branch-free; sets `bad` when not provably exact */
LRPT_DEV float sqrt_to_float_fast(double s, bool &bad)
{
	const float sf = __double2float_rn(s);
	float y0;
	asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(sf));
	const double y = (double)y0;
	const double g = s*y;                       /* ~ sqrt(s), relative error < 2^-21.9 */
	const double h = 0.5*y;
	const double g1 = fma(g, fma(-g, h, 0.5), g);   /* one Newton step: relative error < 2^-43 */
	const float lo = __double2float_rn(g1*(1.0 - 0x1p-40));
	const float hi = __double2float_rn(g1*(1.0 + 0x1p-40));
	bad |= !(lo == hi) || !(sf > 1e-30f) || !(sf < 1e30f);
	return lo;
}

template <bool OQ>
LRPT_DEV bool symbol_fast(Loop &r, const lrpt_consts_t &c, const float *lut, int half,
                          float re, float im, float &out_re, float &out_im, bool &emitted)
{
	bool bad = false;

	/* pll_mix's oscillator values for the CURRENT phase (pll.c:53-54) */
	const float nph = -r.p_phase;
	const float s = sincos_poly(turn_fraction_fast(nph, bad));
	const float co = sincos_poly(turn_fraction_fast(__double2float_rn(__dadd_rn((double)nph, kHalfPiD)), bad));

	/* agc_apply (agc.c:13-25) */
	const float keep = 1.0f - 0.001f;
	r.bias_re = __fadd_rn(__fmul_rn(r.bias_re, keep), __fmul_rn(0.001f, re));
	r.bias_im = __fadd_rn(__fmul_rn(r.bias_im, keep), __fmul_rn(0.001f, im));
	const float sr = __fmul_rn(__fsub_rn(re, r.bias_re), r.gain);
	const float si = __fmul_rn(__fsub_rn(im, r.bias_im), r.gain);
	const double da = (double)sr, db = (double)si;
	const float mag = sqrt_to_float_fast(__dadd_rn(__dmul_rn(da, da), __dmul_rn(db, db)), bad);
	const float g = __fadd_rn(r.gain, __fmul_rn(0.0001f, __fsub_rn(190.0f, mag)));
	r.gain = (0.0f > g) ? 0.0f : g;

	/* NCO advance (pll.c:61-62), wrap by select */
	const float p1 = __fadd_rn(r.p_phase, r.p_freq);
	const float pw = __double2float_rn(__dsub_rn((double)p1, kTwoPiD));
	const float p2 = (p1 >= kTwoPiF) ? pw : p1;

	if (OQ && half == 1) {                                  /* demod.c:66-71: I arm only */
		r.oq_inphase = __fsub_rn(__fmul_rn(sr, co), __fmul_rn(si, s));
		r.p_phase = p2;
		emitted = false;
		return !bad;
	}
	if (OQ) {                                               /* demod.c:72-83 */
		out_im = __fadd_rn(__fmul_rn(sr, s), __fmul_rn(si, co));
		out_re = r.oq_inphase;
	} else {                                                /* pll_mix, pll.c:60 */
		out_re = __fsub_rn(__fmul_rn(sr, co), __fmul_rn(si, s));
		out_im = __fadd_rn(__fmul_rn(sr, s), __fmul_rn(si, co));
	}

	retime(r, c, out_im);                                   /* timing.c:60-95, already branch-free */

	/* pll_update_estimate (pll.c:100-130) */
	const float error = __fsub_rn(__fmul_rn(lut_tanh(lut, out_re), out_im), __fmul_rn(lut_tanh(lut, out_im), out_re));
	const float ph = __fadd_rn(p2, __fmul_rn(c.p_alpha, error));
	bad |= !(fabsf(ph) < kTwoPiF);                          /* else fmod really reduces */
	r.p_phase = ph;
	const float f1 = __fadd_rn(r.p_freq, __fmul_rn(c.p_beta, error));
	r.p_err = __double2float_rn(__dadd_rn((double)__fmul_rn(r.p_err, 1.0f - 0.001f),
	                                      __dmul_rn(fabs((double)error), (double)0.001f)));
	const int was = r.locked;
	const int acquire = (r.p_err < 85.0f && !was) ? 1 : 0;
	const int now = acquire ? 1 : ((r.p_err > 105.0f && was) ? 0 : was);
	r.locked = now;
	r.locked_once |= acquire;
	const float fsw = __double2float_rn(__dadd_rn((double)f1, r.updown > 0 ? 0.000001 : -0.000001));
	const float f2 = now ? f1 : fsw;
	const int up1 = (f2 <= -c.p_fmax) ? 1 : r.updown;
	r.updown = (f2 >= c.p_fmax) ? -1 : up1;
	const float f3 = (c.p_fmax < f2) ? c.p_fmax : f2;
	r.p_freq = (-c.p_fmax > f3) ? -c.p_fmax : f3;
	emitted = true;
	return !bad;
}

/* ----------------------------------------------------- fast path, pipelined ----
 *
 * The Costas NCO phase changes only inside a symbol step, so the oscillator values the NEXT
 * step will use (fast_sin/fast_cos of -p_phase, pll.c:53-54) are known as soon as this step
 * has written p_phase. symbol_fast_osc takes them as an input and computes the next pair at
 * its end, where the work overlaps the tail of the PLL update instead of heading the
 * dependency chain of the next symbol. Same values, same proof conditions as symbol_fast.
 */
struct Osc { float s, co; bool bad; };

LRPT_DEV Osc osc_for(float p_phase)
{
	Osc o; o.bad = false;
	const float nph = -p_phase;
	o.s = sincos_poly(turn_fraction_fast(nph, o.bad));
	o.co = sincos_poly(turn_fraction_fast(__double2float_rn(__dadd_rn((double)nph, kHalfPiD)), o.bad));
	return o;
}

template <bool OQ>
LRPT_DEV bool symbol_fast_osc(Loop &r, const lrpt_consts_t &c, const float *lut, int half,
                              float re, float im, const Osc &osc, float &out_re, float &out_im,
                              bool &emitted, Osc &next)
{
	bool bad = osc.bad;
	const float s = osc.s, co = osc.co;

	/* agc_apply (agc.c:13-25) */
	const float keep = 1.0f - 0.001f;
	r.bias_re = __fadd_rn(__fmul_rn(r.bias_re, keep), __fmul_rn(0.001f, re));
	r.bias_im = __fadd_rn(__fmul_rn(r.bias_im, keep), __fmul_rn(0.001f, im));
	const float sr = __fmul_rn(__fsub_rn(re, r.bias_re), r.gain);
	const float si = __fmul_rn(__fsub_rn(im, r.bias_im), r.gain);
	const double da = (double)sr, db = (double)si;
	const float mag = sqrt_to_float_fast(__dadd_rn(__dmul_rn(da, da), __dmul_rn(db, db)), bad);
	const float g = __fadd_rn(r.gain, __fmul_rn(0.0001f, __fsub_rn(190.0f, mag)));
	r.gain = (0.0f > g) ? 0.0f : g;

	/* NCO advance (pll.c:61-62), wrap by select */
	const float p1 = __fadd_rn(r.p_phase, r.p_freq);
	const float pw = __double2float_rn(__dsub_rn((double)p1, kTwoPiD));
	const float p2 = (p1 >= kTwoPiF) ? pw : p1;

	if (OQ && half == 1) {                                  /* demod.c:66-71: I arm only */
		r.oq_inphase = __fsub_rn(__fmul_rn(sr, co), __fmul_rn(si, s));
		r.p_phase = p2;
		emitted = false;
		next = osc_for(p2);
		return !bad;
	}
	if (OQ) {                                               /* demod.c:72-83 */
		out_im = __fadd_rn(__fmul_rn(sr, s), __fmul_rn(si, co));
		out_re = r.oq_inphase;
	} else {                                                /* pll_mix, pll.c:60 */
		out_re = __fsub_rn(__fmul_rn(sr, co), __fmul_rn(si, s));
		out_im = __fadd_rn(__fmul_rn(sr, s), __fmul_rn(si, co));
	}

	/* phase first: the next oscillator pair hangs off it */
	const float error = __fsub_rn(__fmul_rn(lut_tanh(lut, out_re), out_im), __fmul_rn(lut_tanh(lut, out_im), out_re));
	const float ph = __fadd_rn(p2, __fmul_rn(c.p_alpha, error));
	bad |= !(fabsf(ph) < kTwoPiF);                          /* else fmod really reduces */
	r.p_phase = ph;
	next = osc_for(ph);

	retime(r, c, out_im);                                   /* timing.c:60-95, already branch-free */

	/* rest of pll_update_estimate (pll.c:100-130) */
	const float f1 = __fadd_rn(r.p_freq, __fmul_rn(c.p_beta, error));
	r.p_err = __double2float_rn(__dadd_rn((double)__fmul_rn(r.p_err, 1.0f - 0.001f),
	                                      __dmul_rn(fabs((double)error), (double)0.001f)));
	const int was = r.locked;
	const int acquire = (r.p_err < 85.0f && !was) ? 1 : 0;
	const int now = acquire ? 1 : ((r.p_err > 105.0f && was) ? 0 : was);
	r.locked = now;
	r.locked_once |= acquire;
	const float fsw = __double2float_rn(__dadd_rn((double)f1, r.updown > 0 ? 0.000001 : -0.000001));
	const float f2 = now ? f1 : fsw;
	const int up1 = (f2 <= -c.p_fmax) ? 1 : r.updown;
	r.updown = (f2 >= c.p_fmax) ? -1 : up1;
	const float f3 = (c.p_fmax < f2) ? c.p_fmax : f2;
	r.p_freq = (-c.p_fmax > f3) ? -c.p_fmax : f3;
	emitted = true;
	return !bad;
}

/* ------------------------------------------------- fast path, split in two ----
 *
 * Within one stream, what the NEXT timing decision waits for is short: delay-line output ->
 * bias/gain scaling -> NCO mix -> retime (demod.c:35-39). Everything else a symbol step does --
 * the AGC magnitude and gain update (agc.c:20-22), the Costas NCO advance, phase-error, loop
 * and lock-detector update (pll.c:61-62,100-130), the oscillator values of the new phase -- is
 * not needed before the next filter output has been MIXED. demod_lane.cu therefore runs
 * step_critical at once and weaves step_deferred_fast of symbol k into the tap loop of symbol
 * k+1, where its long double-precision chains cost no time. Values and order of evaluation per
 * variable are unchanged; step_deferred_exact is the statement-by-statement version the caller
 * falls back to when a shortcut of the fast one is not provably exact (same rules as symbol_fast).
 */
struct Pend { float sr, si, ore, oim; int half; };

template <bool OQ>
LRPT_DEV void step_critical(Loop &r, const lrpt_consts_t &c, int half, float re, float im,
                            float s, float co, Pend &p)
{
	const float keep = 1.0f - 0.001f;                               /* agc_apply up to the scaled sample */
	r.bias_re = __fadd_rn(__fmul_rn(r.bias_re, keep), __fmul_rn(0.001f, re));
	r.bias_im = __fadd_rn(__fmul_rn(r.bias_im, keep), __fmul_rn(0.001f, im));
	const float sr = __fmul_rn(__fsub_rn(re, r.bias_re), r.gain);
	const float si = __fmul_rn(__fsub_rn(im, r.bias_im), r.gain);
	p.sr = sr; p.si = si; p.half = half;
	if (OQ && half == 1) {                                          /* demod.c:66-71 */
		r.oq_inphase = __fsub_rn(__fmul_rn(sr, co), __fmul_rn(si, s));
		return;
	}
	if (OQ) {                                                       /* demod.c:72-83 */
		p.oim = __fadd_rn(__fmul_rn(sr, s), __fmul_rn(si, co));
		p.ore = r.oq_inphase;
	} else {                                                        /* pll_mix, pll.c:60 */
		p.ore = __fsub_rn(__fmul_rn(sr, co), __fmul_rn(si, s));
		p.oim = __fadd_rn(__fmul_rn(sr, s), __fmul_rn(si, co));
	}
	retime(r, c, p.oim);                                            /* timing.c:60-95 */
}

/* branch-free; returns false when a shortcut was not provably exact (then nothing it wrote may be used) */
template <bool OQ>
LRPT_DEV bool step_deferred_fast(Loop &r, const lrpt_consts_t &c, const float *lut, const Pend &p, Osc &next)
{
	bool bad = false;
	const double da = (double)p.sr, db = (double)p.si;              /* agc.c:20-22 */
	const float mag = sqrt_to_float_fast(__dadd_rn(__dmul_rn(da, da), __dmul_rn(db, db)), bad);
	const float g = __fadd_rn(r.gain, __fmul_rn(0.0001f, __fsub_rn(190.0f, mag)));
	r.gain = (0.0f > g) ? 0.0f : g;

	const float p1 = __fadd_rn(r.p_phase, r.p_freq);                /* pll.c:61-62 */
	const float pw = __double2float_rn(__dsub_rn((double)p1, kTwoPiD));
	const float p2 = (p1 >= kTwoPiF) ? pw : p1;
	const bool arm_i = OQ && p.half == 1;                           /* I arm: NCO advance only */

	const float error = __fsub_rn(__fmul_rn(lut_tanh(lut, p.ore), p.oim), __fmul_rn(lut_tanh(lut, p.oim), p.ore));
	const float ph = arm_i ? p2 : __fadd_rn(p2, __fmul_rn(c.p_alpha, error));
	bad |= !(fabsf(ph) < kTwoPiF);                                  /* else fmod really reduces (pll.c:113) */
	r.p_phase = ph;
	next = osc_for(ph);
	bad |= next.bad;

	const float f1 = __fadd_rn(r.p_freq, __fmul_rn(c.p_beta, error));
	const float e1 = __double2float_rn(__dadd_rn((double)__fmul_rn(r.p_err, 1.0f - 0.001f),
	                                             __dmul_rn(fabs((double)error), (double)0.001f)));
	const int was = r.locked;
	const int acquire = (e1 < 85.0f && !was) ? 1 : 0;
	const int now = acquire ? 1 : ((e1 > 105.0f && was) ? 0 : was);
	const float fsw = __double2float_rn(__dadd_rn((double)f1, r.updown > 0 ? 0.000001 : -0.000001));
	const float f2 = now ? f1 : fsw;
	const int up1 = (f2 <= -c.p_fmax) ? 1 : r.updown;
	const int up2 = (f2 >= c.p_fmax) ? -1 : up1;
	const float f3 = (c.p_fmax < f2) ? c.p_fmax : f2;
	const float f4 = (-c.p_fmax > f3) ? -c.p_fmax : f3;
	r.p_err = arm_i ? r.p_err : e1;
	r.locked = arm_i ? was : now;
	r.locked_once |= arm_i ? 0 : acquire;
	r.updown = arm_i ? r.updown : up2;
	r.p_freq = arm_i ? r.p_freq : f4;
	return !bad;
}

template <bool OQ>
LRPT_DEV void step_deferred_exact(Loop &r, const lrpt_consts_t &c, const float *lut, const Pend &p, Osc &next)
{
	const float g = __fadd_rn(r.gain, __fmul_rn(0.0001f, __fsub_rn(190.0f, cabsf_exact(p.sr, p.si))));
	r.gain = (0.0f > g) ? 0.0f : g;
	pll_advance(r);
	if (!(OQ && p.half == 1)) pll_update(r, c, lut, p.ore, p.oim);
	next.s = fast_sin(-r.p_phase);
	next.co = fast_cos(-r.p_phase);
	next.bad = false;
}

} // namespace lrpt
