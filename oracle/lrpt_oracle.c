/*
 * oracle/lrpt_oracle.c -- TEST INFRASTRUCTURE ONLY (see lrpt_oracle.h).
 *
 * Restatement of the reference hot path as a re-entrant block transform. Every
 * expression keeps the reference's operand types and evaluation order; the
 * comments name the float/double width of each intermediate because that is
 * what the CUDA path has to reproduce bit for bit.
 *
 * Build: gcc -O2 -ffp-contract=off -std=gnu99 (no FMA contraction, IEEE RN).
 * Not handled identically to the reference: NaN/Inf inputs (the reference
 * indexes its tanh table out of bounds on NaN, pll.c:158) -- finite input only.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "lrpt_oracle.h"

#define PI_D      3.14159265358979323846   /* M_PI */
#define TWO_PI_D  (2*PI_D)

/* ---------------------------------------------------------------- taps -- */

/* filter.c:71-94. Types: t, coeff, interm, osf, alpha are float; every product
 * with M_PI and the window expression are double and narrowed on assignment. */
float
lrpt_oracle_rrc_coeff(int stage_no, unsigned taps, float osf, float alpha)
{
	const float norm = (float)(2.0/5.0);
	const int order = (int)((taps - 1)/2);
	float t, coeff, interm, fourat;
	double win;

	if (order == stage_no)
		return (float)((double)norm * ((double)(1 - alpha) + (double)(4*alpha)/PI_D));

	t = (float)abs(order - stage_no) / osf;
	fourat = 4*alpha*t;
	coeff = sinf((float)(PI_D*(double)t*(double)(1 - alpha)))
	      + fourat*cosf((float)(PI_D*(double)t*(double)(1 + alpha)));
	interm = (float)(PI_D*(double)t*(double)(1 - fourat*fourat));

	/* Blackman window (the reference comment says Hamming, filter.c:90-91) */
	win = 0.42 - 0.5*(double)cosf((float)(2*PI_D*stage_no/(double)(taps - 1)))
	           + 0.08*(double)cosf((float)(4*PI_D*stage_no/(double)(taps - 1)));
	coeff = (float)((double)coeff * win);

	return coeff / interm * norm;
}

/* ------------------------------------------------------------- sin/cos -- */

/* sincos.c:13-34. The float->int16 conversion is what gcc/x86-64 does: truncate
 * the double to int32 (cvttsd2si) and keep the low 16 bits. */
float
lrpt_oracle_fast_sin(float fx)
{
	const int32_t a = 1 << 14;
	const int32_t b = (int32_t)((2 - 3.14159/4)*(1 << 14));   /* 19900 */
	const int32_t c = b - (1 << 14);                          /* 3516  */
	double q = (double)(fx * 65536.0f) / TWO_PI_D;
	int32_t wide = (int32_t)q;              /* |q| < 2^31 for every reachable phase */
	int16_t x = (int16_t)(uint16_t)((uint32_t)wide & 0xffffu);
	int16_t sign = x;
	int32_t x2, y;

	x = (int16_t)(x & 0x7fff);
	x = (int16_t)(x - (1 << 14));
	x2 = ((int32_t)x * x) >> 14;
	y = b - ((x2 * c) >> 14);
	y = a - ((x2 * y) >> 14);

	return (float)(sign < 0 ? -y : y) / 16384.0f;
}

/* sincos.c:37-40 : double add, narrowed to the float parameter */
float
lrpt_oracle_fast_cos(float fx)
{
	return lrpt_oracle_fast_sin((float)((double)fx + PI_D/2));
}

/* libm cabsf as glibc >= 2.35 computes it for finite arguments */
float
lrpt_oracle_cabsf(float re, float im)
{
	return (float)sqrt((double)re*(double)re + (double)im*(double)im);
}

/* main.c:305 : MAX(-127, MIN(127, v/2)) in float, then C truncation to int8 */
int8_t
lrpt_oracle_quantise(float v)
{
	float h = v/2;
	float m = (127 < h) ? 127 : h;
	float r = (-127 > m) ? -127 : m;
	return (int8_t)r;
}

/* main.c:136 */
float
lrpt_oracle_freq_delta(float hz, float symrate)
{
	return (float)((double)hz * TWO_PI_D / (double)symrate);
}

/* ---------------------------------------------------------------- init -- */

/* timing.c:98-105 and pll.c:132-140 share this shape (all float) */
static void
loop_gains(float damp, float bw, float *alpha, float *beta)
{
	float denom = (1 + 2*damp*bw + bw*bw);
	*alpha = 4*damp*bw/denom;
	*beta = 4*bw*bw/denom;
}

int
lrpt_oracle_init(lrpt_oracle_t *o, float pll_bw, float sym_bw, int samplerate, int symrate,
                 int interp, int order, int oqpsk, float freq_max, int bps)
{
	const int multiplier = oqpsk ? 1 : 2;                                   /* demod.c:10 */
	float bw, sym_freq, tbw, osf;
	int i, j;

	memset(o, 0, sizeof(*o));
	if (samplerate <= 0 || symrate <= 0 || interp <= 0 || order < 0) return 1;
	if (bps != 8 && bps != 16 && bps != 32) return 1;
	o->pll_bw = pll_bw; o->sym_bw = sym_bw; o->freq_max = freq_max;
	o->samplerate = samplerate; o->symrate = symrate; o->interp = interp;
	o->order = order; o->oqpsk = oqpsk; o->bps = bps;

	/* pll_init, pll.c:25-44 (argument computed in double, demod.c:12) */
	bw = (float)(2*PI_D*(double)pll_bw/(double)(multiplier*symrate));
	if (freq_max < 0) freq_max = 0.3f;
	else freq_max = (1.0f < freq_max) ? 1.0f : freq_max;
	o->p_freq = 0; o->p_phase = 0; o->p_locked = o->p_locked_once = 0;
	o->p_err = 1000; o->p_bw = bw; o->p_updown = 1;
	o->p_fmax = oqpsk ? freq_max/2 : freq_max;
	for (i=0; i<32; i++) o->lut_tanh[i] = (float)tanh(i - 16);
	loop_gains(0.7071067811865475f, bw, &o->p_alpha, &o->p_beta);

	/* timing_init, timing.c:19-28 (arguments: demod.c:13) */
	sym_freq = (float)(2*PI_D*(double)symrate/(double)(samplerate*interp));
	tbw = sym_bw/interp;
	o->t_freq = sym_freq; o->t_center = sym_freq;
	o->t_maxdev = sym_freq/(1 << 12);
	o->t_phase = 0; o->t_prev = 0; o->t_dual_state = 1;
	loop_gains(1, tbw, &o->t_alpha, &o->t_beta);

	/* filter_init_rrc, filter.c:10-29 (arguments: demod.c:14) */
	osf = (float)samplerate/symrate;
	o->taps = 2*order + 1;
	o->h = malloc(sizeof(float)*(size_t)o->taps*interp);
	o->hist = calloc((size_t)(o->taps > 1 ? o->taps - 1 : 1)*2, sizeof(float));
	if (!o->h || !o->hist) { lrpt_oracle_free(o); return 1; }
	for (j=0; j<interp; j++)
		for (i=0; i<o->taps; i++)
			o->h[j*o->taps + i] = lrpt_oracle_rrc_coeff(i*interp + j, (unsigned)(o->taps*interp),
			                                            osf*(unsigned)interp, 0.6f);

	/* agc.c:9-10 */
	o->agc_gain = 1; o->agc_bias_re = 0; o->agc_bias_im = 0;
	o->oq_inphase = 0;
	o->first_lock_symbol = -1;
	return 0;
}

void
lrpt_oracle_free(lrpt_oracle_t *o)
{
	free(o->h); free(o->hist);
	o->h = NULL; o->hist = NULL;
}

/* ------------------------------------------------------ per-symbol ops -- */

/* filter_get, filter.c:46-65 : taps complex*real MACs, oldest sample first,
 * separate multiply and add (float). w points at the oldest sample (re,im). */
static void
fir_point(const float *w, const float *bank, int taps, float *re, float *im)
{
	float ar = 0, ai = 0;
	int k;
	for (k=0; k<taps; k++) {
		ar = ar + w[2*k]*bank[k];
		ai = ai + w[2*k+1]*bank[k];
	}
	*re = ar; *im = ai;
}

/* agc_apply, agc.c:13-25 */
static void
agc(lrpt_oracle_t *o, float *re, float *im)
{
	const float keep = 1 - 0.001f;
	float sr, si, mag, g;

	o->agc_bias_re = o->agc_bias_re*keep + 0.001f*(*re);
	o->agc_bias_im = o->agc_bias_im*keep + 0.001f*(*im);
	sr = (*re - o->agc_bias_re)*o->agc_gain;
	si = (*im - o->agc_bias_im)*o->agc_gain;
	mag = lrpt_oracle_cabsf(sr, si);
	g = o->agc_gain + 0.0001f*(190.0f - mag);
	o->agc_gain = (0 > g) ? 0 : g;
	*re = sr; *im = si;
}

/* NCO step shared by pll_mix / pll_mix_i / pll_mix_q, pll.c:61-62 */
static void
pll_advance(lrpt_oracle_t *o)
{
	o->p_phase = o->p_phase + o->p_freq;                       /* float */
	if ((double)o->p_phase >= TWO_PI_D)
		o->p_phase = (float)((double)o->p_phase - TWO_PI_D);   /* double, narrowed */
}

static float
lut(const lrpt_oracle_t *o, float v)      /* pll.c:154-159 */
{
	if (v > 15) return 1;
	if (v < -16) return -1;
	return o->lut_tanh[(int)v + 16];
}

/* pll_update_estimate, pll.c:100-130 */
static void
pll_update(lrpt_oracle_t *o, float i, float q)
{
	float error = lut(o, i)*q - lut(o, q)*i;                   /* pll.c:147-148 */
	float f;

	o->p_phase = (float)fmod((double)(o->p_phase + o->p_alpha*error), TWO_PI_D);
	o->p_freq = o->p_freq + o->p_beta*error;

	/* lock detector, pll.c:117-123: float product, double sum, narrowed */
	o->p_err = (float)((double)(o->p_err*(1 - 0.001f)) + fabs((double)error)*(double)0.001f);
	if (o->p_err < 85 && !o->p_locked) { o->p_locked = 1; o->p_locked_once = 1; }
	else if (o->p_err > 105 && o->p_locked) o->p_locked = 0;

	/* sweep, pll.c:126-128 */
	if (!o->p_locked) o->p_freq = (float)((double)o->p_freq + 0.000001*o->p_updown);
	o->p_updown = (o->p_freq >= o->p_fmax) ? -1 : (o->p_freq <= -o->p_fmax) ? 1 : o->p_updown;
	f = (o->p_fmax < o->p_freq) ? o->p_fmax : o->p_freq;
	o->p_freq = (-o->p_fmax > f) ? -o->p_fmax : f;
}

/* retime + mm_err + update_estimate, timing.c:60-95 (imaginary part only) */
static void
retime(lrpt_oracle_t *o, float cur)
{
	const float prev = o->t_prev;
	float err = (float)(prev < 0 ? -1 : 1)*cur - (float)(cur < 0 ? -1 : 1)*prev;
	float fd, m;

	o->t_prev = cur;
	fd = o->t_freq - o->t_center;
	o->t_phase = (float)((double)o->t_phase - (TWO_PI_D + (double)(o->t_alpha*err)));
	fd = fd - o->t_beta*err;
	m = (o->t_maxdev < fd) ? o->t_maxdev : fd;
	fd = (-o->t_maxdev > m) ? -o->t_maxdev : m;
	o->t_freq = o->t_center + fd;
}

/* filter_fwd_sample + filter_get (filter.c:39-65) for EVERY (sample, sub-step) of a block that starts from the
 * zeroed delay line of filter_init_rrc: out[(n*L + i)*2 + {0,1}] = re, im of filter_get(flt, i) after sample n was
 * pushed. The checker of the stand-alone FIR stage (csrc/fir_stage.cu); does not touch the oracle's state. */
int
lrpt_oracle_fir_all(const lrpt_oracle_t *o, const void *raw, long nsamples, float *out)
{
	const uint8_t *u8 = raw; const int16_t *s16 = raw; const float *f32 = raw;
	const int taps = o->taps, L = o->interp, H = taps - 1;
	float *work = calloc(2*(size_t)(H + nsamples), sizeof(float));
	long n;
	int i;
	if (!work) return 1;
	for (n=0; n<nsamples; n++) {
		float re, im;
		if (o->bps == 8)       { re = (float)((int)u8[2*n] - 128); im = (float)((int)u8[2*n+1] - 128); }
		else if (o->bps == 16) { re = (float)s16[2*n]; im = (float)s16[2*n+1]; }
		else                   { re = f32[2*n]; im = f32[2*n+1]; }
		work[2*(H+n)] = re; work[2*(H+n)+1] = im;
	}
	for (n=0; n<nsamples; n++)
		for (i=0; i<L; i++)
			fir_point(work + 2*n, o->h + (L-1-i)*taps, taps, out + 2*((size_t)n*L + i), out + 2*((size_t)n*L + i) + 1);
	free(work);
	return 0;
}

/* ------------------------------------------------------------- process -- */

#define BLOCK 32768

long
lrpt_oracle_process(lrpt_oracle_t *o, const void *raw, long nsamples,
                    float *sym, int8_t *soft, long long *sample_idx,
                    uint8_t *lock_once, long cap)
{
	const uint8_t *u8 = raw; const int16_t *s16 = raw; const float *f32 = raw;
	const int taps = o->taps, L = o->interp, H = taps - 1;
	const float two_pi_f = 2*(float)PI_D, pi_f = (float)PI_D;
	float *work = malloc(sizeof(float)*2*(size_t)(H + BLOCK));
	long done = 0, nsym = 0;

	if (!work) return -1;
	while (done < nsamples) {
		const long nb = (nsamples - done < BLOCK) ? nsamples - done : BLOCK;
		long n;

		/* window = history ++ block; ingest as wavfile.c:58-69 */
		memcpy(work, o->hist, sizeof(float)*2*(size_t)H);
		for (n=0; n<nb; n++) {
			const long g = done + n;
			float re, im;
			if (o->bps == 8)       { re = (float)((int)u8[2*g] - 128); im = (float)((int)u8[2*g+1] - 128); }
			else if (o->bps == 16) { re = (float)s16[2*g]; im = (float)s16[2*g+1]; }
			else                   { re = f32[2*g]; im = f32[2*g+1]; }
			work[2*(H+n)] = re; work[2*(H+n)+1] = im;
		}

		for (n=0; n<nb; n++) {
			const float *w = work + 2*n;          /* oldest sample of the window ending at n */
			int i;
			for (i=0; i<L; i++) {
				float re, im, out_re, out_im, s, c;
				int emit = 0;

				o->t_phase = o->t_phase + o->t_freq;               /* timing.c:34 / :48 */
				if (!o->oqpsk) {
					if (!(o->t_phase >= two_pi_f)) continue;       /* timing.c:37 */
					fir_point(w, o->h + (L-1-i)*taps, taps, &re, &im);   /* demod.c:35 */
					agc(o, &re, &im);
					s = lrpt_oracle_fast_sin(-o->p_phase);         /* pll.c:53-54 */
					c = lrpt_oracle_fast_cos(-o->p_phase);
					out_re = re*c - im*s;                          /* pll.c:60 */
					out_im = re*s + im*c;
					pll_advance(o);
					retime(o, out_im);                             /* demod.c:39 */
					pll_update(o, out_re, out_im);                 /* demod.c:40 */
					emit = 1;
				} else {
					const int st = o->t_dual_state;
					if (!(o->t_phase >= (float)st*pi_f)) continue; /* timing.c:51 */
					o->t_dual_state = (st % 2) + 1;
					fir_point(w, o->h + (L-1-i)*taps, taps, &re, &im);
					agc(o, &re, &im);
					s = lrpt_oracle_fast_sin(-o->p_phase);
					c = lrpt_oracle_fast_cos(-o->p_phase);
					if (st == 1) {                                 /* demod.c:66-71 */
						o->oq_inphase = re*c - im*s;
						pll_advance(o);
						continue;
					}
					out_im = re*s + im*c;                          /* demod.c:72-83 */
					pll_advance(o);
					out_re = o->oq_inphase;
					retime(o, out_im);
					pll_update(o, out_re, out_im);
					emit = 1;
				}
				if (emit) {
					if (o->p_locked_once && o->first_lock_symbol < 0)
						o->first_lock_symbol = o->nsymbols;
					if (nsym < cap) {
						if (sym) { sym[2*nsym] = out_re; sym[2*nsym+1] = out_im; }
						if (soft) {
							soft[2*nsym] = lrpt_oracle_quantise(out_re);
							soft[2*nsym+1] = lrpt_oracle_quantise(out_im);
						}
						if (sample_idx) sample_idx[nsym] = done + n;
						if (o->substep_out) o->substep_out[nsym] = (unsigned char)i;
						if (lock_once) lock_once[nsym] = (uint8_t)o->p_locked_once;
					}
					nsym++; o->nsymbols++;
				}
			}
		}
		/* keep the last taps-1 samples as history */
		memmove(o->hist, work + 2*nb, sizeof(float)*2*(size_t)H);
		done += nb;
	}
	o->nsamples += nsamples;
	free(work);
	return nsym;
}
