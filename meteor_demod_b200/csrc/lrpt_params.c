/*
 * lrpt_params.c -- host-side derivation of everything demod_init computes once
 * (reference: demod.c:8-15 -> pll.c:25-44, timing.c:19-28, filter.c:10-29,71-94).
 *
 * Plain C on purpose: the tap generator and loop gains mix float and double
 * sub-expressions and call libm (sinf, cosf, tanh); evaluating them on the host
 * with the reference's operand widths and the same libm is what makes the tap
 * banks and gains bit-identical to the reference's. Build with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "lrpt_internal.h"

static const double kPi = 3.14159265358979323846;

/* One prototype-filter tap; n-th of ntaps at `osf` samples per symbol.
 * filter.c:71-94 (windowed RRC; the window is Blackman despite the comment). */
static float
rrc_tap(int n, unsigned ntaps, float osf, float alpha)
{
	const float norm = (float)(2.0/5.0);
	const int mid = (int)((ntaps - 1)/2);
	float t, num, den, a4t;
	double window;

	if (n == mid)   /* 0/0 limit, filter.c:82-84 (double expression, narrowed on return) */
		return (float)((double)norm*((double)(1 - alpha) + (double)(4*alpha)/kPi));

	t = (float)abs(mid - n)/osf;
	a4t = 4*alpha*t;
	num = sinf((float)(kPi*(double)t*(double)(1 - alpha)))
	    + a4t*cosf((float)(kPi*(double)t*(double)(1 + alpha)));
	den = (float)(kPi*(double)t*(double)(1 - a4t*a4t));

	window = 0.42 - 0.5*(double)cosf((float)(2*kPi*n/(double)(ntaps - 1)))
	       + 0.08*(double)cosf((float)(4*kPi*n/(double)(ntaps - 1)));
	num = (float)((double)num*window);

	return num/den*norm;
}

/* 2nd-order loop gains, identical shape in timing.c:98-105 and pll.c:132-140 */
static void
gains(float damp, float bw, float *alpha, float *beta)
{
	const float denom = (1 + 2*damp*bw + bw*bw);
	*alpha = 4*damp*bw/denom;
	*beta = 4*bw*bw/denom;
}

int
lrpt_derive(const lrpt_params_t *p, lrpt_consts_t *c, lrpt_state_t *s0, float *h)
{
	const int mult = p->oqpsk ? 1 : 2;                    /* demod.c:10 */
	const int taps = 2*p->rrc_order + 1;                  /* filter.c:12 */
	float bw, fmax, nco, osf;
	int i, j;

	if (p->samplerate <= 0 || p->symrate <= 0) return LRPT_ERR_ARG;
	if (p->interp_factor < 1 || p->interp_factor > LRPT_MAX_INTERP) return LRPT_ERR_ARG;
	if (p->rrc_order < 0 || p->rrc_order > LRPT_MAX_ORDER) return LRPT_ERR_ARG;
	if (p->bps != 8 && p->bps != 16 && p->bps != 32) return LRPT_ERR_ARG;

	memset(c, 0, sizeof(*c));
	memset(s0, 0, sizeof(*s0));
	c->taps = taps; c->interp = p->interp_factor; c->oqpsk = !!p->oqpsk; c->bps = p->bps;

	/* Costas loop, pll.c:25-44; bandwidth argument evaluated in double (demod.c:12) */
	bw = (float)(2*kPi*(double)p->pll_bw/(double)(mult*p->symrate));
	fmax = p->freq_max;
	if (fmax < 0) fmax = 0.3f; else fmax = (1.0f < fmax) ? 1.0f : fmax;
	c->p_fmax = p->oqpsk ? fmax/2 : fmax;
	for (i=0; i<32; i++) c->lut_tanh[i] = (float)tanh(i - 16);
	gains(0.7071067811865475f, bw, &c->p_alpha, &c->p_beta);

	/* symbol clock NCO, timing.c:19-28; step evaluated in double (demod.c:13) */
	nco = (float)(2*kPi*(double)p->symrate/(double)(p->samplerate*p->interp_factor));
	c->t_center = nco;
	c->t_maxdev = nco/(1 << 12);
	gains(1, p->sym_bw/p->interp_factor, &c->t_alpha, &c->t_beta);

	/* polyphase banks, filter.c:18-22: bank j holds prototype taps j, j+L, j+2L, ... */
	osf = (float)p->samplerate/p->symrate;
	for (j=0; j<p->interp_factor; j++)
		for (i=0; i<taps; i++)
			h[j*taps + i] = rrc_tap(i*p->interp_factor + j, (unsigned)(taps*p->interp_factor),
			                        osf*(unsigned)p->interp_factor, 0.6f);

	/* power-on state: the reference's static initialisers + pll_init/timing_init */
	s0->magic = LRPT_STATE_MAGIC; s0->taps = (uint32_t)taps;
	s0->t_phase = 0; s0->t_freq = nco; s0->t_prev = 0; s0->t_dual_state = 1;   /* timing.c:13,21,43 */
	s0->oq_inphase = 0;                                                          /* demod.c:54 */
	s0->agc_gain = 1; s0->agc_bias_re = 0; s0->agc_bias_im = 0;                  /* agc.c:9-10 */
	s0->p_phase = 0; s0->p_freq = 0; s0->p_err = 1000;                           /* pll.c:33-36 */
	s0->p_locked = 0; s0->p_locked_once = 0; s0->p_updown = 1;                   /* pll.c:35,112 */
	s0->nsamples = 0; s0->nsymbols = 0; s0->first_lock_symbol = -1;
	return LRPT_OK;
}

/* main.c:136 */
float
lrpt_freq_delta_from_hz(float hz, float symrate)
{
	return (float)((double)hz*(2*kPi)/(double)symrate);
}
