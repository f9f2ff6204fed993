/*
 * oracle/lrpt_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C99, re-entrant, block based) of the per-sample LRPT
 * demodulator hot path of dbdexter-dev/meteor_demod. It is the checker the
 * CUDA path is compared with; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may use it. It is pinned against the compiled
 * reference itself (oracle/_ref/libref_strict.so, see oracle/Makefile and
 * tests/test_oracle_vs_ref.py) and against committed golden vectors
 * (tests/golden/).
 *
 * Floating-point meaning: the reference sources evaluated in source order with
 * IEEE-754 round-to-nearest and NO fused multiply-add contraction ("ORACLE-STRICT",
 * SURVEY.md section 8c). Build with -ffp-contract=off.
 */
#ifndef LRPT_ORACLE_H
#define LRPT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

typedef struct {
	/* configuration (demod.h:29 arguments + wavfile.c bps) */
	float pll_bw, sym_bw, freq_max;
	int samplerate, symrate, interp, order, oqpsk, bps;

	/* derived constants */
	int taps;                 /* 2*order+1                          filter.c:12        */
	float *h;                 /* taps*interp coefficients, bank j at h[j*taps..]  filter.c:18-22 */
	float lut_tanh[32];       /* pll.c:40-42 */
	float t_center, t_maxdev, t_alpha, t_beta;     /* timing.c:21-27 */
	float p_alpha, p_beta, p_fmax, p_bw;           /* pll.c:37-43    */

	/* running state */
	float t_phase, t_freq, t_prev;                 /* timing.c:13-14 */
	int t_dual_state;                              /* timing.c:43    */
	float oq_inphase;                              /* demod.c:54     */
	float agc_gain, agc_bias_re, agc_bias_im;      /* agc.c:9-10     */
	float p_phase, p_freq, p_err;                  /* pll.c:16,19    */
	int p_locked, p_locked_once, p_updown;         /* pll.c:20, :112 */
	float *hist;              /* last taps-1 input samples (re,im), oldest first: filter.h:6 in linear order */

	/* bookkeeping (main.c:291,298,312) */
	long long nsamples, nsymbols, first_lock_symbol;

	/* optional per-call side output: timing sub-step i (demod.c:33) of every stored symbol */
	unsigned char *substep_out;
} lrpt_oracle_t;

/* mirrors demod_init (demod.c:8-15). Returns 0, or 1 on allocation failure / bad arguments. */
int  lrpt_oracle_init(lrpt_oracle_t *o, float pll_bw, float sym_bw, int samplerate, int symrate,
                      int interp, int order, int oqpsk, float freq_max, int bps);
void lrpt_oracle_free(lrpt_oracle_t *o);

/*
 * Push nsamples raw interleaved I/Q samples (u8 / s16 / f32 as wavfile.c:58-69).
 * Outputs (any may be NULL), one entry per symbol, at most cap entries stored:
 *   sym        2 floats (re, im)                          demod.c:42 / :76
 *   soft       2 int8, quantised as main.c:305-306
 *   sample_idx index (within this call) of the input sample that produced the symbol
 *   lock_once  pll_did_lock_once() right after the symbol   main.c:312
 * Returns the number of symbols produced by this call.
 */
long lrpt_oracle_process(lrpt_oracle_t *o, const void *raw, long nsamples,
                         float *sym, int8_t *soft, long long *sample_idx,
                         uint8_t *lock_once, long cap);

/* filter_get (filter.c:46-65) at every (sample, sub-step) of a block from a zeroed delay line:
 * out[(n*interp + i)*2 + {0,1}]; checker of the stand-alone FIR stage. 0 on success. */
int  lrpt_oracle_fir_all(const lrpt_oracle_t *o, const void *raw, long nsamples, float *out);

/* building blocks, exported for known-answer tests */
float   lrpt_oracle_rrc_coeff(int stage_no, unsigned taps, float osf, float alpha);  /* filter.c:71-94 */
float   lrpt_oracle_fast_sin(float x);                                                /* sincos.c:13-34 */
float   lrpt_oracle_fast_cos(float x);                                                /* sincos.c:37-40 */
float   lrpt_oracle_cabsf(float re, float im);                                        /* libm cabsf model */
int8_t  lrpt_oracle_quantise(float v);                                                /* main.c:305     */
/* main.c:136 : -d <Hz> to rad/symbol (negative stays negative => default FREQ_MAX, pll.c:30) */
float   lrpt_oracle_freq_delta(float freq_max_delta_hz, float symrate);

#endif
