/*
 * lrpt_internal.h -- shared between the C host code (lrpt_params.c) and the CUDA
 * translation units. Not part of the public ABI.
 */
#ifndef LRPT_INTERNAL_H
#define LRPT_INTERNAL_H

#include <stdint.h>
#include "lrpt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define LRPT_STATE_MAGIC 0x5350524cu   /* 'LRPS' little endian */
#define LRPT_MAX_INTERP 16
#define LRPT_MAX_ORDER  512

/* Loop constants derived once on the host exactly as demod_init does (demod.c:8-15);
 * passed to kernels by value (lives in the constant bank). */
typedef struct lrpt_consts {
	int32_t taps, interp, oqpsk, bps;
	float   t_center, t_maxdev, t_alpha, t_beta;   /* timing.c:21-27 */
	float   p_alpha, p_beta, p_fmax;               /* pll.c:38,43    */
	float   lut_tanh[32];                          /* pll.c:40-42    */
} lrpt_consts_t;

/* Fills `c`, the initial per-stream state `s0` and the tap banks `h`
 * (taps*interp floats, caller allocated, bank j at h[j*taps]). 0 on success. */
int lrpt_derive(const lrpt_params_t *p, lrpt_consts_t *c, lrpt_state_t *s0, float *h);

#ifdef __cplusplus
}
#endif
#endif
