/* kernels.h -- launch interface between the C-ABI layer (lrpt_api.cu) and the kernels. */
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "lrpt_internal.h"

namespace lrpt {

struct LaunchArgs {
	const lrpt_consts_t *c;      /* host pointer; passed to the kernel by value */
	const float  *d_taps;        /* [interp][taps] as filter.c:18-22 lays them out */
	lrpt_state_t *d_states;      /* [all streams of the handle] */
	float2       *d_hist;        /* [all streams][taps-1], oldest first */
	const void   *d_raw;         /* stream `first_stream + s` reads at d_raw + s*raw_stride */
	size_t        raw_stride;
	size_t        nsamples;
	int8_t       *d_soft;        /* + s*soft_stride, 2 int8 per symbol */
	size_t        soft_stride;
	float        *d_symf;        /* optional float symbols, + s*symf_stride bytes */
	size_t        symf_stride;
	uint32_t     *d_symq;        /* optional: timing sub-step index (sample*interp + phase, counted from the
	                                start of this call) of every symbol, + s*symq_stride bytes */
	size_t        symq_stride;
	unsigned      cap;           /* symbols per stream that fit */
	uint32_t     *d_nsym;        /* [nstreams] symbols produced by this launch (may be NULL) */
	uint32_t     *d_out_off;     /* [nstreams] append cursor: symbols already in d_soft; advanced by the
	                                kernel (may be NULL = write from 0) */
	int           first_stream;
	int           nstreams;
};

cudaError_t launch_simple(const LaunchArgs &a, cudaStream_t st);

/* warp-specialised kernel (demod_ws.cu) */
bool        ws_supported(const lrpt_consts_t &c);
cudaError_t ws_prepare(int device);                       /* one-time attribute setup */
cudaError_t launch_ws(const LaunchArgs &a, cudaStream_t st, int *launches);

/* speculative-FIR kernel (demod_spec.cu) */
bool        spec_supported(const lrpt_consts_t &c);
cudaError_t spec_prepare(int device);
cudaError_t launch_spec(const LaunchArgs &a, cudaStream_t st, int *launches, unsigned long long *d_fallbacks);

/* lane-per-stream lazy-FIR kernel (demod_lane.cu) */
bool        lane_supported(const lrpt_consts_t &c);
cudaError_t lane_prepare(int device);
cudaError_t launch_lane(const LaunchArgs &a, cudaStream_t st, int *launches);

} // namespace lrpt
