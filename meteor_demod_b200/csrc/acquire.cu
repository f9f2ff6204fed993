/*
 * acquire.cu -- coarse carrier estimate of a batch of rows on the device (the acquisition accelerator
 * of the time-sharded path; meteor_demod_b200/acquire.py is the same estimator in torch ops and its test model).
 *
 * The reference finds the carrier by sweeping its Costas NCO at 1e-6 rad/symbol^2 until the lock detector fires
 * (pll.c:117-128). A time-sharded chunk that starts cold would have to repeat that inside its warm-up; instead every
 * chunk but the first starts with the NCO AT the carrier. Estimator (not in the reference; it only replaces the wait):
 * remove the DC term, strip the modulation by a power law -- x^4 has a line at 4*f_c for QPSK; x^2 has two lines at
 * 2*f_c -+ symrate for OQPSK, whose I and Q pulse trains are half a symbol apart -- Hann window, FFT, strongest
 * candidate within +-fmax, three-point parabola through the log magnitudes around it.
 *
 * One CTA per row: the row's first n samples (n a power of two, 256..16384) are converted, windowed and written to
 * shared memory in bit-reversed order, log2(n) radix-2 stages run in place (twiddles by sincospif), the candidate
 * bins are scanned by the whole CTA and thread 0 refines the peak. 8*n bytes of shared memory, n/2*log2(n)
 * butterflies: at n = 4096 and 32768 rows about a millisecond -- plumbing next to the passes it shortens.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "lrpt_b200.h"

namespace {

constexpr int AQ_THREADS = 256;

struct AqArgs {
	const uint8_t *raw; size_t raw_stride;
	int nrows, n, bits, bps, oqpsk;
	int kmax, shift, order;      /* candidates -kmax..kmax (bins of order*f_c), OQPSK line offset in bins */
	double df;                   /* bin spacing fs/n */
	double *out;                 /* [nrows] carrier offset in Hz */
};

/* wavfile.c:58-69: the float pair wav_read produces for sample i of the row */
__device__ __forceinline__ float2 aq_load(const uint8_t *row, int i, int bps)
{
	if (bps == 16) { const short2 v = reinterpret_cast<const short2 *>(row)[i]; return make_float2((float)v.x, (float)v.y); }
	if (bps == 8)  { const uchar2 v = reinterpret_cast<const uchar2 *>(row)[i]; return make_float2((float)v.x - 128.0f, (float)v.y - 128.0f); }
	return reinterpret_cast<const float2 *>(row)[i];
}

__device__ __forceinline__ float aq_mag(const float2 v) { return sqrtf(v.x*v.x + v.y*v.y); }

__device__ __forceinline__ int aq_mod(int k, int n) { k %= n; return k < 0 ? k + n : k; }

__device__ __forceinline__ float aq_score(const float2 *S, const AqArgs &a, int c)
{
	const int k = c - a.kmax;
	if (a.oqpsk) return aq_mag(S[aq_mod(k - a.shift, a.n)]) + aq_mag(S[aq_mod(k + a.shift, a.n)]);
	return aq_mag(S[aq_mod(k, a.n)]);
}

__global__ void __launch_bounds__(AQ_THREADS)
carrier_estimate_kernel(const AqArgs a)
{
	extern __shared__ float2 S[];                                    /* [n] */
	__shared__ float red_a[AQ_THREADS/32], red_b[AQ_THREADS/32];
	__shared__ int red_i[AQ_THREADS/32];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = a.n;
	const uint8_t *row = a.raw + (size_t)blockIdx.x*a.raw_stride;

	/* 1. DC term */
	float sr = 0.0f, si = 0.0f;
	for (int i = tid; i < n; i += AQ_THREADS) { const float2 v = aq_load(row, i, a.bps); sr += v.x; si += v.y; }
	for (int d = 16; d > 0; d >>= 1) { sr += __shfl_xor_sync(0xffffffffu, sr, d); si += __shfl_xor_sync(0xffffffffu, si, d); }
	if (lane == 0) { red_a[warp] = sr; red_b[warp] = si; }
	__syncthreads();
	sr = 0.0f; si = 0.0f;
	for (int w = 0; w < AQ_THREADS/32; w++) { sr += red_a[w]; si += red_b[w]; }
	const float mr = sr/(float)n, mi = si/(float)n;
	__syncthreads();

	/* 2. power law, Hann window (torch.hann_window(n, periodic=False)), bit-reversed placement */
	for (int i = tid; i < n; i += AQ_THREADS) {
		float2 v = aq_load(row, i, a.bps);
		v.x -= mr; v.y -= mi;
		float2 y = make_float2(v.x*v.x - v.y*v.y, 2.0f*v.x*v.y);
		if (!a.oqpsk) y = make_float2(y.x*y.x - y.y*y.y, 2.0f*y.x*y.y);
		const float w = 0.5f - 0.5f*cospif(2.0f*(float)i/(float)(n - 1));
		S[__brev((unsigned)i) >> (32 - a.bits)] = make_float2(y.x*w, y.y*w);
	}
	__syncthreads();

	/* 3. radix-2 decimation in time, in place */
	for (int s = 0; s < a.bits; s++) {
		const int half = 1 << s;
		for (int k = tid; k < n/2; k += AQ_THREADS) {
			const int j = k & (half - 1);
			const int i0 = ((k >> s) << (s + 1)) + j, i1 = i0 + half;
			float sn, cs;
			sincospif(-(float)j/(float)half, &sn, &cs);
			const float2 u = S[i0], v = S[i1];
			const float2 t = make_float2(cs*v.x - sn*v.y, cs*v.y + sn*v.x);
			S[i0] = make_float2(u.x + t.x, u.y + t.y);
			S[i1] = make_float2(u.x - t.x, u.y - t.y);
		}
		__syncthreads();
	}

	/* 4. strongest candidate, lowest index on ties */
	const int ncand = 2*a.kmax + 1;
	float best = -1.0f; int bi = 0;
	for (int c = tid; c < ncand; c += AQ_THREADS) {
		const float v = aq_score(S, a, c);
		if (v > best) { best = v; bi = c; }
	}
	for (int d = 16; d > 0; d >>= 1) {
		const float ob = __shfl_xor_sync(0xffffffffu, best, d);
		const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
		if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
	}
	if (lane == 0) { red_a[warp] = best; red_i[warp] = bi; }
	__syncthreads();
	if (tid == 0) {
		for (int w = 1; w < AQ_THREADS/32; w++)
			if (red_a[w] > best || (red_a[w] == best && red_i[w] < bi)) { best = red_a[w]; bi = red_i[w]; }
		/* 5. parabola through the log magnitudes of the peak and its neighbours */
		double frac = 0.0;
		if (bi >= 1 && bi <= ncand - 2) {
			const double la = log((double)aq_score(S, a, bi - 1) + 1e-30), lb = log((double)best + 1e-30),
			             lc = log((double)aq_score(S, a, bi + 1) + 1e-30);
			const double den = la - 2.0*lb + lc;
			if (fabs(den) > 1e-12) frac = fmin(0.5, fmax(-0.5, 0.5*(la - lc)/den));
		}
		a.out[blockIdx.x] = ((double)(bi - a.kmax) + frac)*a.df/(double)a.order;
	}
}

} // namespace

extern "C" int lrpt_carrier_estimate_device(const lrpt_params_t *p, const void *d_raw, size_t raw_stride, int nrows, int nfft,
                                            double fmax_hz, double *d_cfo_hz, void *cuda_stream)
{
	if (!p || !d_raw || !d_cfo_hz || nrows < 1) return LRPT_ERR_ARG;
	if (p->bps != 8 && p->bps != 16 && p->bps != 32) return LRPT_ERR_ARG;
	int bits = 0;
	while ((1 << bits) < nfft) bits++;
	if (nfft < 256 || nfft > 16384 || (1 << bits) != nfft) return LRPT_ERR_ARG;
	const size_t pair = (size_t)p->bps/4;
	if (((uintptr_t)d_raw | raw_stride) & (pair - 1)) return LRPT_ERR_ARG;
	if (!(fmax_hz > 0.0) || p->samplerate <= 0 || p->symrate <= 0) return LRPT_ERR_ARG;
	AqArgs a;
	a.raw = static_cast<const uint8_t *>(d_raw); a.raw_stride = raw_stride;
	a.nrows = nrows; a.n = nfft; a.bits = bits; a.bps = p->bps; a.oqpsk = p->oqpsk ? 1 : 0;
	a.order = a.oqpsk ? 2 : 4;
	a.df = (double)p->samplerate/(double)nfft;
	a.kmax = (int)((double)a.order*fmax_hz/a.df);
	a.shift = (int)llround((double)p->symrate/a.df);
	if (a.kmax < 1 || 2*a.kmax + 1 > nfft) return LRPT_ERR_ARG;      /* candidates must not wrap onto each other */
	a.out = d_cfo_hz;
	if (cudaSetDevice(p->device) != cudaSuccess) return LRPT_ERR_CUDA;
	const size_t smem = (size_t)nfft*sizeof(float2);
	if (cudaFuncSetAttribute(carrier_estimate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return LRPT_ERR_CUDA;
	carrier_estimate_kernel<<<(unsigned)nrows, AQ_THREADS, smem, (cudaStream_t)cuda_stream>>>(a);
	return cudaGetLastError() == cudaSuccess ? LRPT_OK : LRPT_ERR_CUDA;
}
