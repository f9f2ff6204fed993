/*
 * demod_simple.cu -- one thread per stream, everything in source order.
 *
 * The plainest possible exact execution of the reference loop
 * (main.c:303-306 -> demod.c:24-91): per input sample push, L timing sub-steps,
 * lazy single-phase FIR at a crossing (filter.c:46-65), AGC, PLL, retime.
 * It is the fallback for configurations the warp-specialised kernel does not
 * cover and the in-tree cross-check for it; it is not the fast path.
 */
#include "demod_core.cuh"
#include "kernels.h"

namespace lrpt {

__global__ void __launch_bounds__(64)
demod_simple_kernel(const lrpt_consts_t c, const float *__restrict__ h, lrpt_state_t *states,
                    float2 *hist, const uint8_t *__restrict__ raw, size_t raw_stride, long long nsamples,
                    int8_t *soft, size_t soft_stride, float *symf, size_t symf_stride,
                    uint32_t *symq, size_t symq_stride, unsigned cap, uint32_t *nsym_out, uint32_t *out_off, int first_stream, int nstreams)
{
	__shared__ float lut[32];
	if (threadIdx.x < 32) lut[threadIdx.x] = c.lut_tanh[threadIdx.x];
	__syncthreads();

	const int local = blockIdx.x*blockDim.x + threadIdx.x;
	if (local >= nstreams) return;
	const int sid = first_stream + local;
	const int taps = c.taps, L = c.interp, H = taps - 1;
	const void *x = raw + (size_t)local*raw_stride;
	float2 *hs = hist + (size_t)sid*H;
	char2 *out = reinterpret_cast<char2 *>(soft + (size_t)local*soft_stride);
	float2 *outf = symf ? reinterpret_cast<float2 *>(reinterpret_cast<char *>(symf) + (size_t)local*symf_stride) : nullptr;
	uint32_t *outq = symq ? reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(symq) + (size_t)local*symq_stride) : nullptr;

	Loop r;
	loop_load(r, states[sid]);
	long long nsymbols = states[sid].nsymbols;
	long long first_lock = states[sid].first_lock_symbol;
	const unsigned off = out_off ? out_off[local] : 0u;
	unsigned nsym = 0;

	for (long long n = 0; n < nsamples; n++) {
		for (int i = 0; i < L; i++) {
			int half = 0;
			r.t_phase = __fadd_rn(r.t_phase, r.t_freq);                 /* timing.c:34 / :48 */
			if (!c.oqpsk) {
				if (!(r.t_phase >= kTwoPiF)) continue;                  /* timing.c:37 */
			} else {
				if (!(r.t_phase >= __fmul_rn((float)r.t_dual, kPiF))) continue;   /* timing.c:51 */
				half = r.t_dual;
				r.t_dual = (r.t_dual % 2) + 1;
			}
			/* filter_get(flt, i): bank L-1-i, oldest sample first, mul then add */
			const float *bank = h + (size_t)(L - 1 - i)*taps;
			float ar = 0.0f, ai = 0.0f;
			for (int k = 0; k < taps; k++) {
				const long long idx = n - H + k;
				const float2 v = (idx < 0) ? hs[H + idx] : ingest(x, c.bps, idx);
				const float hk = bank[k];
				ar = __fadd_rn(ar, __fmul_rn(v.x, hk));
				ai = __fadd_rn(ai, __fmul_rn(v.y, hk));
			}
			float ore, oim;
			if (symbol_event(r, c, lut, half, ar, ai, ore, oim)) {
				if (r.locked_once && first_lock < 0) first_lock = nsymbols;
				if (off + nsym < cap) {
					out[off + nsym] = make_char2((signed char)quantise(ore), (signed char)quantise(oim));
					if (outf) outf[off + nsym] = make_float2(ore, oim);
					if (outq) outq[off + nsym] = (uint32_t)(n*L + i);
				}
				nsym++; nsymbols++;
			}
		}
	}

	/* delay line for the next call: the last taps-1 samples, oldest first */
	for (int j = 0; j < H; j++) {
		const long long idx = nsamples - H + j;
		hs[j] = (idx < 0) ? hs[H + idx] : ingest(x, c.bps, idx);
	}
	loop_store(r, states[sid]);
	states[sid].nsamples += nsamples;
	states[sid].nsymbols = nsymbols;
	states[sid].first_lock_symbol = first_lock;
	if (nsym_out) nsym_out[local] = nsym;
	if (out_off) out_off[local] = off + nsym;
}

cudaError_t launch_simple(const LaunchArgs &a, cudaStream_t st)
{
	const int threads = 32;
	const int blocks = (a.nstreams + threads - 1)/threads;
	demod_simple_kernel<<<blocks, threads, 0, st>>>(*a.c, a.d_taps, a.d_states, a.d_hist,
		reinterpret_cast<const uint8_t *>(a.d_raw), a.raw_stride, (long long)a.nsamples,
		a.d_soft, a.soft_stride, a.d_symf, a.symf_stride, a.d_symq, a.symq_stride, a.cap, a.d_nsym, a.d_out_off, a.first_stream, a.nstreams);
	return cudaGetLastError();
}

} // namespace lrpt
