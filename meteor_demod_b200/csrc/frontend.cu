/*
 * frontend.cu -- the LRPT decoder FRONT-END on the device: frame synchronisation and Viterbi decoding of the
 * soft-symbol stream the demodulator writes (main.c:305-313: int8 I, int8 Q per symbol; the step after this
 * path in the reference's pipeline, README.md:6-9,87-91: `meteor_demod ... | meteor_decode`). SURVEY.md 8(f1).
 *
 * The algorithm is the published one of the link layer (CCSDS 131.0-B as used by Meteor-M LRPT: ASM 0x1ACFFC1D,
 * rate 1/2 K = 7 code with G1 = 171 on I and G2 = 133 on Q, 8192 symbols per 1024-byte CADU), restated in
 * oracle/frontend_oracle.c, whose functions these kernels reproduce bit for bit:
 *
 *   fe_pack_kernel      hard decisions (soft >= 0) of 16 symbols -> one 32-bit word           (2 B in, 0.25 B out / symbol)
 *   fe_score_kernel     per symbol offset: 64-bit window, XOR + popcount against the encoded ASM under the 8
 *                       symmetries of the constellation (4 quarter turns x I/Q swap) -> best score, symmetry
 *   fe_peak_kernel      first maximum per window of one CADU
 *   fe_viterbi_kernel   one warp per CADU, persistent over the frame list: lane l holds the path metrics of
 *                       states l and l + 32; the butterfly's operands come from lanes l>>1 and 16 + (l>>1) by
 *                       shuffles, branch metrics are +-I +-Q with per-lane signs fixed at start, decisions leave as
 *                       two ballots per step into an L2-resident scratch ring, traceback walks it backwards 32 steps
 *                       at a time (one coalesced load per batch, the words handed round by shuffles).
 *
 * Byte / integer work, HBM-bound for the synchroniser (2 bytes read and 2 written per symbol) and latency-bound
 * per frame for the decoder (8320 dependent add-compare-select steps), which is why it runs one warp per frame
 * and as many frames as fit side by side. No tensor cores: there is no GEMM here.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "lrpt_b200.h"

namespace {

constexpr int FE_G1 = 0x4F, FE_G2 = 0x6D;
constexpr int FE_CADU = 1024, FE_CADU_SYMS = 8192, FE_HEAD = 64, FE_TAIL = 64;
constexpr int FE_STEPS = FE_HEAD + FE_CADU_SYMS + FE_TAIL;

__constant__ unsigned long long c_pat[8];

__host__ __device__ inline int parity7(unsigned x) { x ^= x >> 4; x ^= x >> 2; x ^= x >> 1; return (int)(x & 1u); }

unsigned long long sync_pattern(int h)
{
	const unsigned asm_word = 0x1ACFFC1Du;
	unsigned reg = 0;
	unsigned long long v = 0;
	for (int i = 0; i < 32; i++) {
		reg = ((reg << 1) | ((asm_word >> (31 - i)) & 1u)) & 0x7Fu;
		unsigned bi = (unsigned)parity7(reg & FE_G1), bq = (unsigned)parity7(reg & FE_G2), t;
		if (h & 4) { t = bi; bi = bq; bq = t; }
		for (int k = 0; k < (h & 3); k++) { t = bi; bi = bq ^ 1u; bq = t; }
		v = (v << 2) | (unsigned long long)(bi << 1) | bq;
	}
	return v;
}

/* words[j] = hard bits of symbols 16j .. 16j+15, symbol 16j's I in bit 31 (symbols beyond nsym read as -1: bit 0) */
__global__ void fe_pack_kernel(const int8_t *__restrict__ soft, size_t nsym, uint32_t *__restrict__ words, size_t nwords)
{
	const size_t j = (size_t)blockIdx.x*blockDim.x + threadIdx.x;
	if (j >= nwords) return;
	uint32_t w = 0;
	const size_t s0 = 16*j;
	if (s0 + 16 <= nsym) {
		const uint4 *p = reinterpret_cast<const uint4 *>(soft + 2*s0);   /* 32 bytes, 16-byte aligned */
		const uint4 a = __ldg(p), b = __ldg(p + 1);
		const uint32_t v[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
#pragma unroll
		for (int k = 0; k < 8; k++) {
			/* bytes little endian: I0 Q0 I1 Q1; sign bit clear <=> value >= 0 <=> hard bit 1 */
			const uint32_t nonneg = ~v[k] & 0x80808080u;
			const uint32_t four = ((nonneg >> 7) & 1u) << 3 | ((nonneg >> 15) & 1u) << 2 | ((nonneg >> 23) & 1u) << 1 | ((nonneg >> 31) & 1u);
			w = (w << 4) | four;
		}
	} else {
		for (int k = 0; k < 16; k++) {
			const size_t s = s0 + k;
			const uint32_t bi = (s < nsym && soft[2*s] >= 0) ? 1u : 0u, bq = (s < nsym && soft[2*s + 1] >= 0) ? 1u : 0u;
			w = (w << 2) | (bi << 1) | bq;
		}
	}
	words[j] = w;
}

/* four offsets per thread: score/hyp bytes leave as one 32-bit store each */
__global__ void fe_score_kernel(const uint32_t *__restrict__ words, size_t nsym, uint32_t *__restrict__ score4, uint32_t *__restrict__ hyp4)
{
	const size_t q = (size_t)blockIdx.x*blockDim.x + threadIdx.x;   /* offsets 4q .. 4q+3 */
	if (4*q >= nsym) return;
	const size_t wi = (4*q) >> 4;
	const uint32_t w0 = words[wi], w1 = words[wi + 1], w2 = words[wi + 2];   /* the array is padded by two words */
	uint32_t sc = 0, hy = 0;
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const size_t o = 4*q + k;
		const unsigned sh = 2u*(unsigned)(o & 15);
		const unsigned long long win = ((unsigned long long)__funnelshift_l(w1, w0, sh) << 32) | __funnelshift_l(w2, w1, sh);
		int best = -1, bh = 0;
#pragma unroll
		for (int h = 0; h < 8; h++) {
			const int s = 64 - __popcll(win ^ c_pat[h]);
			if (s > best) { best = s; bh = h; }
		}
		if (o + 32 > nsym) { best = 0; bh = 0; }                    /* fewer than 32 symbols left: no window */
		sc |= (uint32_t)best << (8*k); hy |= (uint32_t)bh << (8*k);
	}
	score4[q] = sc; hyp4[q] = hy;
}

/* one block per window: first offset of the maximum score */
__global__ void fe_peak_kernel(const uint8_t *__restrict__ score, const uint8_t *__restrict__ hyp, size_t nsym, unsigned window,
                               uint32_t *__restrict__ off, uint8_t *__restrict__ ohyp, uint8_t *__restrict__ oscore)
{
	__shared__ unsigned long long best[32];
	const size_t lo = (size_t)blockIdx.x*window;
	const size_t hi = lo + window < nsym ? lo + window : nsym;
	/* key = score << 32 | (0xffffffff - offset): the maximum key is the highest score at the LOWEST offset */
	unsigned long long k = 0;
	for (size_t o = lo + threadIdx.x; o < hi; o += blockDim.x) {
		const unsigned long long c = ((unsigned long long)score[o] << 32) | (0xffffffffu - (uint32_t)(o - lo));
		k = c > k ? c : k;
	}
	for (int d = 16; d > 0; d >>= 1) { const unsigned long long c = __shfl_xor_sync(0xffffffffu, k, d); k = c > k ? c : k; }
	if ((threadIdx.x & 31) == 0) best[threadIdx.x >> 5] = k;
	__syncthreads();
	if (threadIdx.x == 0) {
		for (unsigned w = 1; w < blockDim.x/32; w++) k = best[w] > k ? best[w] : k;
		const size_t o = lo + (0xffffffffu - (uint32_t)(k & 0xffffffffu));
		off[blockIdx.x] = (uint32_t)o; ohyp[blockIdx.x] = hyp[o]; oscore[blockIdx.x] = score[o];
	}
}

/* ------------------------------------------------------------------ Viterbi ---- */

__device__ __forceinline__ void unturn(int h, int i, int q, int &oi, int &oq)
{
	for (int k = 0; k < (h & 3); k++) { const int t = i; i = q; q = -t; }
	if (h & 4) { const int t = i; i = q; q = t; }
	oi = i; oq = q;
}

__global__ void __launch_bounds__(128)
fe_viterbi_kernel(const int8_t *__restrict__ soft, size_t nsym, const uint32_t *__restrict__ frame_off, const uint8_t *__restrict__ frame_hyp,
                  int nframes, uint8_t *__restrict__ cadu, int32_t *__restrict__ metric, uint2 *__restrict__ scratch)
{
	const int lane = threadIdx.x & 31;
	const int warp = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
	const int nwarps = (gridDim.x*blockDim.x) >> 5;
	uint2 *dec = scratch + (size_t)warp*FE_STEPS;                       /* this warp's decision words, one per step */

	/* state lane (low) and lane + 32 (high); both are entered by input bit b = lane & 1 from predecessors
	 * p0 = s >> 1 (older bit 0) and p1 = p0 | 32. Signs of the branch metric (c ? +r : -r), fixed per lane. */
	const int b = lane & 1;
	int sgn[2][2][2];                                                   /* [state low/high][pred 0/1][I/Q] */
#pragma unroll
	for (int hs = 0; hs < 2; hs++) {
		const int s = lane + 32*hs;
#pragma unroll
		for (int p = 0; p < 2; p++) {
			const unsigned r = (unsigned)((((s >> 1) | (p ? 32 : 0)) << 1) | b);
			sgn[hs][p][0] = parity7(r & FE_G1) ? 1 : -1;
			sgn[hs][p][1] = parity7(r & FE_G2) ? 1 : -1;
		}
	}
	const int src_lo = lane >> 1, src_hi = 16 + (lane >> 1);           /* lanes holding the predecessors' metrics */

	for (int f = warp; f < nframes; f += nwarps) {
		const long long start = (long long)frame_off[f];
		const int h = frame_hyp[f];
		const long long lo = start - FE_HEAD < 0 ? 0 : start - FE_HEAD;
		const long long hi = start + FE_CADU_SYMS + FE_TAIL > (long long)nsym ? (long long)nsym : start + FE_CADU_SYMS + FE_TAIL;
		const int n = (int)(hi - lo);
		int A = 0, B = 0;                                               /* pm[lane], pm[lane + 32] */
		const char2 *sym = reinterpret_cast<const char2 *>(soft) + lo;
		/* forward pass, 32 steps per batch: lane j fetches the symbol of step t0 + j (one coalesced load) */
		for (int t0 = 0; t0 < n; t0 += 32) {
			char2 mine = make_char2(0, 0);
			if (t0 + lane < n) mine = sym[t0 + lane];
			int mi, mq;
			unturn(h, mine.x, mine.y, mi, mq);
			const int steps = min(32, n - t0);
			uint32_t d_lo_mine = 0, d_hi_mine = 0;                      /* decisions of step t0 + lane */
			for (int j = 0; j < steps; j++) {
				const int ri = __shfl_sync(0xffffffffu, mi, j), rq = __shfl_sync(0xffffffffu, mq, j);
				const int a0 = __shfl_sync(0xffffffffu, A, src_lo), b0 = __shfl_sync(0xffffffffu, B, src_lo);
				const int a1 = __shfl_sync(0xffffffffu, A, src_hi), b1 = __shfl_sync(0xffffffffu, B, src_hi);
				/* low state: predecessors pm[lane>>1] (= A of src_lo) and pm[(lane>>1)+32] (= B of src_lo) */
				const int l0 = a0 + sgn[0][0][0]*ri + sgn[0][0][1]*rq;
				const int l1 = b0 + sgn[0][1][0]*ri + sgn[0][1][1]*rq;
				const int h0 = a1 + sgn[1][0][0]*ri + sgn[1][0][1]*rq;
				const int h1 = b1 + sgn[1][1][0]*ri + sgn[1][1][1]*rq;
				const bool dl = l1 > l0, dh = h1 > h0;                  /* ties to the predecessor with the older bit 0 */
				A = dl ? l1 : l0; B = dh ? h1 : h0;
				const uint32_t wl = __ballot_sync(0xffffffffu, dl), wh = __ballot_sync(0xffffffffu, dh);
				if (lane == j) { d_lo_mine = wl; d_hi_mine = wh; }
			}
			if (t0 + lane < n) dec[t0 + lane] = make_uint2(d_lo_mine, d_hi_mine);   /* coalesced: 256 bytes per batch */
		}
		/* best end state: highest metric, lowest state on ties */
		long long key = (long long)A*256 + (63 - lane);
		{
			const long long kb = (long long)B*256 + (63 - (lane + 32));
			key = kb > key ? kb : key;
		}
		for (int d = 16; d > 0; d >>= 1) { const long long o = __shfl_xor_sync(0xffffffffu, key, d); key = o > key ? o : key; }
		int s = 63 - (int)(key & 0xff);
		if (lane == 0 && metric) metric[f] = (int32_t)((key - (key & 0xff))/256);
		__syncwarp();
		/* traceback, 32 steps per batch, newest first; every lane follows the state, lane k collects byte k's bits */
		uint8_t *out = cadu + (size_t)f*FE_CADU;
		for (int i = lane; i < FE_CADU/4; i += 32) reinterpret_cast<uint32_t *>(out)[i] = 0u;
		__syncwarp();
		const int nb = (n + 31)/32;
		for (int bt = nb - 1; bt >= 0; bt--) {
			const int t0 = 32*bt;
			uint2 w = make_uint2(0u, 0u);
			if (t0 + lane < n) w = dec[t0 + lane];
			const int steps = min(32, n - t0);
			uint32_t bits = 0;                                          /* decoded input bits of this batch, step j in bit j */
			for (int j = steps - 1; j >= 0; j--) {
				const uint32_t wl = __shfl_sync(0xffffffffu, w.x, j), wh = __shfl_sync(0xffffffffu, w.y, j);
				bits |= (uint32_t)(s & 1) << j;
				const uint32_t d = (s < 32 ? (wl >> s) : (wh >> (s - 32))) & 1u;
				s = (s >> 1) | (d ? 32 : 0);
			}
			/* steps t0 .. t0+31 are frame symbols lo + t - start; lane 0 writes the four bytes they cover */
			if (lane == 0) {
				for (int j = 0; j < steps; j++) {
					const long long symi = lo + t0 + j - start;
					if (symi >= 0 && symi < FE_CADU_SYMS && ((bits >> j) & 1u)) out[symi >> 3] |= (uint8_t)(0x80u >> (symi & 7));
				}
			}
		}
		__syncwarp();
	}
}

} // namespace

#define FE_CK(x) do { if ((x) != cudaSuccess) return LRPT_ERR_CUDA; } while (0)

extern "C" int lrpt_fe_sync_device(const int8_t *d_soft, size_t nsym, uint8_t *d_score, uint8_t *d_hyp, uint32_t *d_words,
                                   void *cuda_stream)
{
	if (!d_soft || !d_score || !d_hyp || !d_words || nsym < 32 || nsym > ((size_t)1 << 32) - 64) return LRPT_ERR_ARG;
	if (((uintptr_t)d_soft & 15) || ((uintptr_t)d_score & 3) || ((uintptr_t)d_hyp & 3)) return LRPT_ERR_ARG;
	static bool have_pat = false;
	cudaStream_t st = (cudaStream_t)cuda_stream;
	{
		unsigned long long pat[8];
		for (int h = 0; h < 8; h++) pat[h] = sync_pattern(h);
		FE_CK(cudaMemcpyToSymbolAsync(c_pat, pat, sizeof(pat), 0, cudaMemcpyHostToDevice, st));   /* per device, cheap */
		have_pat = true; (void)have_pat;
	}
	const size_t nwords = (nsym + 15)/16 + 2;
	fe_pack_kernel<<<(unsigned)((nwords + 255)/256), 256, 0, st>>>(d_soft, nsym, d_words, nwords);
	const size_t nq = (nsym + 3)/4;
	fe_score_kernel<<<(unsigned)((nq + 255)/256), 256, 0, st>>>(d_words, nsym, reinterpret_cast<uint32_t *>(d_score),
	                                                          reinterpret_cast<uint32_t *>(d_hyp));
	return cudaGetLastError() == cudaSuccess ? LRPT_OK : LRPT_ERR_CUDA;
}

extern "C" size_t lrpt_fe_sync_words(size_t nsym) { return (nsym + 15)/16 + 2; }

extern "C" int lrpt_fe_peaks_device(const uint8_t *d_score, const uint8_t *d_hyp, size_t nsym, uint32_t window, uint32_t *d_off,
                                    uint8_t *d_ohyp, uint8_t *d_oscore, void *cuda_stream)
{
	if (!d_score || !d_hyp || !d_off || !d_ohyp || !d_oscore || !nsym || !window) return LRPT_ERR_ARG;
	const size_t nw = (nsym + window - 1)/window;
	if (nw > 0x7fffffffu) return LRPT_ERR_ARG;
	fe_peak_kernel<<<(unsigned)nw, 256, 0, (cudaStream_t)cuda_stream>>>(d_score, d_hyp, nsym, window, d_off, d_ohyp, d_oscore);
	return cudaGetLastError() == cudaSuccess ? LRPT_OK : LRPT_ERR_CUDA;
}

extern "C" size_t lrpt_fe_viterbi_scratch_bytes(int device)
{
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
	return (size_t)sms*8*4*FE_STEPS*sizeof(uint2);                     /* 8 CTAs of 4 warps per SM */
}

extern "C" int lrpt_fe_viterbi_device(const int8_t *d_soft, size_t nsym, const uint32_t *d_frame_off, const uint8_t *d_frame_hyp,
                                      int nframes, uint8_t *d_cadu, int32_t *d_metric, void *d_scratch, size_t scratch_bytes,
                                      void *cuda_stream)
{
	if (!d_soft || !d_frame_off || !d_frame_hyp || !d_cadu || !d_scratch || nframes < 0 || ((uintptr_t)d_cadu & 3)) return LRPT_ERR_ARG;
	if (!nframes) return LRPT_OK;
	size_t warps = scratch_bytes/((size_t)FE_STEPS*sizeof(uint2));
	if (warps < 4) return LRPT_ERR_ARG;
	if (warps > (size_t)nframes + 3) warps = (size_t)nframes + 3;
	const unsigned blocks = (unsigned)(warps/4);
	fe_viterbi_kernel<<<blocks, 128, 0, (cudaStream_t)cuda_stream>>>(d_soft, nsym, d_frame_off, d_frame_hyp, nframes, d_cadu, d_metric,
	                                                                   static_cast<uint2 *>(d_scratch));
	return cudaGetLastError() == cudaSuccess ? LRPT_OK : LRPT_ERR_CUDA;
}
