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
 *                       states l and l + 32 as two 16-bit fields of one register; the butterfly's operands come
 *                       from lanes l>>1 and 16 + (l>>1) by two shuffles, the branch metrics of both states are one
 *                       packed integer I*cI + Q*cQ with per-lane constants, add-compare-select of both states is
 *                       two adds and one VIMNMX.S16x2 (DPX); a lane keeps the decisions of its own two states for
 *                       32 steps in two registers and writes them coalesced into a scratch ring; traceback walks
 *                       it backwards 32 steps at a time with one shuffle per step and stores four bytes per batch.
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

struct SyncPatterns { unsigned long long v[8]; };   /* the encoded ASM under the 8 symmetries: a kernel argument (constant bank) */

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
__global__ void fe_score_kernel(const uint32_t *__restrict__ words, size_t nsym, uint32_t *__restrict__ score4, uint32_t *__restrict__ hyp4,
                                const SyncPatterns pat)
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
			const int s = 64 - __popcll(win ^ pat.v[h]);
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

/*
 * Path metrics travel as two signed 16-bit fields of one register, P = pm[lane] | pm[lane + 32] << 16, so that one
 * shuffle moves two metrics and one VIMNMX.S16x2 (the DPX max with per-field predicates) does the compare-select of
 * both states of a lane. Every field stays inside [0, 32767] -- then the packed fields add and subtract as one
 * 32-bit integer without borrows between them: the metrics of the 64 states of this code never differ by more than
 * 12 branch metrics (any state is reached from any other in K - 1 = 6 steps; |branch metric| <= 256) = 3072, and at
 * the start of every batch of 32 steps the smallest is moved to 8192 (it can fall by 256 a step), the moved amount
 * kept in `base`. Decisions depend on metric differences only, so they are the oracle's (32-bit metrics, no
 * normalisation) bit for bit, and field + base is its metric.
 */
static_assert(FE_STEPS % 32 == 0, "the scratch ring holds whole batches");
static_assert((FE_G1 & 0x40) && (FE_G2 & 0x40), "both generators tap the oldest bit: the two branches into a state carry opposite symbols");
constexpr int FE_BIAS = 8192;

__global__ void __launch_bounds__(128)
fe_viterbi_kernel(const int8_t *__restrict__ soft, size_t nsym, const uint32_t *__restrict__ frame_off, const uint8_t *__restrict__ frame_hyp,
                  int nframes, uint8_t *__restrict__ cadu, int32_t *__restrict__ metric, uint2 *__restrict__ scratch)
{
	__shared__ __align__(16) int2 s_sym[4][32];                         /* per warp: (I, Q) of the 32 steps of a batch, symmetry undone */
	const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
	const int warp = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
	const int nwarps = (gridDim.x*blockDim.x) >> 5;
	uint2 *dec = scratch + (size_t)warp*FE_STEPS;                       /* decision words: [batch][lane] = 32 steps of states lane, lane + 32 */

	/* Lane l holds states l (low field) and l + 32 (high field). Both are entered by input bit b = l & 1 from
	 * p0 = s >> 1 (older bit 0) and p1 = p0 | 32; the encoder register of the p0 branch is (p0 << 1) | b = s itself,
	 * the p1 branch carries the opposite symbol pair. Branch metric of the p0 branch: sI*I + sQ*Q. */
	const int sI_lo = parity7((unsigned)lane & FE_G1) ? 1 : -1, sQ_lo = parity7((unsigned)lane & FE_G2) ? 1 : -1;
	const int sI_hi = parity7((unsigned)(lane + 32) & FE_G1) ? 1 : -1, sQ_hi = parity7((unsigned)(lane + 32) & FE_G2) ? 1 : -1;
	const int cI = sI_lo + sI_hi*65536, cQ = sQ_lo + sQ_hi*65536;       /* bm_lo + 65536*bm_hi = I*cI + Q*cQ */
	const int src_lo = lane >> 1, src_hi = 16 + (lane >> 1);           /* lanes holding the predecessors' metrics */

	for (int f = warp; f < nframes; f += nwarps) {
		const long long start = (long long)frame_off[f];
		const int h = frame_hyp[f];
		const long long lo = start - FE_HEAD < 0 ? 0 : start - FE_HEAD;
		const long long hi = start + FE_CADU_SYMS + FE_TAIL > (long long)nsym ? (long long)nsym : start + FE_CADU_SYMS + FE_TAIL;
		const int n = (int)(hi - lo);
		unsigned P = (unsigned)FE_BIAS*65537u;                          /* all metrics 0 */
		int base = -FE_BIAS;
		const char2 *sym = reinterpret_cast<const char2 *>(soft) + lo;

		/* one add-compare-select step of both states of the lane; the survivor bits ("came from p0") shift into acc */
#define FE_ACS(ri, rq) do { \
			const unsigned v0 = __shfl_sync(0xffffffffu, P, src_lo), v1 = __shfl_sync(0xffffffffu, P, src_hi); \
			const unsigned bm = (unsigned)((ri)*cI + (rq)*cQ); \
			const unsigned l0 = __byte_perm(v0, v1, 0x5410) + bm;       /* pm[p0] + bm, fields (low state, high state) */ \
			const unsigned l1 = __byte_perm(v0, v1, 0x7632) - bm;       /* pm[p1] - bm */ \
			bool keep_hi, keep_lo;                                      /* l0 >= l1: ties to the predecessor with the older bit 0 */ \
			P = __vibmax_s16x2(l0, l1, &keep_hi, &keep_lo); \
			acc_lo = (acc_lo << 1) | (keep_lo ? 1u : 0u); acc_hi = (acc_hi << 1) | (keep_hi ? 1u : 0u); \
		} while (0)

		for (int t0 = 0; t0 < n; t0 += 32) {
			char2 mine = make_char2(0, 0);
			if (t0 + lane < n) mine = sym[t0 + lane];                   /* one coalesced load per batch */
			int mi, mq;
			unturn(h, mine.x, mine.y, mi, mq);
			s_sym[wq][lane] = make_int2(mi, mq);
			/* smallest metric back to FE_BIAS */
			unsigned m = P;
#pragma unroll
			for (int d = 16; d > 0; d >>= 1) m = __vmins2(m, __shfl_xor_sync(0xffffffffu, m, d));
			{
				const int f0 = (int)(m & 0xffffu), f1 = (int)(m >> 16);
				const int delta = (f0 < f1 ? f0 : f1) - FE_BIAS;
				P -= (unsigned)delta*65537u; base += delta;
			}
			__syncwarp();
			const int steps = min(32, n - t0);
			unsigned acc_lo = 0, acc_hi = 0;
			if (steps == 32) {
#pragma unroll
				for (int j = 0; j < 32; j += 2) {
					const int4 r = *reinterpret_cast<const int4 *>(&s_sym[wq][j]);   /* two steps per broadcast load */
					FE_ACS(r.x, r.y);
					FE_ACS(r.z, r.w);
				}
			} else {
				for (int j = 0; j < steps; j++) { const int2 r = s_sym[wq][j]; FE_ACS(r.x, r.y); }
				acc_lo <<= 32 - steps; acc_hi <<= 32 - steps;
			}
			/* decision bit = 1 when the survivor came from p1; step j of the batch at bit 31 - j */
			dec[t0 + lane] = make_uint2(~acc_lo, ~acc_hi);              /* coalesced: 256 bytes per batch */
			__syncwarp();                                               /* s_sym is rewritten by the next batch */
		}
#undef FE_ACS
		/* best end state: highest metric, lowest state on ties */
		const int A = (int)(P & 0xffffu) + base, B = (int)(P >> 16) + base;
		long long key = (long long)A*256 + (63 - lane);
		{
			const long long kb = (long long)B*256 + (63 - (lane + 32));
			key = kb > key ? kb : key;
		}
		for (int d = 16; d > 0; d >>= 1) { const long long o = __shfl_xor_sync(0xffffffffu, key, d); key = o > key ? o : key; }
		int s = 63 - (int)(key & 0xff);
		if (lane == 0 && metric) metric[f] = (int32_t)((key - (key & 0xff))/256);
		__syncwarp();
		/* traceback, 32 steps per batch, newest first. Every lane follows the state; the decision of state s at a step
		 * sits in lane s & 31 (field s >> 5): one shuffle per step. */
		uint8_t *out = cadu + (size_t)f*FE_CADU;
		for (int i = lane; i < FE_CADU/4; i += 32) reinterpret_cast<uint32_t *>(out)[i] = 0u;
		__syncwarp();
		const bool word_aligned = ((lo - start) & 31) == 0;             /* a batch = four whole bytes of the frame */
		const int nb = (n + 31)/32;
		for (int bt = nb - 1; bt >= 0; bt--) {
			const int t0 = 32*bt;
			const uint2 w = dec[t0 + lane];                             /* a word per STATE pair: all 32 lanes, also in a short batch */
			const int steps = min(32, n - t0);
			uint32_t bits = 0;                                          /* decoded input bits of this batch, step j in bit j */
#define FE_BACK(k) do { \
				const unsigned c = ((w.x >> (k)) & 1u) | (((w.y >> (k)) & 1u) << 1); \
				const unsigned v = __shfl_sync(0xffffffffu, c, s & 31); \
				bits = (bits << 1) | (unsigned)(s & 1); \
				s = (s >> 1) | (int)(((v >> (s >> 5)) & 1u) << 5); \
			} while (0)
			if (steps == 32) {
#pragma unroll
				for (int k = 0; k < 32; k++) FE_BACK(k);                /* step j = 31 - k */
			} else {
				for (int j = steps - 1; j >= 0; j--) FE_BACK(31 - j);
			}
#undef FE_BACK
			if (lane == 0) {
				const long long sym0 = lo + t0 - start;                 /* frame symbol of step 0 of the batch */
				if (word_aligned && steps == 32) {
					/* symbol sym0 + j is bit 7 - (j & 7) of byte (sym0 + j) >> 3: reverse the bits, then the bytes */
					if (sym0 >= 0 && sym0 < FE_CADU_SYMS)
						*reinterpret_cast<uint32_t *>(out + (sym0 >> 3)) = __byte_perm(__brev(bits), 0u, 0x0123);
				} else {
					for (int j = 0; j < steps; j++) {
						const long long symi = sym0 + j;
						if (symi >= 0 && symi < FE_CADU_SYMS && ((bits >> j) & 1u)) out[symi >> 3] |= (uint8_t)(0x80u >> (symi & 7));
					}
				}
			}
		}
		__syncwarp();
	}
}

} // namespace

extern "C" int lrpt_fe_sync_device(const int8_t *d_soft, size_t nsym, uint8_t *d_score, uint8_t *d_hyp, uint32_t *d_words,
                                   void *cuda_stream)
{
	if (!d_soft || !d_score || !d_hyp || !d_words || nsym < 32 || nsym > ((size_t)1 << 32) - 64) return LRPT_ERR_ARG;
	if (((uintptr_t)d_soft & 15) || ((uintptr_t)d_score & 3) || ((uintptr_t)d_hyp & 3)) return LRPT_ERR_ARG;
	cudaStream_t st = (cudaStream_t)cuda_stream;
	SyncPatterns pat;
	for (int h = 0; h < 8; h++) pat.v[h] = sync_pattern(h);
	const size_t nwords = (nsym + 15)/16 + 2;
	fe_pack_kernel<<<(unsigned)((nwords + 255)/256), 256, 0, st>>>(d_soft, nsym, d_words, nwords);
	const size_t nq = (nsym + 3)/4;
	fe_score_kernel<<<(unsigned)((nq + 255)/256), 256, 0, st>>>(d_words, nsym, reinterpret_cast<uint32_t *>(d_score),
	                                                          reinterpret_cast<uint32_t *>(d_hyp), pat);
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
	return (size_t)sms*16*4*FE_STEPS*sizeof(uint2);                    /* 16 CTAs of 4 warps per SM */
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
