/*
 * fir_stage.cu -- the feed-forward stage of the path, alone: ingest + ALL-PHASE RRC polyphase FIR -> HBM.
 *
 * Of the reference's chain only filter_fwd_sample / filter_get (filter.c:39-65) is feed-forward: the value
 * of filter_get(flt, i) after sample n has been pushed depends on samples and taps only, never on loop
 * state (AGC, Costas loop and timing run at SYMBOL rate behind it, demod.c:33-43). This kernel computes that
 * value for EVERY (sample n, sub-step i) of a batch of rows and writes it to HBM -- the "RRC matched filter
 * stage" BASELINE.json's north_star wants measured on its own: per input
 * sample 2*bps/8 bytes in, 8*L bytes out, 4*taps*L flops. The product path never materialises this
 * (demod_lane.cu evaluates the one output per symbol the timing loop picks; demod_ws.cu keeps the all-phase
 * outputs in shared memory); the entry point exists for the roofline measurement and as a building block.
 *
 * B200 mapping: persistent CTAs (grid = a multiple of the SM count) loop over (row, tile) work items. The tap
 * table (2-8 KB) and every tile of raw samples (+ taps-1 samples of left halo) arrive in shared memory by
 * TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) issued by one thread, the next
 * tile's copy in flight while this tile is filtered; raw samples are converted once per tile to a float2
 * window (wavfile.c:58-69); thread = output sample, L packed-f32x2 accumulators (I and Q of one phase ride
 * one FMUL2 + FFMA2 pair) for each of TWO samples half a tile apart, so every coefficient load feeds 2*L
 * multiply-accumulates (one sample per thread for short filters); the window is read conflict-free (consecutive threads, consecutive samples), the taps as
 * broadcasts; outputs go out as 8-byte stores, L consecutive per thread.
 *
 * mode 0 ("exact"): acc = RN(acc + RN(x*h)), oldest tap first -- bit-identical to filter_get at every
 * (n, i); two instructions per tap and phase. mode 1 ("fma"): acc = RN(acc + x*h), one instruction per tap
 * and phase, what the reference's own -march=native build does in parts (Tier-S).
 * Either way the stage is bound by the FP32 pipe, not by HBM: at 65 taps x 5 phases the exact form needs
 * 1300 fp32 lane-cycles per sample against 44 bytes.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "demod_core.cuh"
#include "lrpt_b200.h"
#include "lrpt_internal.h"

namespace lrpt {

constexpr int FS_THREADS = 256;
constexpr int FS_TILE    = 1024;     /* samples per work item */
constexpr int FS_MAX_TAPS = 257;

struct FirStageArgs {
	const float   *hT;          /* [taps][LP] transposed tap table, bank p of tap k at hT[k*LP + p], zero padded */
	const uint8_t *raw; size_t raw_stride;
	float2        *out; size_t out_stride;   /* bytes */
	int            nrows, nsamples, taps, HP; /* HP: halo rounded up so that every bulk copy starts 16-byte aligned */
	float          one;          /* 1.0f, opaque to the compiler (packed mul+add must not contract, demod_lane.cu) */
};

LRPT_DEV uint32_t fs_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

LRPT_DEV void fs_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(fs_smem(dst)), "l"(src), "r"(bytes), "r"(fs_smem(bar)) : "memory");
}

LRPT_DEV void fs_expect(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(fs_smem(bar)), "r"(bytes) : "memory");
}

LRPT_DEV void fs_wait(uint64_t *bar, unsigned parity)
{
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"W_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra D_%=;\n\t"
		"bra W_%=;\n\t"
		"D_%=:\n\t}"
		:: "r"(fs_smem(bar)), "r"(parity) : "memory");
}

typedef unsigned long long fs2_t;
LRPT_DEV fs2_t fs_pk(float a, float b) { fs2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
LRPT_DEV float2 fs_upk(fs2_t v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
LRPT_DEV fs2_t fs_mul(fs2_t a, fs2_t b) { fs2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
LRPT_DEV fs2_t fs_fma(fs2_t a, fs2_t b, fs2_t c) { fs2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int BPS> LRPT_DEV float2 fs_cvt(const uint8_t *raw, int i);
template <> LRPT_DEV float2 fs_cvt<16>(const uint8_t *raw, int i)
{
	const short2 v = reinterpret_cast<const short2 *>(raw)[i];
	return make_float2((float)v.x, (float)v.y);
}
template <> LRPT_DEV float2 fs_cvt<8>(const uint8_t *raw, int i)
{
	const uchar2 v = reinterpret_cast<const uchar2 *>(raw)[i];
	return make_float2((float)((int)v.x - 128), (float)((int)v.y - 128));
}
template <> LRPT_DEV float2 fs_cvt<32>(const uint8_t *raw, int i) { return reinterpret_cast<const float2 *>(raw)[i]; }

template <int L, int BPS, bool FMA, int NS>
__global__ void __launch_bounds__(FS_THREADS, 3)
fir_stage_kernel(const FirStageArgs a)
{
	constexpr int LP = (L <= 4) ? 4 : 8;
	constexpr int T = FS_TILE, BYTES = BPS/4;
	extern __shared__ __align__(128) unsigned char smem[];
	const int taps = a.taps, HP = a.HP;
	uint64_t *bar = reinterpret_cast<uint64_t *>(smem);                 /* [2]: taps, raw tile */
	float *hT = reinterpret_cast<float *>(smem + 128);                  /* [taps][LP] */
	const int hbytes = ((taps*LP*4 + 127)/128)*128;
	uint8_t *stage = smem + 128 + hbytes;                               /* [(HP + T)*BYTES] raw bytes, TMA destination */
	const int sbytes = (((HP + T)*BYTES + 127)/128)*128;
	float2 *xf = reinterpret_cast<float2 *>(stage + sbytes);            /* [HP + T] converted window */

	const int tiles = (a.nsamples + T - 1)/T;
	const int items = a.nrows*tiles;
	const int tid = threadIdx.x;

	/* first sample the copy of item `it` starts at (may be negative: power-on delay line = zeros, filter.c:16) */
	auto issue = [&](int it) {
		const int row = it/tiles, t = it - row*tiles;
		const int s0 = t*T - HP;
		const int lo = s0 < 0 ? 0 : s0;
		const int hi = min(a.nsamples, (t + 1)*T);
		const int n16 = ((hi - lo)*BYTES) & ~15;                        /* whole 16-byte units by TMA, the tail by threads */
		fs_expect(&bar[1], (uint32_t)n16);
		if (n16) fs_bulk_g2s(stage + (size_t)(lo - s0)*BYTES, a.raw + (size_t)row*a.raw_stride + (size_t)lo*BYTES, (uint32_t)n16, &bar[1]);
	};

	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(fs_smem(&bar[0])));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(fs_smem(&bar[1])));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (tid == 0) {
		fs_expect(&bar[0], (uint32_t)(taps*LP*4));
		fs_bulk_g2s(hT, a.hT, (uint32_t)(taps*LP*4), &bar[0]);          /* taps TMA-staged to shared memory */
		if ((int)blockIdx.x < items) issue(blockIdx.x);
	}
	fs_wait(&bar[0], 0);

	const fs2_t one2 = fs_pk(a.one, a.one);
	unsigned phase = 0;
	for (int it = blockIdx.x; it < items; it += gridDim.x, phase ^= 1) {
		const int row = it/tiles, t = it - row*tiles;
		const int s0 = t*T - HP;
		const int lo = s0 < 0 ? 0 : s0;
		const int hi = min(a.nsamples, (t + 1)*T);
		const int n16 = ((hi - lo)*BYTES) & ~15;
		fs_wait(&bar[1], phase);
		/* the bytes TMA did not bring: zeros in front of the stream, the unaligned tail, zeros behind the end */
		{
			const uint8_t *src = a.raw + (size_t)row*a.raw_stride;
			const int tail0 = lo + n16/BYTES;                           /* first sample not covered by the bulk copy */
			for (int i = tid; i < HP + T; i += FS_THREADS) {
				const int m = s0 + i;
				float2 v;
				if (m < 0 || m >= hi) v = make_float2(0.f, 0.f);
				else if (m >= tail0) v = fs_cvt<BPS>(src, m);
				else v = fs_cvt<BPS>(stage, i);
				xf[i] = v;
			}
		}
		__syncthreads();                                               /* window complete, raw stage free again */
		if (tid == 0 && it + (int)gridDim.x < items) issue(it + gridDim.x);   /* next tile's copy overlaps this tile's FIR */

		float2 *orow = reinterpret_cast<float2 *>(reinterpret_cast<char *>(a.out) + (size_t)row*a.out_stride);
		const int off = HP - (taps - 1);                                /* window entry of the oldest tap of local sample 0 */
		/* NS = 2: two output samples per thread (tile halves), so that every coefficient load feeds 2*L
		 * multiply-accumulates (long filters: the FMA pipe is the bound); NS = 1: one sample per thread (short
		 * filters: more independent threads per tile to cover the stores) */
#pragma unroll 1
		for (int nl = tid; nl < T/NS; nl += FS_THREADS) {
			const int n0 = t*T + nl, n1 = n0 + T/2;
			if (n0 >= a.nsamples) break;
			const float2 *w0 = xf + off + nl, *w1 = w0 + T/2;
			fs2_t acc0[L], acc1[NS == 2 ? L : 1];
#pragma unroll
			for (int p = 0; p < L; p++) { acc0[p] = fs_pk(0.0f, 0.0f); if (NS == 2) acc1[p] = fs_pk(0.0f, 0.0f); }
#pragma unroll 4
			for (int k = 0; k < taps; k++) {
				const float2 xa = w0[k];
				const fs2_t xa2 = fs_pk(xa.x, xa.y);
				fs2_t xb2 = xa2;
				if (NS == 2) { const float2 xb = w1[k]; xb2 = fs_pk(xb.x, xb.y); }
				float hv[LP];
				*reinterpret_cast<float4 *>(hv) = *reinterpret_cast<const float4 *>(hT + k*LP);
				if (LP == 8) *reinterpret_cast<float4 *>(hv + 4) = *reinterpret_cast<const float4 *>(hT + k*LP + 4);
#pragma unroll
				for (int p = 0; p < L; p++) {
					const fs2_t h2 = fs_pk(hv[p], hv[p]);
					if (FMA) {
						acc0[p] = fs_fma(xa2, h2, acc0[p]);
						if (NS == 2) acc1[p] = fs_fma(xb2, h2, acc1[p]);
					} else {                                        /* RN(acc + RN(x*h)): filter.c:57-62 */
						acc0[p] = fs_fma(fs_mul(xa2, h2), one2, acc0[p]);
						if (NS == 2) acc1[p] = fs_fma(fs_mul(xb2, h2), one2, acc1[p]);
					}
				}
			}
			/* sub-step i reads bank L-1-i (filter.c:52) */
#pragma unroll
			for (int i = 0; i < L; i++) orow[(size_t)n0*L + i] = fs_upk(acc0[L - 1 - i]);
			if (NS == 2 && n1 < a.nsamples) {
#pragma unroll
				for (int i = 0; i < L; i++) orow[(size_t)n1*L + i] = fs_upk(acc1[L - 1 - i]);
			}
		}
		__syncthreads();                                               /* xf is rewritten by the next item */
	}
}

template <int L, int BPS, bool FMA, int NS> static cudaError_t fs_launch3(const FirStageArgs &a, int blocks, size_t smem, cudaStream_t st)
{
	cudaError_t e;
	if ((e = cudaFuncSetAttribute(fir_stage_kernel<L, BPS, FMA, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
	fir_stage_kernel<L, BPS, FMA, NS><<<blocks, FS_THREADS, smem, st>>>(a);
	return cudaGetLastError();
}

template <int L, int BPS> static cudaError_t fs_launch2(const FirStageArgs &a, int mode, int blocks, size_t smem, cudaStream_t st)
{
	int ns = (mode >> 1) & 3;                                           /* 0 = by filter length */
	if (ns == 0) ns = a.taps*L >= 160 ? 2 : 1;
	if (mode & 1) return ns == 2 ? fs_launch3<L, BPS, true, 2>(a, blocks, smem, st) : fs_launch3<L, BPS, true, 1>(a, blocks, smem, st);
	return ns == 2 ? fs_launch3<L, BPS, false, 2>(a, blocks, smem, st) : fs_launch3<L, BPS, false, 1>(a, blocks, smem, st);
}

template <int L> static cudaError_t fs_launch1(const FirStageArgs &a, int bps, int mode, int blocks, size_t smem, cudaStream_t st)
{
	if (bps == 16) return fs_launch2<L, 16>(a, mode, blocks, smem, st);
	if (bps == 8)  return fs_launch2<L, 8>(a, mode, blocks, smem, st);
	return fs_launch2<L, 32>(a, mode, blocks, smem, st);
}

} // namespace lrpt

using namespace lrpt;

extern "C" int lrpt_fir_stage_device(const lrpt_params_t *p, const void *d_raw, size_t raw_stride, int nrows, size_t nsamples,
                                     float *d_out, size_t out_stride, int mode, void *cuda_stream)
{
	if (!p || !d_raw || !d_out || nrows < 1 || nsamples < 1 || nsamples > ((size_t)1 << 30)) return LRPT_ERR_ARG;
	const int L = p->interp_factor, taps = 2*p->rrc_order + 1;
	if (L < 1 || L > 8 || taps < 1 || taps > FS_MAX_TAPS || (p->bps != 8 && p->bps != 16 && p->bps != 32)) return LRPT_ERR_ARG;
	if (((uintptr_t)d_raw | raw_stride | (uintptr_t)d_out | out_stride) & 15) return LRPT_ERR_ARG;
	if (out_stride < nsamples*(size_t)L*8) return LRPT_ERR_ARG;
	std::vector<float> h((size_t)taps*L);
	lrpt_consts_t c; lrpt_state_t s0;
	int rc = lrpt_derive(p, &c, &s0, h.data());                         /* the reference's own taps (filter.c:10-29) */
	if (rc) return rc;
	const int LP = (L <= 4) ? 4 : 8;
	std::vector<float> hT((size_t)taps*LP, 0.0f);
	for (int k = 0; k < taps; k++) for (int b = 0; b < L; b++) hT[(size_t)k*LP + b] = h[(size_t)b*taps + k];
	if (cudaSetDevice(p->device) != cudaSuccess) return LRPT_ERR_CUDA;
	cudaStream_t st = (cudaStream_t)cuda_stream;
	float *d_hT = nullptr;
	if (cudaMallocAsync((void **)&d_hT, hT.size()*4, st) != cudaSuccess) return LRPT_ERR_NOMEM;
	if (cudaMemcpyAsync(d_hT, hT.data(), hT.size()*4, cudaMemcpyHostToDevice, st) != cudaSuccess) { cudaFreeAsync(d_hT, st); return LRPT_ERR_CUDA; }
	if (cudaStreamSynchronize(st) != cudaSuccess) { cudaFreeAsync(d_hT, st); return LRPT_ERR_CUDA; }   /* hT[] is pageable and local */
	const int bytes = p->bps/4, per16 = 16/bytes;
	FirStageArgs a;
	a.hT = d_hT; a.raw = (const uint8_t *)d_raw; a.raw_stride = raw_stride;
	a.out = (float2 *)d_out; a.out_stride = out_stride;
	a.nrows = nrows; a.nsamples = (int)nsamples; a.taps = taps;
	a.HP = ((taps - 1 + per16 - 1)/per16)*per16;
	a.one = 1.0f;
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
	const int tiles = (int)((nsamples + FS_TILE - 1)/FS_TILE);
	const long long items = (long long)nrows*tiles;
	const int blocks = (int)(items < 3LL*sms ? items : 3LL*sms);      /* persistent: three CTAs per SM */
	const size_t smem = 128 + (size_t)((taps*LP*4 + 127)/128)*128 + (size_t)(((a.HP + FS_TILE)*bytes + 127)/128)*128 +
	                    (size_t)(a.HP + FS_TILE)*8;
	cudaError_t e;
	switch (L) {
		case 1: e = fs_launch1<1>(a, p->bps, mode, blocks, smem, st); break;
		case 2: e = fs_launch1<2>(a, p->bps, mode, blocks, smem, st); break;
		case 3: e = fs_launch1<3>(a, p->bps, mode, blocks, smem, st); break;
		case 4: e = fs_launch1<4>(a, p->bps, mode, blocks, smem, st); break;
		case 5: e = fs_launch1<5>(a, p->bps, mode, blocks, smem, st); break;
		case 6: e = fs_launch1<6>(a, p->bps, mode, blocks, smem, st); break;
		case 7: e = fs_launch1<7>(a, p->bps, mode, blocks, smem, st); break;
		default: e = fs_launch1<8>(a, p->bps, mode, blocks, smem, st); break;
	}
	cudaFreeAsync(d_hT, st);                                            /* stream-ordered: after the kernel */
	return e == cudaSuccess ? LRPT_OK : LRPT_ERR_CUDA;
}
