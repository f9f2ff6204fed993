/*
 * shard_stitch.cu -- joining the chunks of ONE time-sharded stream on the device (Tier-S, DESIGN.md
 * section 7): the rows are consecutive chunks demodulated as independent streams, each with its int8
 * symbols and, per symbol, the timing sub-step that produced it (lrpt_set_symbol_index_output; ascending
 * within a row, counted from the row's own first sample -- `base[r]` makes it absolute). Nothing of this
 * exists in the reference (one sequential stream, demod.c:24-48); it is the multi-GPU side of the path.
 *
 *   find_cuts   per boundary (row r | row r+1): the cut point mid-way between two symbols of row r around
 *               the boundary's target, and where the paired overlap symbols start in either row
 *   quadrants   per boundary: the quarter-turn count k that maps row r+1 onto row r (a QPSK Costas loop
 *               locks with a k*90 degree ambiguity, pll.c:143-152), from exact integer correlations of
 *               the paired soft symbols, and how many pairs agree under it
 *   ranges      per row: the contiguous run of symbols between its two cut points
 *   gather      the rows' runs, each turned back by its cumulative quarter turns (exact on int8 pairs),
 *               into one contiguous symbol stream
 *
 * The same arithmetic in torch ops is meteor_demod_b200/sharded.py (used on CPU tensors by the tests);
 * these kernels touch only the symbols they need instead of whole [rows x capacity] tables.
 */
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include "lrpt_b200.h"

namespace lrpt {

/* first index in [0, n) whose absolute sub-step q[i] + base is >= target; n if there is none */
__device__ __forceinline__ int first_at_or_after(const uint32_t *q, int n, long long base, long long target)
{
	int lo = 0, hi = n;
	while (lo < hi) {
		const int mid = (lo + hi) >> 1;
		if ((long long)q[mid] + base < target) lo = mid + 1; else hi = mid;
	}
	return lo;
}

__device__ __forceinline__ const uint32_t *qrow(const uint32_t *q, size_t stride, int r)
{
	return reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(q) + (size_t)r*stride);
}

__global__ void shard_find_cuts_kernel(const uint32_t *q, size_t q_stride, const int32_t *count, const long long *base,
                                       int nb, const long long *target, long long *cut, int32_t *ia_out,
                                       int32_t *ib_out, int32_t *navail)
{
	const int b = blockIdx.x*blockDim.x + threadIdx.x;
	if (b >= nb) return;
	const uint32_t *qa = qrow(q, q_stride, b), *qb = qrow(q, q_stride, b + 1);
	const int na = count[b], nbn = count[b + 1];
	const long long ba = base[b], bb = base[b + 1], B = target[b];
	const int ia = first_at_or_after(qa, na, ba, B);
	long long c = B;
	if (ia > 0 && ia < na) c = ((long long)qa[ia - 1] + ba + (long long)qa[ia] + ba)/2;   /* mid-way between two symbols */
	const int ib = first_at_or_after(qb, nbn, bb, c + 1);
	cut[b] = c; ia_out[b] = ia; ib_out[b] = ib;
	const int av = min(na - ia, nbn - ib);
	navail[b] = av > 0 ? av : 0;
}

__device__ __forceinline__ int sign3(int v) { return (v > 0) - (v < 0); }

/* one block per boundary */
__global__ void shard_quadrants_kernel(const int8_t *soft, size_t soft_stride, const uint32_t *q, size_t q_stride,
                                       const long long *base, const int32_t *ia_in, const int32_t *ib_in, int npairs,
                                       int32_t *k_out, int32_t *same_out)
{
	const int b = blockIdx.x;
	const char2 *sa = reinterpret_cast<const char2 *>(soft + (size_t)b*soft_stride) + ia_in[b];
	const char2 *sb = reinterpret_cast<const char2 *>(soft + (size_t)(b + 1)*soft_stride) + ib_in[b];
	const uint32_t *qa = qrow(q, q_stride, b) + ia_in[b], *qb = qrow(q, q_stride, b + 1) + ib_in[b];
	const long long dbase = base[b] - base[b + 1];
	__shared__ long long red0[32], red1[32];
	__shared__ int kk;
	/* correlations of row b with row b+1 turned by 0 and by 1 quarter turn ((I,Q) -> (-Q,I)); turns 2 and 3 are their negatives */
	long long s0 = 0, s1 = 0;
	for (int j = threadIdx.x; j < npairs; j += blockDim.x) {
		const char2 a = sa[j], c = sb[j];
		s0 += (int)a.x*c.x + (int)a.y*c.y;
		s1 += -(int)a.x*c.y + (int)a.y*c.x;
	}
	for (int o = 16; o; o >>= 1) { s0 += __shfl_down_sync(0xffffffffu, s0, o); s1 += __shfl_down_sync(0xffffffffu, s1, o); }
	if ((threadIdx.x & 31) == 0) { red0[threadIdx.x >> 5] = s0; red1[threadIdx.x >> 5] = s1; }
	__syncthreads();
	if (threadIdx.x == 0) {
		long long t0 = 0, t1 = 0;
		for (int w = 0; w < (int)(blockDim.x + 31)/32; w++) { t0 += red0[w]; t1 += red1[w]; }
		const long long sc[4] = { t0, t1, -t0, -t1 };
		int k = 0;
		for (int i = 1; i < 4; i++) if (sc[i] > sc[k]) k = i;      /* first maximum */
		kk = k;
	}
	__syncthreads();
	const int k = kk;
	int same = 0;
	for (int j = threadIdx.x; j < npairs; j += blockDim.x) {
		const char2 a = sa[j], c = sb[j];
		int ri, rq;
		switch (k) { case 0: ri = c.x; rq = c.y; break; case 1: ri = -c.y; rq = c.x; break;
		             case 2: ri = -c.x; rq = -c.y; break; default: ri = c.y; rq = -c.x; }
		long long dq = (long long)qa[j] - (long long)qb[j] + dbase;
		if (dq < 0) dq = -dq;
		same += (sign3(ri) == sign3(a.x)) && (sign3(rq) == sign3(a.y)) && dq <= 2;
	}
	for (int o = 16; o; o >>= 1) same += __shfl_down_sync(0xffffffffu, same, o);
	__shared__ int reds[32];
	if ((threadIdx.x & 31) == 0) reds[threadIdx.x >> 5] = same;
	__syncthreads();
	if (threadIdx.x == 0) {
		int t = 0;
		for (int w = 0; w < (int)(blockDim.x + 31)/32; w++) t += reds[w];
		k_out[b] = k; same_out[b] = t;
	}
}

/* OQPSK boundaries (meteor_demod_b200/sharded.py::boundary_quadrants_oqpsk, same arithmetic in integers): the I arm is
 * sampled half a symbol before the Q arm (demod.c:66-83) and the timing detector reads only Q (timing.c:65), so at an
 * ODD number of quarter turns the later row takes its symbols `half` sub-steps off, its Q arm carrying the earlier
 * row's I stream and its I arm the earlier row's Q stream of the symbol before: b.Q_j = -+a.I_{m+1}, b.I_j = +-a.Q_m
 * with q_b[j] ~ q_a[m] + half. The median timing offset of the paired symbols tells even from odd, a correlation
 * picks the sign. One block per boundary; npairs <= what find_cuts reported minus one symbol of row a in reserve. */
__global__ void shard_quadrants_oqpsk_kernel(const int8_t *soft, size_t soft_stride, const uint32_t *q, size_t q_stride,
                                             const long long *base, const int32_t *ia_in, const int32_t *ib_in, int npairs,
                                             float half_substeps, int32_t *k_out, int32_t *same_out)
{
	const int b = blockIdx.x;
	const int ia = ia_in[b], ib = ib_in[b];
	const char2 *ra = reinterpret_cast<const char2 *>(soft + (size_t)b*soft_stride);
	const char2 *sb = reinterpret_cast<const char2 *>(soft + (size_t)(b + 1)*soft_stride) + ib;
	const uint32_t *qa = qrow(q, q_stride, b) + ia, *qb = qrow(q, q_stride, b + 1) + ib;
	const long long dbase = base[b + 1] - base[b];                  /* d = q_b - q_a in absolute sub-steps */
	__shared__ int hist[257];
	__shared__ long long red0[32], red1[32];
	__shared__ int s_dm, s_k, reds[32];
	for (int i = threadIdx.x; i < 257; i += blockDim.x) hist[i] = 0;
	__syncthreads();
	for (int j = threadIdx.x; j < npairs; j += blockDim.x) {
		long long d = (long long)qb[j] - (long long)qa[j] + dbase;
		d = d < -128 ? -128 : d > 128 ? 128 : d;
		atomicAdd(&hist[(int)d + 128], 1);
	}
	__syncthreads();
	if (threadIdx.x == 0) {                                         /* lower median, as torch.median */
		const int want = (npairs - 1)/2 + 1;
		int cum = 0, v = 0;
		for (v = 0; v < 257; v++) { cum += hist[v]; if (cum >= want) break; }
		s_dm = v - 128;
	}
	__syncthreads();
	const int dm = s_dm;
	const bool odd = fabsf((float)dm) > half_substeps*0.5f;
	const int off = dm > 0 ? 0 : -1;
	long long s0 = 0, s1 = 0;                                       /* even score, odd score */
	for (int j = threadIdx.x; j < npairs; j += blockDim.x) {
		const char2 a = ra[ia + j], c = sb[j];
		int m = ia + j + off;
		m = m < 0 ? 0 : m;
		const int aQm = ra[m].y, aIm1 = ra[m + 1].x;
		s0 += (int)a.x*c.x + (int)a.y*c.y;
		s1 += -aIm1*(int)c.y + aQm*(int)c.x;
	}
	for (int o = 16; o; o >>= 1) { s0 += __shfl_down_sync(0xffffffffu, s0, o); s1 += __shfl_down_sync(0xffffffffu, s1, o); }
	if ((threadIdx.x & 31) == 0) { red0[threadIdx.x >> 5] = s0; red1[threadIdx.x >> 5] = s1; }
	__syncthreads();
	if (threadIdx.x == 0) {
		long long t0 = 0, t1 = 0;
		for (int w = 0; w < (int)(blockDim.x + 31)/32; w++) { t0 += red0[w]; t1 += red1[w]; }
		s_k = odd ? (t1 >= 0 ? 1 : 3) : (t0 >= 0 ? 0 : 2);
	}
	__syncthreads();
	const int k = s_k;
	const int sg = (k == 0 || k == 1) ? 1 : -1;
	int same = 0;
	for (int j = threadIdx.x; j < npairs; j += blockDim.x) {
		const char2 a = ra[ia + j], c = sb[j];
		long long d = (long long)qb[j] - (long long)qa[j] + dbase;
		if (!odd) {
			const long long ad = d < 0 ? -d : d;
			same += (sign3(sg*c.x) == sign3(a.x)) && (sign3(sg*c.y) == sign3(a.y)) && ad <= 2;
		} else {
			int m = ia + j + off;
			m = m < 0 ? 0 : m;
			const int aQm = ra[m].y, aIm1 = ra[m + 1].x;
			long long dd = d - dm;
			dd = dd < 0 ? -dd : dd;
			same += (sign3(-sg*c.y) == sign3(aIm1)) && (sign3(sg*c.x) == sign3(aQm)) && dd <= 2;
		}
	}
	for (int o = 16; o; o >>= 1) same += __shfl_down_sync(0xffffffffu, same, o);
	if ((threadIdx.x & 31) == 0) reds[threadIdx.x >> 5] = same;
	__syncthreads();
	if (threadIdx.x == 0) {
		int t = 0;
		for (int w = 0; w < (int)(blockDim.x + 31)/32; w++) t += reds[w];
		k_out[b] = k; same_out[b] = t;
	}
}

/* symbols of row r with lo < q <= hi: a contiguous run, because q ascends */
__global__ void shard_ranges_kernel(const uint32_t *q, size_t q_stride, const int32_t *count, const long long *base, int M,
                                    const long long *lo, const long long *hi, int32_t *start, int32_t *len)
{
	const int r = blockIdx.x*blockDim.x + threadIdx.x;
	if (r >= M) return;
	const uint32_t *qr = qrow(q, q_stride, r);
	const int n = count[r];
	const int i0 = first_at_or_after(qr, n, base[r], lo[r] + 1);
	const int i1 = hi[r] == LLONG_MAX ? n : first_at_or_after(qr, n, base[r], hi[r] + 1);
	start[r] = i0; len[r] = i1 > i0 ? i1 - i0 : 0;
}

/* grid (tiles, rows): out[off[r] + j] = soft[r][start[r] + j] turned by turns[r] quarter turns, (I,Q) -> (-Q,I) each */
__global__ void shard_gather_kernel(const int8_t *soft, size_t soft_stride, const int32_t *start, const int32_t *len,
                                    const long long *off, const int32_t *turns, int8_t *out, int row0)
{
	const int r = row0 + blockIdx.y;
	const int n = len[r];
	const int k = turns[r] & 3;
	const char2 *src = reinterpret_cast<const char2 *>(soft + (size_t)r*soft_stride) + start[r];
	char2 *dst = reinterpret_cast<char2 *>(out) + off[r];
	for (int j = blockIdx.x*blockDim.x + threadIdx.x; j < n; j += gridDim.x*blockDim.x) {
		const char2 v = src[j];
		char2 o;
		switch (k) { case 0: o = v; break; case 1: o = make_char2((signed char)-v.y, v.x); break;
		             case 2: o = make_char2((signed char)-v.x, (signed char)-v.y); break;
		             default: o = make_char2(v.y, (signed char)-v.x); }
		dst[j] = o;
	}
}

static int done(cudaError_t e) { return e == cudaSuccess ? LRPT_OK : LRPT_ERR_CUDA; }

} // namespace lrpt

using namespace lrpt;

extern "C" int lrpt_shard_find_cuts_device(const uint32_t *d_q, size_t q_stride, const int32_t *d_count, const int64_t *d_base,
                                           int nrows, const int64_t *d_target, int64_t *d_cut, int32_t *d_ia, int32_t *d_ib,
                                           int32_t *d_navail, void *cuda_stream)
{
	if (!d_q || !d_count || !d_base || !d_target || !d_cut || !d_ia || !d_ib || !d_navail || nrows < 1 || (q_stride & 3)) return LRPT_ERR_ARG;
	if (nrows < 2) return LRPT_OK;
	const int nb = nrows - 1;
	shard_find_cuts_kernel<<<(nb + 127)/128, 128, 0, (cudaStream_t)cuda_stream>>>(
		d_q, q_stride, d_count, reinterpret_cast<const long long *>(d_base), nb, reinterpret_cast<const long long *>(d_target),
		reinterpret_cast<long long *>(d_cut), d_ia, d_ib, d_navail);
	return done(cudaGetLastError());
}

extern "C" int lrpt_shard_quadrants_device(const int8_t *d_soft, size_t soft_stride, const uint32_t *d_q, size_t q_stride,
                                           const int64_t *d_base, int nrows, const int32_t *d_ia, const int32_t *d_ib, int npairs,
                                           int32_t *d_k, int32_t *d_same, void *cuda_stream)
{
	if (!d_soft || !d_q || !d_base || !d_ia || !d_ib || !d_k || !d_same || nrows < 1 || npairs < 0 || (q_stride & 3) || (soft_stride & 1))
		return LRPT_ERR_ARG;
	if (nrows < 2) return LRPT_OK;
	shard_quadrants_kernel<<<nrows - 1, 256, 0, (cudaStream_t)cuda_stream>>>(
		d_soft, soft_stride, d_q, q_stride, reinterpret_cast<const long long *>(d_base), d_ia, d_ib, npairs, d_k, d_same);
	return done(cudaGetLastError());
}

extern "C" int lrpt_shard_quadrants_oqpsk_device(const int8_t *d_soft, size_t soft_stride, const uint32_t *d_q, size_t q_stride,
                                                 const int64_t *d_base, int nrows, const int32_t *d_ia, const int32_t *d_ib, int npairs,
                                                 float half_substeps, int32_t *d_k, int32_t *d_same, void *cuda_stream)
{
	if (!d_soft || !d_q || !d_base || !d_ia || !d_ib || !d_k || !d_same || nrows < 1 || npairs < 0 || (q_stride & 3) || (soft_stride & 1))
		return LRPT_ERR_ARG;
	if (nrows < 2) return LRPT_OK;
	shard_quadrants_oqpsk_kernel<<<nrows - 1, 256, 0, (cudaStream_t)cuda_stream>>>(
		d_soft, soft_stride, d_q, q_stride, reinterpret_cast<const long long *>(d_base), d_ia, d_ib, npairs, half_substeps, d_k, d_same);
	return done(cudaGetLastError());
}

extern "C" int lrpt_shard_ranges_device(const uint32_t *d_q, size_t q_stride, const int32_t *d_count, const int64_t *d_base, int nrows,
                                        const int64_t *d_lo, const int64_t *d_hi, int32_t *d_start, int32_t *d_len, void *cuda_stream)
{
	if (!d_q || !d_count || !d_base || !d_lo || !d_hi || !d_start || !d_len || nrows < 1 || (q_stride & 3)) return LRPT_ERR_ARG;
	shard_ranges_kernel<<<(nrows + 127)/128, 128, 0, (cudaStream_t)cuda_stream>>>(
		d_q, q_stride, d_count, reinterpret_cast<const long long *>(d_base), nrows, reinterpret_cast<const long long *>(d_lo),
		reinterpret_cast<const long long *>(d_hi), d_start, d_len);
	return done(cudaGetLastError());
}

extern "C" int lrpt_shard_gather_device(const int8_t *d_soft, size_t soft_stride, int nrows, size_t max_len, const int32_t *d_start,
                                        const int32_t *d_len, const int64_t *d_off, const int32_t *d_turns, int8_t *d_out,
                                        void *cuda_stream)
{
	if (!d_soft || !d_start || !d_len || !d_off || !d_turns || !d_out || nrows < 1 || (soft_stride & 1)) return LRPT_ERR_ARG;
	if (!max_len) return LRPT_OK;
	size_t tiles = (max_len + 1023)/1024;                            /* 4 symbols per thread of a 256-thread block */
	if (tiles > 1024) tiles = 1024;
	for (int r0 = 0; r0 < nrows; r0 += 65535) {                      /* grid.y limit */
		const int nr = nrows - r0 < 65535 ? nrows - r0 : 65535;
		shard_gather_kernel<<<dim3((unsigned)tiles, (unsigned)nr), 256, 0, (cudaStream_t)cuda_stream>>>(
			d_soft, soft_stride, d_start, d_len, reinterpret_cast<const long long *>(d_off), d_turns, d_out, r0);
	}
	return done(cudaGetLastError());
}
