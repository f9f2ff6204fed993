/*
 * lrpt_api.cu -- the C ABI declared in include/lrpt_b200.h.
 *
 * Thin host layer: derives the loop constants (lrpt_params.c), owns the device
 * copies of taps / per-stream state / delay lines, pipelines host<->device copies
 * for the HOST-buffer entry points, and launches the kernels. There is no CPU
 * implementation of the DSP behind this file.
 */
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>
#include "kernels.h"
#include "lrpt_b200.h"
#include "lrpt_internal.h"

using namespace lrpt;

struct lrpt_demod {
	lrpt_params_t p;
	lrpt_consts_t c;
	lrpt_state_t  s0;
	std::vector<float> taps;
	int H;                          /* taps-1 */
	int kernel;                     /* resolved LRPT_KERNEL_* */
	float        *d_taps   = nullptr;
	lrpt_state_t *d_states = nullptr;
	float2       *d_hist   = nullptr;
	uint32_t     *d_nsym   = nullptr;   /* [nstreams] last launch */
	uint32_t     *d_off    = nullptr;   /* [nstreams] append cursors (host-buffer path) */
	unsigned long long *d_fallbacks = nullptr;   /* spec kernel: FIR evaluations done by the recurrence warp */
	uint32_t     *h_counts = nullptr;   /* pinned [nstreams] */
	lrpt_state_t *d_init   = nullptr;   /* [nstreams] power-on states, source of resets */
	lrpt_state_t *d_snap   = nullptr;   /* [nstreams] lrpt_snapshot */
	float2       *d_snap_hist = nullptr;
	lrpt_state_t *h_state  = nullptr;   /* pinned scratch, one state */
	cudaStream_t  stream = nullptr, copy_stream = nullptr, out_stream = nullptr;
	cudaEvent_t   ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_cnt[2] = {nullptr, nullptr};
	cudaEvent_t   ev_order = nullptr;   /* lrpt_stream_wait / lrpt_stream_release */
	uint32_t     *d_mm = nullptr;       /* [2][2] min/max append cursor after a slab */
	uint32_t     *h_mm = nullptr;       /* pinned copy */
	/* staging for the host-buffer entry points (grown on demand) */
	void   *d_raw[2] = {nullptr, nullptr}; size_t d_raw_bytes = 0;
	int8_t *d_soft = nullptr; size_t d_soft_bytes = 0;
	float  *d_symf = nullptr; size_t d_symf_bytes = 0;
	uint32_t *d_symq_user = nullptr; size_t symq_stride_user = 0;   /* optional side output, device path */
	unsigned long long launches = 0;
	char err[256];
};

static int fail(lrpt_demod *h, int code, const char *fmt, ...)
{
	if (h) {
		va_list ap; va_start(ap, fmt);
		vsnprintf(h->err, sizeof(h->err), fmt, ap);
		va_end(ap);
	}
	return code;
}

#define CU(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return fail(h, LRPT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

static size_t bytes_per_sample(const lrpt_demod *h) { return (size_t)h->p.bps/4; }   /* I+Q */

static size_t round_up(size_t v, size_t a) { return (v + a - 1)/a*a; }

extern "C" int lrpt_abi_version(void) { return LRPT_ABI_VERSION; }

extern "C" const char *lrpt_strerror(int code)
{
	switch (code) {
		case LRPT_OK: return "ok";
		case LRPT_ERR_ARG: return "bad argument or unsupported configuration";
		case LRPT_ERR_CUDA: return "CUDA failure";
		case LRPT_ERR_NOMEM: return "out of memory";
		case LRPT_ERR_CAP: return "symbol capacity exceeded";
		case LRPT_ERR_STATE: return "state blob does not match the handle";
		default: return "unknown error";
	}
}

extern "C" const char *lrpt_last_error(const lrpt_demod_t *h) { return h ? h->err : "null handle"; }

static int reset_on(lrpt_demod *h, cudaStream_t st)
{
	CU(h, cudaMemcpyAsync(h->d_states, h->d_init, sizeof(lrpt_state_t)*(size_t)h->p.nstreams,
	                      cudaMemcpyDeviceToDevice, st));
	if (h->H > 0)
		CU(h, cudaMemsetAsync(h->d_hist, 0, sizeof(float2)*(size_t)h->H*h->p.nstreams, st));   /* calloc, filter.c:16 */
	CU(h, cudaMemsetAsync(h->d_nsym, 0, sizeof(uint32_t)*h->p.nstreams, st));
	return LRPT_OK;
}

static int upload_initial_state(lrpt_demod *h)
{
	std::vector<lrpt_state_t> init((size_t)h->p.nstreams, h->s0);
	/* on the handle's own (non-blocking) stream: a pageable cudaMemcpy on the legacy stream is not ordered with it */
	CU(h, cudaMemcpyAsync(h->d_init, init.data(), sizeof(lrpt_state_t)*init.size(), cudaMemcpyHostToDevice, h->stream));
	int rc = reset_on(h, h->stream);
	if (rc) return rc;
	CU(h, cudaStreamSynchronize(h->stream));                       /* init[] may go out of scope now */
	memset(h->h_counts, 0, sizeof(uint32_t)*h->p.nstreams);
	return LRPT_OK;
}

extern "C" int lrpt_create(lrpt_demod_t **out, const lrpt_params_t *p)
{
	if (!out || !p || p->nstreams < 1) return LRPT_ERR_ARG;
	*out = nullptr;
	lrpt_demod *h = new (std::nothrow) lrpt_demod();
	if (!h) return LRPT_ERR_NOMEM;
	h->err[0] = 0;
	h->p = *p;
	const int taps = 2*p->rrc_order + 1;
	if (p->rrc_order < 0 || p->rrc_order > LRPT_MAX_ORDER || p->interp_factor < 1 ||
	    p->interp_factor > LRPT_MAX_INTERP) { delete h; return LRPT_ERR_ARG; }
	h->taps.resize((size_t)taps*p->interp_factor);
	int rc = lrpt_derive(p, &h->c, &h->s0, h->taps.data());
	if (rc) { delete h; return rc; }
	h->H = taps - 1;

	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev < 1 || p->device < 0 || p->device >= ndev) {
		fprintf(stderr, "lrpt_create: no usable CUDA device %d (%s); there is no CPU fallback\n",
		        p->device, e == cudaSuccess ? "not present" : cudaGetErrorString(e));
		delete h; return LRPT_ERR_CUDA;
	}
#define CUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	fprintf(stderr, "lrpt_create: %s: %s\n", #call, cudaGetErrorString(e_)); lrpt_destroy(h); return LRPT_ERR_CUDA; } } while (0)
	CUC(cudaSetDevice(p->device));
	CUC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
	CUC(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
	CUC(cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking));
	for (int i = 0; i < 2; i++) {
		CUC(cudaEventCreateWithFlags(&h->ev_copy[i], cudaEventDisableTiming));
		CUC(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
		CUC(cudaEventCreateWithFlags(&h->ev_cnt[i], cudaEventDisableTiming));
	}
	CUC(cudaMalloc(&h->d_mm, 4*sizeof(uint32_t)));
	CUC(cudaMallocHost(&h->h_mm, 4*sizeof(uint32_t)));
	CUC(cudaMalloc(&h->d_taps, sizeof(float)*h->taps.size()));
	CUC(cudaMalloc(&h->d_states, sizeof(lrpt_state_t)*p->nstreams));
	CUC(cudaMalloc(&h->d_init, sizeof(lrpt_state_t)*p->nstreams));
	CUC(cudaMalloc(&h->d_hist, sizeof(float2)*(size_t)(h->H > 0 ? h->H : 1)*p->nstreams));
	CUC(cudaMalloc(&h->d_nsym, sizeof(uint32_t)*p->nstreams));
	CUC(cudaMalloc(&h->d_off, sizeof(uint32_t)*p->nstreams));
	CUC(cudaMalloc(&h->d_fallbacks, sizeof(unsigned long long)));
	CUC(cudaMemset(h->d_fallbacks, 0, sizeof(unsigned long long)));
	CUC(cudaMallocHost(&h->h_counts, sizeof(uint32_t)*p->nstreams));
	CUC(cudaMallocHost(&h->h_state, sizeof(lrpt_state_t)));
	CUC(cudaMemcpyAsync(h->d_taps, h->taps.data(), sizeof(float)*h->taps.size(), cudaMemcpyHostToDevice, h->stream));   /* synchronised in upload_initial_state */
#undef CUC
	/* Kernel choice. Explicit requests must be supported by that kernel. AUTO goes by streams per SM:
	 * few streams -> the warp-specialised kernels (shortest time per symbol of one stream: all-phase
	 * FIR below ~16 streams per SM, speculative FIR up to ~36), many streams -> the lane kernel
	 * (least work per symbol; needs many warps per SM to hide its latency). All are bit-identical. */
	int want = p->kernel;
	if (want == LRPT_KERNEL_AUTO) {
		int sms = 1;
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
		const double per_sm = (double)p->nstreams/(double)(sms > 0 ? sms : 1);
		if (per_sm > 36.0 && lane_supported(h->c)) want = LRPT_KERNEL_LANE;
		else if (per_sm > 8.0 && spec_supported(h->c)) want = LRPT_KERNEL_SPEC;   /* ws gives 5 of its 12 FIR warps to the split recurrence */
		else if (ws_supported(h->c)) want = LRPT_KERNEL_WS;
		else if (lane_supported(h->c)) want = LRPT_KERNEL_LANE;
		else want = LRPT_KERNEL_SIMPLE;
	}
	const bool ok = want == LRPT_KERNEL_SIMPLE || (want == LRPT_KERNEL_WS && ws_supported(h->c)) ||
	                (want == LRPT_KERNEL_SPEC && spec_supported(h->c)) || (want == LRPT_KERNEL_LANE && lane_supported(h->c));
	if (!ok) {
		fprintf(stderr, "lrpt_create: configuration not supported by the requested kernel (%d)\n", want);
		lrpt_destroy(h); return LRPT_ERR_ARG;
	}
	cudaError_t pe = cudaSuccess;
	if (want == LRPT_KERNEL_WS) pe = ws_prepare(p->device);
	else if (want == LRPT_KERNEL_SPEC) pe = spec_prepare(p->device);
	else if (want == LRPT_KERNEL_LANE) pe = lane_prepare(p->device);
	if (pe != cudaSuccess) {
		fprintf(stderr, "lrpt_create: kernel setup: %s\n", cudaGetErrorString(pe));
		lrpt_destroy(h); return LRPT_ERR_CUDA;
	}
	h->kernel = want;
	rc = upload_initial_state(h);
	if (rc) { lrpt_destroy(h); return rc; }
	*out = h;
	return LRPT_OK;
}

extern "C" void lrpt_destroy(lrpt_demod_t *h)
{
	if (!h) return;
	cudaSetDevice(h->p.device);
	if (h->stream) cudaStreamSynchronize(h->stream);
	if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
	if (h->out_stream) cudaStreamSynchronize(h->out_stream);
	cudaFree(h->d_mm);
	if (h->h_mm) cudaFreeHost(h->h_mm);
	cudaFree(h->d_taps); cudaFree(h->d_states); cudaFree(h->d_init); cudaFree(h->d_snap); cudaFree(h->d_snap_hist); cudaFree(h->d_hist); cudaFree(h->d_nsym); cudaFree(h->d_off); cudaFree(h->d_fallbacks);
	cudaFree(h->d_raw[0]); cudaFree(h->d_raw[1]); cudaFree(h->d_soft); cudaFree(h->d_symf);
	if (h->h_counts) cudaFreeHost(h->h_counts);
	if (h->h_state) cudaFreeHost(h->h_state);
	for (int i = 0; i < 2; i++) {
		if (h->ev_copy[i]) cudaEventDestroy(h->ev_copy[i]);
		if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
		if (h->ev_cnt[i]) cudaEventDestroy(h->ev_cnt[i]);
	}
	if (h->ev_order) cudaEventDestroy(h->ev_order);
	if (h->out_stream) cudaStreamDestroy(h->out_stream);
	if (h->stream) cudaStreamDestroy(h->stream);
	if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
	delete h;
}

extern "C" int lrpt_reset(lrpt_demod_t *h)
{
	if (!h) return LRPT_ERR_ARG;
	CU(h, cudaSetDevice(h->p.device));
	CU(h, cudaDeviceSynchronize());          /* nothing of this handle may still be in flight */
	int rc = reset_on(h, h->stream);
	if (rc) return rc;
	CU(h, cudaStreamSynchronize(h->stream));
	return LRPT_OK;
}

extern "C" int lrpt_reset_async(lrpt_demod_t *h, void *cuda_stream)
{
	if (!h) return LRPT_ERR_ARG;
	CU(h, cudaSetDevice(h->p.device));
	return reset_on(h, cuda_stream ? (cudaStream_t)cuda_stream : h->stream);
}

/* ------------------------------------------------------------------ launch -- */

static int launch(lrpt_demod *h, LaunchArgs &a, cudaStream_t st)
{
	a.c = &h->c; a.d_taps = h->d_taps; a.d_states = h->d_states; a.d_hist = h->d_hist;
	cudaError_t e;
	if (h->kernel == LRPT_KERNEL_SPEC) {
		int n = 0;
		e = launch_spec(a, st, &n, h->d_fallbacks);
		h->launches += (unsigned long long)n;
	} else if (h->kernel == LRPT_KERNEL_LANE) {
		int n = 0;
		e = launch_lane(a, st, &n);
		h->launches += (unsigned long long)n;
	} else if (h->kernel == LRPT_KERNEL_WS) {
		int n = 0;
		e = launch_ws(a, st, &n);
		h->launches += (unsigned long long)n;
	} else {
		e = launch_simple(a, st);
		h->launches += 1;
	}
	if (e != cudaSuccess) return fail(h, LRPT_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(e));
	return LRPT_OK;
}

extern "C" int lrpt_process_batch_device(lrpt_demod_t *h, const void *d_raw_iq, size_t raw_stride,
                                         size_t nsamples, int8_t *d_soft, size_t soft_stride, size_t cap,
                                         uint32_t *d_nsym, float *d_sym_f32, size_t symf_stride,
                                         void *cuda_stream)
{
	if (!h || !d_raw_iq || !d_soft) return LRPT_ERR_ARG;
	if (((uintptr_t)d_raw_iq | (uintptr_t)d_soft | (uintptr_t)d_sym_f32 | raw_stride | soft_stride | symf_stride) & 15)
		return fail(h, LRPT_ERR_ARG, "device pointers and strides must be 16-byte aligned");
	if (cap > 0xffffffffu || nsamples > ((size_t)1 << 40)) return LRPT_ERR_ARG;
	CU(h, cudaSetDevice(h->p.device));
	cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
	LaunchArgs a{};
	a.d_raw = d_raw_iq; a.raw_stride = raw_stride; a.nsamples = nsamples;
	a.d_soft = d_soft; a.soft_stride = soft_stride; a.cap = (unsigned)cap;
	a.d_symf = d_sym_f32; a.symf_stride = symf_stride;
	a.d_symq = h->d_symq_user; a.symq_stride = h->symq_stride_user;
	a.d_nsym = nullptr; a.d_out_off = h->d_off;              /* cursor doubles as the per-call count */
	a.first_stream = 0; a.nstreams = h->p.nstreams;
	CU(h, cudaMemsetAsync(h->d_off, 0, sizeof(uint32_t)*h->p.nstreams, st));
	if (nsamples) {
		int rc = launch(h, a, st);
		if (rc) return rc;
	}
	if (d_nsym)
		CU(h, cudaMemcpyAsync(d_nsym, h->d_off, sizeof(uint32_t)*h->p.nstreams, cudaMemcpyDeviceToDevice, st));
	CU(h, cudaMemcpyAsync(h->h_counts, h->d_off, sizeof(uint32_t)*h->p.nstreams, cudaMemcpyDeviceToHost, st));
	return LRPT_OK;
}

extern "C" int lrpt_set_symbol_index_output(lrpt_demod_t *h, uint32_t *d_index, size_t stride)
{
	if (!h || (((uintptr_t)d_index | stride) & 3)) return LRPT_ERR_ARG;
	h->d_symq_user = d_index; h->symq_stride_user = stride;
	return LRPT_OK;
}

extern "C" int lrpt_sync(lrpt_demod_t *h, void *cuda_stream)
{
	if (!h) return LRPT_ERR_ARG;
	CU(h, cudaSetDevice(h->p.device));
	CU(h, cudaStreamSynchronize(cuda_stream ? (cudaStream_t)cuda_stream : h->stream));
	return LRPT_OK;
}

/* Stream ordering without a host synchronisation: the handle's own stream is created non-blocking, so work a
 * caller enqueued on ANOTHER stream (the legacy default stream included: pass NULL) is not ordered with it. */
static int order_streams(lrpt_demod *h, cudaStream_t first, cudaStream_t then)
{
	CU(h, cudaSetDevice(h->p.device));
	if (!h->ev_order) CU(h, cudaEventCreateWithFlags(&h->ev_order, cudaEventDisableTiming));
	CU(h, cudaEventRecord(h->ev_order, first));
	CU(h, cudaStreamWaitEvent(then, h->ev_order, 0));
	return LRPT_OK;
}

extern "C" int lrpt_stream_wait(lrpt_demod_t *h, void *producer_stream)
{
	if (!h) return LRPT_ERR_ARG;
	return order_streams(h, (cudaStream_t)producer_stream, h->stream);
}

extern "C" int lrpt_stream_release(lrpt_demod_t *h, void *consumer_stream)
{
	if (!h) return LRPT_ERR_ARG;
	return order_streams(h, h->stream, (cudaStream_t)consumer_stream);
}

extern "C" int lrpt_get_counts(lrpt_demod_t *h, uint32_t *nsym, int nstreams)
{
	if (!h || !nsym || nstreams < 0 || nstreams > h->p.nstreams) return LRPT_ERR_ARG;
	memcpy(nsym, h->h_counts, sizeof(uint32_t)*(size_t)nstreams);
	return LRPT_OK;
}

/* Grow a device staging buffer. */
static int ensure(lrpt_demod *h, void **ptr, size_t *have, size_t need)
{
	if (*have >= need) return LRPT_OK;
	if (*ptr) { cudaFree(*ptr); *ptr = nullptr; *have = 0; }
	cudaError_t e = cudaMalloc(ptr, need);
	if (e != cudaSuccess) return fail(h, LRPT_ERR_NOMEM, "cudaMalloc(%zu): %s", need, cudaGetErrorString(e));
	*have = need;
	return LRPT_OK;
}

/* smallest and largest append cursor of a batch: bounds the columns a slab added to d_soft */
__global__ void cursor_minmax_kernel(const uint32_t *off, int n, uint32_t *mm)
{
	__shared__ uint32_t smin[32], smax[32];
	uint32_t lo = 0xffffffffu, hi = 0u;
	for (int i = threadIdx.x; i < n; i += blockDim.x) { const uint32_t v = off[i]; lo = min(lo, v); hi = max(hi, v); }
	for (int d = 16; d > 0; d >>= 1) {
		lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
		hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
	}
	if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
	__syncthreads();
	if (threadIdx.x == 0) {
		for (unsigned w = 1; w < blockDim.x/32; w++) { lo = min(lo, smin[w]); hi = max(hi, smax[w]); }
		mm[0] = lo; mm[1] = hi;
	}
}

/*
 * Host-buffer batch over streams [first, first+count): time slabs are copied to the
 * device on copy_stream while the previous slab is demodulated on stream (two raw
 * staging buffers); symbols are appended in a device buffer through per-stream
 * cursors and copied back once at the end.
 */
static int process_host(lrpt_demod *h, int first, int count, const void *raw_iq, size_t raw_stride,
                        size_t nsamples, int8_t *soft, size_t soft_stride, size_t cap, uint32_t *nsym,
                        float *sym_f32, size_t symf_stride)
{
	if (!h || (!raw_iq && nsamples) || !soft || cap > 0xffffffffu) return LRPT_ERR_ARG;
	CU(h, cudaSetDevice(h->p.device));
	const size_t bpsm = bytes_per_sample(h);
	/* slab: about 256 MiB of raw input across the batch, at least 16 Ki samples per stream: small enough
	 * that the first copy (nothing to overlap with) and the last kernel are short, large enough that a
	 * launch still runs hundreds of tiles */
	size_t slab = ((size_t)256 << 20)/(bpsm*(size_t)count);
	const size_t slab_min = h->kernel == LRPT_KERNEL_LANE ? 4096 : 16384;   /* lane: short launches cost little */
	if (slab < slab_min) slab = slab_min;
	slab = round_up(slab, 4096);
	if (slab > nsamples) slab = round_up(nsamples ? nsamples : 1, 16);
	const size_t d_raw_pitch = round_up(slab*bpsm, 256);
	const size_t d_soft_pitch = round_up(2*cap + 16, 256);
	const size_t d_symf_pitch = round_up(8*cap + 16, 256);
	int rc;
	for (int i = 0; i < 2; i++) {
		size_t have = h->d_raw_bytes;
		if ((rc = ensure(h, &h->d_raw[i], &have, d_raw_pitch*count))) return rc;
		if (i == 1) h->d_raw_bytes = have;
	}
	if ((rc = ensure(h, (void **)&h->d_soft, &h->d_soft_bytes, d_soft_pitch*count))) return rc;
	if (sym_f32 && (rc = ensure(h, (void **)&h->d_symf, &h->d_symf_bytes, d_symf_pitch*count))) return rc;
	CU(h, cudaMemsetAsync(h->d_off, 0, sizeof(uint32_t)*count, h->stream));

	/* Symbols go back to the host slab by slab, on a third stream, while the next slabs are still being
	 * copied in and demodulated (the link is full duplex): after slab j every stream's cursor lies in
	 * [min_j, max_j], so the columns [min_{j-1}, max_j) of d_soft hold everything slab j appended.
	 * Columns past a stream's own cursor are copied too and are rewritten by the next slab's copy. */
	size_t done = 0; int j = 0;
	uint32_t col_from = 0;
	auto flush_slab = [&](int b) -> int {                           /* symbols of the slab that used buffer b */
		CU(h, cudaEventSynchronize(h->ev_cnt[b]));
		const uint32_t lo = h->h_mm[2*b], hi = h->h_mm[2*b + 1] < cap ? h->h_mm[2*b + 1] : (uint32_t)cap;
		if (hi > col_from) {
			CU(h, cudaStreamWaitEvent(h->out_stream, h->ev_cnt[b], 0));
			CU(h, cudaMemcpy2DAsync(soft + 2*(size_t)col_from, soft_stride, h->d_soft + 2*(size_t)col_from, d_soft_pitch,
			                        2*(size_t)(hi - col_from), count, cudaMemcpyDeviceToHost, h->out_stream));
		}
		if (lo > col_from) col_from = lo < cap ? lo : (uint32_t)cap;
		return LRPT_OK;
	};
	while (done < nsamples) {
		const size_t n = (nsamples - done < slab) ? nsamples - done : slab;
		const int b = j & 1;
		if (j >= 2) CU(h, cudaStreamWaitEvent(h->copy_stream, h->ev_done[b], 0));   /* buffer free again */
		if (count == 1)
			CU(h, cudaMemcpyAsync(h->d_raw[b], (const char *)raw_iq + done*bpsm, n*bpsm,
			                      cudaMemcpyHostToDevice, h->copy_stream));
		else
			CU(h, cudaMemcpy2DAsync(h->d_raw[b], d_raw_pitch, (const char *)raw_iq + done*bpsm, raw_stride,
			                        n*bpsm, count, cudaMemcpyHostToDevice, h->copy_stream));
		CU(h, cudaEventRecord(h->ev_copy[b], h->copy_stream));
		CU(h, cudaStreamWaitEvent(h->stream, h->ev_copy[b], 0));
		if (j >= 1 && (rc = flush_slab(b ^ 1))) return rc;          /* previous slab out while this one comes in;
		                                                               also frees ev_cnt[b]/h_mm[b] (slab j-2 was flushed at j-1) */
		LaunchArgs a{};
		a.d_raw = h->d_raw[b]; a.raw_stride = d_raw_pitch; a.nsamples = n;
		a.d_soft = h->d_soft; a.soft_stride = d_soft_pitch; a.cap = (unsigned)cap;
		a.d_symf = sym_f32 ? h->d_symf : nullptr; a.symf_stride = d_symf_pitch;
		a.d_nsym = nullptr; a.d_out_off = h->d_off;
		a.first_stream = first; a.nstreams = count;
		if ((rc = launch(h, a, h->stream))) return rc;
		CU(h, cudaEventRecord(h->ev_done[b], h->stream));
		cursor_minmax_kernel<<<1, 1024, 0, h->stream>>>(h->d_off, count, h->d_mm + 2*b);
		CU(h, cudaMemcpyAsync(h->h_mm + 2*b, h->d_mm + 2*b, 2*sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
		CU(h, cudaEventRecord(h->ev_cnt[b], h->stream));
		done += n; j++;
	}
	if (j >= 1 && (rc = flush_slab((j - 1) & 1))) return rc;
	CU(h, cudaMemcpyAsync(h->h_counts, h->d_off, sizeof(uint32_t)*count, cudaMemcpyDeviceToHost, h->stream));
	CU(h, cudaStreamSynchronize(h->stream));
	uint32_t most = 0; int over = 0;
	for (int s = 0; s < count; s++) {
		if (nsym) nsym[s] = h->h_counts[s];
		uint32_t stored = h->h_counts[s];
		if (stored > cap) { stored = (uint32_t)cap; over = 1; }
		if (stored > most) most = stored;
	}
	if (most && sym_f32)
		CU(h, cudaMemcpy2DAsync(sym_f32, symf_stride, h->d_symf, d_symf_pitch, 8*(size_t)most, count,
		                        cudaMemcpyDeviceToHost, h->stream));
	CU(h, cudaStreamSynchronize(h->stream));
	CU(h, cudaStreamSynchronize(h->out_stream));
	return over ? fail(h, LRPT_ERR_CAP, "a stream produced more than cap=%zu symbols", cap) : LRPT_OK;
}

extern "C" int lrpt_process_batch(lrpt_demod_t *h, const void *raw_iq, size_t raw_stride, size_t nsamples,
                                  int8_t *soft, size_t soft_stride, size_t cap, uint32_t *nsym,
                                  float *sym_f32, size_t symf_stride)
{
	if (!h) return LRPT_ERR_ARG;
	if (h->p.nstreams > 1 && (soft_stride < 2*cap || (sym_f32 && symf_stride < 8*cap)))
		return fail(h, LRPT_ERR_ARG, "output strides smaller than cap symbols");
	return process_host(h, 0, h->p.nstreams, raw_iq, raw_stride, nsamples, soft, soft_stride, cap, nsym,
	                    sym_f32, symf_stride);
}

extern "C" int lrpt_process(lrpt_demod_t *h, const void *raw_iq, size_t nsamples, int8_t *soft, size_t cap,
                            size_t *nsym, long long *first_lock_symbol)
{
	if (!h) return LRPT_ERR_ARG;
	uint32_t n32 = 0;
	int rc = process_host(h, 0, 1, raw_iq, 0, nsamples, soft, 2*cap, cap, &n32, nullptr, 0);
	if (nsym) *nsym = n32;
	if (first_lock_symbol && (rc == LRPT_OK || rc == LRPT_ERR_CAP)) {
		lrpt_status_t st;
		int rc2 = lrpt_status(h, 0, &st);
		if (rc2) return rc2;
		*first_lock_symbol = st.first_lock_symbol;
	}
	return rc;
}

/* ------------------------------------------------------------ status/state -- */

static int fetch_state(lrpt_demod *h, int stream, lrpt_state_t *s)
{
	if (!h || stream < 0 || stream >= h->p.nstreams) return LRPT_ERR_ARG;
	CU(h, cudaSetDevice(h->p.device));
	CU(h, cudaStreamSynchronize(h->stream));
	CU(h, cudaMemcpy(s, h->d_states + stream, sizeof(*s), cudaMemcpyDeviceToHost));
	return LRPT_OK;
}

extern "C" int lrpt_status(lrpt_demod_t *h, int stream, lrpt_status_t *st)
{
	lrpt_state_t s;
	if (!st) return LRPT_ERR_ARG;
	int rc = fetch_state(h, stream, &s);
	if (rc) return rc;
	st->pll_freq = s.p_freq; st->mm_omega = s.t_freq; st->agc_gain = s.agc_gain;
	st->locked = s.p_locked; st->locked_once = s.p_locked_once;
	st->nsamples = s.nsamples; st->nsymbols = s.nsymbols; st->first_lock_symbol = s.first_lock_symbol;
	return LRPT_OK;
}

extern "C" size_t lrpt_state_size(const lrpt_demod_t *h)
{
	return h ? sizeof(lrpt_state_t) + sizeof(float2)*(size_t)h->H : 0;
}

extern "C" int lrpt_export_state(lrpt_demod_t *h, int stream, void *buf, size_t *len)
{
	if (!h || !len) return LRPT_ERR_ARG;
	const size_t need = lrpt_state_size(h);
	if (!buf || *len < need) { *len = need; return buf ? LRPT_ERR_ARG : LRPT_OK; }
	int rc = fetch_state(h, stream, (lrpt_state_t *)buf);
	if (rc) return rc;
	if (h->H > 0)
		CU(h, cudaMemcpy((char *)buf + sizeof(lrpt_state_t), h->d_hist + (size_t)stream*h->H,
		                 sizeof(float2)*(size_t)h->H, cudaMemcpyDeviceToHost));
	*len = need;
	return LRPT_OK;
}

extern "C" int lrpt_import_state(lrpt_demod_t *h, int stream, const void *buf, size_t len)
{
	if (!h || !buf || stream < 0 || stream >= h->p.nstreams) return LRPT_ERR_ARG;
	const lrpt_state_t *s = (const lrpt_state_t *)buf;
	if (len != lrpt_state_size(h) || s->magic != LRPT_STATE_MAGIC || s->taps != (uint32_t)h->c.taps)
		return fail(h, LRPT_ERR_STATE, "state blob: size/magic/taps mismatch");
	if (h->p.bps != 32) {
		/* the delay line of an 8/16-bit handle holds converted input samples (wavfile.c:58-66): integers of
		 * that range. The lane kernel keeps them in the input's own type, so anything else is refused. */
		const float *hv = (const float *)((const char *)buf + sizeof(lrpt_state_t));
		const float lo = h->p.bps == 8 ? -128.0f : -32768.0f, hi = h->p.bps == 8 ? 127.0f : 32767.0f;
		for (int i = 0; i < 2*h->H; i++)
			if (!(hv[i] >= lo && hv[i] <= hi) || hv[i] != (float)(int)hv[i])
				return fail(h, LRPT_ERR_STATE, "state blob: delay line holds a value no %d-bit input produces", h->p.bps);
	}
	CU(h, cudaSetDevice(h->p.device));
	CU(h, cudaMemcpyAsync(h->d_states + stream, s, sizeof(*s), cudaMemcpyHostToDevice, h->stream));
	if (h->H > 0)
		CU(h, cudaMemcpyAsync(h->d_hist + (size_t)stream*h->H, (const char *)buf + sizeof(lrpt_state_t),
		                      sizeof(float2)*(size_t)h->H, cudaMemcpyHostToDevice, h->stream));
	CU(h, cudaStreamSynchronize(h->stream));                       /* the caller's buffer is free again */
	return LRPT_OK;
}

extern "C" size_t lrpt_states_size(const lrpt_demod_t *h)
{
	return h ? (size_t)h->p.nstreams*lrpt_state_size(h) : 0;
}

extern "C" int lrpt_export_states_device(lrpt_demod_t *h, void *d_buf, size_t len, void *cuda_stream)
{
	if (!h || !d_buf || len != lrpt_states_size(h)) return LRPT_ERR_ARG;
	CU(h, cudaSetDevice(h->p.device));
	cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
	const size_t ns = (size_t)h->p.nstreams, sb = sizeof(lrpt_state_t)*ns;
	CU(h, cudaMemcpyAsync(d_buf, h->d_states, sb, cudaMemcpyDeviceToDevice, st));
	if (h->H > 0)
		CU(h, cudaMemcpyAsync((char *)d_buf + sb, h->d_hist, sizeof(float2)*(size_t)h->H*ns, cudaMemcpyDeviceToDevice, st));
	return LRPT_OK;
}

extern "C" int lrpt_import_states_device(lrpt_demod_t *h, const void *d_buf, size_t len, int check, void *cuda_stream)
{
	if (!h || !d_buf || len != lrpt_states_size(h)) return LRPT_ERR_ARG;
	CU(h, cudaSetDevice(h->p.device));
	cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
	if (check) {
		CU(h, cudaMemcpyAsync(h->h_state, d_buf, sizeof(lrpt_state_t), cudaMemcpyDeviceToHost, st));
		CU(h, cudaStreamSynchronize(st));
		if (h->h_state->magic != LRPT_STATE_MAGIC || h->h_state->taps != (uint32_t)h->c.taps)
			return fail(h, LRPT_ERR_STATE, "state buffer: magic/taps mismatch");
	}
	const size_t ns = (size_t)h->p.nstreams, sb = sizeof(lrpt_state_t)*ns;
	CU(h, cudaMemcpyAsync(h->d_states, d_buf, sb, cudaMemcpyDeviceToDevice, st));
	if (h->H > 0)
		CU(h, cudaMemcpyAsync(h->d_hist, (const char *)d_buf + sb, sizeof(float2)*(size_t)h->H*ns, cudaMemcpyDeviceToDevice, st));
	return LRPT_OK;
}

extern "C" int lrpt_snapshot(lrpt_demod_t *h)
{
	if (!h) return LRPT_ERR_ARG;
	CU(h, cudaSetDevice(h->p.device));
	CU(h, cudaDeviceSynchronize());
	const size_t ns = (size_t)h->p.nstreams, nh = (size_t)(h->H > 0 ? h->H : 1)*ns;
	if (!h->d_snap) {
		CU(h, cudaMalloc(&h->d_snap, sizeof(lrpt_state_t)*ns));
		CU(h, cudaMalloc(&h->d_snap_hist, sizeof(float2)*nh));
	}
	/* device-to-device copies do not block the host and the legacy stream is not ordered with the handle's
	 * non-blocking stream: issue them on that stream */
	CU(h, cudaMemcpyAsync(h->d_snap, h->d_states, sizeof(lrpt_state_t)*ns, cudaMemcpyDeviceToDevice, h->stream));
	CU(h, cudaMemcpyAsync(h->d_snap_hist, h->d_hist, sizeof(float2)*nh, cudaMemcpyDeviceToDevice, h->stream));
	CU(h, cudaStreamSynchronize(h->stream));
	return LRPT_OK;
}

extern "C" int lrpt_restore(lrpt_demod_t *h, const int32_t *quarter_turns)
{
	if (!h || !h->d_snap) return LRPT_ERR_ARG;
	CU(h, cudaSetDevice(h->p.device));
	CU(h, cudaDeviceSynchronize());
	const size_t ns = (size_t)h->p.nstreams, nh = (size_t)(h->H > 0 ? h->H : 1)*ns;
	CU(h, cudaMemcpyAsync(h->d_hist, h->d_snap_hist, sizeof(float2)*nh, cudaMemcpyDeviceToDevice, h->stream));
	if (!quarter_turns) {
		CU(h, cudaMemcpyAsync(h->d_states, h->d_snap, sizeof(lrpt_state_t)*ns, cudaMemcpyDeviceToDevice, h->stream));
		CU(h, cudaStreamSynchronize(h->stream));
		return LRPT_OK;
	}
	std::vector<lrpt_state_t> st(ns);
	CU(h, cudaMemcpyAsync(st.data(), h->d_snap, sizeof(lrpt_state_t)*ns, cudaMemcpyDeviceToHost, h->stream));
	CU(h, cudaStreamSynchronize(h->stream));
	for (size_t s = 0; s < ns; s++)
		st[s].p_phase = (float)((double)st[s].p_phase - (double)(quarter_turns[s] & 3)*1.57079632679489661923);
	CU(h, cudaMemcpyAsync(h->d_states, st.data(), sizeof(lrpt_state_t)*ns, cudaMemcpyHostToDevice, h->stream));
	CU(h, cudaStreamSynchronize(h->stream));
	return LRPT_OK;
}

/* ------------------------------------------------------------ host memory ---- */

extern "C" void *lrpt_alloc_host(size_t bytes)
{
	void *p = nullptr;
	if (!bytes || cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	return p;
}

extern "C" void lrpt_free_host(void *p) { if (p) cudaFreeHost(p); }

extern "C" int lrpt_pin_host(void *p, size_t bytes)
{
	if (!p || !bytes) return LRPT_ERR_ARG;
	if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return LRPT_ERR_CUDA; }
	return LRPT_OK;
}

extern "C" int lrpt_unpin_host(void *p)
{
	if (!p) return LRPT_ERR_ARG;
	if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return LRPT_ERR_CUDA; }
	return LRPT_OK;
}

/* ----------------------------------------------------------- introspection -- */

extern "C" int lrpt_describe(const lrpt_params_t *p, lrpt_state_t *initial_state, float *taps, int taps_cap,
                             float loop_consts[7], float lut[32])
{
	if (!p) return LRPT_ERR_ARG;
	if (p->rrc_order < 0 || p->rrc_order > LRPT_MAX_ORDER || p->interp_factor < 1 ||
	    p->interp_factor > LRPT_MAX_INTERP) return LRPT_ERR_ARG;
	const int n = (2*p->rrc_order + 1)*p->interp_factor;
	std::vector<float> h((size_t)n);
	lrpt_consts_t c; lrpt_state_t s0;
	int rc = lrpt_derive(p, &c, &s0, h.data());
	if (rc) return rc;
	if (initial_state) *initial_state = s0;
	if (taps && taps_cap >= n) memcpy(taps, h.data(), sizeof(float)*(size_t)n);
	if (loop_consts) {
		loop_consts[0] = c.t_center; loop_consts[1] = c.t_maxdev; loop_consts[2] = c.t_alpha;
		loop_consts[3] = c.t_beta; loop_consts[4] = c.p_alpha; loop_consts[5] = c.p_beta; loop_consts[6] = c.p_fmax;
	}
	if (lut) memcpy(lut, c.lut_tanh, sizeof(float)*32);
	return n;
}

extern "C" int lrpt_get_taps(const lrpt_demod_t *h, float *dst, int cap)
{
	if (!h) return LRPT_ERR_ARG;
	const int n = (int)h->taps.size();
	if (dst && cap >= n) memcpy(dst, h->taps.data(), sizeof(float)*(size_t)n);
	return n;
}

extern "C" int lrpt_get_tanh_lut(const lrpt_demod_t *h, float dst[32])
{
	if (!h || !dst) return LRPT_ERR_ARG;
	memcpy(dst, h->c.lut_tanh, sizeof(float)*32);
	return LRPT_OK;
}

extern "C" unsigned long long lrpt_launch_count(const lrpt_demod_t *h) { return h ? h->launches : 0; }

extern "C" const char *lrpt_kernel_name(const lrpt_demod_t *h)
{
	return !h ? "" : h->kernel == LRPT_KERNEL_WS ? "ws" : h->kernel == LRPT_KERNEL_SPEC ? "spec" :
	       h->kernel == LRPT_KERNEL_LANE ? "lane" : "simple";
}

extern "C" unsigned long long lrpt_fir_fallbacks(lrpt_demod_t *h)
{
	unsigned long long v = 0;
	if (!h) return 0;
	cudaSetDevice(h->p.device);
	cudaDeviceSynchronize();
	cudaMemcpy(&v, h->d_fallbacks, sizeof(v), cudaMemcpyDeviceToHost);
	return v;
}
