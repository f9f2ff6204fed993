/*
 * shard_run.cu -- ONE recording, time-sharded over the lanes of one GPU or of several, behind a single C call
 * (lrpt_sharded_process / lrpt_sharded_process_multi): what meteor_demod_b200/sharded.py::run_handoff does, for
 * hosts without Python (host/lrpt_demod --shard [--gpus N]). A client of the public ABI only: the chunks run as
 * the streams of ordinary batch handles (one per GPU), the join is csrc/shard_stitch.cu.
 *
 * The reference demodulates a recording as one sequential recurrence (demod.c:24-48, main.c:303-317);
 * its result cannot be reproduced bit for bit by anything that starts in the middle (DESIGN.md section
 * 1). This path is therefore STATISTICAL parity (Tier-S): the first two chunks are bit-exact, later ones
 * differ from the sequential run on a fraction of a percent of the symbols by more than one LSB --
 * about what separates the reference's own FMA and strict builds.
 *
 *   pass A  every chunk c demodulates W warm-up samples before its boundary B_c = W + c*C from power-on state
 *   pass B  ... then its C owned samples and V of overlap into the successor; on the overlap the
 *           quarter-turn k between neighbours is read off the paired symbols (a Costas loop locks with
 *           a k*90 degree ambiguity, pll.c:143-152), prefix-summed to K_c
 *   hand-off every chunk's end state (V samples past its successor's boundary), its Costas NCO turned
 *           back K_c quarter turns, becomes the START state of the successor
 *   pass C  every chunk again, C + V samples from B_c + V, now on a trajectory with W + C + V samples
 *           of history at the sequential run's lock point; chunk 1 continues chunk 0 exactly
 *   join    cut points mid-way between symbols, runs gathered in stream order
 *
 * Several GPUs (one host thread per device = "rank", consecutive runs of chunks, every rank holding only its time
 * slice of the recording plus the warm-up and overlap it over-reads): what crosses a rank boundary is exactly the
 * reference's state vector (pll.c:16-20,112, timing.c:13-14,43, agc.c:9-10, filter.h:5-11,
 * demod.c:54) of the rank's last chunk -- lrpt_export_states_device bytes, by ncclSend / ncclRecv over NVLink --
 * and that chunk's overlap symbols for the quadrant scan and the cut, also by ncclSend / ncclRecv; the quarter-turn
 * prefix and the symbol counts are integers the threads share in host memory. NCCL is loaded at run time
 * (libnccl.so.2) and only when more than one device is asked for. The result is byte-identical to the one-GPU run.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <vector>
#include "lrpt_b200.h"

namespace {

struct DevBuf {
	void *p = nullptr;
	~DevBuf() { if (p) cudaFree(p); }
	cudaError_t alloc(size_t n) { if (p) { cudaFree(p); p = nullptr; } return cudaMalloc(&p, n ? n : 1); }
	template <class T> T *as() const { return static_cast<T *>(p); }
};

/* Small device arrays (per-boundary / per-row tables of the joins) come out of one allocation per rank: a cudaMalloc +
 * cudaFree pair costs up to a millisecond next to multi-GB buffers, and the joins need some thirty of them. */
struct Arena {
	DevBuf buf;
	size_t cap = 0, used = 0;
	cudaError_t init(size_t bytes) { cap = bytes; used = 0; return buf.alloc(bytes); }
	void reset() { used = 0; }
	template <class T> T *take(size_t count)
	{
		const size_t at = (used + 255) & ~(size_t)255, need = count*sizeof(T);
		if (!buf.p || at + need > cap) return nullptr;
		used = at + need;
		return reinterpret_cast<T *>(buf.as<char>() + at);
	}
};

struct Handle {
	lrpt_demod_t *h = nullptr;
	~Handle() { if (h) lrpt_destroy(h); }
};

size_t symbol_capacity(size_t nsamples, const lrpt_params_t &p)
{
	return (size_t)((double)nsamples*((double)p.symrate/(double)p.samplerate)*1.02) + 64;
}

/* LRPT_SHARD_TIMING=1: wall time of every phase on stderr (device work is synchronised at each mark) */
struct Phases {
	bool on;
	std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
	explicit Phases(bool enable) : on(enable && getenv("LRPT_SHARD_TIMING") != nullptr) {}
	void mark(const char *what)
	{
		if (!on) return;
		cudaDeviceSynchronize();
		const auto n = std::chrono::steady_clock::now();
		fprintf(stderr, "lrpt_sharded_process: %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
		t = n;
	}
};

#define CK(x) do { if ((x) != cudaSuccess) return LRPT_ERR_CUDA; } while (0)
#define RC(x) do { const int rc_ = (x); if (rc_) return rc_; } while (0)

/* boundaries between consecutive rows [r0, r0 + n): k, agreement and cut per boundary (host vectors) */
int scan_rows(Arena &ar, const int8_t *d_soft, size_t soft_stride, const uint32_t *d_q, size_t q_stride, const int32_t *d_count,
              const int64_t *d_base, int n, const std::vector<int64_t> &target, std::vector<int32_t> &k,
              std::vector<float> &agree, std::vector<int64_t> &cut, float oqpsk_half = 0.0f)
{
	const int nb = n - 1;
	k.assign(nb > 0 ? nb : 0, 0); agree.assign(nb > 0 ? nb : 0, 0.0f); cut.assign(nb > 0 ? nb : 0, 0);
	if (nb <= 0) return LRPT_OK;
	ar.reset();
	int64_t *tgt = ar.take<int64_t>(nb), *dcut = ar.take<int64_t>(nb);
	int32_t *ia = ar.take<int32_t>(nb), *ib = ar.take<int32_t>(nb), *nav = ar.take<int32_t>(nb), *dk = ar.take<int32_t>(nb),
	        *same = ar.take<int32_t>(nb);
	if (!tgt || !dcut || !ia || !ib || !nav || !dk || !same) return LRPT_ERR_NOMEM;
	CK(cudaMemcpy(tgt, target.data(), 8*(size_t)nb, cudaMemcpyHostToDevice));
	RC(lrpt_shard_find_cuts_device(d_q, q_stride, d_count, d_base, n, tgt, dcut, ia, ib, nav, nullptr));
	std::vector<int32_t> navail(nb);
	CK(cudaMemcpy(navail.data(), nav, 4*(size_t)nb, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(cut.data(), dcut, 8*(size_t)nb, cudaMemcpyDeviceToHost));
	int npairs = INT_MAX;
	for (int v : navail) npairs = v < npairs ? v : npairs;
	if (oqpsk_half > 0.0f) {
		/* one symbol of the earlier row in reserve (the odd pairing reads a.I of the NEXT symbol), and every boundary
		 * needs a symbol of that row in front of its first pair (sharded.py::boundary_quadrants_oqpsk) */
		std::vector<int32_t> via(nb);
		CK(cudaMemcpy(via.data(), ia, 4*(size_t)nb, cudaMemcpyDeviceToHost));
		npairs = npairs > 0 ? npairs - 1 : 0;
		for (int v : via) if (v <= 0) npairs = 0;
	}
	if (npairs < 8) return LRPT_OK;                                 /* k = 0, agreement = 0: the caller reports it */
	if (oqpsk_half > 0.0f)
		RC(lrpt_shard_quadrants_oqpsk_device(d_soft, soft_stride, d_q, q_stride, d_base, n, ia, ib, npairs, oqpsk_half, dk, same, nullptr));
	else
		RC(lrpt_shard_quadrants_device(d_soft, soft_stride, d_q, q_stride, d_base, n, ia, ib, npairs, dk, same, nullptr));
	std::vector<int32_t> s(nb);
	CK(cudaMemcpy(k.data(), dk, 4*(size_t)nb, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(s.data(), same, 4*(size_t)nb, cudaMemcpyDeviceToHost));
	for (int b = 0; b < nb; b++) agree[b] = (float)s[b]/(float)npairs;
	return LRPT_OK;
}

/* ------------------------------------------------------------------ NCCL, loaded at run time ---- */

typedef struct ncclComm *ncclComm_t;
struct Nccl {
	void *lib = nullptr;
	int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
	int (*CommDestroy)(ncclComm_t) = nullptr;
	int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*GroupStart)(void) = nullptr;
	int (*GroupEnd)(void) = nullptr;
	bool load()
	{
		lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if (!lib) return false;
#define SYM(field, name) *(void **)(&field) = dlsym(lib, name)
		SYM(CommInitAll, "ncclCommInitAll"); SYM(CommDestroy, "ncclCommDestroy"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv");
		SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd");
#undef SYM
		return CommInitAll && CommDestroy && Send && Recv && GroupStart && GroupEnd;
	}
};
constexpr int NCCL_CHAR = 0;                                    /* ncclInt8 / ncclChar */

/* Communicators are kept between calls: ncclCommInitAll and the first exchange on a fresh communicator cost ~0.5 s,
 * more than the whole demodulation of a 2-GSample recording. One multi-GPU call at a time uses them (the mutex is held
 * for the whole call); a different device list replaces them; lrpt_sharded_release() frees them. */
struct CommCache {
	pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
	Nccl nccl;
	bool loaded = false;
	std::vector<int> devs;
	std::vector<ncclComm_t> comms;
	void drop()
	{
		for (ncclComm_t c : comms) if (c) nccl.CommDestroy(c);
		comms.clear(); devs.clear();
	}
	/* 0 on success; the caller holds mu */
	int get(const int *devices, int world)
	{
		if (!loaded) {
			if (!nccl.load()) { fprintf(stderr, "lrpt_sharded_process_multi: libnccl.so.2 not found\n"); return LRPT_ERR_CUDA; }
			loaded = true;
		}
		std::vector<int> want(devices, devices + world);
		if (want == devs && (int)comms.size() == world) return LRPT_OK;
		drop();
		comms.assign(world, nullptr);
		if (nccl.CommInitAll(comms.data(), world, devices)) { comms.clear(); return LRPT_ERR_CUDA; }
		devs = want;
		return LRPT_OK;
	}
};
CommCache g_comms;

/* What the rank threads share (host memory): integers only. */
struct Shared {
	int world = 1;
	pthread_barrier_t bar;
	pthread_mutex_t gate_mu = PTHREAD_MUTEX_INITIALIZER;        /* start gate: 0 = wait, 1 = go, -1 = a thread could not be created */
	pthread_cond_t gate_cv = PTHREAD_COND_INITIALIZER;
	int gate = 0;
	std::vector<int> rc, ksum, ksum2, aligned, launches;
	std::vector<float> agree_scan, agree_final;
	std::vector<long long> nsym_rank, first_lock_rank;          /* symbols a rank contributes; its first locked output symbol (-1) */
	long long head_lock = -1;                                   /* chunk 0's own first lock (exact) */
	Nccl *nccl = nullptr;
	std::vector<ncclComm_t> comms;
	bool failed() const { for (int v : rc) if (v) return true; return false; }
};

struct Job {
	const lrpt_params_t *params; const lrpt_shard_plan_t *plan;
	const uint8_t *raw; size_t nsamples;
	int8_t *soft; size_t cap;
	size_t M;                                                   /* chunks in the whole stream */
};

/* A row's loop state moved from the lock point it acquired to the one `turns` quarter turns back (the sequential run's).
 * QPSK: the Costas NCO alone, p_phase = (float)((double)p_phase - turns*pi/2) (pll.c:16, as lrpt_restore does). OQPSK
 * (sharded.py::turn_oqpsk_state): for an odd count the arms also change roles -- what was sampled as I at the pi crossing
 * (timing.c:47-50) is the new Q and vice versa -- so the timing NCO moves by pi, the dual-threshold state toggles and the
 * two remembered samples swap with the signs of the turn (demod.c:54, timing.c:13); half a turn flips both signs. */
void turn_state(lrpt_state_t &s, int turns, bool oqpsk)
{
	const int kk = turns & 3;
	s.p_phase = (float)((double)s.p_phase - (double)kk*1.57079632679489661923);
	if (!oqpsk) return;
	if (kk & 1) {
		const double pi = 3.14159265358979323846, s_q = kk == 1 ? 1.0 : -1.0;
		const float prev = s.t_prev, inph = s.oq_inphase;
		s.t_phase = (float)((double)s.t_phase + (s.t_dual_state == 1 ? pi : -pi));
		s.t_dual_state = 3 - s.t_dual_state;
		s.t_prev = (float)(s_q*(double)inph);
		s.oq_inphase = (float)(-s_q*(double)prev);
	} else if (kk == 2) {
		s.t_prev = -s.t_prev;
		s.oq_inphase = -s.oq_inphase;
	}
}

void split_rows(size_t M, int world, int rank, size_t &c0, size_t &c1)
{
	const size_t base = M/(size_t)world, extra = M%(size_t)world;
	c0 = (size_t)rank*base + ((size_t)rank < extra ? (size_t)rank : extra);
	c1 = c0 + base + ((size_t)rank < extra ? 1 : 0);
}

/* One rank: rows [c0, c1) of the stream on device `dev`. Every rank walks the same sequence of barriers and
 * NCCL exchanges; an error is recorded and the walk goes on with whatever it has, so nobody waits for ever. */
struct Rank {
	Job job; Shared *sh; int rank, dev;
	cudaStream_t st = nullptr;

	int xfer(const void *send, void *recv, size_t bytes)        /* my buffer -> next rank, previous rank -> my buffer */
	{
		if (sh->world < 2) return LRPT_OK;
		Nccl &n = *sh->nccl;
		int e = n.GroupStart();
		if (!e && send && rank + 1 < sh->world) e = n.Send(send, bytes, NCCL_CHAR, rank + 1, sh->comms[rank], st);
		if (!e && recv && rank > 0) e = n.Recv(recv, bytes, NCCL_CHAR, rank - 1, sh->comms[rank], st);
		const int e2 = n.GroupEnd();
		if (e || e2) return LRPT_ERR_CUDA;
		CK(cudaStreamSynchronize(st));
		return LRPT_OK;
	}
	void sync() { pthread_barrier_wait(&sh->bar); }

	int run()
	{
		const lrpt_params_t &params = *job.params;
		const size_t C = job.plan->chunk, W = job.plan->warm, V = job.plan->overlap, nsamples = job.nsamples;
		const size_t bytes = (size_t)params.bps/4;
		const long long L = params.interp_factor;
		const int world = sh->world;
		size_t c0, c1;
		split_rows(job.M, world, rank, c0, c1);
		const size_t M = c1 - c0;
		const bool first = rank == 0, last = rank == world - 1;
		/* OQPSK: timing sub-steps in half a symbol; 0 = QPSK join */
		const float oq_half = params.oqpsk ? (float)((double)params.samplerate*(double)params.interp_factor/(2.0*(double)params.symrate)) : 0.0f;
		int rc = LRPT_OK;
		const auto t_start = std::chrono::steady_clock::now();
		Phases ph(first);
		auto fail = [&](int code) { if (!rc) rc = code; sh->rc[rank] = rc; };
#define TRY(x) do { if (!rc) { const int rc_ = (x); if (rc_) fail(rc_); } } while (0)
#define TRYCU(x) do { if (!rc && (x) != cudaSuccess) fail(LRPT_ERR_CUDA); } while (0)

		TRYCU(cudaSetDevice(dev));
		TRYCU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
		lrpt_params_t p = params;
		p.nstreams = (int32_t)M; p.device = dev;
		Handle hd;
		TRY(lrpt_create(&hd.h, &p));
		lrpt_demod_t *h = hd.h;

		/* this rank's time slice on its device: samples [c0*C, (c1-1)*C + W + C + 2V), zero padded behind the stream */
		const size_t s0 = c0*C, span = (M - 1)*C + W + C + 2*V;
		const size_t have = s0 < nsamples ? (nsamples - s0 < span ? nsamples - s0 : span) : 0;
		const size_t n_row = C + V;
		const size_t cap_row = (symbol_capacity(W > n_row ? W : n_row, p) + 7)/8*8;
		const size_t soft_stride = 2*cap_row, q_stride = 4*cap_row;
		DevBuf d_raw, d_soft, d_q, d_nsym, d_cnt, d_base, d_states, d_pack, d_two_soft, d_two_q, d_two_cnt, d_two_base, d_st_io;
		Arena ar;
		TRYCU(ar.init(64*M + 65536));
		TRYCU(d_raw.alloc(span*bytes)); TRYCU(d_soft.alloc(M*soft_stride)); TRYCU(d_q.alloc(M*q_stride)); TRYCU(d_nsym.alloc(4*M));
		TRYCU(d_cnt.alloc(4*M)); TRYCU(d_base.alloc(8*M));
		/* a boundary row on its way to the next rank: symbols, sub-step indices, then count (int32) and base (int64) */
		const size_t pack_bytes = soft_stride + q_stride + 16;
		TRYCU(d_pack.alloc(2*pack_bytes));                          /* [0]: outgoing, [1]: incoming */
		TRYCU(d_two_soft.alloc(2*soft_stride)); TRYCU(d_two_q.alloc(2*q_stride)); TRYCU(d_two_cnt.alloc(8)); TRYCU(d_two_base.alloc(16));
		ph.mark("handle + device buffers");
		if (!rc && have < span) TRYCU(cudaMemsetAsync(d_raw.as<char>() + have*bytes, 0, (span - have)*bytes, st));
		if (!rc && have) TRYCU(cudaMemcpyAsync(d_raw.p, job.raw + s0*bytes, have*bytes, cudaMemcpyHostToDevice, st));
		TRYCU(cudaStreamSynchronize(st));
		ph.mark("H2D of the recording");
		if (h) TRY(lrpt_set_symbol_index_output(h, d_q.as<uint32_t>(), q_stride));
		const char *raw0 = d_raw.as<char>();
		std::vector<uint32_t> counts(M, 0);
		std::vector<int64_t> base(M, 0);
		auto run_pass = [&](size_t first_sample, size_t n) {        /* first_sample: offset within every row's own span */
			if (h) TRY(lrpt_process_batch_device(h, raw0 + first_sample*bytes, C*bytes, n, d_soft.as<int8_t>(), soft_stride, cap_row,
			                                     d_nsym.as<uint32_t>(), nullptr, 0, nullptr));
			if (h) TRY(lrpt_sync(h, nullptr));
			if (h) TRY(lrpt_get_counts(h, counts.data(), (int)M));
			for (size_t c = 0; c < M && !rc; c++) if (counts[c] > cap_row) fail(LRPT_ERR_CAP);   /* the stitch kernels trust the counts */
			for (size_t c = 0; c < M; c++) base[c] = (long long)((c0 + c)*C + first_sample)*L;
			TRYCU(cudaMemcpy(d_cnt.p, counts.data(), 4*M, cudaMemcpyHostToDevice));   /* uint32 counts < 2^31: read as int32 */
			TRYCU(cudaMemcpy(d_base.p, base.data(), 8*M, cudaMemcpyHostToDevice));
		};
		auto cut_target = [&](size_t c, size_t shift) { return (long long)(W + c*C + shift + (V/4 < 64 ? V/4 : 64))*L; };

		/* my last row -> the next rank, the previous rank's last row -> rows (prev | my first) as a two-row problem.
		 * Returns k, agreement and cut of the boundary in front of my first row (rank > 0). */
		auto boundary_with_prev = [&](size_t shift, int32_t &k_prev, float &agree_prev, int64_t &cut_prev) {
			k_prev = 0; agree_prev = 1.0f; cut_prev = -1;
			if (world < 2) return;
			char *out = d_pack.as<char>(), *in = out + pack_bytes;
			if (!rc && !last) {
				const size_t r = M - 1;
				TRYCU(cudaMemcpyAsync(out, d_soft.as<char>() + r*soft_stride, soft_stride, cudaMemcpyDeviceToDevice, st));
				TRYCU(cudaMemcpyAsync(out + soft_stride, d_q.as<char>() + r*q_stride, q_stride, cudaMemcpyDeviceToDevice, st));
				TRYCU(cudaMemcpyAsync(out + soft_stride + q_stride, d_cnt.as<char>() + 4*r, 4, cudaMemcpyDeviceToDevice, st));
				TRYCU(cudaMemcpyAsync(out + soft_stride + q_stride + 8, d_base.as<char>() + 8*r, 8, cudaMemcpyDeviceToDevice, st));
			}
			const int e = xfer(last ? nullptr : out, first ? nullptr : in, pack_bytes);   /* always entered: the peers wait for it */
			if (e) fail(e);
			if (first || rc) return;
			TRYCU(cudaMemcpyAsync(d_two_soft.p, in, soft_stride, cudaMemcpyDeviceToDevice, st));
			TRYCU(cudaMemcpyAsync(d_two_soft.as<char>() + soft_stride, d_soft.p, soft_stride, cudaMemcpyDeviceToDevice, st));
			TRYCU(cudaMemcpyAsync(d_two_q.p, in + soft_stride, q_stride, cudaMemcpyDeviceToDevice, st));
			TRYCU(cudaMemcpyAsync(d_two_q.as<char>() + q_stride, d_q.p, q_stride, cudaMemcpyDeviceToDevice, st));
			TRYCU(cudaMemcpyAsync(d_two_cnt.p, in + soft_stride + q_stride, 4, cudaMemcpyDeviceToDevice, st));
			TRYCU(cudaMemcpyAsync(d_two_cnt.as<char>() + 4, d_cnt.p, 4, cudaMemcpyDeviceToDevice, st));
			TRYCU(cudaMemcpyAsync(d_two_base.p, in + soft_stride + q_stride + 8, 8, cudaMemcpyDeviceToDevice, st));
			TRYCU(cudaMemcpyAsync(d_two_base.as<char>() + 8, d_base.p, 8, cudaMemcpyDeviceToDevice, st));
			TRYCU(cudaStreamSynchronize(st));
			std::vector<int64_t> tgt(1, cut_target(c0, shift)), cut;
			std::vector<int32_t> k;
			std::vector<float> ag;
			TRY(scan_rows(ar, d_two_soft.as<int8_t>(), soft_stride, d_two_q.as<uint32_t>(), q_stride, d_two_cnt.as<int32_t>(),
			              d_two_base.as<int64_t>(), 2, tgt, k, ag, cut, oq_half));
			if (!rc) { k_prev = k[0]; agree_prev = ag[0]; cut_prev = cut[0]; }
		};

		/* ---- optional: chunks after the stream's first start with their Costas NCO at a coarse carrier estimate instead
		 * of sweeping to the carrier (sharded.py::GpuEngine::seed_carrier: p_freq, everything else power-on) ---- */
		if (job.plan->seed_nfft && h && !rc) {
			const int nfft = (int)job.plan->seed_nfft;
			const size_t total = lrpt_states_size(h);
			DevBuf d_cfo, d_pow;
			std::vector<double> cfo(M, 0.0);
			std::vector<char> st0(total);
			TRYCU(d_cfo.alloc(8*M)); TRYCU(d_pow.alloc(total));
			TRY(lrpt_carrier_estimate_device(&p, raw0, C*bytes, (int)M, nfft, 4000.0, d_cfo.as<double>(), nullptr));
			TRYCU(cudaMemcpy(cfo.data(), d_cfo.p, 8*M, cudaMemcpyDeviceToHost));
			TRY(lrpt_export_states_device(h, d_pow.p, total, nullptr));
			TRY(lrpt_sync(h, nullptr));
			TRYCU(cudaMemcpy(st0.data(), d_pow.p, total, cudaMemcpyDeviceToHost));
			if (!rc) {
				lrpt_state_t *s0 = reinterpret_cast<lrpt_state_t *>(st0.data());
				const double per = (double)((long long)params.symrate*(params.oqpsk ? 2 : 1));
				for (size_t c = first ? 1 : 0; c < M; c++)
					s0[c].p_freq = (float)(2.0*3.14159265358979323846*cfo[c]/per);   /* acquire.py::p_freq_for; main.c:250 backwards */
				TRYCU(cudaMemcpy(d_pow.p, st0.data(), total, cudaMemcpyHostToDevice));
				TRYCU(cudaDeviceSynchronize());
				TRY(lrpt_import_states_device(h, d_pow.p, total, 1, nullptr));
				TRY(lrpt_sync(h, nullptr));
			}
			ph.mark("carrier estimates");
		}
		/* ---- pass A: warm-up; chunk 0's warm-up symbols open the output (they are the sequential run's) ---- */
		std::vector<int8_t> head;                                   /* rank 0: chunk 0's own symbols (passes A and B) */
		if (W) {
			run_pass(0, W);
			if (first && !rc) {
				head.resize(2*(size_t)counts[0]);
				TRYCU(cudaMemcpy(head.data(), d_soft.p, head.size(), cudaMemcpyDeviceToHost));
			}
		}
		ph.mark("pass A (warm-up)");
		/* ---- pass B: owned + overlap, from the live state ---- */
		run_pass(W, n_row);
		ph.mark("pass B");
		std::vector<int64_t> target(M > 1 ? M - 1 : 0), cut;
		std::vector<int32_t> k;
		std::vector<float> agree;
		for (size_t c = 1; c < M; c++) target[c - 1] = cut_target(c0 + c, 0);
		TRY(scan_rows(ar, d_soft.as<int8_t>(), soft_stride, d_q.as<uint32_t>(), q_stride, d_cnt.as<int32_t>(), d_base.as<int64_t>(), (int)M,
		              target, k, agree, cut, oq_half));
		if (rc) { k.assign(M > 1 ? M - 1 : 0, 0); agree.assign(k.size(), 0.0f); }
		int32_t k_prev = 0; float agree_prev = 1.0f; int64_t cut_prev = -1;
		boundary_with_prev(0, k_prev, agree_prev, cut_prev);
		float min_scan = agree_prev;
		for (float a : agree) min_scan = a < min_scan ? a : min_scan;
		sh->agree_scan[rank] = min_scan;
		{
			int s = k_prev;
			for (int32_t v : k) s += v;
			sh->ksum[rank] = s & 3;
		}
		sync();                                                     /* everybody's quarter-turn sums are in */
		std::vector<int32_t> K(M, 0);
		{
			int kf = k_prev;
			for (int j = 0; j < rank; j++) kf += sh->ksum[j];
			K[0] = kf & 3;
			for (size_t c = 1; c < M; c++) K[c] = (K[c - 1] + k[c - 1]) & 3;
		}
		if (first && !rc) {
			/* chunk 0's pass-B symbols continue the head, up to the end of the stream */
			const int64_t l = -1, hh = (int64_t)nsamples*L - 1;
			int32_t s_0 = 0, n_0 = 0;
			ar.reset();
			int64_t *lo = ar.take<int64_t>(1), *hi = ar.take<int64_t>(1);
			int32_t *stt = ar.take<int32_t>(1), *ln = ar.take<int32_t>(1);
			if (!lo || !hi || !stt || !ln) fail(LRPT_ERR_NOMEM);
			TRYCU(cudaMemcpy(lo, &l, 8, cudaMemcpyHostToDevice)); TRYCU(cudaMemcpy(hi, &hh, 8, cudaMemcpyHostToDevice));
			TRY(lrpt_shard_ranges_device(d_q.as<uint32_t>(), q_stride, d_cnt.as<int32_t>(), d_base.as<int64_t>(), 1, lo, hi, stt, ln, nullptr));
			TRYCU(cudaMemcpy(&s_0, stt, 4, cudaMemcpyDeviceToHost)); TRYCU(cudaMemcpy(&n_0, ln, 4, cudaMemcpyDeviceToHost));
			if (!rc) {
				const size_t at = head.size();
				head.resize(at + 2*(size_t)n_0);
				TRYCU(cudaMemcpy(head.data() + at, d_soft.as<int8_t>() + 2*(size_t)s_0, 2*(size_t)n_0, cudaMemcpyDeviceToHost));
			}
			lrpt_status_t s;
			memset(&s, 0, sizeof(s));
			s.first_lock_symbol = -1;
			if (h) TRY(lrpt_status(h, 0, &s));
			if (!rc) sh->head_lock = s.first_lock_symbol;            /* chunk 0 counts its symbols as the stream does */
		}
		ph.mark("quadrant scan + chunk 0");

		/* ---- hand-off: row c starts pass C from row c-1's end state, its Costas NCO turned back K[c-1] quarter turns
		 * (p_phase = (float)((double)p_phase - K*pi/2), as lrpt_restore does; pll.c:16); across a rank boundary that
		 * one state travels by ncclSend / ncclRecv ---- */
		std::vector<long long> nsym_start(M, 0);                    /* symbol count row c inherits at the start of pass C */
		{
			const size_t total = h ? lrpt_states_size(h) : 0, sb = sizeof(lrpt_state_t);
			const size_t hb = M && total ? (total - M*sb)/M : 0;    /* delay line bytes per row */
			const size_t one = sb + hb;
			TRYCU(d_states.alloc(total)); TRYCU(d_st_io.alloc(2*one));
			if (h) TRY(lrpt_export_states_device(h, d_states.p, total, nullptr));
			if (h) TRY(lrpt_sync(h, nullptr));
			std::vector<char> cur(total), nxt(total), io(2*one);
			TRYCU(cudaMemcpy(cur.data(), d_states.p, total, cudaMemcpyDeviceToHost));
			lrpt_state_t *sc = reinterpret_cast<lrpt_state_t *>(cur.data()), *sn = reinterpret_cast<lrpt_state_t *>(nxt.data());
			if (!rc) {
				for (size_t c = 0; c < M; c++) turn_state(sc[c], K[c], params.oqpsk != 0);
				memcpy(io.data(), &sc[M - 1], sb);                  /* my last row's turned state -> the next rank */
				memcpy(io.data() + sb, cur.data() + M*sb + (M - 1)*hb, hb);
				TRYCU(cudaMemcpy(d_st_io.p, io.data(), one, cudaMemcpyHostToDevice));
				TRYCU(cudaDeviceSynchronize());
			}
			const int e = xfer(last ? nullptr : d_st_io.p, first ? nullptr : d_st_io.as<char>() + one, one);
			if (e) fail(e);
			if (!rc) {
				if (!first) {
					TRYCU(cudaMemcpy(io.data() + one, d_st_io.as<char>() + one, one, cudaMemcpyDeviceToHost));
					memcpy(&sn[0], io.data() + one, sb);
					memcpy(nxt.data() + M*sb, io.data() + one + sb, hb);
				} else {
					sn[0] = sc[0];                                  /* chunk 0 has no predecessor; its pass C is not used */
					memcpy(nxt.data() + M*sb, cur.data() + M*sb, hb);
				}
				for (size_t c = 1; c < M; c++) {
					sn[c] = sc[c - 1];
					memcpy(nxt.data() + M*sb + c*hb, cur.data() + M*sb + (c - 1)*hb, hb);
				}
				for (size_t c = 0; c < M; c++) nsym_start[c] = sn[c].nsymbols;
				TRYCU(cudaMemcpy(d_states.p, nxt.data(), total, cudaMemcpyHostToDevice));
				TRYCU(cudaDeviceSynchronize());                     /* pageable upload on the legacy stream: not ordered with the handle's stream */
				TRY(lrpt_import_states_device(h, d_states.p, total, 1, nullptr));
				TRY(lrpt_sync(h, nullptr));
			}
		}
		ph.mark("state hand-off");

		/* ---- pass C, then the join. Rank 0: rows 0 and 1 are one exact trajectory, so its table is rows 1.. with
		 * nothing cut off the front of row 1; other ranks: all their rows, the first one cut against the previous rank's last ---- */
		run_pass(W + V, n_row);
		ph.mark("pass C");
		const size_t skip = first ? 1 : 0;
		const int n = (int)(M - skip);
		const int8_t *soft1 = d_soft.as<int8_t>() + skip*soft_stride;
		const uint32_t *q1 = reinterpret_cast<const uint32_t *>(d_q.as<char>() + skip*q_stride);
		const int32_t *cnt1 = d_cnt.as<int32_t>() + skip;
		const int64_t *base1 = d_base.as<int64_t>() + skip;
		std::vector<int64_t> target2(n > 1 ? n - 1 : 0), cut2;
		std::vector<int32_t> k2;
		std::vector<float> agree2;
		for (int b = 0; b + 1 < n; b++) target2[b] = cut_target(c0 + skip + (size_t)b + 1, V);
		TRY(scan_rows(ar, soft1, soft_stride, q1, q_stride, cnt1, base1, n, target2, k2, agree2, cut2, oq_half));
		if (rc) { k2.assign(n > 1 ? n - 1 : 0, 0); agree2.assign(k2.size(), 0.0f); cut2.assign(k2.size(), 0); }
		int32_t k_prev2 = 0; float agree_prev2 = 1.0f; int64_t cut_prev2 = -1;
		boundary_with_prev(V, k_prev2, agree_prev2, cut_prev2);
		float min_final = agree_prev2;
		for (float a : agree2) min_final = a < min_final ? a : min_final;
		sh->agree_final[rank] = min_final;
		int al = k_prev2 ? 0 : 1;
		{
			int s = k_prev2;
			for (int32_t v : k2) { s += v; if (v) al = 0; }
			sh->ksum2[rank] = s & 3;
		}
		sh->aligned[rank] = al;
		/* where my last row ends: the cut against the next rank's first row, placed by MY row's symbols alone
		 * (mid-way between two of them around the target), so both sides agree without talking */
		int64_t hi_last = (int64_t)nsamples*L - 1;                  /* last rank: nothing from the zero padding */
		if (!last && !rc && n >= 1) {
			const size_t r = M - 1;
			TRYCU(cudaMemcpy(d_two_q.p, d_q.as<char>() + r*q_stride, q_stride, cudaMemcpyDeviceToDevice));
			TRYCU(cudaMemcpy(d_two_q.as<char>() + q_stride, d_q.as<char>() + r*q_stride, q_stride, cudaMemcpyDeviceToDevice));
			const int32_t two_c[2] = { (int32_t)counts[r], (int32_t)counts[r] };
			const int64_t two_b[2] = { base[r], base[r] };
			TRYCU(cudaMemcpy(d_two_cnt.p, two_c, 8, cudaMemcpyHostToDevice));
			TRYCU(cudaMemcpy(d_two_base.p, two_b, 16, cudaMemcpyHostToDevice));
			const int64_t t = cut_target(c1, V);
			ar.reset();
			int64_t *tgt = ar.take<int64_t>(1), *dcut = ar.take<int64_t>(1);
			int32_t *ia = ar.take<int32_t>(1), *ib = ar.take<int32_t>(1), *nav = ar.take<int32_t>(1);
			if (!tgt || !dcut || !ia || !ib || !nav) fail(LRPT_ERR_NOMEM);
			TRYCU(cudaMemcpy(tgt, &t, 8, cudaMemcpyHostToDevice));
			TRY(lrpt_shard_find_cuts_device(d_two_q.as<uint32_t>(), q_stride, d_two_cnt.as<int32_t>(), d_two_base.as<int64_t>(), 2,
			                                tgt, dcut, ia, ib, nav, nullptr));
			TRYCU(cudaMemcpy(&hi_last, dcut, 8, cudaMemcpyDeviceToHost));
		}
		sync();                                                     /* second round of quarter-turn sums */
		std::vector<int32_t> turns(n > 0 ? n : 0, 0);
		std::vector<int64_t> lo(n > 0 ? n : 0, -1), hi(n > 0 ? n : 0, LLONG_MAX);
		if (n > 0) {
			int kf = k_prev2;
			for (int j = 0; j < rank; j++) kf += sh->ksum2[j];
			turns[0] = kf & 3;
			if (!first) lo[0] = cut_prev2;
			for (int b = 0; b + 1 < n; b++) {
				turns[b + 1] = (turns[b] + k2[b]) & 3;
				hi[b] = cut2[b]; lo[b + 1] = cut2[b];
			}
			hi[n - 1] = hi_last;
		}
		DevBuf dout;
		std::vector<int32_t> len(n > 0 ? n : 0, 0), start(n > 0 ? n : 0, 0);
		std::vector<int64_t> off(n > 0 ? n : 0, 0);
		size_t total = 0, longest = 0;
		if (n > 0) {
			ar.reset();
			int64_t *dlo = ar.take<int64_t>(n), *dhi = ar.take<int64_t>(n), *doff = ar.take<int64_t>(n);
			int32_t *dst = ar.take<int32_t>(n), *dln = ar.take<int32_t>(n), *dturn = ar.take<int32_t>(n);
			if (!dlo || !dhi || !doff || !dst || !dln || !dturn) fail(LRPT_ERR_NOMEM);
			TRYCU(cudaMemcpy(dlo, lo.data(), 8*(size_t)n, cudaMemcpyHostToDevice));
			TRYCU(cudaMemcpy(dhi, hi.data(), 8*(size_t)n, cudaMemcpyHostToDevice));
			TRY(lrpt_shard_ranges_device(q1, q_stride, cnt1, base1, n, dlo, dhi, dst, dln, nullptr));
			TRYCU(cudaMemcpy(len.data(), dln, 4*(size_t)n, cudaMemcpyDeviceToHost));
			TRYCU(cudaMemcpy(start.data(), dst, 4*(size_t)n, cudaMemcpyDeviceToHost));
			if (!rc) for (int b = 0; b < n; b++) { off[b] = (int64_t)total; total += (size_t)len[b]; longest = (size_t)len[b] > longest ? (size_t)len[b] : longest; }
			TRYCU(dout.alloc(2*total));
			TRYCU(cudaMemcpy(doff, off.data(), 8*(size_t)n, cudaMemcpyHostToDevice));
			TRYCU(cudaMemcpy(dturn, turns.data(), 4*(size_t)n, cudaMemcpyHostToDevice));
			TRY(lrpt_shard_gather_device(soft1, soft_stride, n, longest, dst, dln, doff, dturn, dout.as<int8_t>(), nullptr));
		}
		const size_t out_n = head.size()/2;                         /* rank 0: chunk 0's own symbols come first */
		sh->nsym_rank[rank] = rc ? 0 : (long long)(out_n + total);

		/* Stream-wide first lock (main.c:312 gates the output on pll_did_lock_once()). Chunk 0 may not have locked by the
		 * end of its passes -- a recording that starts before the signal is up -- so the answer is the first OUTPUT symbol
		 * that came from a loop which had locked once: row c's pass C inherits its predecessor's flag and symbol count
		 * (exact for row 1, which continues chunk 0 bit for bit), so the row-local index at which it had locked is
		 * first_lock_symbol - inherited count. Here: relative to this rank's own share. */
		long long my_lock = -1;
		if (!rc && n > 0 && h) {
			const size_t tot_states = lrpt_states_size(h);
			TRY(lrpt_export_states_device(h, d_states.p, tot_states, nullptr));
			TRY(lrpt_sync(h, nullptr));
			std::vector<lrpt_state_t> fin(M);
			TRYCU(cudaMemcpy(fin.data(), d_states.p, M*sizeof(lrpt_state_t), cudaMemcpyDeviceToHost));
			for (int b = 0; b < n && my_lock < 0 && !rc; b++) {
				const lrpt_state_t &s = fin[(size_t)b + skip];
				if (s.first_lock_symbol < 0 || len[b] <= 0) continue;
				long long j = s.first_lock_symbol - nsym_start[(size_t)b + skip];   /* < 0: locked before this pass began */
				if (j < start[b]) j = start[b];
				if (j < (long long)start[b] + len[b]) my_lock = (long long)out_n + off[b] + (j - start[b]);
			}
		}
		sh->first_lock_rank[rank] = my_lock;
		sh->launches[rank] = h ? (int)lrpt_launch_count(h) : 0;
		ph.mark("join (scan, ranges, gather)");
		sync();                                                     /* every rank's symbol count is known: output offsets */
		long long at = 0, all = 0;
		for (int j = 0; j < world; j++) { if (j < rank) at += sh->nsym_rank[j]; all += sh->nsym_rank[j]; }
		if (!rc && (size_t)all > job.cap) fail(LRPT_ERR_CAP);
		if (!rc && !sh->failed()) {
			if (!head.empty()) memcpy(job.soft + 2*(size_t)at, head.data(), head.size());
			if (total) TRYCU(cudaMemcpy(job.soft + 2*((size_t)at + out_n), dout.p, 2*total, cudaMemcpyDeviceToHost));
		}
		ph.mark("D2H of the symbols");
		if (st) cudaStreamDestroy(st);
		sh->rc[rank] = rc;
		if (getenv("LRPT_SHARD_TIMING"))
			fprintf(stderr, "lrpt_sharded_process: rank %d of %d, rows %zu, %.2f ms from thread start to results\n", rank, world, M,
			        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
		return rc;
#undef TRY
#undef TRYCU
	}
};

void *rank_main(void *arg)
{
	Rank *r = static_cast<Rank *>(arg);
	pthread_mutex_lock(&r->sh->gate_mu);
	while (!r->sh->gate) pthread_cond_wait(&r->sh->gate_cv, &r->sh->gate_mu);
	const int go = r->sh->gate;
	pthread_mutex_unlock(&r->sh->gate_mu);
	if (go > 0) r->run();
	return nullptr;
}

} // namespace

extern "C" int lrpt_sharded_process_multi(const lrpt_params_t *params, const lrpt_shard_plan_t *plan, const void *raw_iq,
                                          size_t nsamples, int8_t *soft, size_t cap, size_t *nsym, lrpt_shard_report_t *rep,
                                          const int *devices, int ndev)
{
	if (!params || !plan || !raw_iq || !soft || !nsym || ndev < 1 || ndev > 64 || (ndev > 1 && !devices)) return LRPT_ERR_ARG;
	const size_t C = plan->chunk, W = plan->warm, V = plan->overlap;
	if (!C || !V || (C & 7) || (W & 7) || (V & 7)) return LRPT_ERR_ARG;             /* 16-byte aligned rows for every sample format */
	if (params->bps != 8 && params->bps != 16 && params->bps != 32) return LRPT_ERR_ARG;
	if (plan->seed_nfft) {
		const uint64_t f = plan->seed_nfft;
		if (f < 256 || f > 16384 || (f & (f - 1)) || f > W + C) return LRPT_ERR_ARG;
	}
	const size_t M = nsamples > W ? (nsamples - W + C - 1)/C : 1;
	if (M > (size_t)INT_MAX/2) return LRPT_ERR_ARG;
	/* the rows' sub-step indices (sample*interp + sub-step, uint32 side output) must not wrap: the joins
	 * rely on their ascending order */
	if ((double)(W > C + V ? W : C + V)*(double)params->interp_factor >= 4294967296.0) return LRPT_ERR_ARG;
	lrpt_shard_report_t r;
	memset(&r, 0, sizeof(r));
	r.nchunks = (int32_t)M; r.min_agreement_scan = r.min_agreement_final = 1.0f; r.first_lock_symbol = -1; r.aligned = 1;
	*nsym = 0;
	const int dev0 = devices ? devices[0] : params->device;

	if (M < 2) {                                                    /* one chunk: the sequential run itself, exact */
		CK(cudaSetDevice(dev0));
		lrpt_params_t p = *params;
		p.nstreams = 1; p.device = dev0;
		Handle hd;
		RC(lrpt_create(&hd.h, &p));
		size_t n = 0; long long fl = -1;
		RC(lrpt_process(hd.h, raw_iq, nsamples, soft, cap, &n, &fl));
		*nsym = n; r.first_lock_symbol = fl; r.launches = (int32_t)lrpt_launch_count(hd.h);
		if (rep) *rep = r;
		return LRPT_OK;
	}
	/* every rank needs a chunk, the one holding chunk 0 two (chunk 1 continues chunk 0 exactly) */
	int world = ndev;
	while (world > 1 && M/(size_t)world < 2) world--;

	Shared sh;
	sh.world = world;
	sh.rc.assign(world, 0); sh.ksum.assign(world, 0); sh.ksum2.assign(world, 0); sh.aligned.assign(world, 1); sh.launches.assign(world, 0);
	sh.agree_scan.assign(world, 1.0f); sh.agree_final.assign(world, 1.0f);
	sh.nsym_rank.assign(world, 0); sh.first_lock_rank.assign(world, -1);
	const auto t_call = std::chrono::steady_clock::now();
	if (world > 1) {
		pthread_mutex_lock(&g_comms.mu);
		const int rc = g_comms.get(devices, world);
		if (rc) { pthread_mutex_unlock(&g_comms.mu); return rc; }
		sh.nccl = &g_comms.nccl;
		sh.comms = g_comms.comms;
	}
	pthread_barrier_init(&sh.bar, nullptr, (unsigned)world);
	std::vector<Rank> ranks(world);
	std::vector<pthread_t> tids(world);
	for (int i = 0; i < world; i++) {
		ranks[i].job = Job{ params, plan, static_cast<const uint8_t *>(raw_iq), nsamples, soft, cap, M };
		ranks[i].sh = &sh; ranks[i].rank = i; ranks[i].dev = devices ? devices[i] : params->device;
	}
	/* the ranks meet at barriers sized for `world`: nobody starts before every thread exists */
	int started = 1;
	for (int i = 1; i < world; i++, started++) if (pthread_create(&tids[i], nullptr, rank_main, &ranks[i])) break;
	pthread_mutex_lock(&sh.gate_mu);
	sh.gate = started == world ? 1 : -1;
	pthread_cond_broadcast(&sh.gate_cv);
	pthread_mutex_unlock(&sh.gate_mu);
	if (started == world) ranks[0].run();
	for (int i = 1; i < started; i++) pthread_join(tids[i], nullptr);
	pthread_barrier_destroy(&sh.bar);
	if (world > 1) {
		if (sh.failed() || started != world) g_comms.drop();         /* a failed exchange may have left them unusable */
		pthread_mutex_unlock(&g_comms.mu);
	}
	if (getenv("LRPT_SHARD_TIMING"))
		fprintf(stderr, "lrpt_sharded_process: %d device(s), %.2f ms in the call so far (threads joined, buffers freed)\n", world,
		        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count());
	if (started != world) return LRPT_ERR_NOMEM;
	for (int i = 0; i < world; i++) if (sh.rc[i]) return sh.rc[i];

	long long total = 0, lock = sh.head_lock;
	for (int i = 0; i < world; i++) {
		if (lock < 0 && sh.first_lock_rank[i] >= 0) lock = total + sh.first_lock_rank[i];
		total += sh.nsym_rank[i];
		r.min_agreement_scan = sh.agree_scan[i] < r.min_agreement_scan ? sh.agree_scan[i] : r.min_agreement_scan;
		r.min_agreement_final = sh.agree_final[i] < r.min_agreement_final ? sh.agree_final[i] : r.min_agreement_final;
		if (!sh.aligned[i]) r.aligned = 0;
		r.launches += sh.launches[i];
	}
	r.first_lock_symbol = lock;
	*nsym = (size_t)total;
	r.launches /= world;                                            /* launches per GPU: the three passes */
	if (rep) *rep = r;
	return LRPT_OK;
}

extern "C" void lrpt_sharded_release(void)
{
	pthread_mutex_lock(&g_comms.mu);
	g_comms.drop();
	pthread_mutex_unlock(&g_comms.mu);
}

extern "C" int lrpt_sharded_process(const lrpt_params_t *params, const lrpt_shard_plan_t *plan, const void *raw_iq,
                                    size_t nsamples, int8_t *soft, size_t cap, size_t *nsym, lrpt_shard_report_t *rep)
{
	if (!params) return LRPT_ERR_ARG;
	return lrpt_sharded_process_multi(params, plan, raw_iq, nsamples, soft, cap, nsym, rep, &params->device, 1);
}
