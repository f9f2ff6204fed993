/*
 * shard_run.cu -- ONE recording, time-sharded over the lanes of one GPU, behind a single C call
 * (lrpt_sharded_process): what meteor_demod_b200/sharded.py::run_handoff does, for hosts without
 * Python (host/lrpt_demod --shards). A client of the public ABI only: the chunks run as the streams
 * of an ordinary batch handle, the join is csrc/shard_stitch.cu.
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
 */
#include <cuda_runtime.h>
#include <chrono>
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "lrpt_b200.h"

namespace {

struct DevBuf {
	void *p = nullptr;
	~DevBuf() { if (p) cudaFree(p); }
	cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 1); }
	template <class T> T *as() const { return static_cast<T *>(p); }
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
	bool on = getenv("LRPT_SHARD_TIMING") != nullptr;
	std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
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
int scan_rows(const int8_t *d_soft, size_t soft_stride, const uint32_t *d_q, size_t q_stride, const int32_t *d_count,
              const int64_t *d_base, int n, const std::vector<int64_t> &target, std::vector<int32_t> &k,
              std::vector<float> &agree, std::vector<int64_t> &cut)
{
	const int nb = n - 1;
	k.assign(nb > 0 ? nb : 0, 0); agree.assign(nb > 0 ? nb : 0, 0.0f); cut.assign(nb > 0 ? nb : 0, 0);
	if (nb <= 0) return LRPT_OK;
	DevBuf tgt, dcut, ia, ib, nav, dk, same;
	CK(tgt.alloc(8*(size_t)nb)); CK(dcut.alloc(8*(size_t)nb)); CK(ia.alloc(4*(size_t)nb)); CK(ib.alloc(4*(size_t)nb));
	CK(nav.alloc(4*(size_t)nb)); CK(dk.alloc(4*(size_t)nb)); CK(same.alloc(4*(size_t)nb));
	CK(cudaMemcpy(tgt.p, target.data(), 8*(size_t)nb, cudaMemcpyHostToDevice));
	RC(lrpt_shard_find_cuts_device(d_q, q_stride, d_count, d_base, n, tgt.as<int64_t>(), dcut.as<int64_t>(), ia.as<int32_t>(),
	                               ib.as<int32_t>(), nav.as<int32_t>(), nullptr));
	std::vector<int32_t> navail(nb);
	CK(cudaMemcpy(navail.data(), nav.p, 4*(size_t)nb, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(cut.data(), dcut.p, 8*(size_t)nb, cudaMemcpyDeviceToHost));
	int npairs = INT_MAX;
	for (int v : navail) npairs = v < npairs ? v : npairs;
	if (npairs < 8) return LRPT_OK;                                 /* k = 0, agreement = 0: the caller reports it */
	RC(lrpt_shard_quadrants_device(d_soft, soft_stride, d_q, q_stride, d_base, n, ia.as<int32_t>(), ib.as<int32_t>(), npairs,
	                               dk.as<int32_t>(), same.as<int32_t>(), nullptr));
	std::vector<int32_t> s(nb);
	CK(cudaMemcpy(k.data(), dk.p, 4*(size_t)nb, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(s.data(), same.p, 4*(size_t)nb, cudaMemcpyDeviceToHost));
	for (int b = 0; b < nb; b++) agree[b] = (float)s[b]/(float)npairs;
	return LRPT_OK;
}

} // namespace

extern "C" int lrpt_sharded_process(const lrpt_params_t *params, const lrpt_shard_plan_t *plan, const void *raw_iq,
                                    size_t nsamples, int8_t *soft, size_t cap, size_t *nsym, lrpt_shard_report_t *rep)
{
	if (!params || !plan || !raw_iq || !soft || !nsym) return LRPT_ERR_ARG;
	const size_t C = plan->chunk, W = plan->warm, V = plan->overlap;
	if (!C || !V || (C & 7) || (W & 7) || (V & 7) || params->oqpsk) return LRPT_ERR_ARG;   /* 16-byte aligned rows; QPSK ambiguity only */
	if (params->bps != 8 && params->bps != 16 && params->bps != 32) return LRPT_ERR_ARG;
	const size_t bytes = (size_t)params->bps/4;
	const long long L = params->interp_factor;
	const size_t M = nsamples > W ? (nsamples - W + C - 1)/C : 1;
	if (M > (size_t)INT_MAX/2) return LRPT_ERR_ARG;
	/* the rows' sub-step indices (sample*interp + sub-step, uint32 side output) must not wrap: the joins
	 * rely on their ascending order */
	if ((double)(W > C + V ? W : C + V)*(double)params->interp_factor >= 4294967296.0) return LRPT_ERR_ARG;
	lrpt_shard_report_t r;
	memset(&r, 0, sizeof(r));
	r.nchunks = (int32_t)M; r.min_agreement_scan = r.min_agreement_final = 1.0f; r.first_lock_symbol = -1; r.aligned = 1;
	*nsym = 0;
	CK(cudaSetDevice(params->device));

	Phases ph;
	lrpt_params_t p = *params;
	p.nstreams = (int32_t)M;
	Handle hd;
	RC(lrpt_create(&hd.h, &p));
	lrpt_demod_t *h = hd.h;

	if (M < 2) {                                                    /* one chunk: the sequential run itself, exact */
		size_t n = 0; long long fl = -1;
		RC(lrpt_process(h, raw_iq, nsamples, soft, cap, &n, &fl));
		*nsym = n; r.first_lock_symbol = fl; r.launches = (int32_t)lrpt_launch_count(h);
		if (rep) *rep = r;
		return LRPT_OK;
	}

	/* the stream on the device, zero padded so that every row can read W + C + 2V samples */
	const size_t padded = (M - 1)*C + W + C + 2*V;
	const size_t n_row = C + V;                                     /* samples of pass B / pass C */
	const size_t cap_row = (symbol_capacity(W > n_row ? W : n_row, p) + 7)/8*8;
	DevBuf d_raw, d_soft, d_q, d_nsym, d_cnt, d_base, d_states;
	CK(d_raw.alloc(padded*bytes)); CK(d_soft.alloc(M*2*cap_row)); CK(d_q.alloc(M*4*cap_row)); CK(d_nsym.alloc(4*M));
	CK(d_cnt.alloc(4*M)); CK(d_base.alloc(8*M));
	ph.mark("handle + device buffers");
	CK(cudaMemset(d_raw.as<char>() + nsamples*bytes, 0, (padded - nsamples)*bytes));
	CK(cudaMemcpy(d_raw.p, raw_iq, nsamples*bytes, cudaMemcpyHostToDevice));
	ph.mark("H2D of the recording");
	RC(lrpt_set_symbol_index_output(h, d_q.as<uint32_t>(), 4*cap_row));
	const char *raw0 = d_raw.as<char>();
	const size_t soft_stride = 2*cap_row, q_stride = 4*cap_row;
	std::vector<uint32_t> counts(M);
	std::vector<int64_t> base(M);
	auto run_pass = [&](size_t first_sample, size_t n) -> int {
		RC(lrpt_process_batch_device(h, raw0 + first_sample*bytes, C*bytes, n, d_soft.as<int8_t>(), soft_stride, cap_row,
		                             d_nsym.as<uint32_t>(), nullptr, 0, nullptr));
		RC(lrpt_sync(h, nullptr));
		RC(lrpt_get_counts(h, counts.data(), (int)M));
		for (size_t c = 0; c < M; c++) if (counts[c] > cap_row) return LRPT_ERR_CAP;   /* the stitch kernels trust the counts */
		for (size_t c = 0; c < M; c++) base[c] = (long long)(c*C + first_sample)*L;
		CK(cudaMemcpy(d_cnt.p, counts.data(), 4*M, cudaMemcpyHostToDevice));   /* uint32 counts < 2^31: read as int32 */
		CK(cudaMemcpy(d_base.p, base.data(), 8*M, cudaMemcpyHostToDevice));
		return LRPT_OK;
	};
	auto cut_target = [&](size_t c, size_t shift) { return (long long)(W + c*C + shift + (V/4 < 64 ? V/4 : 64))*L; };

	/* pass A: warm-up; chunk 0's warm-up symbols open the output (they are the sequential run's) */
	size_t out_n = 0;
	if (W) {
		RC(run_pass(0, W));
		if (counts[0] > cap) return LRPT_ERR_CAP;
		CK(cudaMemcpy(soft, d_soft.p, 2*(size_t)counts[0], cudaMemcpyDeviceToHost));
		out_n = counts[0];
	}
	ph.mark("pass A (warm-up)");
	/* pass B: owned + overlap, from the live state */
	RC(run_pass(W, n_row));
	ph.mark("pass B");
	std::vector<int64_t> target(M - 1), cut;
	std::vector<int32_t> k;
	std::vector<float> agree;
	for (size_t c = 1; c < M; c++) target[c - 1] = cut_target(c, 0);
	RC(scan_rows(d_soft.as<int8_t>(), soft_stride, d_q.as<uint32_t>(), q_stride, d_cnt.as<int32_t>(), d_base.as<int64_t>(), (int)M,
	             target, k, agree, cut));
	for (float a : agree) r.min_agreement_scan = a < r.min_agreement_scan ? a : r.min_agreement_scan;
	std::vector<int32_t> K(M, 0);
	for (size_t c = 1; c < M; c++) K[c] = (K[c - 1] + k[c - 1]) & 3;
	{
		/* chunk 0's pass-B symbols continue the output, up to the end of the stream */
		DevBuf lo, hi, st, ln;
		const int64_t l = -1, hh = (int64_t)nsamples*L - 1;
		int32_t s0 = 0, n0 = 0;
		CK(lo.alloc(8)); CK(hi.alloc(8)); CK(st.alloc(4)); CK(ln.alloc(4));
		CK(cudaMemcpy(lo.p, &l, 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(hi.p, &hh, 8, cudaMemcpyHostToDevice));
		RC(lrpt_shard_ranges_device(d_q.as<uint32_t>(), q_stride, d_cnt.as<int32_t>(), d_base.as<int64_t>(), 1, lo.as<int64_t>(),
		                            hi.as<int64_t>(), st.as<int32_t>(), ln.as<int32_t>(), nullptr));
		CK(cudaMemcpy(&s0, st.p, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&n0, ln.p, 4, cudaMemcpyDeviceToHost));
		if (out_n + (size_t)n0 > cap) return LRPT_ERR_CAP;
		CK(cudaMemcpy(soft + 2*out_n, d_soft.as<int8_t>() + 2*(size_t)s0, 2*(size_t)n0, cudaMemcpyDeviceToHost));
		out_n += (size_t)n0;
		lrpt_status_t s;
		RC(lrpt_status(h, 0, &s));
		r.first_lock_symbol = s.first_lock_symbol;                  /* chunk 0 counts its symbols as the stream does */
	}

	ph.mark("quadrant scan + chunk 0");
	const size_t out_head = out_n;                                  /* symbols of chunk 0's own trajectory (passes A + B) */
	std::vector<long long> nsym_start(M, 0);                        /* symbol count row c inherits at the start of pass C */
	/* hand-off: row c starts pass C from row c-1's end state, its Costas NCO turned back K[c-1] quarter turns
	 * (p_phase = (float)((double)p_phase - K*pi/2), as lrpt_restore does; pll.c:16) */
	{
		const size_t total = lrpt_states_size(h), sb = sizeof(lrpt_state_t);
		const size_t hb = (total - M*sb)/M;                         /* delay line bytes per row */
		CK(d_states.alloc(total));
		RC(lrpt_export_states_device(h, d_states.p, total, nullptr));
		RC(lrpt_sync(h, nullptr));
		std::vector<char> cur(total), nxt(total);
		CK(cudaMemcpy(cur.data(), d_states.p, total, cudaMemcpyDeviceToHost));
		lrpt_state_t *sc = reinterpret_cast<lrpt_state_t *>(cur.data()), *sn = reinterpret_cast<lrpt_state_t *>(nxt.data());
		for (size_t c = 0; c < M; c++)
			sc[c].p_phase = (float)((double)sc[c].p_phase - (double)(K[c] & 3)*1.57079632679489661923);
		sn[0] = sc[0];                                              /* row 0 has no predecessor; its pass C is not used */
		memcpy(nxt.data() + M*sb, cur.data() + M*sb, hb);
		for (size_t c = 1; c < M; c++) {
			sn[c] = sc[c - 1];
			memcpy(nxt.data() + M*sb + c*hb, cur.data() + M*sb + (c - 1)*hb, hb);
		}
		for (size_t c = 0; c < M; c++) nsym_start[c] = sn[c].nsymbols;
		CK(cudaMemcpy(d_states.p, nxt.data(), total, cudaMemcpyHostToDevice));
		CK(cudaDeviceSynchronize());                                /* pageable upload on the legacy stream: not ordered with the handle's stream */
		RC(lrpt_import_states_device(h, d_states.p, total, 1, nullptr));
		RC(lrpt_sync(h, nullptr));
	}

	ph.mark("state hand-off");
	/* pass C, then the join of rows 1 .. M-1 (rows 0 and 1 are one exact trajectory: nothing is cut off row 1's front) */
	RC(run_pass(W + V, n_row));
	ph.mark("pass C");
	const int n = (int)M - 1;
	const int8_t *soft1 = d_soft.as<int8_t>() + soft_stride;
	const uint32_t *q1 = reinterpret_cast<const uint32_t *>(d_q.as<char>() + q_stride);
	const int32_t *cnt1 = d_cnt.as<int32_t>() + 1;
	const int64_t *base1 = d_base.as<int64_t>() + 1;
	std::vector<int64_t> target2(n > 1 ? n - 1 : 0), cut2;
	std::vector<int32_t> k2;
	std::vector<float> agree2;
	for (int b = 0; b + 1 < n; b++) target2[b] = cut_target((size_t)b + 2, V);
	RC(scan_rows(soft1, soft_stride, q1, q_stride, cnt1, base1, n, target2, k2, agree2, cut2));
	for (float a : agree2) r.min_agreement_final = a < r.min_agreement_final ? a : r.min_agreement_final;
	std::vector<int32_t> turns(n, 0);
	std::vector<int64_t> lo(n, -1), hi(n, LLONG_MAX);
	for (int b = 0; b + 1 < n; b++) {
		turns[b + 1] = (turns[b] + k2[b]) & 3;
		if (k2[b]) r.aligned = 0;
		hi[b] = cut2[b]; lo[b + 1] = cut2[b];
	}
	hi[n - 1] = (int64_t)nsamples*L - 1;                            /* nothing from the zero padding */
	DevBuf dlo, dhi, dst, dln, doff, dturn, dout;
	CK(dlo.alloc(8*(size_t)n)); CK(dhi.alloc(8*(size_t)n)); CK(dst.alloc(4*(size_t)n)); CK(dln.alloc(4*(size_t)n));
	CK(doff.alloc(8*(size_t)n)); CK(dturn.alloc(4*(size_t)n));
	CK(cudaMemcpy(dlo.p, lo.data(), 8*(size_t)n, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(dhi.p, hi.data(), 8*(size_t)n, cudaMemcpyHostToDevice));
	RC(lrpt_shard_ranges_device(q1, q_stride, cnt1, base1, n, dlo.as<int64_t>(), dhi.as<int64_t>(), dst.as<int32_t>(),
	                            dln.as<int32_t>(), nullptr));
	std::vector<int32_t> len(n);
	CK(cudaMemcpy(len.data(), dln.p, 4*(size_t)n, cudaMemcpyDeviceToHost));
	std::vector<int64_t> off(n);
	size_t total = 0, longest = 0;
	for (int b = 0; b < n; b++) { off[b] = (int64_t)total; total += (size_t)len[b]; longest = (size_t)len[b] > longest ? (size_t)len[b] : longest; }
	if (out_n + total > cap) return LRPT_ERR_CAP;
	CK(dout.alloc(2*total));
	CK(cudaMemcpy(doff.p, off.data(), 8*(size_t)n, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(dturn.p, turns.data(), 4*(size_t)n, cudaMemcpyHostToDevice));
	RC(lrpt_shard_gather_device(soft1, soft_stride, n, longest, dst.as<int32_t>(), dln.as<int32_t>(), doff.as<int64_t>(),
	                            dturn.as<int32_t>(), dout.as<int8_t>(), nullptr));
	/* Stream-wide first lock (main.c:312 gates the output on pll_did_lock_once()). Chunk 0 may not have locked
	 * by the end of its passes -- a recording that starts before the signal is up, the normal case for a
	 * satellite pass -- so the answer is the first OUTPUT symbol that came from a loop which had locked once:
	 * row c's pass C inherits its predecessor's flag and symbol count (exact for row 1, which continues chunk 0
	 * bit for bit), so the row-local index at which it had locked is first_lock_symbol - inherited count. */
	if (r.first_lock_symbol < 0) {
		std::vector<int32_t> start(n);
		CK(cudaMemcpy(start.data(), dst.p, 4*(size_t)n, cudaMemcpyDeviceToHost));
		const size_t tot_states = lrpt_states_size(h);
		RC(lrpt_export_states_device(h, d_states.p, tot_states, nullptr));
		RC(lrpt_sync(h, nullptr));
		std::vector<lrpt_state_t> fin(M);
		CK(cudaMemcpy(fin.data(), d_states.p, M*sizeof(lrpt_state_t), cudaMemcpyDeviceToHost));
		for (int b = 0; b < n && r.first_lock_symbol < 0; b++) {
			const lrpt_state_t &st = fin[(size_t)b + 1];
			if (st.first_lock_symbol < 0 || len[b] <= 0) continue;
			long long j = st.first_lock_symbol - nsym_start[(size_t)b + 1];     /* < 0: locked before this pass began */
			if (j < start[b]) j = start[b];
			if (j < (long long)start[b] + len[b]) r.first_lock_symbol = (long long)out_head + off[b] + (j - start[b]);
		}
	}
	ph.mark("join (scan, ranges, gather)");
	CK(cudaMemcpy(soft + 2*out_n, dout.p, 2*total, cudaMemcpyDeviceToHost));   /* default stream: ordered after the gather */
	ph.mark("D2H of the symbols");
	out_n += total;
	*nsym = out_n;
	r.launches = (int32_t)lrpt_launch_count(h);
	if (rep) *rep = r;
	return LRPT_OK;
}
