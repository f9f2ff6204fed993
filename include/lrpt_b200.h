/*
 * lrpt_b200.h -- C ABI of the B200-native LRPT demodulator hot path.
 *
 * Drop-in boundary. The reference (dbdexter-dev/meteor_demod)
 * has no plugin/FFI interface; its narrowest seam is the per-sample push API
 *     demod_init / demod_deinit / demod_qpsk / demod_oqpsk         demod.h:29-50
 * plus the status getters
 *     pll_get_freq / pll_get_locked / pll_did_lock_once            pll.h:20,27,34
 *     mm_omega                                                     timing.h:32
 *     agc_get_gain                                                 agc.h:18
 * all over process-wide statics, called from the loop in thread_process
 * (main.c:303-317). A per-sample call into a GPU is untenable, so this ABI replaces
 * the BODY of that loop with a block call: raw interleaved I/Q bytes in (exactly what
 * wav_read consumes, wavfile.c:51-80), int8 soft symbols out (exactly what
 * main.c:305-306 puts in the ring), every symbol, ungated; the host applies the
 * 512-symbol lock gating of main.c:308-316 using first_lock_symbol.
 *
 * Plain C: opaque handle, POD structs, pointers + sizes, int return codes
 * (0 = ok, negative = error), no global state, no exceptions, no torch types.
 * A handle owns `nstreams` independent demodulators ("streams" = recordings or
 * time shards); each keeps the complete state of the reference's statics, so a
 * stream continues seamlessly across calls and can be checkpointed / handed to
 * another process or GPU with lrpt_export_state / lrpt_import_state.
 *
 * There is no CPU fallback: every entry point that computes needs a CUDA device
 * and fails with LRPT_ERR_CUDA otherwise.
 */
#ifndef LRPT_B200_H
#define LRPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LRPT_ABI_VERSION 1

enum {
	LRPT_OK        =  0,
	LRPT_ERR_ARG   = -1,   /* bad argument / unsupported configuration          */
	LRPT_ERR_CUDA  = -2,   /* CUDA runtime failure (see lrpt_last_error)          */
	LRPT_ERR_NOMEM = -3,   /* host or device allocation failed                    */
	LRPT_ERR_CAP   = -4,   /* a stream produced more symbols than `cap`; state is
	                          still advanced, surplus symbols were counted, not stored */
	LRPT_ERR_STATE = -5    /* state blob does not match this handle's configuration */
};

/* Which kernel runs the recurrence; results are bit-identical. AUTO picks by streams per SM:
 * WS (all-phase FIR producers + one recurrence warp) for few streams, SPEC (speculative FIR)
 * in between, LANE (one lane per stream, lazy FIR, many warps per SM) for large batches;
 * SIMPLE (one thread per stream, source order) covers whatever the others do not. */
enum { LRPT_KERNEL_AUTO = 0, LRPT_KERNEL_SIMPLE = 1, LRPT_KERNEL_WS = 2, LRPT_KERNEL_SPEC = 3, LRPT_KERNEL_LANE = 4 };

typedef struct lrpt_demod lrpt_demod_t;       /* opaque; owns device buffers + a CUDA stream */

/* Exactly demod_init's arguments (demod.h:29) + the ingest format (wavfile.c:57-72). */
typedef struct lrpt_params {
	float   pll_bw;          /* -b, default 1         (demod.h:15)                 */
	float   sym_bw;          /* SYM_BW 0.00005        (demod.h:14)                 */
	float   freq_max;        /* rad/symbol; < 0 => FREQ_MAX 0.3 (pll.c:30); main.c:136 converts -d Hz */
	int32_t samplerate;      /* -s / wav header                                    */
	int32_t symrate;         /* -r, truncated to int as main.c:187 does            */
	int32_t interp_factor;   /* -O, default 5                                      */
	int32_t rrc_order;       /* -f, default 32 (taps = 2*order+1)                  */
	int32_t oqpsk;           /* -m oqpsk                                           */
	int32_t bps;             /* 8 (u8 offset-128) | 16 (s16) | 32 (f32)            */
	int32_t device;          /* CUDA ordinal                                       */
	int32_t nstreams;        /* independent demodulators in this handle (>= 1)     */
	int32_t kernel;          /* LRPT_KERNEL_*                                      */
} lrpt_params_t;

/*
 * Complete per-stream state = every static of the reference's hot path
 * (SURVEY.md section 8e). A state blob is this header followed by the FIR delay
 * line: (taps-1) complex float samples (re,im), oldest first (filter.h:6 in
 * chronological order).
 */
typedef struct lrpt_state {
	uint32_t magic;              /* 'LRPS' */
	uint32_t taps;               /* 2*order+1, for validation                    */
	/* symbol timing: timing.c:13-14, dual-threshold flag timing.c:43            */
	float    t_phase, t_freq, t_prev;
	int32_t  t_dual_state;       /* 1 | 2 (OQPSK only)                            */
	float    oq_inphase;         /* demod.c:54                                    */
	/* AGC: agc.c:9-10 */
	float    agc_gain, agc_bias_re, agc_bias_im;
	/* Costas PLL: pll.c:16-20, sweep direction pll.c:112 */
	float    p_phase, p_freq, p_err;
	int32_t  p_locked, p_locked_once, p_updown;
	/* bookkeeping: totals since creation / import (main.c:291,298,312) */
	int64_t  nsamples, nsymbols, first_lock_symbol;   /* first_lock_symbol = -1 until locked once */
} lrpt_state_t;

/* Host-visible status, the values main.c:250-258 prints. */
typedef struct lrpt_status {
	float   pll_freq;        /* pll_get_freq(), rad/symbol (rad/half-symbol for OQPSK) */
	float   mm_omega;        /* mm_omega()                                          */
	float   agc_gain;        /* agc_get_gain()                                      */
	int32_t locked;          /* pll_get_locked()                                    */
	int32_t locked_once;     /* pll_did_lock_once()                                 */
	int64_t nsamples, nsymbols, first_lock_symbol;
} lrpt_status_t;

/* ---- lifecycle: replaces demod_init / demod_deinit (demod.h:29,34) ------------- */
int  lrpt_create(lrpt_demod_t **out, const lrpt_params_t *p);
void lrpt_destroy(lrpt_demod_t *h);
/* re-initialise every stream to the state demod_init + fresh statics give */
int  lrpt_reset(lrpt_demod_t *h);
/* same, enqueued on `cuda_stream` (NULL = the handle's stream) without host synchronisation, so it
 * orders with lrpt_process_batch_device calls on that stream */
int  lrpt_reset_async(lrpt_demod_t *h, void *cuda_stream);

/* ---- the hot path: replaces the body of main.c:303-317 ------------------------- */
/*
 * Single-stream push (stream 0), HOST buffers. raw_iq: nsamples interleaved I,Q
 * items of the configured format. soft: 2 int8 per symbol, capacity `cap` symbols.
 * *nsym = symbols produced by this call. *first_lock_symbol = index (counted from
 * the stream's first symbol ever) of the first symbol after which
 * pll_did_lock_once() was true, or -1. Synchronous.
 */
int  lrpt_process(lrpt_demod_t *h, const void *raw_iq, size_t nsamples,
                  int8_t *soft, size_t cap, size_t *nsym, long long *first_lock_symbol);

/*
 * Batch push, HOST buffers: stream s reads nsamples items at
 * (char*)raw_iq + s*raw_stride and writes symbols at soft + s*soft_stride
 * (capacity cap symbols each); nsym[s] receives the per-stream count. sym_f32
 * (may be NULL) receives the unquantised float symbols (2 floats per symbol,
 * stride symf_stride bytes) for state-level parity tests. Synchronous; host<->device
 * copies are pipelined with the kernel in slabs.
 */
int  lrpt_process_batch(lrpt_demod_t *h, const void *raw_iq, size_t raw_stride, size_t nsamples,
                        int8_t *soft, size_t soft_stride, size_t cap, uint32_t *nsym,
                        float *sym_f32, size_t symf_stride);

/*
 * Batch push, DEVICE buffers (HBM-resident input, e.g. torch tensors' data_ptr()).
 * Asynchronous on `cuda_stream` (a cudaStream_t passed as void*; NULL = the handle's
 * own stream). d_nsym: device uint32[nstreams] (may be NULL). Pointers must be
 * 16-byte aligned and strides multiples of 16. Use lrpt_sync + lrpt_get_counts to
 * read the counts on the host.
 */
int  lrpt_process_batch_device(lrpt_demod_t *h, const void *d_raw_iq, size_t raw_stride,
                               size_t nsamples, int8_t *d_soft, size_t soft_stride, size_t cap,
                               uint32_t *d_nsym, float *d_sym_f32, size_t symf_stride,
                               void *cuda_stream);
/*
 * Optional side output of lrpt_process_batch_device: for every symbol the timing sub-step that produced
 * it (input sample index * interp_factor + sub-step, counted from the start of the call; demod.c:33-35),
 * as uint32 at d_index + stream*stride (bytes), same capacity as the soft output. Time-shard stitching
 * needs it. NULL switches it off again.
 */
int  lrpt_set_symbol_index_output(lrpt_demod_t *h, uint32_t *d_index, size_t stride);
int  lrpt_sync(lrpt_demod_t *h, void *cuda_stream);
/*
 * Ordering against the caller's streams without blocking the host. The handle's own stream (the one a NULL
 * cuda_stream argument selects) is non-blocking: it is NOT ordered with the legacy default stream or any other.
 *   lrpt_stream_wait:    everything enqueued on producer_stream so far (NULL = the legacy default stream)
 *                        completes before anything enqueued on the handle's stream from now on starts -- call it
 *                        after filling the raw buffer / clearing the soft buffer on your stream;
 *   lrpt_stream_release: the reverse, for a consumer_stream that reads the symbols.
 * (The reference is synchronous, demod.c:25-58; these have no counterpart there.)
 */
int  lrpt_stream_wait(lrpt_demod_t *h, void *producer_stream);
int  lrpt_stream_release(lrpt_demod_t *h, void *consumer_stream);
/* per-stream symbol counts of the most recent call (host copy; call after lrpt_sync) */
int  lrpt_get_counts(lrpt_demod_t *h, uint32_t *nsym, int nstreams);

/* ---- status: replaces pll_get_freq / pll_get_locked / pll_did_lock_once /
 *      mm_omega / agc_get_gain (pll.h:20-34, timing.h:32, agc.h:18) -------------- */
int  lrpt_status(lrpt_demod_t *h, int stream, lrpt_status_t *st);

/* ---- state hand-off / checkpoint ------------------------ */
size_t lrpt_state_size(const lrpt_demod_t *h);       /* bytes of one state blob */
int  lrpt_export_state(lrpt_demod_t *h, int stream, void *buf, size_t *len);
int  lrpt_import_state(lrpt_demod_t *h, int stream, const void *buf, size_t len);

/*
 * All streams at once, DEVICE buffers, asynchronous on `cuda_stream` (NULL = the handle's stream): the
 * rank-to-rank hand-off of a time-sharded batch (sharded.py, relay). Layout of `d_buf`
 * (lrpt_states_size bytes): nstreams x lrpt_state_t, then nstreams x (taps-1) complex float delay
 * lines. The bytes can go straight into ncclSend / ncclRecv. Import trusts the buffer (it checks the
 * header of stream 0 when `check` is non-zero, which costs a synchronisation).
 */
size_t lrpt_states_size(const lrpt_demod_t *h);
int  lrpt_export_states_device(lrpt_demod_t *h, void *d_buf, size_t len, void *cuda_stream);
int  lrpt_import_states_device(lrpt_demod_t *h, const void *d_buf, size_t len, int check, void *cuda_stream);

/*
 * Snapshot / restore of ALL streams' states on the device (time-shard two-pass scheme, sharded.py).
 * lrpt_restore copies the snapshot back; quarter_turns (host int32[nstreams], may be NULL) then turns
 * every stream's Costas NCO back by that many quarter turns: p_phase -= turns*pi/2 (pll.c:16), which
 * moves a locked QPSK loop to another of its four equivalent lock points without losing lock.
 */
int  lrpt_snapshot(lrpt_demod_t *h);
int  lrpt_restore(lrpt_demod_t *h, const int32_t *quarter_turns);

/* ---- joining the chunks of one time-sharded stream (device buffers) ----------
 * No counterpart in the reference (one sequential stream, demod.c:24-48). Rows = consecutive chunks of
 * ONE stream demodulated as the streams of a batch (lrpt_process_batch_device with a raw_stride of one
 * chunk), with the symbol index side output: d_q[r][i] = sub-step of symbol i counted from the row's
 * first sample, ascending; d_base[r] makes it absolute in the stream; d_count[r] symbols are valid.
 * Strides in bytes. All four are asynchronous on `cuda_stream` and need no handle. */
/* boundary b = rows (b | b+1), b < nrows-1: cut point mid-way between the two symbols of row b around
 * d_target[b] (absolute sub-step), first paired symbol in either row, pairs available */
int  lrpt_shard_find_cuts_device(const uint32_t *d_q, size_t q_stride, const int32_t *d_count, const int64_t *d_base,
                                 int nrows, const int64_t *d_target, int64_t *d_cut, int32_t *d_ia, int32_t *d_ib,
                                 int32_t *d_navail, void *cuda_stream);
/* per boundary: quarter turns k (0..3) that map row b+1 onto row b -- a QPSK Costas loop locks with a
 * k*90 degree ambiguity (pll.c:143-152) -- from exact integer correlations of `npairs` paired symbols,
 * and the number of pairs whose hard decisions agree under k with instants at most 2 sub-steps apart */
int  lrpt_shard_quadrants_device(const int8_t *d_soft, size_t soft_stride, const uint32_t *d_q, size_t q_stride,
                                 const int64_t *d_base, int nrows, const int32_t *d_ia, const int32_t *d_ib, int npairs,
                                 int32_t *d_k, int32_t *d_same, void *cuda_stream);
/* the same for OQPSK rows: the I arm is sampled half a symbol before the Q arm (demod.c:66-83), so an ODD quarter turn
 * shows as a timing offset of half_substeps = fs*interp/(2*symrate) between paired symbols with the arms re-paired
 * (meteor_demod_b200/sharded.py::boundary_quadrants_oqpsk); npairs must leave one symbol of row b in reserve */
int  lrpt_shard_quadrants_oqpsk_device(const int8_t *d_soft, size_t soft_stride, const uint32_t *d_q, size_t q_stride,
                                       const int64_t *d_base, int nrows, const int32_t *d_ia, const int32_t *d_ib, int npairs,
                                       float half_substeps, int32_t *d_k, int32_t *d_same, void *cuda_stream);
/* per row: the run of symbols with d_lo[r] < absolute sub-step <= d_hi[r] (INT64_MAX = to the end) */
int  lrpt_shard_ranges_device(const uint32_t *d_q, size_t q_stride, const int32_t *d_count, const int64_t *d_base, int nrows,
                              const int64_t *d_lo, const int64_t *d_hi, int32_t *d_start, int32_t *d_len, void *cuda_stream);
/* d_out[d_off[r] + j] = row r's symbol d_start[r] + j turned by d_turns[r] quarter turns ((I,Q) -> (-Q,I)
 * each; exact on int8 pairs), j < d_len[r] <= max_len; d_off in symbols */
int  lrpt_shard_gather_device(const int8_t *d_soft, size_t soft_stride, int nrows, size_t max_len, const int32_t *d_start,
                              const int32_t *d_len, const int64_t *d_off, const int32_t *d_turns, int8_t *d_out,
                              void *cuda_stream);

/*
 * ONE recording time-sharded over the lanes of one GPU in a single call (csrc/shard_run.cu; HOST buffers;
 * replaces the whole loop main.c:303-317 for an offline file, at STATISTICAL parity): the stream is cut into
 * chunks of `chunk` samples that run as the streams of a batch -- `warm` samples of warm-up before each
 * chunk, `overlap` samples into its successor (all multiples of 8; chunk + warm + overlap >= ~400 k samples
 * keeps the share of symbols off by more than one LSB at the 0.3-0.4 % of the reference's own FMA/strict
 * builds). The first two chunks are bit-exact. QPSK and OQPSK (whose warm-up has to cover the
 * reference's slow carrier pull-in: ~250 k samples at 700 Hz); p->nstreams is ignored.
 * soft: 2 int8 per symbol, ALL symbols in stream order (the host applies the 512-symbol lock gating with
 * rep->first_lock_symbol, which is chunk 0's: -1 if the loop had not locked by the end of chunk 0).
 */
typedef struct lrpt_shard_plan {
	uint64_t chunk, warm, overlap;
	uint64_t seed_nfft;            /* 0: every chunk acquires the carrier as the reference does (pll.c:117-128; needs
	                                * ~150 k samples of warm-up at 700 Hz). A power of two in 256..16384: chunks after the
	                                * first start with their Costas NCO at lrpt_carrier_estimate_device's estimate over
	                                * their first seed_nfft samples (32 Ki samples of warm-up are then enough)          */
} lrpt_shard_plan_t;
typedef struct lrpt_shard_report {
	int32_t nchunks, launches;
	float   min_agreement_scan;    /* worst boundary of the quadrant scan: share of overlap symbols agreeing  */
	float   min_agreement_final;   /* worst boundary of the final join                                          */
	int32_t aligned;               /* 1: every final boundary needed no turn (all chunks at one lock point)   */
	int32_t reserved;
	int64_t first_lock_symbol;
} lrpt_shard_report_t;
int  lrpt_sharded_process(const lrpt_params_t *p, const lrpt_shard_plan_t *plan, const void *raw_iq, size_t nsamples,
                          int8_t *soft, size_t cap, size_t *nsym, lrpt_shard_report_t *rep);
/*
 * The same over `ndev` GPUs of one node, driven from this one process (one host thread per device, no Python):
 * device i takes a consecutive run of chunks and only its time slice of the recording (+ the warm-up and overlap
 * it over-reads). What crosses a device boundary is the reference's state vector of the boundary chunk
 * (lrpt_state_t + delay line: pll.c:16-20,112, timing.c:13-14,43, agc.c:9-10, filter.h:5-11, demod.c:54) and that
 * chunk's overlap symbols for the quadrant scan, by ncclSend / ncclRecv (NCCL is loaded at run time, libnccl.so.2);
 * quarter-turn sums and symbol counts are integers in host memory. Byte-identical to lrpt_sharded_process.
 * devices: CUDA ordinals (p->device is ignored when ndev > 1); fewer devices are used when the recording has fewer
 * than two chunks per device. rep->launches is per device.
 */
int  lrpt_sharded_process_multi(const lrpt_params_t *p, const lrpt_shard_plan_t *plan, const void *raw_iq, size_t nsamples,
                                int8_t *soft, size_t cap, size_t *nsym, lrpt_shard_report_t *rep, const int *devices, int ndev);
/* The NCCL communicators of lrpt_sharded_process_multi are kept for the next call with the same device list (creating
 * them costs more than demodulating a 2-GSample recording; one multi-GPU call runs at a time). This frees them. */
void lrpt_sharded_release(void);

/* ---- decoder front-end (csrc/frontend.cu) ----------------------------------------
 * The consumer of this path's output in the reference's pipeline (README.md:6-9,87-91: the `.s` soft-symbol file
 * or pipe goes to meteor_decode): frame synchronisation and Viterbi decoding of the int8 soft-symbol stream exactly
 * as main.c:305-313 lays it out (I, Q per symbol, no header). The algorithm is the link layer's published one --
 * CCSDS 131.0-B as Meteor-M LRPT uses it: ASM 0x1ACFFC1D, rate 1/2 K = 7 code, G1 = 171 (I) / G2 = 133 (Q), 8192
 * symbols per 1024-byte CADU -- restated for the CPU in oracle/frontend_oracle.c, which these entry points
 * reproduce bit for bit. Device buffers, asynchronous on `cuda_stream`, no handle.
 *
 * lrpt_fe_sync_device: per symbol offset o <= nsym - 32 the best match (0..64 equal bits) of the 64 hard
 * decisions from o on with the encoded ASM under the 8 symmetries of the constellation, hyp = quarter turns
 * (0..3) + 4 if I and Q are swapped (lowest on ties); later offsets get 0. d_score / d_hyp: nsym bytes rounded
 * up to a multiple of 4; d_words: scratch of lrpt_fe_sync_words(nsym) uint32; d_soft 16-byte aligned.
 * lrpt_fe_peaks_device: first offset of the highest score in every window of `window` offsets
 * (ceil(nsym/window) results).
 * lrpt_fe_viterbi_device: every frame f = the CADU starting at symbol d_frame_off[f] under symmetry
 * d_frame_hyp[f] -> d_cadu[f*1024 .. +1024) (the first four bytes are the ASM when a frame really starts there)
 * and the winning path metric; 64 symbols before and after the frame are decoded along with it. d_scratch:
 * any multiple of 66560 bytes >= 4 of them (lrpt_fe_viterbi_scratch_bytes = enough for every SM). */
int    lrpt_fe_sync_device(const int8_t *d_soft, size_t nsym, uint8_t *d_score, uint8_t *d_hyp, uint32_t *d_words,
                           void *cuda_stream);
size_t lrpt_fe_sync_words(size_t nsym);
int    lrpt_fe_peaks_device(const uint8_t *d_score, const uint8_t *d_hyp, size_t nsym, uint32_t window, uint32_t *d_off,
                            uint8_t *d_ohyp, uint8_t *d_oscore, void *cuda_stream);
size_t lrpt_fe_viterbi_scratch_bytes(int device);
int    lrpt_fe_viterbi_device(const int8_t *d_soft, size_t nsym, const uint32_t *d_frame_off, const uint8_t *d_frame_hyp,
                              int nframes, uint8_t *d_cadu, int32_t *d_metric, void *d_scratch, size_t scratch_bytes,
                              void *cuda_stream);

/* ---- page-locked host buffers ---------------------------------------------------------------------
 * The host-buffer entry points (lrpt_process, lrpt_process_batch, lrpt_sharded_process) copy at the full
 * speed of the host link only from / to page-locked memory; from ordinary malloc memory the driver stages
 * every copy through its own bounce buffer (measured: 470 ms instead of 80 ms for a 4.3 GB recording).
 * The reference reads into a static 32 KiB buffer (wavfile.c:8,55); a host that wants the link's speed
 * allocates its slabs here (or pins memory it already has). NULL / LRPT_ERR_CUDA when the driver refuses. */
void *lrpt_alloc_host(size_t bytes);
void  lrpt_free_host(void *p);
int   lrpt_pin_host(void *p, size_t bytes);       /* cudaHostRegister on caller-owned memory */
int   lrpt_unpin_host(void *p);

/* ---- the feed-forward stage alone (csrc/fir_stage.cu) ----------------------------------------------
 * filter_fwd_sample + filter_get (filter.c:39-65) for EVERY (sample n, timing sub-step i) of `nrows` rows,
 * device buffers: d_out[row][n*L + i] = the complex value filter_get(flt, i) returns after sample n of the
 * row has been pushed (float2: re, im), starting from the zeroed delay line of filter_init_rrc (filter.c:16).
 * The reference evaluates this lazily, once per symbol (demod.c:35); the product kernels do too. This entry
 * point materialises the whole stage so that it can be measured against its own roofline (2*bps/8 bytes in,
 * 8*L bytes out and 4*taps*L flops per sample) and compared value by value with filter_get.
 * mode 0: multiply and add rounded separately, oldest tap first -- bit-identical to filter_get;
 * mode 1: fused multiply-add (one rounding per tap; statistical parity only).
 * Bits 1-2 of mode are a tuning knob: output samples per thread (0 = chosen by filter length, 1 or 2);
 * the results do not depend on it.
 * Strides in bytes, pointers and strides 16-byte aligned, interp_factor <= 8, taps <= 257; asynchronous on
 * `cuda_stream` except for a short synchronisation while the tap table is uploaded. */
int  lrpt_fir_stage_device(const lrpt_params_t *p, const void *d_raw, size_t raw_stride, int nrows, size_t nsamples,
                           float *d_out, size_t out_stride, int mode, void *cuda_stream);

/* ---- coarse carrier estimate (csrc/acquire.cu) ------------------------------------
 * Not in the reference: it finds the carrier by sweeping its Costas NCO at 1e-6 rad/symbol^2 until the lock detector
 * fires (pll.c:117-128). Time-sharded chunks skip that wait: every chunk but the first starts with p_freq =
 * 2*pi*f_c/(symrate * (oqpsk ? 2 : 1)) (main.c:250 read backwards) for the f_c this returns.
 * d_cfo_hz[row] = carrier offset in Hz of the first nfft samples (a power of two, 256..16384) of each of `nrows` rows,
 * searched within +-fmax_hz: DC removed, x^4 (QPSK: line at 4 f_c) or x^2 (OQPSK: lines at 2 f_c -+ symrate), Hann
 * window, FFT, strongest candidate, three-point parabola. Strides in bytes; rows aligned to one I/Q pair.
 * Asynchronous on cuda_stream. */
int  lrpt_carrier_estimate_device(const lrpt_params_t *p, const void *d_raw, size_t raw_stride, int nrows, int nfft,
                                  double fmax_hz, double *d_cfo_hz, void *cuda_stream);

/* ---- introspection -------------------------------------------------------------- */
/*
 * Host-only (needs no CUDA device): what lrpt_create derives from `p`, exactly as
 * demod_init does (demod.c:8-15): the power-on state, the polyphase tap banks
 * (returns taps*interp, the count of floats), the loop constants
 * {t_center, t_maxdev, t_alpha, t_beta, p_alpha, p_beta, p_fmax} (timing.c:21-27,
 * pll.c:38,43) and the tanh table (pll.c:40-42). Any output pointer may be NULL.
 */
int  lrpt_describe(const lrpt_params_t *p, lrpt_state_t *initial_state, float *taps, int taps_cap,
                   float loop_consts[7], float lut[32]);
/* polyphase tap banks as filter_init_rrc lays them out (filter.c:18-22): bank j at [j*taps, (j+1)*taps) */
int  lrpt_get_taps(const lrpt_demod_t *h, float *dst, int cap);   /* returns taps*interp */
int  lrpt_get_tanh_lut(const lrpt_demod_t *h, float dst[32]);
/* number of CUDA kernels this handle has launched so far */
unsigned long long lrpt_launch_count(const lrpt_demod_t *h);
/* FIR outputs the recurrence had to evaluate itself because the speculative FIR warps had not (spec kernel;
 * a cost indicator, never a correctness matter) */
unsigned long long lrpt_fir_fallbacks(lrpt_demod_t *h);
/* name of the kernel AUTO resolved to ("simple" | "ws" | "spec" | "lane") */
const char *lrpt_kernel_name(const lrpt_demod_t *h);
const char *lrpt_last_error(const lrpt_demod_t *h);  /* human-readable detail of the last failure */
const char *lrpt_strerror(int code);
int  lrpt_abi_version(void);

/* main.c:136 : -d <Hz> to the rad/symbol freq_max argument (negative stays negative) */
float lrpt_freq_delta_from_hz(float freq_max_delta_hz, float symrate);

#ifdef __cplusplus
}
#endif
#endif /* LRPT_B200_H */
