/*
 * lrpt_demod -- C host with the reference's command line (meteor_demod main.c) in front
 * of the B200 demodulator (liblrpt_b200.so, include/lrpt_b200.h).
 *
 * What is kept from the reference, because it decides WHICH bytes come out:
 *   - flags and defaults                         main.c:19,35-51,66-79,82-134
 *   - -d Hz -> rad/symbol                        main.c:136
 *   - --stdout implies -B -q; "-" is stdin       main.c:145-148,155-157
 *   - canonical 44-byte WAV header; its sample rate / bits override -s / --bps; rewind on
 *     failure (which silently fails on a pipe)   wavfile.c:34-49, main.c:164-165
 *   - numbers are parsed like human_to_float: k/M suffix, truncated to an integer
 *                                                utils.c:62-85
 *   - only whole 32 KiB input blocks are used    wavfile.c:55
 *   - soft symbols leave in 512-symbol blocks, a block only if the PLL had locked once
 *     when it completed; the last partial block always   main.c:308-323
 *   - status line text                           main.c:253-258
 * What is not: ncurses (the TUI panes of tui.c:139-247 -- input position, bytes out, lock / gain / carrier /
 * symbol rate, constellation of the last 512 symbols -- are redrawn with plain ANSI escapes from the state
 * snapshot taken after every block), the per-sample demod calls
 * (one lrpt_process call per input slab instead), and the final-flush length bug
 * (main.c:321 writes 2*ring_idx bytes, the second half stale or out of bounds) unless
 * --ref-compatible-tail asks for the reference's byte count.
 * Beyond the reference: several input files on one command line are demodulated TOGETHER as the
 * streams of one batch (lrpt_process_batch; archive reprocessing) -- every file's output is what a
 * separate run on it writes, into <input>.s.
 */
#include <errno.h>
#include <getopt.h>
#include <math.h>
#include <poll.h>
#include <pthread.h>
#include <unistd.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "lrpt_b200.h"

#define SHORTOPTS "a:Bb:d:f:hm:o:O:qR:r:s:S:v"
#define RINGSIZE 512                      /* symbols, main.c:20 */
#define FILE_BLOCK 32768                  /* wavfile.c:8 */
#define SLAB_BLOCKS 512                   /* 16 MiB of input per lrpt_process call */

#ifndef VERSION
#define VERSION "1.0-b200"
#endif

enum { OPT_STDOUT = 0x100, OPT_DEVICE, OPT_REFTAIL, OPT_SHARD, OPT_TUI, OPT_GPUS, OPT_SEED };

static struct option longopts[] = {
	{ "batch",        0, NULL, 'B' }, { "pll-bw",       1, NULL, 'b' },
	{ "freq-delta",   1, NULL, 'd' }, { "fir-order",    1, NULL, 'f' },
	{ "help",         0, NULL, 'h' }, { "mode",         1, NULL, 'm' },
	{ "output",       1, NULL, 'o' }, { "oversamp",     1, NULL, 'O' },
	{ "quiet",        0, NULL, 'q' }, { "refresh-rate", 1, NULL, 'R' },
	{ "symrate",      1, NULL, 'r' }, { "stdout",       0, NULL, OPT_STDOUT },
	{ "samplerate",   1, NULL, 's' }, { "bps",          1, NULL, 'S' },
	{ "version",      0, NULL, 'v' }, { "device",       1, NULL, OPT_DEVICE },
	{ "ref-compatible-tail", 0, NULL, OPT_REFTAIL }, { "shard", 1, NULL, OPT_SHARD },
	{ "tui",          0, NULL, OPT_TUI },          { "gpus",         1, NULL, OPT_GPUS },
	{ "seed",         1, NULL, OPT_SEED },
	{ NULL, 0, NULL, 0 }
};

static void
usage(const char *pname)
{
	fprintf(stderr, "Usage: %s [options] file_in [file_in2 ...]\n", pname);
	fprintf(stderr,
	        "   -B, --batch             Script-friendly status output (no control characters)\n"
	        "   -m, --mode <mode>       Modulation scheme (default: qpsk, valid modes: qpsk, oqpsk)\n"
	        "   -o, --output <file>     Output decoded symbols to <file>\n"
	        "   -q, --quiet             Do not print status information\n"
	        "   -r, --symrate <rate>    Set the symbol rate to <rate> (default: 72000)\n"
	        "   -R, --refresh-rate <ms> Refresh the status line every <ms> ms (default: 50ms, 2000ms in batch mode)\n"
	        "   -s, --samplerate <samp> Force the input samplerate to <samp> (default: auto)\n"
	        "       --bps <bps>         Force the input bits per sample to <bps> (default: 16)\n"
	        "       --stdout            Write output symbols to stdout (implies -B, -q)\n"
	        "       --device <n>        CUDA device ordinal (default: 0)\n"
	        "       --ref-compatible-tail  Final flush writes the reference's byte count (main.c:321)\n"
	        "       --tui               Full-screen status with constellation plot (default on a terminal without -B)\n"
	        "       --shard <samples>   Offline speed-up for ONE long recording: cut it into chunks of <samples>\n"
	        "                           (e.g. 256k) demodulated side by side on the GPU and joined;\n"
	        "                           statistical parity (the first two chunks are exact), see DESIGN.md\n"
	        "       --gpus <n>          With --shard: spread the chunks over CUDA devices <device> .. <device>+n-1 (state and\n"
	        "                           overlap symbols of the boundary chunks travel over NCCL); same bytes as one GPU\n"
	        "       --seed <n>          With --shard: chunks after the first start at a coarse carrier estimate over their\n"
	        "                           first <n> samples (a power of two, 256..16384) and warm up over 32768 samples\n"
	        "                           instead of repeating the acquisition sweep over 150000\n"
	        "   Several input files are demodulated together as one batch; output goes to <file_in>.s each\n"
	        "\n"
	        "   -h, --help              Print this help screen\n"
	        "   -v, --version           Print version info\n"
	        "\n"
	        "Advanced options:\n"
	        "   -b, --pll-bw <bw>       Set the PLL bandwidth to <bw> (default: 1)\n"
	        "   -d, --freq-delta <freq> Set the maximum carrier deviation to <freq> (default: +-3.5kHz)\n"
	        "   -f, --fir-order <ord>   Set the RRC filter order to <ord> (default: 32)\n"
	        "   -O, --oversamp <mult>   Set the interpolation factor to <mult> (default: 5)\n");
}

/* utils.c:62-85: "137.1M" -> 137100000, result truncated to int before it becomes a float */
static float
human_to_float(const char *human)
{
	const char *suffix;
	float tmp = atof(human);
	int ret;
	for (suffix = human; (*suffix >= '0' && *suffix <= '9') || *suffix == '.'; suffix++)
		;
	switch (*suffix) {
		case 'k': case 'K': ret = tmp*1000; break;
		case 'M': ret = tmp*1000000; break;
		default: ret = tmp; break;
	}
	return ret;
}

static char *
gen_fname(void)                               /* utils.c:8-19 */
{
	static char name[sizeof("LRPT_YYYY_MM_DD_HH_MM.s") + 1];
	time_t t = time(NULL);
	strftime(name, sizeof(name), "LRPT_%Y_%m_%d-%H_%M.s", localtime(&t));
	return name;
}

struct wave_header {                          /* wavfile.c:16-31 */
	char riff[4]; uint32_t chunk_size; char wave[4];
	char fmt[4]; uint32_t subchunk_size; uint16_t audio_format, num_channels;
	uint32_t sample_rate, byte_rate; uint16_t block_align, bits_per_sample;
	char data[4]; uint32_t subchunk2_size;
};

/* wavfile.c:34-49 for the canonical 44-byte header -- the only form the reference understands, and for it
 * this function consumes exactly the bytes the reference consumes. Anything else that IS a RIFF/WAVE file
 * (an 18- or 40-byte `fmt ` chunk, WAVE_FORMAT_EXTENSIBLE, `LIST`/`fact`/`bext` chunks before `data`) the
 * reference would take for canonical and demodulate the rest of the header as samples, shifting I against Q;
 * here the chunks are walked properly (sequential reads only, so it works on a pipe) and the stream is left
 * at the first byte of the `data` payload. Returns 0 on success, 1 if this is not a 2-channel WAV. */
static int
wav_parse(FILE *fd, int *samplerate, int *bps)
{
	struct wave_header h;
	uint8_t ck[8], fmt[40];
	uint32_t size;
	int have_fmt = 1;
	if (!fread(&h, sizeof(h), 1, fd)) return 1;
	if (strncmp(h.riff, "RIFF", 4) || strncmp(h.wave, "WAVE", 4)) return 1;
	if (!strncmp(h.fmt, "fmt ", 4) && h.subchunk_size == 16 && !strncmp(h.data, "data", 4)) {   /* canonical */
		if (h.num_channels != 2 || !(*bps = h.bits_per_sample)) return 1;
		*samplerate = (int)h.sample_rate;
		return 0;
	}
	/* general RIFF walk; 44 bytes are already consumed: re-read them from the struct */
	{
		const uint8_t *raw = (const uint8_t *)&h;
		size_t pos = 12;                                       /* first chunk header */
		uint16_t channels = 0, bits = 0; uint32_t rate = 0;
		have_fmt = 0;
		for (;;) {
			size_t i;
			for (i = 0; i < 8; i++, pos++) {                   /* chunk id + size, from the buffer or the stream */
				if (pos < sizeof(h)) ck[i] = raw[pos];
				else { int c = fgetc(fd); if (c == EOF) return 1; ck[i] = (uint8_t)c; }
			}
			size = (uint32_t)ck[4] | (uint32_t)ck[5] << 8 | (uint32_t)ck[6] << 16 | (uint32_t)ck[7] << 24;
			if (!memcmp(ck, "data", 4)) {
				if (pos < sizeof(h)) return 1;                 /* payload would start inside what was read as header: not handled */
				break;
			}
			{
				const size_t padded = (size_t)size + (size & 1);
				for (i = 0; i < padded; i++, pos++) {
					int c;
					if (pos < sizeof(h)) c = raw[pos];
					else if ((c = fgetc(fd)) == EOF) return 1;
					if (!memcmp(ck, "fmt ", 4) && i < sizeof(fmt)) fmt[i] = (uint8_t)c;
				}
			}
			if (!memcmp(ck, "fmt ", 4) && size >= 16) {
				channels = (uint16_t)(fmt[2] | fmt[3] << 8);
				rate = (uint32_t)fmt[4] | (uint32_t)fmt[5] << 8 | (uint32_t)fmt[6] << 16 | (uint32_t)fmt[7] << 24;
				bits = (uint16_t)(fmt[14] | fmt[15] << 8);
				have_fmt = 1;
			}
		}
		if (!have_fmt || channels != 2 || !bits) return 1;
		*bps = bits; *samplerate = (int)rate;
	}
	return 0;
}

static double
now_ms(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec*1e3 + ts.tv_nsec*1e-6;
}

/* Soft symbols leave in 512-symbol blocks, a block only if the PLL had locked once when it completed;
 * the last partial block always (main.c:305-323). One of these per output file. */
struct egress {
	FILE *out;
	int8_t ring[2*RINGSIZE];                               /* the partial block carried between calls */
	int8_t prev_block[2*RINGSIZE];                         /* last complete block, for --ref-compatible-tail */
	size_t ring_idx;                                       /* int8 values in `ring`, as in main.c:291 */
	unsigned long long nsym_total, bytes_out;
};

static void
egress_push(struct egress *e, const int8_t *soft, size_t nsym, long long first_lock)
{
	size_t i = 0;
	while (i < 2*nsym) {
		size_t take = 2*RINGSIZE - e->ring_idx;
		if (take > 2*nsym - i) take = 2*nsym - i;
		memcpy(e->ring + e->ring_idx, soft + i, take);
		e->ring_idx += take; i += take;
		if (e->ring_idx == 2*RINGSIZE) {
			const unsigned long long block_last = e->nsym_total + i/2 - 1;      /* index of the block's last symbol */
			if (first_lock >= 0 && (unsigned long long)first_lock <= block_last) {
				fwrite(e->ring, RINGSIZE, 2, e->out);
				e->bytes_out += 2*RINGSIZE;
			}
			memcpy(e->prev_block, e->ring, sizeof(e->ring));
			e->ring_idx = 0;
		}
	}
	e->nsym_total += nsym;
}

static void
egress_finish(struct egress *e, int ref_tail)
{
	/* final flush: not lock-gated (main.c:321) */
	fwrite(e->ring, 1, e->ring_idx, e->out);
	e->bytes_out += e->ring_idx;
	if (ref_tail && e->ring_idx) {
		/* the reference writes ring_idx more bytes: stale ring content, then (beyond the ring) out of bounds */
		size_t k;
		for (k = e->ring_idx; k < 2*e->ring_idx; k++) fputc(k < 2*RINGSIZE ? e->prev_block[k] : 0, e->out);
	}
}

/* Whole 32 KiB blocks only: a trailing partial block is never consumed (wavfile.c:55). Returns the bytes
 * usable, at most `want` (a multiple of the block size). The first block is waited for; after that only
 * what has ALREADY arrived is taken (a partial block waits in `carry` for the next call), so a live source
 * (rtl_sdr | lrpt_demod -, README.md:75 of the reference) is demodulated block by block as it comes in --
 * 71 ms of signal at 230 kS/s 8-bit, like the reference's own 32 KiB reads (wavfile.c:8,55) -- while a file
 * or a fast pipe still fills whole slabs. The stream must be unbuffered (setvbuf _IONBF): read() is used. */
struct reader {
	FILE *in;
	int eof;
	size_t ncarry;
	uint8_t carry[FILE_BLOCK];
};

static size_t
read_blocks(struct reader *r, uint8_t *dst, size_t want)
{
	const int fd = fileno(r->in);
	size_t got = r->ncarry;
	memcpy(dst, r->carry, got);
	r->ncarry = 0;
	while (!r->eof && got < want) {
		if (got >= FILE_BLOCK) {
			struct pollfd p = { fd, POLLIN, 0 };
			if (poll(&p, 1, 0) <= 0) break;                     /* nothing more right now */
		}
		ssize_t n = read(fd, dst + got, want - got);
		if (n < 0 && errno == EINTR) continue;
		if (n <= 0) { r->eof = 1; break; }
		got += (size_t)n;
	}
	const size_t whole = got/FILE_BLOCK*FILE_BLOCK;
	if (!r->eof) { r->ncarry = got - whole; memcpy(r->carry, dst + whole, r->ncarry); }   /* else: the partial block is dropped */
	return whole;
}

/* ------------------------------------------------------------------ status panes ----
 * The reference's interactive mode (main.c:222-245) redraws four ncurses panes every -R ms from the demodulator's
 * statics and the symbol ring: tui_update_file_in, tui_update_data_out, tui_update_pll, tui_draw_constellation
 * (tui.c:139-247). Here the same panes are fed from the snapshot lrpt_status returns after every block and from
 * the last 512 symbols the host holds anyway (the egress ring), and drawn with ANSI escapes -- ncurses is not a
 * dependency of this host. The constellation follows tui.c:165-201: cell (rows/2 - Q*rows/255, cols/2 + I*cols/255),
 * density marks . - + # for 1, 2, 3, 4+ symbols in a cell, axes through the middle. */
#define TUI_ROWS 21
#define TUI_COLS 43

static void
seconds_to_str(char *dst, size_t n, unsigned long long secs)
{
	snprintf(dst, n, "%02llu:%02llu:%02llu", secs/3600, secs/60%60, secs%60);
}

static void
tui_render(FILE *f, int ansi, const char *in_name, unsigned in_rate_bytes, unsigned long long done, unsigned long long total,
           unsigned long long bytes_out, const lrpt_status_t *st, float freq_hz, float rate_hz,
           const int8_t *dots, size_t ndots)
{
	static const char marks[] = " .-+#";
	unsigned char grid[TUI_ROWS][TUI_COLS];
	char t_done[32], t_total[32];
	size_t i;
	int r, c;
	memset(grid, 0, sizeof(grid));
	for (i = 0; i + 1 < 2*ndots; i += 2) {
		const int x = dots[i]*TUI_COLS/255, y = dots[i + 1]*TUI_ROWS/255;
		r = TUI_ROWS/2 - y; c = x + TUI_COLS/2;
		if (r >= 0 && r < TUI_ROWS && c >= 0 && c < TUI_COLS && grid[r][c] < 4) grid[r][c]++;
	}
	seconds_to_str(t_done, sizeof(t_done), in_rate_bytes ? done/in_rate_bytes : 0);
	seconds_to_str(t_total, sizeof(t_total), in_rate_bytes ? total/in_rate_bytes : 0);
	if (ansi) fputs("\033[H\033[J", f);                     /* home, clear */
	fprintf(f, "+- File in ------------------------------+\n");
	fprintf(f, "  %s\n  %s / %s  (%5.1f%%)\n", in_name, t_done, t_total, total ? 100.0*done/total : 0.0);
	fprintf(f, "+- Data out -----------------------------+\n");
	fprintf(f, "  %llu bytes\n", bytes_out);
	fprintf(f, "+- PLL ----------------------------------+\n");
	fprintf(f, "  %s\n", st->locked ? "Locked" : "Acquiring...");
	fprintf(f, "  Gain\tCarrier freq\tSymbol rate\n  %.3f\t%+7.1f Hz\t%7.1f Hz\n", st->agc_gain, freq_hz, rate_hz);
	fprintf(f, "+- Constellation ------------------------+\n");
	for (r = 0; r < TUI_ROWS; r++) {
		fputs("  ", f);
		for (c = 0; c < TUI_COLS; c++) {
			char ch = marks[grid[r][c]];
			if (ch == ' ') ch = (r == TUI_ROWS/2 && c == TUI_COLS/2) ? '+' : r == TUI_ROWS/2 ? '-' : c == TUI_COLS/2 ? '|' : ' ';
			fputc(ch, f);
		}
		fputc('\n', f);
	}
	fflush(f);
}

/* ------------------------------------------------------------------ double-buffered ingest ----
 * A reader thread fills one page-locked slab (read_blocks: whole 32 KiB blocks, live sources block by block)
 * while the demodulator works on the other, so file / pipe reads overlap the GPU instead of alternating with it. */
struct ingest {
	struct reader rd;
	uint8_t *slab[2];
	size_t have[2], slab_bytes;
	int filled[2], done;                                    /* filled[i]: slab i holds data the consumer has not taken */
	pthread_mutex_t mu;
	pthread_cond_t cv;
	pthread_t tid;
};

static void *
ingest_thread(void *arg)
{
	struct ingest *g = arg;
	int i = 0;
	for (;;) {
		pthread_mutex_lock(&g->mu);
		while (g->filled[i]) pthread_cond_wait(&g->cv, &g->mu);
		pthread_mutex_unlock(&g->mu);
		const size_t n = read_blocks(&g->rd, g->slab[i], g->slab_bytes);
		pthread_mutex_lock(&g->mu);
		g->have[i] = n; g->filled[i] = 1;
		if (!n || g->rd.eof) g->done = 1;
		pthread_cond_broadcast(&g->cv);
		pthread_mutex_unlock(&g->mu);
		if (!n || g->rd.eof) return NULL;
		i ^= 1;
	}
}

/* next filled slab (blocks until the reader has one); 0 bytes = end of input */
static size_t
ingest_take(struct ingest *g, int i, uint8_t **data)
{
	pthread_mutex_lock(&g->mu);
	while (!g->filled[i]) pthread_cond_wait(&g->cv, &g->mu);
	pthread_mutex_unlock(&g->mu);
	*data = g->slab[i];
	return g->have[i];
}

/* hands slab i back to the reader; returns 1 when no further slab will come (the reader has stopped and the
 * other slab holds nothing) */
static int
ingest_release(struct ingest *g, int i)
{
	int last;
	pthread_mutex_lock(&g->mu);
	last = g->done && !g->filled[i ^ 1];
	g->filled[i] = 0;
	pthread_cond_broadcast(&g->cv);
	pthread_mutex_unlock(&g->mu);
	return last;
}

static void *
host_alloc(size_t n, int *pinned)
{
	void *p = lrpt_alloc_host(n);                          /* page-locked: the host link's full speed */
	*pinned = p != NULL;
	return p ? p : malloc(n);
}

static void
host_free(void *p, int pinned)
{
	if (pinned) lrpt_free_host(p); else free(p);
}

/* Several recordings as the streams of one batch. All share the command line's settings and must agree
 * on sample rate and sample format; lengths may differ (a finished stream idles on silence, its output
 * is no longer written). Output of <input> goes to <input>.s. */
static int
run_batch(int nfiles, char **names, lrpt_params_t p, int samplerate_opt, int bps_opt, float symrate, int quiet, int ref_tail)
{
	struct input { struct reader rd; int open; size_t avail; struct egress eg; lrpt_status_t last; } *f = calloc((size_t)nfiles, sizeof(*f));
	int samplerate = -1, bps = 0, i, rc, active = nfiles;
	char path[4096];
	if (!f) { fprintf(stderr, "out of memory\n"); return 1; }
	for (i = 0; i < nfiles; i++) {
		int sr = samplerate_opt, b = bps_opt;
		if (!(f[i].rd.in = fopen(names[i], "rb"))) { fprintf(stderr, "Could not open input file\n"); return 1; }
		f[i].open = 1;
		setvbuf(f[i].rd.in, NULL, _IONBF, 0);
		if (wav_parse(f[i].rd.in, &sr, &b)) fseek(f[i].rd.in, 0, SEEK_SET);
		if (sr < 0) { fprintf(stderr, "Could not auto-detect sample rate. Please specify it with -s <samplerate>\n"); return 1; }
		if (!b) { fprintf(stderr, "Could not auto-detect bits per sample, assuming 16\n"); b = 16; }
		if (i && (sr != samplerate || b != bps)) {
			fprintf(stderr, "%s: %d Hz / %d bits differs from %s (%d Hz / %d bits); one batch needs one format\n",
			        names[i], sr, b, names[0], samplerate, bps);
			return 1;
		}
		samplerate = sr; bps = b;
		snprintf(path, sizeof(path), "%s.s", names[i]);
		if (!(f[i].eg.out = fopen(path, "wb"))) { fprintf(stderr, "Could not open output file\n"); return 1; }
	}
	if (bps != 8 && bps != 16 && bps != 32) { fprintf(stderr, "Unsupported bits per sample: %d\n", bps); return 1; }
	p.samplerate = samplerate; p.bps = bps; p.nstreams = nfiles;
	lrpt_demod_t *h = NULL;
	if ((rc = lrpt_create(&h, &p))) { fprintf(stderr, "lrpt_create failed: %s\n", lrpt_strerror(rc)); return 1; }

	const size_t bytes_per_sample = (size_t)bps/4;
	size_t slab_bytes = (size_t)SLAB_BLOCKS*FILE_BLOCK;
	while (slab_bytes > FILE_BLOCK && slab_bytes*(size_t)nfiles > ((size_t)1 << 31)) slab_bytes /= 2;   /* <= 2 GiB of input per call */
	const size_t cap = slab_bytes/bytes_per_sample;
	uint8_t *raw = calloc((size_t)nfiles, slab_bytes);
	int8_t *soft = malloc((size_t)nfiles*2*cap);
	uint32_t *nsym = calloc((size_t)nfiles, sizeof(*nsym));
	if (!raw || !soft || !nsym) { fprintf(stderr, "out of memory\n"); return 1; }

	while (active) {
		/* refill: every stream still open reads up to one slab; the call covers what ALL of them have */
		size_t common = slab_bytes;
		for (i = 0; i < nfiles; i++) {
			if (!f[i].open) continue;
			if (!f[i].avail && !f[i].rd.eof) f[i].avail = read_blocks(&f[i].rd, raw + (size_t)i*slab_bytes, slab_bytes);
			if (!f[i].avail) {                                  /* finished: final flush, then silence */
				egress_finish(&f[i].eg, ref_tail);
				fclose(f[i].eg.out); fclose(f[i].rd.in); f[i].open = 0; active--;
				memset(raw + (size_t)i*slab_bytes, bps == 8 ? 128 : 0, slab_bytes);
				continue;
			}
			if (f[i].avail < common) common = f[i].avail;
		}
		if (!active) break;
		rc = lrpt_process_batch(h, raw, slab_bytes, common/bytes_per_sample, soft, 2*cap, cap, nsym, NULL, 0);
		if (rc) { fprintf(stderr, "lrpt_process_batch failed: %s (%s)\n", lrpt_strerror(rc), lrpt_last_error(h)); return 1; }
		for (i = 0; i < nfiles; i++) {
			if (!f[i].open) continue;
			lrpt_status(h, i, &f[i].last);                      /* the stream's state while it still has input */
			egress_push(&f[i].eg, soft + (size_t)i*2*cap, nsym[i], f[i].last.first_lock_symbol);
			f[i].avail -= common;                               /* keep the unconsumed rest at the front of the row */
			if (f[i].avail) memmove(raw + (size_t)i*slab_bytes, raw + (size_t)i*slab_bytes + common, f[i].avail);
		}
	}
	if (!quiet) {
		for (i = 0; i < nfiles; i++)
			printf("%s -> %s.s: %llu symbols, %llu bytes, Carrier: %+7.1f Hz, Locked: %s\n", names[i], names[i],
			       f[i].eg.nsym_total, f[i].eg.bytes_out, f[i].last.pll_freq*symrate/(2*M_PI)*(p.oqpsk ? 2 : 1),
			       f[i].last.locked ? "Yes" : "No");
	}
	lrpt_destroy(h);
	free(raw); free(soft); free(nsym); free(f);
	return 0;
}

/* --shard: the whole recording (whole 32 KiB blocks of it, wavfile.c:55) in memory, one lrpt_sharded_process
 * call, then the reference's egress rules (main.c:305-323) on the joined symbols. */
static int
run_sharded(FILE *in, FILE *out, const lrpt_params_t *p, size_t chunk, float symrate, int quiet, int ref_tail, int gpus, int seed)
{
	size_t have = 0, room = (size_t)64 << 20;
	int pinned = 0, pin_soft = 0;
	uint8_t *raw = NULL;
	struct reader *rd = calloc(1, sizeof(*rd));
	if (!rd) { fprintf(stderr, "out of memory\n"); return 1; }
	rd->in = in;
	{
		/* a seekable input's length is known: read it straight into ONE page-locked buffer (the H2D copy of a
		 * 4.3 GB recording takes 80 ms from page-locked memory, 470 ms from malloc memory) */
		const long pos = ftell(in);
		if (pos >= 0 && !fseek(in, 0, SEEK_END)) {
			const long end = ftell(in);
			fseek(in, pos, SEEK_SET);
			if (end > pos) {
				room = (size_t)(end - pos) + FILE_BLOCK;
				raw = lrpt_alloc_host(room);
				pinned = raw != NULL;
			}
		}
	}
	if (!raw) raw = malloc(room);
	while (raw && !rd->eof) {
		if (room - have < (size_t)FILE_BLOCK) {
			if (pinned) break;                                  /* cannot happen: the buffer holds the whole file */
			uint8_t *grown = realloc(raw, room *= 2);
			if (!grown) { free(raw); raw = NULL; break; }
			raw = grown;
		}
		size_t want = room - have;
		if (want > (size_t)FILE_BLOCK*SLAB_BLOCKS) want = (size_t)FILE_BLOCK*SLAB_BLOCKS;
		want = want/FILE_BLOCK*FILE_BLOCK;
		const size_t got = read_blocks(rd, raw + have, want);
		if (!got && !rd->eof) continue;
		have += got;
	}
	if (!raw) { fprintf(stderr, "out of memory\n"); return 1; }
	const size_t nsamples = have/((size_t)p->bps/4);
	const size_t cap = (size_t)((double)nsamples*p->symrate/p->samplerate*1.02) + 64;
	int8_t *soft = host_alloc(2*cap, &pin_soft);
	if (!soft) { fprintf(stderr, "out of memory\n"); return 1; }
	lrpt_shard_plan_t plan = { (chunk + 7)/8*8, seed ? 32768 : 150000, 8192, (uint64_t)(seed > 0 ? seed : 0) };
	lrpt_shard_report_t rep;
	size_t nsym = 0;
	int rc, devs[64], g;
	if (gpus < 1) gpus = 1;
	if (gpus > 64) gpus = 64;
	for (g = 0; g < gpus; g++) devs[g] = p->device + g;
	rc = gpus > 1 ? lrpt_sharded_process_multi(p, &plan, raw, nsamples, soft, cap, &nsym, &rep, devs, gpus)
	              : lrpt_sharded_process(p, &plan, raw, nsamples, soft, cap, &nsym, &rep);
	if (rc) { fprintf(stderr, "lrpt_sharded_process failed: %s\n", lrpt_strerror(rc)); return 1; }
	struct egress eg;
	memset(&eg, 0, sizeof(eg));
	eg.out = out;
	egress_push(&eg, soft, nsym, rep.first_lock_symbol);
	egress_finish(&eg, ref_tail);
	if (!quiet)
		printf("(100.0%%) %zu samples in %d chunks, %zu symbols (%.1f Hz nominal), worst boundary agreement %.4f%s, Locked: %s\n",
		       nsamples, rep.nchunks, nsym, symrate, rep.min_agreement_final, rep.aligned ? "" : " (a chunk kept another lock point)",
		       rep.first_lock_symbol >= 0 ? "Yes" : "No");
	host_free(raw, pinned); host_free(soft, pin_soft);
	free(rd);
	return 0;
}

int
main(int argc, char *argv[])
{
	float pll_bw = 1, symrate = 72000.0f, freq_max_delta = -1;
	int rrc_order = 32, interp_factor = 5, quiet = 0, oqpsk = 0, batch = 0;
	int update_interval = -1, bps = 0, samplerate = -1, stdout_mode = 0, device = 0, ref_tail = 0, tui = -1, gpus = 1, seed = 0;
	size_t shard = 0;
	char *output_fname = NULL;
	FILE *in, *out;
	int c;

	while ((c = getopt_long(argc, argv, SHORTOPTS, longopts, NULL)) != -1) {
		switch (c) {
			case OPT_STDOUT: stdout_mode = 1; break;
			case OPT_DEVICE: device = atoi(optarg); break;
			case OPT_REFTAIL: ref_tail = 1; break;
			case OPT_TUI: tui = 1; break;
			case OPT_GPUS: gpus = atoi(optarg); break;
			case OPT_SEED: seed = atoi(optarg); break;
			case OPT_SHARD: shard = (size_t)human_to_float(optarg); break;
			case 'b': pll_bw = human_to_float(optarg); break;
			case 'B': batch = 1; break;
			case 'd': freq_max_delta = human_to_float(optarg); break;
			case 'f': rrc_order = atoi(optarg); break;
			case 'h': usage(argv[0]); return 0;
			case 'm': if (!strcmp(optarg, "oqpsk")) oqpsk = 1; break;      /* anything else: qpsk, main.c:103-105 */
			case 'o': output_fname = optarg; break;
			case 'O': interp_factor = atoi(optarg); break;
			case 'q': quiet = 1; break;
			case 'R': update_interval = atoi(optarg); break;
			case 'r': symrate = human_to_float(optarg); break;
			case 's': samplerate = human_to_float(optarg); break;
			case 'S': bps = atoi(optarg); break;
			case 'v': fprintf(stderr, "lrpt_demod (meteor_demod compatible) v" VERSION "\n"); return 0;
			default: usage(argv[0]); return 1;
		}
	}
	freq_max_delta = lrpt_freq_delta_from_hz(freq_max_delta, symrate);           /* main.c:136 */
	if (argc - optind < 1) { usage(argv[0]); return 1; }
	if (!output_fname) output_fname = gen_fname();
	if (update_interval < 0) update_interval = batch ? 2000 : 50;
	if (stdout_mode) { batch = 1; quiet = 1; }

	if (argc - optind > 1) {                                 /* several recordings: one batch, <input>.s each */
		lrpt_params_t bp;
		if (stdout_mode) { fprintf(stderr, "--stdout needs a single input\n"); return 1; }
		memset(&bp, 0, sizeof(bp));
		bp.pll_bw = pll_bw; bp.sym_bw = 0.00005f; bp.freq_max = freq_max_delta;
		bp.symrate = (int)symrate; bp.interp_factor = interp_factor; bp.rrc_order = rrc_order; bp.oqpsk = oqpsk;
		bp.device = device; bp.kernel = LRPT_KERNEL_AUTO;
		return run_batch(argc - optind, argv + optind, bp, samplerate, bps, symrate, quiet, ref_tail);
	}

	if (!strcmp(argv[optind], "-")) { in = stdin; batch = 1; }
	else if (!(in = fopen(argv[optind], "rb"))) { fprintf(stderr, "Could not open input file\n"); return 1; }

	setvbuf(in, NULL, _IONBF, 0);                                    /* no read-ahead: poll() in read_blocks sees the pipe itself */
	if (wav_parse(in, &samplerate, &bps)) fseek(in, 0, SEEK_SET);    /* fails silently on a pipe, as in the reference */
	if (samplerate < 0) {
		fprintf(stderr, "Could not auto-detect sample rate. Please specify it with -s <samplerate>\n");
		usage(argv[0]);
		return 1;
	}
	if (!bps) { fprintf(stderr, "Could not auto-detect bits per sample, assuming 16\n"); bps = 16; }
	if (bps != 8 && bps != 16 && bps != 32) { fprintf(stderr, "Unsupported bits per sample: %d\n", bps); return 1; }

	if (stdout_mode) out = stdout;
	else if (!(out = fopen(output_fname, "wb"))) { fprintf(stderr, "Could not open output file\n"); return 1; }

	lrpt_params_t p;
	memset(&p, 0, sizeof(p));
	p.pll_bw = pll_bw; p.sym_bw = 0.00005f; p.freq_max = freq_max_delta;
	p.samplerate = samplerate; p.symrate = (int)symrate;                     /* float -> int, main.c:187 */
	p.interp_factor = interp_factor; p.rrc_order = rrc_order; p.oqpsk = oqpsk; p.bps = bps;
	p.device = device; p.nstreams = 1; p.kernel = LRPT_KERNEL_AUTO;
	if (shard) {
		if (!quiet) printf("Input: %s, output: %s\n", argv[optind], output_fname);
		int src = run_sharded(in, out, &p, shard, symrate, quiet, ref_tail, gpus, seed);
		if (out != stdout) fclose(out);
		if (in != stdin) fclose(in);
		return src;
	}
	lrpt_demod_t *h = NULL;
	int rc = lrpt_create(&h, &p);
	if (rc) { fprintf(stderr, "lrpt_create failed: %s\n", lrpt_strerror(rc)); return 1; }

	unsigned long file_len = 0;
	{
		long pos = ftell(in);
		if (pos >= 0 && !fseek(in, 0, SEEK_END)) { long e = ftell(in); file_len = e > 0 ? (unsigned long)e : 0; fseek(in, pos, SEEK_SET); }
	}
	if (!quiet) { printf("Input: %s, output: %s\n", argv[optind], output_fname); printf("Demodulator initialized\n"); }

	const size_t slab_bytes = (size_t)SLAB_BLOCKS*FILE_BLOCK;
	const size_t bytes_per_sample = (size_t)bps/4;
	const size_t slab_samples = slab_bytes/bytes_per_sample;
	const size_t cap = slab_samples;                       /* a symbol needs at least one sample in any sane setup */
	int pin_soft = 0, pin_raw[2] = {0, 0};
	int8_t *soft = host_alloc(2*cap, &pin_soft);
	struct egress eg;
	unsigned long long bytes_in = 0;
	long long first_lock = -1;
	double last_status = now_ms();
	memset(&eg, 0, sizeof(eg));
	eg.out = out;
	if (tui < 0) tui = !batch && !quiet && isatty(fileno(stdout));   /* main.c:222: interactive unless -B */
	if (tui && out == stdout) tui = 0;

	/* double-buffered ingest into page-locked slabs (reader thread) */
	struct ingest *ing = calloc(1, sizeof(*ing));
	if (!ing || !soft) { fprintf(stderr, "out of memory\n"); return 1; }
	ing->rd.in = in;
	ing->slab_bytes = slab_bytes;
	ing->slab[0] = host_alloc(slab_bytes, &pin_raw[0]);
	ing->slab[1] = host_alloc(slab_bytes, &pin_raw[1]);
	if (!ing->slab[0] || !ing->slab[1]) { fprintf(stderr, "out of memory\n"); return 1; }
	pthread_mutex_init(&ing->mu, NULL);
	pthread_cond_init(&ing->cv, NULL);
	if (pthread_create(&ing->tid, NULL, ingest_thread, ing)) { fprintf(stderr, "could not start the reader thread\n"); return 1; }
	for (int cur = 0;; cur ^= 1) {
		uint8_t *raw = NULL;
		const size_t use = ingest_take(ing, cur, &raw);
		if (!use) break;
		size_t nsym = 0;
		rc = lrpt_process(h, raw, use/bytes_per_sample, soft, cap, &nsym, &first_lock);
		const int last = ingest_release(ing, cur);                  /* the reader stopped after this slab */
		if (rc) { fprintf(stderr, "lrpt_process failed: %s (%s)\n", lrpt_strerror(rc), lrpt_last_error(h)); return 1; }
		bytes_in += use;
		egress_push(&eg, soft, nsym, first_lock);

		if (!quiet && (now_ms() - last_status >= update_interval || (tui && last))) {
			lrpt_status_t st;
			lrpt_status(h, 0, &st);                                  /* the snapshot after this block */
			const float freq_hz = st.pll_freq*symrate/(2*M_PI)*(oqpsk ? 2 : 1);          /* main.c:250 */
			const float rate_hz = st.mm_omega*(samplerate*interp_factor)/(2*M_PI);       /* main.c:251 */
			if (tui) {
				/* the constellation ring: the last 512 symbols (main.c:34,243) = end of this block's output */
				const size_t nd = nsym < RINGSIZE ? nsym : RINGSIZE;
				tui_render(stdout, isatty(fileno(stdout)), argv[optind], 2*samplerate*bps/8, bytes_in, file_len, eg.bytes_out,
				           &st, freq_hz, rate_hz, soft + 2*(nsym - nd), nd);
			} else {
				printf(batch ? "\n" : "\033[1K\r");
				printf("(%5.1f%%) Carrier: %+7.1f Hz, Symbol rate: %.1f Hz, Locked: %s",
				       file_len ? 100.0*bytes_in/file_len : 0, freq_hz, rate_hz, st.locked ? "Yes" : "No");
				fflush(stdout);
			}
			last_status = now_ms();
		}
		if (last) break;
	}
	pthread_join(ing->tid, NULL);

	egress_finish(&eg, ref_tail);
	if (!quiet) {
		lrpt_status_t st;
		lrpt_status(h, 0, &st);
		printf(batch ? "\n" : "\033[1K\r");
		printf("(%5.1f%%) Carrier: %+7.1f Hz, Symbol rate: %.1f Hz, Locked: %s\n",
		       file_len ? 100.0*bytes_in/file_len : 0, st.pll_freq*symrate/(2*M_PI)*(oqpsk ? 2 : 1),
		       st.mm_omega*(samplerate*interp_factor)/(2*M_PI), st.locked ? "Yes" : "No");
	}

	lrpt_destroy(h);
	host_free(ing->slab[0], pin_raw[0]); host_free(ing->slab[1], pin_raw[1]); host_free(soft, pin_soft);
	free(ing);
	if (out != stdout) fclose(out);
	if (in != stdin) fclose(in);
	return 0;
}
