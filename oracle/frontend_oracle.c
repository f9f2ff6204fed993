/*
 * oracle/frontend_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the LRPT decoder FRONT-END, the consumer of the demodulator's output: the 8-bit soft-symbol
 * stream meteor_demod writes (main.c:305-313: int8 I, int8 Q per symbol, no header) is what
 * dbdexter-dev/meteor_decode (and artlav's medet before it) reads (README.md:6-9,87-91 of the reference).
 * That program is NOT part of /root/reference (a separate repository, no pinned version in the reference; the
 * reference only names it), so this file restates the PUBLISHED algorithm of that step, from the standards the
 * signal follows, and is pinned by their known answers rather than by that program's bytes:
 *
 *   CCSDS 131.0-B (TM synchronization and channel coding) as used by Meteor-M LRPT:
 *     - attached sync marker ASM = 0x1ACFFC1D in front of every 1020-byte transfer frame (CADU = 1024 bytes);
 *     - rate 1/2, constraint length 7 convolutional code, G1 = 171 (octal) on the I arm, G2 = 133 (octal) on the
 *       Q arm, no symbol inversion, one QPSK symbol per input bit (I carries G1's output, Q carries G2's):
 *       a CADU is 8192 symbols = 16384 soft values;
 *     - known answer: the ASM encoded from the all-zero state is the 64-bit pattern 0x035D49C24FF2686B, and its
 *       four quarter-turn images {0xFCA2B63DB00D9794, 0x56FBD394DAA4C1C2, 0x035D49C24FF2686B,
 *       0xA9042C6B255B3E3D} are the constants LRPT decoders correlate against (tests/test_frontend.py).
 *
 *   1. frame synchronisation: hard decisions of the soft stream, 64-bit window per symbol offset, XOR + popcount
 *      against the encoded ASM under the 8 symmetries of the constellation (4 quarter turns x I/Q swap, which is
 *      what a Costas loop's phase ambiguity and a spectrally inverted receiver leave open): score = matching
 *      bits (0..64), best symmetry per offset; the best offset per window of one CADU;
 *   2. Viterbi decoding of one CADU from its symbol offset under its symmetry: 64 states, correlation branch
 *      metrics on the int8 soft values (saturating nothing: int32 path metrics), add-compare-select with ties to
 *      the predecessor with the OLDER bit 0, full traceback from the best end state over the frame plus
 *      FE_TAIL symbols of the following frame; the first FE_HEAD symbols before the frame warm the metrics up.
 *
 * The CUDA path (csrc/frontend.cu) must reproduce these functions bit for bit (integer work).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FE_ASM        0x1ACFFC1Du
#define FE_G1         0x4F                  /* 171 octal with the newest bit in the LSB: 1001111 */
#define FE_G2         0x6D                  /* 133 octal: 1101101 */
#define FE_CADU       1024                  /* bytes */
#define FE_CADU_SYMS  8192                  /* QPSK symbols = input bits of one CADU */
#define FE_HEAD       64                    /* symbols decoded before the frame (metric warm-up) */
#define FE_TAIL       64                    /* symbols decoded after the frame (traceback convergence) */

static int parity7(unsigned x) { x ^= x >> 4; x ^= x >> 2; x ^= x >> 1; return (int)(x & 1u); }

/* Convolutional encoder: nbits input bits (MSB first in each byte) -> 2*nbits output bits, one byte each (0/1),
 * I arm (G1) first. *state carries the 7-bit register (newest bit = LSB) across calls. */
void
fe_conv_encode(const uint8_t *in, long nbits, uint8_t *out, unsigned *state)
{
	unsigned reg = *state;
	long i;
	for (i=0; i<nbits; i++) {
		const unsigned b = (in[i >> 3] >> (7 - (i & 7))) & 1u;
		reg = ((reg << 1) | b) & 0x7Fu;
		out[2*i]   = (uint8_t)parity7(reg & FE_G1);
		out[2*i+1] = (uint8_t)parity7(reg & FE_G2);
	}
	*state = reg;
}

/* The encoded ASM (64 bits, first output bit = MSB) under symmetry h = swap*4 + quarter turns k:
 * a received symbol (I, Q) of a stream that was turned by k quarter turns (and I/Q-swapped first when swap = 1)
 * carries hard bits (1 <=> soft value >= 0) that equal this pattern where the untouched stream would carry the
 * plain one. */
uint64_t
fe_sync_pattern(int h)
{
	uint8_t asm_bytes[4] = { 0x1A, 0xCF, 0xFC, 0x1D }, enc[64];
	unsigned st = 0;
	uint64_t v = 0;
	int i, k;
	fe_conv_encode(asm_bytes, 32, enc, &st);
	for (i=0; i<32; i++) {
		unsigned bi = enc[2*i], bq = enc[2*i+1], t;
		if (h & 4) { t = bi; bi = bq; bq = t; }                 /* I/Q swap */
		for (k=0; k<(h & 3); k++) { t = bi; bi = bq ^ 1u; bq = t; }   /* quarter turn: (I,Q) -> (-Q, I) */
		v = (v << 2) | (uint64_t)(bi << 1) | bq;
	}
	return v;
}

static int popcount64(uint64_t x) { int n = 0; while (x) { x &= x - 1; n++; } return n; }

/* Per symbol offset o in [0, nsym-32]: score[o] = most matching bits over the 8 symmetries, hyp[o] = the symmetry
 * (lowest index on ties). Hard bit = 1 when the soft value is >= 0. Offsets beyond nsym-32 get score 0, hyp 0. */
void
fe_sync_scores(const int8_t *soft, long nsym, uint8_t *score, uint8_t *hyp)
{
	uint64_t pat[8], w = 0;
	long o;
	int h;
	for (h=0; h<8; h++) pat[h] = fe_sync_pattern(h);
	for (o=0; o<nsym; o++) { score[o] = 0; hyp[o] = 0; }
	for (o=0; o<nsym; o++) {
		w = (w << 2) | (uint64_t)((soft[2*o] >= 0) << 1) | (uint64_t)(soft[2*o+1] >= 0);
		if (o >= 31) {
			int best = -1, bh = 0;
			for (h=0; h<8; h++) {
				const int s = 64 - popcount64(w ^ pat[h]);
				if (s > best) { best = s; bh = h; }
			}
			score[o-31] = (uint8_t)best; hyp[o-31] = (uint8_t)bh;
		}
	}
}

/* Best offset per window of `window` offsets: off[w] = first offset of the maximum score in
 * [w*window, min((w+1)*window, nsym)). Returns the number of windows. */
long
fe_window_peaks(const uint8_t *score, const uint8_t *hyp, long nsym, long window, uint32_t *off, uint8_t *ohyp, uint8_t *oscore)
{
	long w, nw = (nsym + window - 1)/window;
	for (w=0; w<nw; w++) {
		long o, lo = w*window, hi = lo + window < nsym ? lo + window : nsym, bo = lo;
		for (o=lo; o<hi; o++) if (score[o] > score[bo]) bo = o;
		off[w] = (uint32_t)bo; ohyp[w] = hyp[bo]; oscore[w] = score[bo];
	}
	return nw;
}

/* Undo symmetry h on one received symbol: returns the (I, Q) the untouched stream would carry. */
static void
fe_unturn(int h, int i, int q, int *oi, int *oq)
{
	int k, t;
	for (k=0; k<(h & 3); k++) { t = i; i = q; q = -t; }          /* inverse quarter turn: (I,Q) -> (Q, -I) */
	if (h & 4) { t = i; i = q; q = t; }
	*oi = i; *oq = q;
}

/*
 * Viterbi decoder for ONE CADU that starts at symbol `start` of `soft` (nsym symbols available) under symmetry h.
 * Decodes symbols [start - FE_HEAD, start + 8192 + FE_TAIL) clipped to the stream, all start states equally
 * likely, traceback from the best end state; writes the 1024 bytes of the frame (its first four are the ASM when
 * the frame is really there) and returns the winning path metric. Branch metric of a transition that emits
 * (c1, c2) for received (I, Q): (c1 ? I : -I) + (c2 ? Q : -Q) -- an encoded 1 is a POSITIVE amplitude, the
 * convention of fe_sync_scores (hard bit 1 <=> soft >= 0); a transmitter that maps 1 to a negative amplitude is a
 * stream turned by two quarter turns, which symmetry 2 absorbs.
 */
int32_t
fe_viterbi_cadu(const int8_t *soft, long nsym, long start, int h, uint8_t *cadu)
{
	const long lo = start - FE_HEAD < 0 ? 0 : start - FE_HEAD;
	const long hi = start + FE_CADU_SYMS + FE_TAIL > nsym ? nsym : start + FE_CADU_SYMS + FE_TAIL;
	const long n = hi - lo;
	int32_t pm[64], nm[64];
	uint64_t *dec;
	long t;
	int s, best;
	memset(cadu, 0, FE_CADU);
	if (n <= 0 || start < 0) return 0;
	dec = malloc(sizeof(uint64_t)*(size_t)n);
	if (!dec) return 0;
	for (s=0; s<64; s++) pm[s] = 0;
	for (t=0; t<n; t++) {
		int ri, rq;
		uint64_t d = 0;
		fe_unturn(h, soft[2*(lo+t)], soft[2*(lo+t)+1], &ri, &rq);
		/* state = the last six input bits, newest in bit 0; next = ((state << 1) | b) & 63; the encoder register of
		 * that transition is (state << 1) | b (7 bits) */
		for (s=0; s<64; s++) {
			const int b = s & 1;                                 /* input bit that leads INTO state s */
			const int p0 = s >> 1, p1 = (s >> 1) | 32;           /* predecessors: oldest bit 0 / 1 */
			const unsigned r0 = (unsigned)((p0 << 1) | b), r1 = (unsigned)((p1 << 1) | b);
			const int m0 = pm[p0] + (parity7(r0 & FE_G1) ? ri : -ri) + (parity7(r0 & FE_G2) ? rq : -rq);
			const int m1 = pm[p1] + (parity7(r1 & FE_G1) ? ri : -ri) + (parity7(r1 & FE_G2) ? rq : -rq);
			if (m1 > m0) { nm[s] = m1; d |= (uint64_t)1 << s; } else nm[s] = m0;
		}
		dec[t] = d;
		memcpy(pm, nm, sizeof(pm));
	}
	best = 0;
	for (s=1; s<64; s++) if (pm[s] > pm[best]) best = s;
	{
		const int32_t metric = pm[best];
		s = best;
		for (t=n-1; t>=0; t--) {
			const long sym = lo + t - start;                     /* index of this input bit within the frame */
			if (sym >= 0 && sym < FE_CADU_SYMS && (s & 1)) cadu[sym >> 3] |= (uint8_t)(0x80u >> (sym & 7));
			s = (s >> 1) | (((dec[t] >> s) & 1u) ? 32 : 0);
		}
		free(dec);
		return metric;
	}
}
