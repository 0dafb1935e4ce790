/*
 * ksw.c -- CPU ORACLE (test infrastructure only; see orc.h): the three dynamic
 * programming primitives of bwa-mem, restated in scalar C from their published
 * recurrences (un-vendored lh3/bwa ksw.c reached
 * through jnibwa.c:214).
 *
 *   orc_ksw_extend2 : banded, z-dropped, origin-anchored affine extension
 *   orc_ksw_global2 : banded global alignment with traceback
 *   orc_ksw_align2  : local alignment with the observable semantics of the
 *                     striped SSE2 kernels (segment-local F in the first pass,
 *                     E derived from the pre-correction H, zero-score padding
 *                     columns, u8 saturation, per-row maxima list, reverse pass)
 */
#include "orc_internal.h"

#define MINUS_INF (-0x40000000)

int orc_ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                    int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                    int *qle_, int *tle_, int *gtle_, int *gscore_, int *max_off_)
{
	int32_t *H, *E; /* H[j] = H(i-1, j-1) on entry to row i, E[j] = E(i, j) */
	int i, j, k, oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	int beg, end, max, max_i, max_j, max_ins, max_del, max_ie, gscore, max_off;
	H = calloc(qlen + 1, sizeof(int32_t));
	E = calloc(qlen + 1, sizeof(int32_t));
	H[0] = h0;
	if (qlen >= 1) H[1] = h0 > oe_ins ? h0 - oe_ins : 0;
	for (j = 2; j <= qlen && H[j - 1] > e_ins; ++j) H[j] = H[j - 1] - e_ins;
	for (k = 0, max = 0; k < 25; ++k) max = max > mat[k] ? max : mat[k];
	max_ins = (int)((double)(qlen * max + end_bonus - o_ins) / e_ins + 1.);
	max_ins = max_ins > 1 ? max_ins : 1;
	w = w < max_ins ? w : max_ins;
	max_del = (int)((double)(qlen * max + end_bonus - o_del) / e_del + 1.);
	max_del = max_del > 1 ? max_del : 1;
	w = w < max_del ? w : max_del;
	max = h0; max_i = max_j = -1; max_ie = -1; gscore = -1; max_off = 0;
	beg = 0; end = qlen;
	++orc_tls_cnt.n_ext_calls;
	for (i = 0; i < tlen; ++i) {
		int f = 0, h1, m = 0, mj = -1, t;
		const int8_t *srow = &mat[target[i] * 5];
		if (beg < i - w) beg = i - w;
		if (end > i + w + 1) end = i + w + 1;
		if (end > qlen) end = qlen;
		if (beg == 0) { h1 = h0 - (o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; }
		else h1 = 0;
		if (end > beg) orc_tls_cnt.cells_ext += (uint64_t)(end - beg);
		for (j = beg; j < end; ++j) {
			int h, M = H[j], e = E[j];
			H[j] = h1;
			M = M ? M + srow[query[j]] : 0;
			h = M > e ? M : e;
			h = h > f ? h : f;
			h1 = h;
			mj = m > h ? mj : j;
			m = m > h ? m : h;
			t = M - oe_del; t = t > 0 ? t : 0;
			e -= e_del; e = e > t ? e : t;
			E[j] = e;
			t = M - oe_ins; t = t > 0 ? t : 0;
			f -= e_ins; f = f > t ? f : t;
		}
		H[end] = h1; E[end] = 0;
		if (j == qlen) {
			max_ie = gscore > h1 ? max_ie : i;
			gscore = gscore > h1 ? gscore : h1;
		}
		if (m == 0) break;
		if (m > max) {
			max = m; max_i = i; max_j = mj;
			max_off = max_off > abs(mj - i) ? max_off : abs(mj - i);
		} else if (zdrop > 0) {
			if (i - max_i > mj - max_j) {
				if (max - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop) { ++orc_tls_cnt.n_ext_zdrop; break; }
			} else {
				if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) { ++orc_tls_cnt.n_ext_zdrop; break; }
			}
		}
		for (j = beg; j < end && H[j] == 0 && E[j] == 0; ++j) {}
		beg = j;
		for (j = end; j >= beg && H[j] == 0 && E[j] == 0; --j) {}
		end = j + 2 < qlen ? j + 2 : qlen;
	}
	free(H); free(E);
	if (qle_) *qle_ = max_j + 1;
	if (tle_) *tle_ = max_i + 1;
	if (gtle_) *gtle_ = max_ie + 1;
	if (gscore_) *gscore_ = gscore;
	if (max_off_) *max_off_ = max_off;
	return max;
}

static uint32_t *push_cigar(int *n, int *m, uint32_t *cigar, int op, int len)
{
	if (*n == 0 || (int)(cigar[*n - 1] & 0xf) != op) {
		if (*n == *m) { *m = *m ? *m << 1 : 4; cigar = realloc(cigar, (size_t)*m * 4); }
		cigar[(*n)++] = (uint32_t)len << 4 | (uint32_t)op;
	} else cigar[*n - 1] += (uint32_t)len << 4;
	return cigar;
}

int orc_ksw_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                    int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar_, uint32_t **cigar_)
{
	int32_t *H, *E;
	int i, j, k, oe_del = o_del + e_del, oe_ins = o_ins + e_ins, score, n_col;
	uint8_t *z = 0;
	int want_bt = n_cigar_ && cigar_;
	if (n_cigar_) *n_cigar_ = 0;
	n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
	if (want_bt) z = malloc((size_t)n_col * tlen + 1);
	H = malloc((qlen + 1) * sizeof(int32_t));
	E = malloc((qlen + 1) * sizeof(int32_t));
	H[0] = 0; E[0] = MINUS_INF;
	for (j = 1; j <= qlen && j <= w; ++j) { H[j] = -(o_ins + e_ins * j); E[j] = MINUS_INF; }
	for (; j <= qlen; ++j) H[j] = E[j] = MINUS_INF;
	++orc_tls_cnt.n_glb_calls;
	for (i = 0; i < tlen; ++i) {
		int32_t f = MINUS_INF, h1, beg, end, t;
		const int8_t *srow = &mat[target[i] * 5];
		uint8_t *zi = want_bt ? &z[(size_t)i * n_col] : 0;
		beg = i > w ? i - w : 0;
		end = i + w + 1 < qlen ? i + w + 1 : qlen;
		h1 = beg == 0 ? -(o_del + e_del * (i + 1)) : MINUS_INF;
		if (end > beg) orc_tls_cnt.cells_glb += (uint64_t)(end - beg);
		for (j = beg; j < end; ++j) {
			int32_t h, m = H[j], e = E[j];
			uint8_t d;
			H[j] = h1;
			m += srow[query[j]];
			d = m >= e ? 0 : 1;
			h = m >= e ? m : e;
			d = h >= f ? d : 2;
			h = h >= f ? h : f;
			h1 = h;
			t = m - oe_del;
			e -= e_del;
			d |= e > t ? 1 << 2 : 0;
			e = e > t ? e : t;
			E[j] = e;
			t = m - oe_ins;
			f -= e_ins;
			d |= f > t ? 2 << 4 : 0;
			f = f > t ? f : t;
			if (zi) zi[j - beg] = d;
		}
		H[end] = h1; E[end] = MINUS_INF;
	}
	score = H[qlen];
	if (want_bt) {
		int n_cigar = 0, m_cigar = 0, which = 0;
		uint32_t *cigar = 0, tmp;
		i = tlen - 1; k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
		while (i >= 0 && k >= 0) {
			which = z[(size_t)i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
			if (which == 0) { cigar = push_cigar(&n_cigar, &m_cigar, cigar, 0, 1); --i; --k; }
			else if (which == 1) { cigar = push_cigar(&n_cigar, &m_cigar, cigar, 2, 1); --i; }
			else { cigar = push_cigar(&n_cigar, &m_cigar, cigar, 1, 1); --k; }
		}
		if (i >= 0) cigar = push_cigar(&n_cigar, &m_cigar, cigar, 2, i + 1);
		if (k >= 0) cigar = push_cigar(&n_cigar, &m_cigar, cigar, 1, k + 1);
		for (i = 0; i < n_cigar >> 1; ++i) { tmp = cigar[i]; cigar[i] = cigar[n_cigar - 1 - i]; cigar[n_cigar - 1 - i] = tmp; }
		*n_cigar_ = n_cigar; *cigar_ = cigar;
	}
	free(H); free(E); free(z);
	return score;
}

/* ------------------------------------------------------------------
 * Local alignment with striped-SIMD-observable semantics.
 *   lanes = 16 (u8 kernel) or 8 (i16 kernel); slen = ceil(qlen/lanes); the
 *   query is padded to lanes*slen columns that score 0 against everything.
 *   Query column q belongs to segment q / slen.  Per target row:
 *     pass 1:  Hp(q) = max(0, H(i-1,q-1)+s, E(q), Fseg(q)) where Fseg restarts
 *              at 0 on every segment boundary and is fed by Hp;
 *              E'(q)  = max(E(q)-e_del, Hp(q)-oe_del)  (>= 0)
 *     pass 2:  H(q)   = max(Hp(q), Fcarry(q)) with Fcarry the gap-extension of
 *              every segment's final F across the following segments
 *   rowmax is taken over Hp only (a carried F never exceeds it).
 * u8: every value saturates at 255-shift..; the row loop stops once
 * gmax+shift >= 255, and the score is reported as 255.
 */
typedef struct { int score, te, qe, score2, te2; } sw_fwd_t;

static sw_fwd_t sw_striped(int lanes, int is_u8, int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                           const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int xtra)
{
	sw_fwd_t r = { 0, -1, -1, -1, -1 };
	int slen = (qlen + lanes - 1) / lanes, plen = slen * lanes;
	int i, q, k, te = -1, gmax = 0, minsc, endsc, shift = 0, mmax = 0, n_b = 0, m_b = 0;
	int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	int32_t *H0, *H1, *E, *Hmax, *tmpp, fend[16];
	uint64_t *b = 0;
	if (slen == 0) return r;
	for (k = 0, shift = 127; k < 25; ++k) { if (mat[k] < shift) shift = mat[k]; if (mat[k] > mmax) mmax = mat[k]; }
	shift = (256 - (shift & 0xff)) & 0xff; /* uint8 arithmetic of the profile builder */
	minsc = (xtra & ORC_KSW_XSUBO) ? xtra & 0xffff : 0x10000;
	endsc = (xtra & ORC_KSW_XSTOP) ? xtra & 0xffff : 0x10000;
	H0 = calloc(plen, 4); H1 = calloc(plen, 4); E = calloc(plen, 4); Hmax = calloc(plen, 4);
	++orc_tls_cnt.n_msw_calls;
	for (i = 0; i < tlen; ++i) {
		const int8_t *srow = &mat[target[i] * 5];
		int imax = 0, f = 0;
		orc_tls_cnt.cells_msw += (uint64_t)qlen;
		for (q = 0; q < plen; ++q) { /* pass 1 */
			int s = q < qlen ? srow[query[q]] : 0;
			int h = q > 0 ? H0[q - 1] : 0, e = E[q], t;
			if (q % slen == 0) f = 0;
			if (is_u8) { h = h + s + shift; if (h > 255) h = 255; h -= shift; if (h < 0) h = 0; }
			else { h = h + s; if (h > 32767) h = 32767; }
			h = h > e ? h : e;
			h = h > f ? h : f;
			imax = imax > h ? imax : h;
			H1[q] = h;
			e -= e_del; if (e < 0) e = 0;
			t = h - oe_del; if (t < 0) t = 0;
			E[q] = e > t ? e : t;
			f -= e_ins; if (f < 0) f = 0;
			t = h - oe_ins; if (t < 0) t = 0;
			f = f > t ? f : t;
			if (q % slen == slen - 1) fend[q / slen] = f; /* F leaving this segment */
		}
		{ /* pass 2: carry each segment's closing F into the following segments */
			int carry = 0;
			for (q = 0; q < plen; ++q) {
				if (q % slen == 0 && q > 0) carry = carry > fend[q / slen - 1] ? carry : fend[q / slen - 1];
				if (carry > H1[q]) H1[q] = carry;
				carry -= e_ins; if (carry < 0) carry = 0;
			}
		}
		if (imax >= minsc) {
			if (n_b == 0 || (int32_t)b[n_b - 1] + 1 != i) {
				if (n_b == m_b) { m_b = m_b ? m_b << 1 : 8; b = realloc(b, 8 * (size_t)m_b); }
				b[n_b++] = (uint64_t)imax << 32 | (uint32_t)i;
			} else if ((int)(b[n_b - 1] >> 32) < imax) b[n_b - 1] = (uint64_t)imax << 32 | (uint32_t)i;
		}
		if (imax > gmax) {
			gmax = imax; te = i;
			memcpy(Hmax, H1, (size_t)plen * 4);
			if (is_u8) { if (gmax + shift >= 255 || gmax >= endsc) break; }
			else if (gmax >= endsc) break;
		}
		tmpp = H1; H1 = H0; H0 = tmpp;
	}
	r.score = is_u8 ? (gmax + shift < 255 ? gmax : 255) : gmax;
	r.te = te;
	if (!is_u8 || r.score != 255) {
		int max = -1, low, high;
		for (q = 0; q < plen; ++q) /* smallest query column holding the row maximum */
			if (Hmax[q] > max) { max = Hmax[q]; r.qe = q; }
		if (b) {
			k = (r.score + mmax - 1) / mmax;
			low = te - k; high = te + k;
			for (k = 0; k < n_b; ++k) {
				int e = (int32_t)b[k];
				if ((e < low || e > high) && (int)(b[k] >> 32) > r.score2) { r.score2 = (int)(b[k] >> 32); r.te2 = e; }
			}
		}
	}
	free(b); free(H0); free(H1); free(E); free(Hmax);
	return r;
}

static void revseq(int l, uint8_t *s)
{
	int i; uint8_t t;
	for (i = 0; i < l >> 1; ++i) { t = s[i]; s[i] = s[l - 1 - i]; s[l - 1 - i] = t; }
}

/* which form of the striped kernel orc_ksw_align2 runs: the SSE2 one (ksw_sse2.c; default where SSE2 exists, as bwa
 * itself is built) or the scalar emulation above.  Results are identical (tests/test_oracle_kat.py). */
static int g_simd = -1;
void orc_set_simd(int on) { g_simd = on ? orc_have_sse2() : 0; }
int orc_get_simd(void) { if (g_simd < 0) g_simd = orc_have_sse2(); return g_simd; }

static sw_fwd_t sw_pass(int lanes, int is_u8, int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                        const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int xtra)
{
	if (orc_get_simd()) {
		int o[5]; sw_fwd_t r;
		orc_sw_striped_sse2(is_u8, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra, o);
		r.score = o[0]; r.te = o[1]; r.qe = o[2]; r.score2 = o[3]; r.te2 = o[4];
		return r;
	}
	return sw_striped(lanes, is_u8, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra);
}

orc_kswr_t orc_ksw_align2(int qlen, uint8_t *query, int tlen, uint8_t *target, const int8_t *mat,
                          int o_del, int e_del, int o_ins, int e_ins, int xtra)
{
	orc_kswr_t r = { 0, -1, -1, -1, -1, -1, -1 };
	int is_u8 = (xtra & ORC_KSW_XBYTE) ? 1 : 0, lanes = is_u8 ? 16 : 8;
	sw_fwd_t f, rr;
	f = sw_pass(lanes, is_u8, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra);
	r.score = f.score; r.te = f.te; r.qe = f.qe; r.score2 = f.score2; r.te2 = f.te2;
	if ((xtra & ORC_KSW_XSTART) == 0 || ((xtra & ORC_KSW_XSUBO) && r.score < (xtra & 0xffff))) return r;
	revseq(r.qe + 1, query); revseq(r.te + 1, target);
	rr = sw_pass(lanes, is_u8, r.qe + 1, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, ORC_KSW_XSTOP | r.score);
	revseq(r.qe + 1, query); revseq(r.te + 1, target);
	if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
	return r;
}
