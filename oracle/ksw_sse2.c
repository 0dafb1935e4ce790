/*
 * ksw_sse2.c -- CPU ORACLE (test infrastructure only; see orc.h): the striped SSE2 form of the local alignment
 * behind ksw_align2 (ksw_u8 / ksw_i16 of un-vendored lh3/bwa), restated from the published
 * algorithm (Farrar's striped Smith-Waterman with the lazy-F loop, as bwa uses it: query profile of 16 byte lanes or
 * 8 short lanes, E fed by the pre-correction H, saturating arithmetic, per-row maxima list for the 2nd-best score).
 *
 * Why it exists: bwa itself runs mate rescue on these SIMD kernels, so a CPU baseline that emulates them cell by cell
 * (ksw.c: sw_striped) would flatter any speed-up quoted against it.  This file is what orc_ksw_align2 runs by default
 * on x86-64; the scalar emulation stays as the readable statement of the semantics and as the cross-check:
 * tests/test_oracle_kat.py compares the two on random inputs, and orc_set_simd(0) switches the whole oracle to it.
 */
#include "orc_internal.h"

#if defined(__SSE2__)
#include <emmintrin.h>

int orc_have_sse2(void) { return 1; }

typedef struct { int score, te, qe, score2, te2; } fwd_t;

static void *aligned16(size_t n) { void *p = 0; if (posix_memalign(&p, 16, n ? n : 16) != 0) return 0; return p; }

static fwd_t finish(int is_u8, int lanes, int slen, const void *Hmax, int gmax, int shift, int mmax, int te, const uint64_t *b, int n_b)
{
	fwd_t r = { 0, -1, -1, -1, -1 };
	int i, k, low, high, max = -1, plen = slen * lanes;
	r.score = is_u8 ? (gmax + shift < 255 ? gmax : 255) : gmax;
	r.te = te;
	if (!is_u8 || r.score != 255) {
		for (i = 0; i < plen; ++i) { /* memory order is (j, lane); the query column is lane * slen + j: smallest column wins ties */
			int v = is_u8 ? ((const uint8_t *)Hmax)[i] : ((const int16_t *)Hmax)[i];
			int q = i / lanes + i % lanes * slen;
			if (v > max) { max = v; r.qe = q; }
			else if (v == max && q < r.qe) r.qe = q;
		}
		if (b) {
			k = (r.score + mmax - 1) / mmax;
			low = te - k; high = te + k;
			for (k = 0; k < n_b; ++k) {
				int e = (int32_t)b[k];
				if ((e < low || e > high) && (int)(b[k] >> 32) > r.score2) { r.score2 = (int)(b[k] >> 32); r.te2 = e; }
			}
		}
	}
	return r;
}

#define PUSH_ROW_MAX() do { \
		if (imax >= minsc) { \
			if (n_b == 0 || (int32_t)b[n_b - 1] + 1 != i) { \
				if (n_b == m_b) { m_b = m_b ? m_b << 1 : 8; b = realloc(b, 8 * (size_t)m_b); } \
				b[n_b++] = (uint64_t)imax << 32 | (uint32_t)i; \
			} else if ((int)(b[n_b - 1] >> 32) < imax) b[n_b - 1] = (uint64_t)imax << 32 | (uint32_t)i; \
		} } while (0)

static fwd_t sse2_u8(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                     int o_del, int e_del_, int o_ins, int e_ins_, int xtra)
{
	fwd_t r = { 0, -1, -1, -1, -1 };
	int slen = (qlen + 15) >> 4, i, j, k, a, te = -1, gmax = 0, minsc, endsc, shift = 127, mmax = 0, n_b = 0, m_b = 0;
	uint64_t *b = 0;
	__m128i zero, oe_del, e_del, oe_ins, e_ins, vshift, *qp, *H0, *H1, *E, *Hmax, *tmpv;
	if (slen == 0) return r;
	for (k = 0; k < 25; ++k) { if (mat[k] < shift) shift = mat[k]; if (mat[k] > mmax) mmax = mat[k]; }
	shift = (256 - (shift & 0xff)) & 0xff;
	minsc = (xtra & ORC_KSW_XSUBO) ? xtra & 0xffff : 0x10000;
	endsc = (xtra & ORC_KSW_XSTOP) ? xtra & 0xffff : 0x10000;
	qp = aligned16((size_t)slen * 16 * 5); H0 = aligned16((size_t)slen * 16); H1 = aligned16((size_t)slen * 16);
	E = aligned16((size_t)slen * 16); Hmax = aligned16((size_t)slen * 16);
	{ /* query profile: lane l of vector j holds column l * slen + j; columns beyond the query score 0 */
		uint8_t *t = (uint8_t *)qp;
		for (a = 0; a < 5; ++a)
			for (j = 0; j < slen; ++j)
				for (k = j; k < slen * 16; k += slen)
					*t++ = (uint8_t)((k >= qlen ? 0 : mat[a * 5 + query[k]]) + shift);
	}
	zero = _mm_set1_epi32(0);
	oe_del = _mm_set1_epi8((char)(o_del + e_del_)); e_del = _mm_set1_epi8((char)e_del_);
	oe_ins = _mm_set1_epi8((char)(o_ins + e_ins_)); e_ins = _mm_set1_epi8((char)e_ins_);
	vshift = _mm_set1_epi8((char)shift);
	for (j = 0; j < slen; ++j) { _mm_store_si128(E + j, zero); _mm_store_si128(H0 + j, zero); _mm_store_si128(Hmax + j, zero); }
	++orc_tls_cnt.n_msw_calls;
	for (i = 0; i < tlen; ++i) {
		int cmp, imax;
		__m128i e, h, t, f = zero, max = zero, *S = qp + (size_t)target[i] * slen;
		orc_tls_cnt.cells_msw += (uint64_t)qlen;
		h = _mm_load_si128(H0 + slen - 1);
		h = _mm_slli_si128(h, 1);
		for (j = 0; j < slen; ++j) {
			h = _mm_adds_epu8(h, _mm_load_si128(S + j));
			h = _mm_subs_epu8(h, vshift);
			e = _mm_load_si128(E + j);
			h = _mm_max_epu8(h, e);
			h = _mm_max_epu8(h, f);
			max = _mm_max_epu8(max, h);
			_mm_store_si128(H1 + j, h);
			e = _mm_subs_epu8(e, e_del);
			t = _mm_subs_epu8(h, oe_del);
			e = _mm_max_epu8(e, t);
			_mm_store_si128(E + j, e);
			f = _mm_subs_epu8(f, e_ins);
			t = _mm_subs_epu8(h, oe_ins);
			f = _mm_max_epu8(f, t);
			h = _mm_load_si128(H0 + j);
		}
		for (k = 0; k < 16; ++k) { /* lazy F: carry the F leaving a segment into the following ones while it still matters */
			f = _mm_slli_si128(f, 1);
			for (j = 0; j < slen; ++j) {
				h = _mm_load_si128(H1 + j);
				h = _mm_max_epu8(h, f);
				_mm_store_si128(H1 + j, h);
				h = _mm_subs_epu8(h, oe_ins);
				f = _mm_subs_epu8(f, e_ins);
				cmp = _mm_movemask_epi8(_mm_cmpeq_epi8(_mm_subs_epu8(f, h), zero));
				if (cmp == 0xffff) goto end_lazy_f;
			}
		}
end_lazy_f:
		max = _mm_max_epu8(max, _mm_srli_si128(max, 8)); max = _mm_max_epu8(max, _mm_srli_si128(max, 4));
		max = _mm_max_epu8(max, _mm_srli_si128(max, 2)); max = _mm_max_epu8(max, _mm_srli_si128(max, 1));
		imax = _mm_extract_epi16(max, 0) & 0xff;
		PUSH_ROW_MAX();
		if (imax > gmax) {
			gmax = imax; te = i;
			for (j = 0; j < slen; ++j) _mm_store_si128(Hmax + j, _mm_load_si128(H1 + j));
			if (gmax + shift >= 255 || gmax >= endsc) break;
		}
		tmpv = H1; H1 = H0; H0 = tmpv;
	}
	r = finish(1, 16, slen, Hmax, gmax, shift, mmax, te, b, n_b);
	free(b); free(qp); free(H0); free(H1); free(E); free(Hmax);
	return r;
}

static fwd_t sse2_i16(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                      int o_del, int e_del_, int o_ins, int e_ins_, int xtra)
{
	fwd_t r = { 0, -1, -1, -1, -1 };
	int slen = (qlen + 7) >> 3, i, j, k, a, te = -1, gmax = 0, minsc, endsc, mmax = 0, n_b = 0, m_b = 0;
	uint64_t *b = 0;
	__m128i zero, oe_del, e_del, oe_ins, e_ins, *qp, *H0, *H1, *E, *Hmax, *tmpv;
	if (slen == 0) return r;
	for (k = 0; k < 25; ++k) if (mat[k] > mmax) mmax = mat[k];
	minsc = (xtra & ORC_KSW_XSUBO) ? xtra & 0xffff : 0x10000;
	endsc = (xtra & ORC_KSW_XSTOP) ? xtra & 0xffff : 0x10000;
	qp = aligned16((size_t)slen * 16 * 5); H0 = aligned16((size_t)slen * 16); H1 = aligned16((size_t)slen * 16);
	E = aligned16((size_t)slen * 16); Hmax = aligned16((size_t)slen * 16);
	{
		int16_t *t = (int16_t *)qp;
		for (a = 0; a < 5; ++a)
			for (j = 0; j < slen; ++j)
				for (k = j; k < slen * 8; k += slen)
					*t++ = (int16_t)(k >= qlen ? 0 : mat[a * 5 + query[k]]);
	}
	zero = _mm_set1_epi32(0);
	oe_del = _mm_set1_epi16((short)(o_del + e_del_)); e_del = _mm_set1_epi16((short)e_del_);
	oe_ins = _mm_set1_epi16((short)(o_ins + e_ins_)); e_ins = _mm_set1_epi16((short)e_ins_);
	for (j = 0; j < slen; ++j) { _mm_store_si128(E + j, zero); _mm_store_si128(H0 + j, zero); _mm_store_si128(Hmax + j, zero); }
	++orc_tls_cnt.n_msw_calls;
	for (i = 0; i < tlen; ++i) {
		int imax;
		__m128i e, h, t, f = zero, max = zero, *S = qp + (size_t)target[i] * slen;
		orc_tls_cnt.cells_msw += (uint64_t)qlen;
		h = _mm_load_si128(H0 + slen - 1);
		h = _mm_slli_si128(h, 2);
		for (j = 0; j < slen; ++j) {
			h = _mm_adds_epi16(h, _mm_load_si128(S + j));
			e = _mm_load_si128(E + j);
			h = _mm_max_epi16(h, e);
			h = _mm_max_epi16(h, f);
			max = _mm_max_epi16(max, h);
			_mm_store_si128(H1 + j, h);
			e = _mm_subs_epu16(e, e_del);
			t = _mm_subs_epu16(h, oe_del);
			e = _mm_max_epi16(e, t);
			_mm_store_si128(E + j, e);
			f = _mm_subs_epu16(f, e_ins);
			t = _mm_subs_epu16(h, oe_ins);
			f = _mm_max_epi16(f, t);
			h = _mm_load_si128(H0 + j);
		}
		for (k = 0; k < 16; ++k) {
			f = _mm_slli_si128(f, 2);
			for (j = 0; j < slen; ++j) {
				h = _mm_load_si128(H1 + j);
				h = _mm_max_epi16(h, f);
				_mm_store_si128(H1 + j, h);
				h = _mm_subs_epu16(h, oe_ins);
				f = _mm_subs_epu16(f, e_ins);
				if (!_mm_movemask_epi8(_mm_cmpgt_epi16(f, h))) goto end_lazy_f;
			}
		}
end_lazy_f:
		max = _mm_max_epi16(max, _mm_srli_si128(max, 8)); max = _mm_max_epi16(max, _mm_srli_si128(max, 4));
		max = _mm_max_epi16(max, _mm_srli_si128(max, 2));
		imax = (int16_t)_mm_extract_epi16(max, 0);
		PUSH_ROW_MAX();
		if (imax > gmax) {
			gmax = imax; te = i;
			for (j = 0; j < slen; ++j) _mm_store_si128(Hmax + j, _mm_load_si128(H1 + j));
			if (gmax >= endsc) break;
		}
		tmpv = H1; H1 = H0; H0 = tmpv;
	}
	r = finish(0, 8, slen, Hmax, gmax, 0, mmax, te, b, n_b);
	free(b); free(qp); free(H0); free(H1); free(E); free(Hmax);
	return r;
}

void orc_sw_striped_sse2(int is_u8, int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                         int o_del, int e_del, int o_ins, int e_ins, int xtra, int out[5])
{
	fwd_t r = is_u8 ? sse2_u8(qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra)
	                : sse2_i16(qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra);
	out[0] = r.score; out[1] = r.te; out[2] = r.qe; out[3] = r.score2; out[4] = r.te2;
}
#else
int orc_have_sse2(void) { return 0; }
void orc_sw_striped_sse2(int is_u8, int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                         int o_del, int e_del, int o_ins, int e_ins, int xtra, int out[5])
{ (void)is_u8; (void)qlen; (void)query; (void)tlen; (void)target; (void)mat; (void)o_del; (void)e_del; (void)o_ins; (void)e_ins; (void)xtra; out[0] = 0; out[1] = out[2] = out[3] = out[4] = -1; }
#endif
