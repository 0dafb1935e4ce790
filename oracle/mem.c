/*
 * mem.c -- CPU ORACLE (test infrastructure only; see orc.h): single-end
 * BWA-MEM pipeline restated from SURVEY.md App. B.4-B.11, B.14 (un-vendored
 * lh3/bwa bwamem.c / bwamem_extra.c / bwa.c @cb95061, entered from
 * /root/reference/src/main/c/jnibwa.c:214) plus the reference's own record
 * encoder fmt_BAMish (/root/reference/src/main/c/jnibwa.c:43-97).
 */
#include <math.h>
#include "orc_internal.h"

/* ------------------------------------------------------------------ chaining (B.4) */

static int test_and_merge(const orc_opt_t *opt, int64_t l_pac, chain_t *c, const seed_t *p, int seed_rid)
{
	int64_t qend, rend, x, y;
	const seed_t *last = &c->seeds[c->n - 1];
	qend = last->qbeg + last->len;
	rend = last->rbeg + last->len;
	if (seed_rid != c->rid) return 0;
	if (p->qbeg >= c->seeds[0].qbeg && p->qbeg + p->len <= qend && p->rbeg >= c->seeds[0].rbeg && p->rbeg + p->len <= rend)
		return 1; /* contained: absorbed without being stored */
	if ((last->rbeg < l_pac || c->seeds[0].rbeg < l_pac) && p->rbeg >= l_pac) return 0;
	x = p->qbeg - last->qbeg;
	y = p->rbeg - last->rbeg;
	if (y >= 0 && x - y <= opt->w && y - x <= opt->w && x - last->len < opt->max_chain_gap && y - last->len < opt->max_chain_gap) {
		if (c->n == c->m) { c->m <<= 1; c->seeds = realloc(c->seeds, (size_t)c->m * sizeof(seed_t)); }
		c->seeds[c->n++] = *p;
		return 1;
	}
	return 0;
}

/*
 * The reference keeps chains in an ordered tree keyed by `pos` and asks for the
 * greatest key <= rbeg.  Here: an array kept sorted by pos.  Equal keys (rare;
 * parity unpinned): a lookup returns the oldest chain of that key and a new
 * chain is placed right after it, which is what a lower-bound search followed
 * by "insert after the hit" does inside one tree leaf.
 */
void orc_mem_chain(const orc_index_t *x, const orc_opt_t *opt, int len, const uint8_t *seq, chain_v *out)
{
	int b, e, l_rep;
	size_t i, j;
	int64_t l_pac = x->l_pac;
	intv_v mem;
	chain_v ch; /* sorted by pos */
	vec_init(*out);
	if (len < opt->min_seed_len) return;
	vec_init(mem); vec_init(ch);
	orc_collect_intv_core(x, opt, len, seq, &mem);
	for (i = 0, b = e = l_rep = 0; i < mem.n; ++i) {
		orc_intv_t *p = &mem.a[i];
		int sb = (int)(p->info >> 32), se = (int)(uint32_t)p->info;
		if (p->size <= (uint64_t)opt->max_occ) continue;
		if (sb > e) { l_rep += e - b; b = sb; e = se; }
		else e = e > se ? e : se;
	}
	l_rep += e - b;
	for (i = 0; i < mem.n; ++i) {
		orc_intv_t *p = &mem.a[i];
		int step, count, slen = (int)((uint32_t)p->info - (uint32_t)(p->info >> 32));
		int64_t k;
		step = p->size > (uint64_t)opt->max_occ ? (int)(p->size / opt->max_occ) : 1;
		for (k = count = 0; (uint64_t)k < p->size && count < opt->max_occ; k += step, ++count) {
			seed_t s;
			int rid, to_add = 0;
			size_t lo, hi, at = 0;
			s.rbeg = (int64_t)orc_bwt_sa(x, p->x0 + k);
			s.qbeg = (int)(p->info >> 32);
			s.score = s.len = slen;
			++orc_tls_cnt.n_seeds;
			rid = orc_intv2rid(x, s.rbeg, s.rbeg + s.len);
			if (rid < 0) continue;
			if (ch.n) {
				chain_t *lower = 0;
				/* first index with pos >= rbeg */
				lo = 0; hi = ch.n;
				while (lo < hi) { size_t mid = (lo + hi) >> 1; if (ch.a[mid].pos < s.rbeg) lo = mid + 1; else hi = mid; }
				if (lo < ch.n && ch.a[lo].pos == s.rbeg) { lower = &ch.a[lo]; at = lo + 1; }
				else if (lo > 0) { lower = &ch.a[lo - 1]; at = lo; }
				else { lower = 0; at = 0; }
				if (!lower || !test_and_merge(opt, l_pac, lower, &s, rid)) to_add = 1;
			} else { to_add = 1; at = 0; }
			if (to_add) {
				chain_t tmp;
				memset(&tmp, 0, sizeof(tmp));
				tmp.n = 1; tmp.m = 4;
				tmp.seeds = calloc(tmp.m, sizeof(seed_t));
				tmp.seeds[0] = s;
				tmp.rid = rid;
				tmp.pos = s.rbeg;
				tmp.is_alt = !!x->anns[rid].is_alt;
				vec_push(ch, tmp); /* grow */
				for (j = ch.n - 1; j > at; --j) ch.a[j] = ch.a[j - 1];
				ch.a[at] = tmp;
			}
		}
	}
	for (i = 0; i < ch.n; ++i) ch.a[i].frac_rep = (float)l_rep / len;
	*out = ch;
	vec_free(mem);
}

/* ------------------------------------------------------------------ chain weight + filter (B.5) */

static int chain_weight(const chain_t *c)
{
	int64_t end;
	int j, w = 0, tmp;
	for (j = 0, end = 0; j < c->n; ++j) {
		const seed_t *s = &c->seeds[j];
		if (s->qbeg >= end) w += s->len;
		else if (s->qbeg + s->len > end) w += (int)(s->qbeg + s->len - end);
		end = end > s->qbeg + s->len ? end : s->qbeg + s->len;
	}
	tmp = w; w = 0;
	for (j = 0, end = 0; j < c->n; ++j) {
		const seed_t *s = &c->seeds[j];
		if (s->rbeg >= end) w += s->len;
		else if (s->rbeg + s->len > end) w += (int)(s->rbeg + s->len - end);
		end = end > s->rbeg + s->len ? end : s->rbeg + s->len;
	}
	w = w < tmp ? w : tmp;
	return w < 1 << 30 ? w : (1 << 30) - 1;
}

#define CHN_BEG(ch) ((ch).seeds->qbeg)
#define CHN_END(ch) ((ch).seeds[(ch).n - 1].qbeg + (ch).seeds[(ch).n - 1].len)
#define FLT_LT(a, b) ((a).w > (b).w)
ORC_SORT_DECL(chain_flt, chain_t, FLT_LT)

int orc_mem_chain_flt(const orc_opt_t *opt, int n_chn, chain_t *a)
{
	int i, k;
	int_v chains;
	if (n_chn == 0) return 0;
	vec_init(chains);
	for (i = k = 0; i < n_chn; ++i) {
		chain_t *c = &a[i];
		c->first = -1; c->kept = 0;
		c->w = (uint32_t)chain_weight(c);
		if ((int)c->w < opt->min_chain_weight) free(c->seeds);
		else a[k++] = *c;
	}
	n_chn = k;
	if (n_chn == 0) return 0;
	orc_sort_chain_flt(n_chn, a);
	a[0].kept = 3;
	vec_push(chains, 0);
	for (i = 1; i < n_chn; ++i) {
		int large_ovlp = 0;
		for (k = 0; k < (int)chains.n; ++k) {
			int j = chains.a[k];
			int b_max = CHN_BEG(a[j]) > CHN_BEG(a[i]) ? CHN_BEG(a[j]) : CHN_BEG(a[i]);
			int e_min = CHN_END(a[j]) < CHN_END(a[i]) ? CHN_END(a[j]) : CHN_END(a[i]);
			if (e_min > b_max && (!a[j].is_alt || a[i].is_alt)) {
				int li = CHN_END(a[i]) - CHN_BEG(a[i]);
				int lj = CHN_END(a[j]) - CHN_BEG(a[j]);
				int min_l = li < lj ? li : lj;
				if (e_min - b_max >= min_l * opt->mask_level && min_l < opt->max_chain_gap) {
					large_ovlp = 1;
					if (a[j].first < 0) a[j].first = i;
					if ((int)a[i].w < (int)a[j].w * opt->drop_ratio && (int)a[j].w - (int)a[i].w >= opt->min_seed_len << 1)
						break;
				}
			}
		}
		if (k == (int)chains.n) {
			vec_push(chains, i);
			a[i].kept = large_ovlp ? 2 : 3;
		}
	}
	for (i = 0; i < (int)chains.n; ++i) {
		chain_t *c = &a[chains.a[i]];
		if (c->first >= 0) a[c->first].kept = 1;
	}
	vec_free(chains);
	for (i = k = 0; i < n_chn; ++i) {
		if (a[i].kept == 0 || a[i].kept == 3) continue;
		if (++k >= opt->max_chain_extend) break;
	}
	for (; i < n_chn; ++i)
		if (a[i].kept < 3) a[i].kept = 0;
	for (i = k = 0; i < n_chn; ++i) {
		chain_t *c = &a[i];
		if (c->kept == 0) free(c->seeds);
		else a[k++] = a[i];
	}
	return k;
}

/* long reads only: score short seeds by a local SW over +-50 bp (B.5 tail) */
static int seed_sw(const orc_index_t *x, const orc_opt_t *opt, int l_query, const uint8_t *query, const seed_t *s)
{
	int qb, qe, rid;
	int64_t rb, re, mid, l_pac = x->l_pac;
	uint8_t *rseq, *qtmp;
	orc_kswr_t r;
	if (s->len >= 200) return -1;
	qb = s->qbeg; qe = s->qbeg + s->len;
	rb = s->rbeg; re = s->rbeg + s->len;
	mid = (rb + re) >> 1;
	qb -= 50; qb = qb > 0 ? qb : 0;
	qe += 50; qe = qe < l_query ? qe : l_query;
	rb -= 50; rb = rb > 0 ? rb : 0;
	re += 50; re = re < l_pac << 1 ? re : l_pac << 1;
	if (rb < l_pac && l_pac < re) { if (mid < l_pac) re = l_pac; else rb = l_pac; }
	if (qe - qb >= 200 || re - rb >= 200) return -1;
	rseq = orc_fetch_seq(x, &rb, mid, &re, &rid);
	qtmp = malloc(qe - qb + 1); memcpy(qtmp, query + qb, qe - qb);
	r = orc_ksw_align2(qe - qb, qtmp, (int)(re - rb), rseq, opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, ORC_KSW_XSTART);
	free(rseq); free(qtmp);
	return r.score;
}

void orc_mem_flt_chained_seeds(const orc_index_t *x, const orc_opt_t *opt, int l_query, const uint8_t *query, int n_chn, chain_t *a)
{
	double min_l = opt->min_chain_weight ? 1.1f * opt->min_chain_weight : 5.5f * log(l_query);
	int i, j, k, min_HSP_score = (int)(opt->a * min_l + .499);
	if (min_l > 0.05f * l_query) return;
	for (i = 0; i < n_chn; ++i) {
		chain_t *c = &a[i];
		for (j = k = 0; j < c->n; ++j) {
			seed_t *s = &c->seeds[j];
			s->score = seed_sw(x, opt, l_query, query, s);
			if (s->score < 0 || s->score >= min_HSP_score) {
				s->score = s->score < 0 ? s->len * opt->a : s->score;
				c->seeds[k++] = *s;
			}
		}
		c->n = k;
	}
}

/* ------------------------------------------------------------------ extension (B.6) */

static inline int cal_max_gap(const orc_opt_t *opt, int qlen)
{
	int l_del = (int)((double)(qlen * opt->a - opt->o_del) / opt->e_del + 1.);
	int l_ins = (int)((double)(qlen * opt->a - opt->o_ins) / opt->e_ins + 1.);
	int l = l_del > l_ins ? l_del : l_ins;
	l = l > 1 ? l : 1;
	return l < opt->w << 1 ? l : opt->w << 1;
}

#define U64_LT(a, b) ((a) < (b))
ORC_SORT_DECL(u64, uint64_t, U64_LT)

#define MAX_BAND_TRY 2

void orc_mem_chain2aln(const orc_index_t *x, const orc_opt_t *opt, int l_query, const uint8_t *query, const chain_t *c, reg_v *av)
{
	int i, k, rid, max_off[2], aw[2];
	int64_t l_pac = x->l_pac, rmax[2], tmp, max = 0;
	const seed_t *s;
	uint8_t *rseq = 0;
	uint64_t *srt;
	if (c->n == 0) return;
	rmax[0] = l_pac << 1; rmax[1] = 0;
	for (i = 0; i < c->n; ++i) {
		int64_t b, e;
		const seed_t *t = &c->seeds[i];
		b = t->rbeg - (t->qbeg + cal_max_gap(opt, t->qbeg));
		e = t->rbeg + t->len + ((l_query - t->qbeg - t->len) + cal_max_gap(opt, l_query - t->qbeg - t->len));
		rmax[0] = rmax[0] < b ? rmax[0] : b;
		rmax[1] = rmax[1] > e ? rmax[1] : e;
		if (t->len > max) max = t->len;
	}
	rmax[0] = rmax[0] > 0 ? rmax[0] : 0;
	rmax[1] = rmax[1] < l_pac << 1 ? rmax[1] : l_pac << 1;
	if (rmax[0] < l_pac && l_pac < rmax[1]) {
		if (c->seeds[0].rbeg < l_pac) rmax[1] = l_pac;
		else rmax[0] = l_pac;
	}
	rseq = orc_fetch_seq(x, &rmax[0], c->seeds[0].rbeg, &rmax[1], &rid);

	srt = malloc((size_t)c->n * 8);
	for (i = 0; i < c->n; ++i) srt[i] = (uint64_t)c->seeds[i].score << 32 | (uint32_t)i;
	orc_sort_u64(c->n, srt);

	for (k = c->n - 1; k >= 0; --k) {
		reg_t *a;
		s = &c->seeds[(uint32_t)srt[k]];
		for (i = 0; i < (int)av->n; ++i) { /* is the seed inside an earlier region, near its diagonal? */
			reg_t *p = &av->a[i];
			int64_t rd;
			int qd, w, max_gap;
			if (s->rbeg < p->rb || s->rbeg + s->len > p->re || s->qbeg < p->qb || s->qbeg + s->len > p->qe) continue;
			if (s->len - p->seedlen0 > .1 * l_query) continue;
			qd = s->qbeg - p->qb; rd = s->rbeg - p->rb;
			max_gap = cal_max_gap(opt, qd < rd ? qd : (int)rd);
			w = max_gap < p->w ? max_gap : p->w;
			if (qd - rd < w && rd - qd < w) break;
			qd = p->qe - (s->qbeg + s->len); rd = p->re - (s->rbeg + s->len);
			max_gap = cal_max_gap(opt, qd < rd ? qd : (int)rd);
			w = max_gap < p->w ? max_gap : p->w;
			if (qd - rd < w && rd - qd < w) break;
		}
		if (i < (int)av->n) { /* contained: extend anyway only if an overlapping seed sits on another diagonal */
			for (i = k + 1; i < c->n; ++i) {
				const seed_t *t;
				if (srt[i] == 0) continue;
				t = &c->seeds[(uint32_t)srt[i]];
				if (t->len < s->len * .95) continue;
				if (s->qbeg <= t->qbeg && s->qbeg + s->len - t->qbeg >= s->len >> 2 && t->qbeg - s->qbeg != t->rbeg - s->rbeg) break;
				if (t->qbeg <= s->qbeg && t->qbeg + t->len - s->qbeg >= s->len >> 2 && s->qbeg - t->qbeg != s->rbeg - t->rbeg) break;
			}
			if (i == c->n) { srt[k] = 0; continue; }
		}
		{
			reg_t z; memset(&z, 0, sizeof(z));
			vec_push(*av, z);
			a = &av->a[av->n - 1];
		}
		a->w = aw[0] = aw[1] = opt->w;
		a->score = a->truesc = -1;
		a->rid = c->rid;

		if (s->qbeg) { /* left extension */
			uint8_t *rs, *qs;
			int qle, tle, gtle, gscore;
			qs = malloc(s->qbeg);
			for (i = 0; i < s->qbeg; ++i) qs[i] = query[s->qbeg - 1 - i];
			tmp = s->rbeg - rmax[0];
			rs = malloc(tmp + 1);
			for (i = 0; i < tmp; ++i) rs[i] = rseq[tmp - 1 - i];
			for (i = 0; i < MAX_BAND_TRY; ++i) {
				int prev = a->score;
				aw[0] = opt->w << i;
				if (i) ++orc_tls_cnt.n_ext_retry;
				a->score = orc_ksw_extend2(s->qbeg, qs, (int)tmp, rs, opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins,
				                           aw[0], opt->pen_clip5, opt->zdrop, s->len * opt->a, &qle, &tle, &gtle, &gscore, &max_off[0]);
				if (a->score == prev || max_off[0] < (aw[0] >> 1) + (aw[0] >> 2)) break;
			}
			if (gscore <= 0 || gscore <= a->score - opt->pen_clip5) {
				a->qb = s->qbeg - qle; a->rb = s->rbeg - tle;
				a->truesc = a->score;
			} else {
				a->qb = 0; a->rb = s->rbeg - gtle;
				a->truesc = gscore;
			}
			free(qs); free(rs);
		} else { a->score = a->truesc = s->len * opt->a; a->qb = 0; a->rb = s->rbeg; }

		if (s->qbeg + s->len != l_query) { /* right extension */
			int qle, tle, qe, re, gtle, gscore, sc0 = a->score;
			qe = s->qbeg + s->len;
			re = (int)(s->rbeg + s->len - rmax[0]);
			for (i = 0; i < MAX_BAND_TRY; ++i) {
				int prev = a->score;
				aw[1] = opt->w << i;
				if (i) ++orc_tls_cnt.n_ext_retry;
				a->score = orc_ksw_extend2(l_query - qe, query + qe, (int)(rmax[1] - rmax[0] - re), rseq + re, opt->mat,
				                           opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, aw[1], opt->pen_clip3, opt->zdrop, sc0,
				                           &qle, &tle, &gtle, &gscore, &max_off[1]);
				if (a->score == prev || max_off[1] < (aw[1] >> 1) + (aw[1] >> 2)) break;
			}
			if (gscore <= 0 || gscore <= a->score - opt->pen_clip3) {
				a->qe = qe + qle; a->re = rmax[0] + re + tle;
				a->truesc += a->score - sc0;
			} else {
				a->qe = l_query; a->re = rmax[0] + re + gtle;
				a->truesc += gscore - sc0;
			}
		} else { a->qe = l_query; a->re = s->rbeg + s->len; }

		for (i = 0, a->seedcov = 0; i < c->n; ++i) {
			const seed_t *t = &c->seeds[i];
			if (t->qbeg >= a->qb && t->qbeg + t->len <= a->qe && t->rbeg >= a->rb && t->rbeg + t->len <= a->re)
				a->seedcov += t->len;
		}
		a->w = aw[0] > aw[1] ? aw[0] : aw[1];
		a->seedlen0 = s->len;
		a->frac_rep = c->frac_rep;
	}
	free(srt); free(rseq);
}

/* ------------------------------------------------------------------ CIGAR / MD / NM (B.9) */

typedef struct { size_t l, m; char *s; } kstr_t;
static void ks_putc(kstr_t *s, int c) { if (s->l + 2 > s->m) { s->m = s->m ? s->m << 1 : 32; s->s = realloc(s->s, s->m); } s->s[s->l++] = (char)c; s->s[s->l] = 0; }
static void ks_putsn(kstr_t *s, const char *p, size_t n) { size_t i; for (i = 0; i < n; ++i) ks_putc(s, p[i]); }
static void ks_puts(kstr_t *s, const char *p) { ks_putsn(s, p, strlen(p)); }
static void ks_putl(kstr_t *s, long v) { char buf[32]; snprintf(buf, sizeof buf, "%ld", v); ks_puts(s, buf); }

uint32_t *orc_gen_cigar2(const int8_t mat[25], int o_del, int e_del, int o_ins, int e_ins, int w_, int64_t l_pac, const uint8_t *pac,
                         int l_query, uint8_t *query, int64_t rb, int64_t re, int *score, int *n_cigar, int *NM)
{
	uint32_t *cigar = 0;
	uint8_t tmp, *rseq;
	int i;
	int64_t rlen;
	if (n_cigar) *n_cigar = 0;
	if (NM) *NM = -1;
	if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return 0;
	rseq = orc_get_seq(l_pac, pac, rb, re, &rlen);
	if (re - rb != rlen) goto done;
	if (rb >= l_pac) { /* reverse (not complement) both, so indels left-align on the forward strand */
		for (i = 0; i < l_query >> 1; ++i) { tmp = query[i]; query[i] = query[l_query - 1 - i]; query[l_query - 1 - i] = tmp; }
		for (i = 0; i < rlen >> 1; ++i) { tmp = rseq[i]; rseq[i] = rseq[rlen - 1 - i]; rseq[rlen - 1 - i] = tmp; }
	}
	if (l_query == re - rb && w_ == 0) {
		if (n_cigar) { cigar = malloc(4); cigar[0] = (uint32_t)l_query << 4 | 0; *n_cigar = 1; }
		for (i = 0, *score = 0; i < l_query; ++i) *score += mat[rseq[i] * 5 + query[i]];
	} else {
		int w, max_gap, max_ins, max_del, min_w;
		max_ins = (int)((double)(((l_query + 1) >> 1) * mat[0] - o_ins) / e_ins + 1.);
		max_del = (int)((double)(((l_query + 1) >> 1) * mat[0] - o_del) / e_del + 1.);
		max_gap = max_ins > max_del ? max_ins : max_del;
		max_gap = max_gap > 1 ? max_gap : 1;
		w = (max_gap + abs((int)rlen - l_query) + 1) >> 1;
		w = w < w_ ? w : w_;
		min_w = abs((int)rlen - l_query) + 3;
		w = w > min_w ? w : min_w;
		*score = orc_ksw_global2(l_query, query, (int)rlen, rseq, mat, o_del, e_del, o_ins, e_ins, w, n_cigar, n_cigar ? &cigar : 0);
	}
	if (NM && n_cigar) { /* MD string is appended behind the CIGAR words */
		int k, xq, y, u, n_mm = 0, n_gap = 0;
		kstr_t str = { 0, 0, 0 };
		const char *int2base = rb < l_pac ? "ACGTN" : "TGCAN";
		for (k = 0, xq = y = u = 0; k < *n_cigar; ++k) {
			int op = cigar[k] & 0xf, len = (int)(cigar[k] >> 4);
			if (op == 0) {
				for (i = 0; i < len; ++i) {
					if (query[xq + i] != rseq[y + i]) {
						ks_putl(&str, u); ks_putc(&str, int2base[rseq[y + i]]);
						++n_mm; u = 0;
					} else ++u;
				}
				xq += len; y += len;
			} else if (op == 2) {
				if (k > 0 && k < *n_cigar - 1) {
					ks_putl(&str, u); ks_putc(&str, '^');
					for (i = 0; i < len; ++i) ks_putc(&str, int2base[rseq[y + i]]);
					u = 0; n_gap += len;
				}
				y += len;
			} else if (op == 1) { xq += len; n_gap += len; }
		}
		ks_putl(&str, u);
		*NM = n_mm + n_gap;
		cigar = realloc(cigar, (size_t)*n_cigar * 4 + str.l + 1);
		memcpy((char *)(cigar + *n_cigar), str.s, str.l + 1);
		free(str.s);
	}
	if (rb >= l_pac)
		for (i = 0; i < l_query >> 1; ++i) { tmp = query[i]; query[i] = query[l_query - 1 - i]; query[l_query - 1 - i] = tmp; }
done:
	free(rseq);
	return cigar;
}

/* ------------------------------------------------------------------ dedup / patch (B.10) */

static int patch_reg(const orc_index_t *x, const orc_opt_t *opt, uint8_t *query, const reg_t *a, const reg_t *b, int *w_)
{
	int w, score, q_s, r_s;
	double r;
	if (query == 0) return 0;
	if (a->rb < x->l_pac && b->rb >= x->l_pac) return 0;
	if (a->qb >= b->qb || a->qe >= b->qe || a->re >= b->re) return 0;
	w = (int)((a->re - b->rb) - (a->qe - b->qb));
	w = w > 0 ? w : -w;
	r = (double)(a->re - b->rb) / (b->re - a->rb) - (double)(a->qe - b->qb) / (b->qe - a->qb);
	r = r > 0. ? r : -r;
	if (a->re < b->rb || a->qe < b->qb) {
		if (w > opt->w << 1 || r >= 0.05f) return 0;
	} else if (w > opt->w << 2 || r >= 0.05f * 2) return 0;
	w += a->w + b->w;
	w = w < opt->w << 2 ? w : opt->w << 2;
	score = 0;
	orc_gen_cigar2(opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, w, x->l_pac, x->pac, b->qe - a->qb, query + a->qb, a->rb, b->re, &score, 0, 0);
	q_s = (int)((double)(b->qe - a->qb) / ((b->qe - b->qb) + (a->qe - a->qb)) * (b->score + a->score) + .499);
	r_s = (int)((double)(b->re - a->rb) / ((b->re - b->rb) + (a->re - a->rb)) * (b->score + a->score) + .499);
	if ((double)score / (q_s > r_s ? q_s : r_s) < 0.90f) return 0;
	*w_ = w;
	return score;
}

#define REG_LT_END(a, b) ((a).re < (b).re)
#define REG_LT_SCORE(a, b) ((a).score > (b).score || ((a).score == (b).score && ((a).rb < (b).rb || ((a).rb == (b).rb && (a).qb < (b).qb))))
#define REG_LT_HASH(a, b) ((a).score > (b).score || ((a).score == (b).score && ((a).is_alt < (b).is_alt || ((a).is_alt == (b).is_alt && (a).hash < (b).hash))))
ORC_SORT_DECL(reg_end, reg_t, REG_LT_END)
ORC_SORT_DECL(reg_score, reg_t, REG_LT_SCORE)
ORC_SORT_DECL(reg_hash, reg_t, REG_LT_HASH)
#define REG_LT_HASH2(a, b) ((a).is_alt < (b).is_alt || ((a).is_alt == (b).is_alt && ((a).score > (b).score || ((a).score == (b).score && (a).hash < (b).hash))))
ORC_SORT_DECL(reg_hash2, reg_t, REG_LT_HASH2)

int orc_mem_sort_dedup_patch(const orc_index_t *x, const orc_opt_t *opt, uint8_t *query, int n, reg_t *a)
{
	int m, i, j;
	if (n <= 1) return n;
	orc_sort_reg_end(n, a);
	for (i = 0; i < n; ++i) a[i].n_comp = 1;
	for (i = 1; i < n; ++i) {
		reg_t *p = &a[i];
		if (p->rid != a[i - 1].rid || p->rb >= a[i - 1].re + opt->max_chain_gap) continue;
		for (j = i - 1; j >= 0 && p->rid == a[j].rid && p->rb < a[j].re + opt->max_chain_gap; --j) {
			reg_t *q = &a[j];
			int64_t or_, oq, mr, mq;
			int score, w;
			if (q->qe == q->qb) continue;
			or_ = q->re - p->rb;
			oq = q->qb < p->qb ? q->qe - p->qb : p->qe - q->qb;
			mr = q->re - q->rb < p->re - p->rb ? q->re - q->rb : p->re - p->rb;
			mq = q->qe - q->qb < p->qe - p->qb ? q->qe - q->qb : p->qe - p->qb;
			if (or_ > opt->mask_level_redun * mr && oq > opt->mask_level_redun * mq) {
				if (p->score < q->score) { p->qe = p->qb; break; }
				else q->qe = q->qb;
			} else if (q->rb < p->rb && (score = patch_reg(x, opt, query, q, p, &w)) > 0) {
				p->n_comp += q->n_comp + 1;
				p->seedcov = p->seedcov > q->seedcov ? p->seedcov : q->seedcov;
				p->sub = p->sub > q->sub ? p->sub : q->sub;
				p->csub = p->csub > q->csub ? p->csub : q->csub;
				p->qb = q->qb; p->rb = q->rb;
				p->truesc = p->score = score;
				p->w = w;
				q->qb = q->qe;
			}
		}
	}
	for (i = 0, m = 0; i < n; ++i)
		if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
	n = m;
	orc_sort_reg_score(n, a);
	for (i = 1; i < n; ++i)
		if (a[i].score == a[i - 1].score && a[i].rb == a[i - 1].rb && a[i].qb == a[i - 1].qb)
			a[i].qe = a[i].qb;
	for (i = 1, m = 1; i < n; ++i)
		if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
	return m;
}

void orc_mem_align1_core(const orc_index_t *x, const orc_opt_t *opt, int l_seq, uint8_t *seq, reg_v *regs)
{
	int i;
	chain_v chn;
	for (i = 0; i < l_seq; ++i) seq[i] = seq[i] < 4 ? seq[i] : orc_nt4_table[seq[i]];
	orc_mem_chain(x, opt, l_seq, seq, &chn);
	chn.n = (size_t)orc_mem_chain_flt(opt, (int)chn.n, chn.a);
	orc_mem_flt_chained_seeds(x, opt, l_seq, seq, (int)chn.n, chn.a);
	orc_tls_cnt.n_chains += chn.n;
	vec_init(*regs);
	for (i = 0; i < (int)chn.n; ++i) {
		orc_mem_chain2aln(x, opt, l_seq, seq, &chn.a[i], regs);
		free(chn.a[i].seeds);
	}
	free(chn.a);
	regs->n = (size_t)orc_mem_sort_dedup_patch(x, opt, seq, (int)regs->n, regs->a);
	for (i = 0; i < (int)regs->n; ++i) {
		reg_t *p = &regs->a[i];
		if (p->rid >= 0 && x->anns[p->rid].is_alt) p->is_alt = 1;
	}
	orc_tls_cnt.n_regs += regs->n;
}

/* ------------------------------------------------------------------ primary marking / MAPQ (B.11) */

uint64_t orc_hash64(uint64_t key)
{
	key += ~(key << 32);
	key ^= (key >> 22);
	key += ~(key << 13);
	key ^= (key >> 8);
	key += (key << 3);
	key ^= (key >> 15);
	key += ~(key << 27);
	key ^= (key >> 31);
	return key;
}

static void mark_primary_core(const orc_opt_t *opt, int n, reg_t *a, int_v *z)
{
	int i, k, tmp;
	tmp = opt->a + opt->b;
	tmp = opt->o_del + opt->e_del > tmp ? opt->o_del + opt->e_del : tmp;
	tmp = opt->o_ins + opt->e_ins > tmp ? opt->o_ins + opt->e_ins : tmp;
	z->n = 0;
	vec_push(*z, 0);
	for (i = 1; i < n; ++i) {
		for (k = 0; k < (int)z->n; ++k) {
			int j = z->a[k];
			int b_max = a[j].qb > a[i].qb ? a[j].qb : a[i].qb;
			int e_min = a[j].qe < a[i].qe ? a[j].qe : a[i].qe;
			if (e_min > b_max) {
				int min_l = a[i].qe - a[i].qb < a[j].qe - a[j].qb ? a[i].qe - a[i].qb : a[j].qe - a[j].qb;
				if (e_min - b_max >= min_l * opt->mask_level) {
					if (a[j].sub == 0) a[j].sub = a[i].score;
					if (a[j].score - a[i].score <= tmp && (a[j].is_alt || !a[i].is_alt)) ++a[j].sub_n;
					break;
				}
			}
		}
		if (k == (int)z->n) vec_push(*z, i);
		else a[i].secondary = z->a[k];
	}
}

/* mem_mark_primary_se (B.11).  With hits on ALT contigs (n_pri < n) the marking runs twice: once over all hits (kept in
 * secondary_all; a primary-assembly hit shadowed by an ALT hit remembers that hit's score in alt_sc), then -- hits
 * re-ordered with the primary-assembly ones first -- over the primary-assembly hits alone (kept in secondary). */
int orc_mem_mark_primary_se(const orc_opt_t *opt, int n, reg_t *a, int64_t id)
{
	int i, n_pri;
	int_v z;
	if (n == 0) return 0;
	vec_init(z);
	for (i = n_pri = 0; i < n; ++i) {
		a[i].sub = a[i].alt_sc = 0; a[i].secondary = a[i].secondary_all = -1; a[i].hash = orc_hash64((uint64_t)(id + i));
		if (!a[i].is_alt) ++n_pri;
	}
	orc_sort_reg_hash(n, a);
	mark_primary_core(opt, n, a, &z);
	for (i = 0; i < n; ++i) {
		reg_t *p = &a[i];
		p->secondary_all = i; /* keep the rank of the first round */
		if (!p->is_alt && p->secondary >= 0 && a[p->secondary].is_alt) p->alt_sc = a[p->secondary].score;
	}
	if (n_pri >= 0 && n_pri < n) {
		int *rank = malloc(sizeof(int) * n);
		if (n_pri > 0) orc_sort_reg_hash2(n, a);
		for (i = 0; i < n; ++i) rank[a[i].secondary_all] = i;
		for (i = 0; i < n; ++i) {
			if (a[i].secondary >= 0) {
				a[i].secondary_all = rank[a[i].secondary];
				if (a[i].is_alt) a[i].secondary = INT_MAX;
			} else a[i].secondary_all = -1;
		}
		free(rank);
		if (n_pri > 0) { /* mark primary for hits to the primary assembly only */
			for (i = 0; i < n_pri; ++i) { a[i].sub = 0; a[i].secondary = -1; }
			mark_primary_core(opt, n_pri, a, &z);
		}
	} else {
		for (i = 0; i < n; ++i) a[i].secondary_all = a[i].secondary;
	}
	vec_free(z);
	return n_pri;
}

void orc_mem_reorder_primary5(int T, reg_v *a)
{
	int k, n_pri = 0, left_st = INT_MAX, left_k = -1;
	reg_t t;
	for (k = 0; k < (int)a->n; ++k)
		if (a->a[k].secondary < 0 && !a->a[k].is_alt && a->a[k].score >= T) ++n_pri;
	if (n_pri <= 1) return;
	for (k = 0; k < (int)a->n; ++k) {
		reg_t *p = &a->a[k];
		if (p->secondary >= 0 || p->is_alt || p->score < T) continue;
		if (p->qb < left_st) { left_st = p->qb; left_k = k; }
	}
	if (left_k == 0) return;
	t = a->a[0]; a->a[0] = a->a[left_k]; a->a[left_k] = t;
	for (k = 1; k < (int)a->n; ++k) {
		reg_t *p = &a->a[k];
		if (p->secondary == 0) p->secondary = left_k;
		else if (p->secondary == left_k) p->secondary = 0;
		if (p->secondary_all == 0) p->secondary_all = left_k;
		else if (p->secondary_all == left_k) p->secondary_all = 0;
	}
}

int orc_mem_approx_mapq_se(const orc_opt_t *opt, const reg_t *a)
{
	int mapq, l, sub = a->sub ? a->sub : opt->min_seed_len * opt->a;
	double identity;
	sub = a->csub > sub ? a->csub : sub;
	if (sub >= a->score) return 0;
	l = a->qe - a->qb > a->re - a->rb ? a->qe - a->qb : (int)(a->re - a->rb);
	identity = 1. - (double)(l * opt->a - a->score) / (opt->a + opt->b) / l;
	if (a->score == 0) mapq = 0;
	else if (opt->mapQ_coef_len > 0) {
		double tmp;
		tmp = l < opt->mapQ_coef_len ? 1. : opt->mapQ_coef_fac / log(l);
		tmp *= identity * identity;
		mapq = (int)(6.02 * (a->score - sub) / opt->a * tmp * tmp + .499);
	} else {
		mapq = (int)(30.0 * (1. - (double)sub / a->score) * log(a->seedcov) + .499);
		mapq = identity < 0.95 ? (int)(mapq * identity * identity + .499) : mapq;
	}
	if (a->sub_n > 0) mapq -= (int)(4.343 * log(a->sub_n + 1) + .499);
	if (mapq > 60) mapq = 60;
	if (mapq < 0) mapq = 0;
	mapq = (int)(mapq * (1. - a->frac_rep) + .499);
	return mapq;
}

/* ------------------------------------------------------------------ reg -> alignment record (B.9) */

static inline int infer_bw(int l1, int l2, int score, int a, int q, int r)
{
	int w;
	if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
	w = (int)(((double)((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.));
	if (w < abs(l1 - l2)) w = abs(l1 - l2);
	return w;
}

aln_t orc_mem_reg2aln(const orc_index_t *x, const orc_opt_t *opt, int l_query, const uint8_t *query_, const reg_t *ar)
{
	aln_t a;
	int i, w2, tmp, qb, qe, NM, score, is_rev, last_sc = -(1 << 30), l_MD;
	int64_t pos, rb, re;
	uint8_t *query;
	memset(&a, 0, sizeof(a));
	if (ar == 0 || ar->rb < 0 || ar->re < 0) { a.rid = -1; a.pos = -1; a.flag |= 0x4; return a; }
	qb = ar->qb; qe = ar->qe; rb = ar->rb; re = ar->re;
	query = malloc(l_query + 1);
	for (i = 0; i < l_query; ++i) query[i] = query_[i] < 5 ? query_[i] : orc_nt4_table[query_[i]];
	a.mapq = ar->secondary < 0 ? (uint32_t)orc_mem_approx_mapq_se(opt, ar) : 0;
	if (ar->secondary >= 0) a.flag |= 0x100;
	tmp = infer_bw(qe - qb, (int)(re - rb), ar->truesc, opt->a, opt->o_del, opt->e_del);
	w2 = infer_bw(qe - qb, (int)(re - rb), ar->truesc, opt->a, opt->o_ins, opt->e_ins);
	w2 = w2 > tmp ? w2 : tmp;
	if (w2 > opt->w) w2 = w2 < ar->w ? w2 : ar->w;
	i = 0; a.cigar = 0;
	do {
		free(a.cigar);
		w2 = w2 < opt->w << 2 ? w2 : opt->w << 2;
		a.cigar = orc_gen_cigar2(opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, w2, x->l_pac, x->pac, qe - qb, &query[qb], rb, re, &score, &a.n_cigar, &NM);
		if (score == last_sc || w2 == opt->w << 2) break;
		last_sc = score;
		w2 <<= 1;
	} while (++i < 3 && score < ar->truesc - opt->a);
	l_MD = (int)strlen((char *)(a.cigar + a.n_cigar)) + 1;
	a.NM = (uint32_t)NM;
	pos = orc_depos(x, rb < x->l_pac ? rb : re - 1, &is_rev);
	a.is_rev = (uint32_t)is_rev;
	if (a.n_cigar > 0) { /* squeeze out a leading or (else) trailing deletion */
		if ((a.cigar[0] & 0xf) == 2) {
			pos += a.cigar[0] >> 4;
			--a.n_cigar;
			memmove(a.cigar, a.cigar + 1, (size_t)a.n_cigar * 4 + l_MD);
		} else if ((a.cigar[a.n_cigar - 1] & 0xf) == 2) {
			--a.n_cigar;
			memmove(a.cigar + a.n_cigar, a.cigar + a.n_cigar + 1, l_MD);
		}
	}
	if (qb != 0 || qe != l_query) {
		int clip5, clip3;
		clip5 = is_rev ? l_query - qe : qb;
		clip3 = is_rev ? qb : l_query - qe;
		a.cigar = realloc(a.cigar, 4 * (size_t)(a.n_cigar + 2) + l_MD);
		if (clip5) {
			memmove(a.cigar + 1, a.cigar, (size_t)a.n_cigar * 4 + l_MD);
			a.cigar[0] = (uint32_t)clip5 << 4 | 3;
			++a.n_cigar;
		}
		if (clip3) {
			memmove(a.cigar + a.n_cigar + 1, a.cigar + a.n_cigar, l_MD);
			a.cigar[a.n_cigar++] = (uint32_t)clip3 << 4 | 3;
		}
	}
	a.rid = orc_pos2rid(x, pos);
	a.pos = pos - x->anns[a.rid].offset;
	a.score = ar->score; a.sub = ar->sub > ar->csub ? ar->sub : ar->csub;
	a.is_alt = (uint32_t)ar->is_alt; a.alt_sc = ar->alt_sc;
	free(query);
	return a;
}

/* ------------------------------------------------------------------ XA (B.14) */

static inline int get_pri_idx(double XA_drop_ratio, const reg_t *a, int i)
{
	int k = a[i].secondary_all;
	if (k >= 0 && a[i].score >= a[k].score * XA_drop_ratio) return k;
	return -1;
}

char **orc_mem_gen_alt(const orc_index_t *x, const orc_opt_t *opt, const reg_v *a, int l_query, const uint8_t *query)
{
	int i, k, r, *cnt, tot;
	kstr_t *aln = 0, str = { 0, 0, 0 };
	char **XA = 0, *has_alt;
	cnt = calloc(a->n + 1, sizeof(int));
	has_alt = calloc(a->n + 1, 1);
	for (i = 0, tot = 0; i < (int)a->n; ++i) {
		r = get_pri_idx(opt->XA_drop_ratio, a->a, i);
		if (r >= 0) { ++cnt[r]; ++tot; if (a->a[i].is_alt) has_alt[r] = 1; }
	}
	if (tot == 0) goto end_gen_alt;
	aln = calloc(a->n, sizeof(kstr_t));
	for (i = 0; i < (int)a->n; ++i) {
		aln_t t;
		if ((r = get_pri_idx(opt->XA_drop_ratio, a->a, i)) < 0) continue;
		if (cnt[r] > opt->max_XA_hits_alt || (!has_alt[r] && cnt[r] > opt->max_XA_hits)) continue;
		t = orc_mem_reg2aln(x, opt, l_query, query, &a->a[i]);
		str.l = 0;
		ks_puts(&str, x->anns[t.rid].name);
		ks_putc(&str, ','); ks_putc(&str, "+-"[t.is_rev]); ks_putl(&str, (long)(t.pos + 1));
		ks_putc(&str, ',');
		for (k = 0; k < t.n_cigar; ++k) { ks_putl(&str, (long)(t.cigar[k] >> 4)); ks_putc(&str, "MIDSHN"[t.cigar[k] & 0xf]); }
		ks_putc(&str, ','); ks_putl(&str, (long)t.NM);
		free(t.cigar);
		ks_putsn(&aln[r], str.s, str.l); ks_putc(&aln[r], ';');
	}
	XA = calloc(a->n, sizeof(char *));
	for (k = 0; k < (int)a->n; ++k) XA[k] = aln[k].s;
end_gen_alt:
	free(has_alt); free(cnt); free(aln); free(str.s);
	return XA;
}

/* ------------------------------------------------------------------ record encoder (jnibwa.c:26-97) */

static void blob_put(blob_t *b, const void *p, size_t n)
{
	if (b->l + n + 1 > b->m) { b->m = b->m ? b->m : 64; while (b->l + n + 1 > b->m) b->m <<= 1; b->s = realloc(b->s, b->m); }
	memcpy(b->s + b->l, p, n); b->l += n;
}
static void blob_put32(blob_t *b, int32_t v) { blob_put(b, &v, 4); }

static int cigar_ref_len(int n_cigar, const uint32_t *cigar)
{
	int k, len = 0;
	for (k = 0; k < n_cigar; ++k) { int op = cigar[k] & 0xf; if (!op || op == 2) len += (int)(cigar[k] >> 4); }
	return len;
}

/* flag finalisation that bwa performs on copies before calling the formatter hook (B.14), then fmt_BAMish */
void orc_aln2blob(blob_t *str, int n, const aln_t *list, int which, const aln_t *m_)
{
	aln_t ptmp = list[which], *p = &ptmp, mtmp, *m = 0;
	uint32_t flag_mapQ;
	if (m_) { mtmp = *m_; m = &mtmp; }
	p->flag |= m ? 0x1 : 0;
	p->flag |= p->rid < 0 ? 0x4 : 0;
	p->flag |= m && m->rid < 0 ? 0x8 : 0;
	if (p->rid < 0 && m && m->rid >= 0) { p->rid = m->rid; p->pos = m->pos; p->is_rev = m->is_rev; p->n_cigar = 0; }
	if (m && m->rid < 0 && p->rid >= 0) { m->rid = p->rid; m->pos = p->pos; m->is_rev = p->is_rev; m->n_cigar = 0; }
	p->flag |= p->is_rev ? 0x10 : 0;
	p->flag |= m && m->is_rev ? 0x20 : 0;
	/* ---- fmt_BAMish ---- */
	if (!which) blob_put32(str, n);
	flag_mapQ = (uint32_t)p->flag;
	if (p->flag & 0x10000) flag_mapQ |= 0x100;
	flag_mapQ = (flag_mapQ << 16) | (p->mapq & 0xff);
	blob_put32(str, (int32_t)flag_mapQ);
	if (!(p->flag & 0x4)) {
		int k;
		int32_t nMD, nXA;
		const char *md = (const char *)(p->cigar + p->n_cigar);
		blob_put32(str, p->rid);
		blob_put32(str, (int32_t)p->pos);
		blob_put32(str, (int32_t)p->NM);
		blob_put32(str, p->score);
		blob_put32(str, p->sub);
		blob_put32(str, p->n_cigar);
		for (k = 0; k < p->n_cigar; ++k) {
			uint32_t lenOp = p->cigar[k];
			if ((lenOp & 0xf) > 2) ++lenOp; /* MIDSH -> BAM's MIDNSH numbering */
			blob_put32(str, (int32_t)lenOp);
		}
		nMD = p->n_cigar ? (int32_t)strlen(md) : 0;
		blob_put32(str, nMD);
		if (nMD) { char pad[4] = { 0, 0, 0, 0 }; blob_put(str, md, nMD); blob_put(str, pad, ((nMD + 3) & ~3) - nMD); }
		nXA = p->XA ? (int32_t)strlen(p->XA) : 0;
		blob_put32(str, nXA);
		if (nXA) { char pad[4] = { 0, 0, 0, 0 }; blob_put(str, p->XA, nXA); blob_put(str, pad, ((nXA + 3) & ~3) - nXA); }
	}
	if ((p->flag & 0x9) == 1) {
		blob_put32(str, m->rid);
		blob_put32(str, (int32_t)m->pos);
		if ((p->flag & 0x4) || p->rid != m->rid) blob_put32(str, 0);
		else {
			long p0 = p->pos, m0 = m->pos;
			if (p->is_rev) p0 += cigar_ref_len(p->n_cigar, p->cigar) - 1;
			if (m->is_rev) m0 += cigar_ref_len(m->n_cigar, m->cigar) - 1;
			blob_put32(str, (int32_t)(m0 - p0 + (p0 > m0 ? -1 : p0 < m0 ? 1 : 0)));
		}
	}
}

/* ------------------------------------------------------------------ SE emission (B.14) */

void orc_mem_reg2sam(const orc_index_t *x, const orc_opt_t *opt, int l_seq, const uint8_t *seq, reg_v *a, int extra_flag, const aln_t *m, blob_t *out)
{
	VEC(aln_t) aa;
	int k, l;
	char **XA = 0;
	if (!(opt->flag & ORC_F_ALL)) XA = orc_mem_gen_alt(x, opt, a, l_seq, seq);
	vec_init(aa);
	for (k = l = 0; k < (int)a->n; ++k) {
		reg_t *p = &a->a[k];
		aln_t q;
		if (p->score < opt->T) continue;
		if (p->secondary >= 0 && (p->is_alt || !(opt->flag & ORC_F_ALL))) continue;
		if (p->secondary >= 0 && p->secondary < INT_MAX && p->score < a->a[p->secondary].score * opt->drop_ratio) continue;
		q = orc_mem_reg2aln(x, opt, l_seq, seq, p);
		q.XA = XA ? XA[k] : 0;
		q.flag |= extra_flag;
		if (p->secondary >= 0) q.sub = -1;
		if (l && p->secondary < 0) q.flag |= (opt->flag & ORC_F_NO_MULTI) ? 0x10000 : 0x800;
		if (l && !p->is_alt && q.mapq > aa.a[0].mapq) q.mapq = aa.a[0].mapq;
		vec_push(aa, q);
		++l;
	}
	if (aa.n == 0) {
		aln_t t = orc_mem_reg2aln(x, opt, l_seq, seq, 0);
		t.flag |= extra_flag;
		orc_aln2blob(out, 1, &t, 0, m);
	} else {
		for (k = 0; k < (int)aa.n; ++k) orc_aln2blob(out, (int)aa.n, aa.a, k, m);
		for (k = 0; k < (int)aa.n; ++k) free(aa.a[k].cigar);
	}
	vec_free(aa);
	if (XA) { for (k = 0; k < (int)a->n; ++k) free(XA[k]); free(XA); }
}
