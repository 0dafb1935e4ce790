/*
 * pair.c -- CPU ORACLE (test infrastructure only; see orc.h): paired-end part of
 * BWA-MEM restated from SURVEY.md App. B.8 (mem_matesw), B.12 (mem_pestat),
 * B.13 (mem_pair, mem_sam_pe) -- un-vendored lh3/bwa bwamem_pair.c @cb95061,
 * entered from /root/reference/src/main/c/jnibwa.c:214.
 */
#include <math.h>
#include "orc_internal.h"

#define MIN_RATIO     0.8
#define MIN_DIR_CNT   10
#define MIN_DIR_RATIO 0.05
#define OUTLIER_BOUND 2.0
#define MAPPING_BOUND 3.0
#define MAX_STDDEV    4.0

static inline int infer_dir(int64_t l_pac, int64_t b1, int64_t b2, int64_t *dist)
{
	int64_t p2;
	int r1 = (b1 >= l_pac), r2 = (b2 >= l_pac);
	p2 = r1 == r2 ? b2 : (l_pac << 1) - 1 - b2;
	*dist = p2 > b1 ? p2 - b1 : b1 - p2;
	return (r1 == r2 ? 0 : 1) ^ (p2 > b1 ? 0 : 3);
}

static int cal_sub(const orc_opt_t *opt, const reg_v *r)
{
	int j;
	for (j = 1; j < (int)r->n; ++j) {
		int b_max = r->a[j].qb > r->a[0].qb ? r->a[j].qb : r->a[0].qb;
		int e_min = r->a[j].qe < r->a[0].qe ? r->a[j].qe : r->a[0].qe;
		if (e_min > b_max) {
			int min_l = r->a[j].qe - r->a[j].qb < r->a[0].qe - r->a[0].qb ? r->a[j].qe - r->a[j].qb : r->a[0].qe - r->a[0].qb;
			if (e_min - b_max >= min_l * opt->mask_level) break;
		}
	}
	return j < (int)r->n ? r->a[j].score : opt->min_seed_len * opt->a;
}

#define U64_LT(a, b) ((a) < (b))
ORC_SORT_DECL(u64p, uint64_t, U64_LT)

void orc_mem_pestat(const orc_opt_t *opt, int64_t l_pac, int n, const reg_v *regs, orc_pestat_t pes[4])
{
	int i, d, max;
	u64_v isize[4];
	memset(pes, 0, 4 * sizeof(orc_pestat_t));
	for (d = 0; d < 4; ++d) vec_init(isize[d]);
	for (i = 0; i < n >> 1; ++i) {
		int dir;
		int64_t is;
		const reg_v *r[2];
		r[0] = &regs[i << 1 | 0];
		r[1] = &regs[i << 1 | 1];
		if (r[0]->n == 0 || r[1]->n == 0) continue;
		if (cal_sub(opt, r[0]) > MIN_RATIO * r[0]->a[0].score) continue;
		if (cal_sub(opt, r[1]) > MIN_RATIO * r[1]->a[0].score) continue;
		if (r[0]->a[0].rid != r[1]->a[0].rid) continue;
		dir = infer_dir(l_pac, r[0]->a[0].rb, r[1]->a[0].rb, &is);
		if (is && is <= opt->max_ins) vec_push(isize[dir], (uint64_t)is);
	}
	for (d = 0; d < 4; ++d) {
		orc_pestat_t *r = &pes[d];
		u64_v *q = &isize[d];
		int p25, p50, p75, x;
		if (q->n < MIN_DIR_CNT) { r->failed = 1; continue; }
		orc_sort_u64p(q->n, q->a);
		p25 = (int)q->a[(int)(.25 * q->n + .499)];
		p50 = (int)q->a[(int)(.50 * q->n + .499)];
		p75 = (int)q->a[(int)(.75 * q->n + .499)];
		(void)p50;
		r->low = (int)(p25 - OUTLIER_BOUND * (p75 - p25) + .499);
		if (r->low < 1) r->low = 1;
		r->high = (int)(p75 + OUTLIER_BOUND * (p75 - p25) + .499);
		for (i = x = 0, r->avg = 0; i < (int)q->n; ++i)
			if (q->a[i] >= (uint64_t)r->low && q->a[i] <= (uint64_t)r->high) { r->avg += q->a[i]; ++x; }
		r->avg /= x;
		for (i = 0, r->std = 0; i < (int)q->n; ++i)
			if (q->a[i] >= (uint64_t)r->low && q->a[i] <= (uint64_t)r->high)
				r->std += (q->a[i] - r->avg) * (q->a[i] - r->avg);
		r->std = sqrt(r->std / x);
		r->low = (int)(p25 - MAPPING_BOUND * (p75 - p25) + .499);
		r->high = (int)(p75 + MAPPING_BOUND * (p75 - p25) + .499);
		if (r->low > r->avg - MAX_STDDEV * r->std) r->low = (int)(r->avg - MAX_STDDEV * r->std + .499);
		if (r->high < r->avg + MAX_STDDEV * r->std) r->high = (int)(r->avg + MAX_STDDEV * r->std + .499);
		if (r->low < 1) r->low = 1;
	}
	for (d = 0, max = 0; d < 4; ++d) max = max > (int)isize[d].n ? max : (int)isize[d].n;
	for (d = 0; d < 4; ++d)
		if (pes[d].failed == 0 && isize[d].n < max * MIN_DIR_RATIO) pes[d].failed = 1;
	for (d = 0; d < 4; ++d) vec_free(isize[d]);
}

/* mate rescue (B.8) */
static int matesw(const orc_index_t *x, const orc_opt_t *opt, const orc_pestat_t pes[4], const reg_t *a, int l_ms, const uint8_t *ms, reg_v *ma)
{
	int64_t l_pac = x->l_pac;
	int i, r, skip[4], n = 0, rid = -1;
	for (r = 0; r < 4; ++r) skip[r] = pes[r].failed ? 1 : 0;
	for (i = 0; i < (int)ma->n; ++i) {
		int64_t dist;
		r = infer_dir(l_pac, a->rb, ma->a[i].rb, &dist);
		if (dist >= pes[r].low && dist <= pes[r].high) skip[r] = 1;
	}
	if (skip[0] + skip[1] + skip[2] + skip[3] == 4) return 0;
	for (r = 0; r < 4; ++r) {
		int is_rev, is_larger;
		uint8_t *seq, *ref = 0;
		int64_t rb, re;
		if (skip[r]) continue;
		is_rev = (r >> 1 != (r & 1));
		is_larger = !(r >> 1);
		seq = malloc(l_ms + 1);
		if (is_rev) { for (i = 0; i < l_ms; ++i) seq[l_ms - 1 - i] = ms[i] < 4 ? 3 - ms[i] : 4; }
		else memcpy(seq, ms, l_ms);
		if (!is_rev) {
			rb = is_larger ? a->rb + pes[r].low : a->rb - pes[r].high;
			re = (is_larger ? a->rb + pes[r].high : a->rb - pes[r].low) + l_ms;
		} else {
			rb = (is_larger ? a->rb + pes[r].low : a->rb - pes[r].high) - l_ms;
			re = is_larger ? a->rb + pes[r].high : a->rb - pes[r].low;
		}
		if (rb < 0) rb = 0;
		if (re > l_pac << 1) re = l_pac << 1;
		rid = -1;
		if (rb < re) ref = orc_fetch_seq(x, &rb, (rb + re) >> 1, &re, &rid);
		if (a->rid == rid && re - rb >= opt->min_seed_len) {
			orc_kswr_t aln;
			reg_t b;
			int tmp, xtra = ORC_KSW_XSUBO | ORC_KSW_XSTART | (l_ms * opt->a < 250 ? ORC_KSW_XBYTE : 0) | (opt->min_seed_len * opt->a);
			aln = orc_ksw_align2(l_ms, seq, (int)(re - rb), ref, opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, xtra);
			memset(&b, 0, sizeof(b));
			if (aln.score >= opt->min_seed_len && aln.qb >= 0) {
				b.rid = a->rid;
				b.is_alt = a->is_alt;
				b.qb = is_rev ? l_ms - (aln.qe + 1) : aln.qb;
				b.qe = is_rev ? l_ms - aln.qb : aln.qe + 1;
				b.rb = is_rev ? (l_pac << 1) - (rb + aln.te + 1) : rb + aln.tb;
				b.re = is_rev ? (l_pac << 1) - (rb + aln.tb) : rb + aln.te + 1;
				b.score = aln.score;
				b.csub = aln.score2;
				b.secondary = -1;
				b.seedcov = (int)((b.re - b.rb < b.qe - b.qb ? b.re - b.rb : b.qe - b.qb) >> 1);
				vec_push(*ma, b);
				for (i = 0; i < (int)ma->n - 1; ++i)
					if (ma->a[i].score < b.score) break;
				tmp = i;
				for (i = (int)ma->n - 1; i > tmp; --i) ma->a[i] = ma->a[i - 1];
				ma->a[i] = b;
			}
			++n;
		}
		if (n) ma->n = (size_t)orc_mem_sort_dedup_patch(x, opt, 0, (int)ma->n, ma->a);
		free(seq); free(ref);
	}
	return n;
}

typedef struct { uint64_t x, y; } pair64_t;
#define PAIR64_LT(a, b) ((a).x < (b).x || ((a).x == (b).x && (a).y < (b).y))
ORC_SORT_DECL(p128, pair64_t, PAIR64_LT)

static int mem_pair(const orc_index_t *x, const orc_opt_t *opt, const orc_pestat_t pes[4], reg_v a[2], int id, int *sub, int *n_sub, int z[2], int n_pri[2])
{
	VEC(pair64_t) v, u;
	int r, i, k, y[4], ret;
	int64_t l_pac = x->l_pac;
	vec_init(v); vec_init(u);
	for (r = 0; r < 2; ++r) {
		for (i = 0; i < n_pri[r]; ++i) {
			pair64_t key;
			reg_t *e = &a[r].a[i];
			key.x = (uint64_t)(e->rb < l_pac ? e->rb : (l_pac << 1) - 1 - e->rb);
			key.x = (uint64_t)e->rid << 32 | (key.x - (uint64_t)x->anns[e->rid].offset);
			key.y = (uint64_t)e->score << 32 | (uint64_t)(i << 2) | (uint64_t)((e->rb >= l_pac) << 1) | (uint64_t)r;
			vec_push(v, key);
		}
	}
	orc_sort_p128(v.n, v.a);
	y[0] = y[1] = y[2] = y[3] = -1;
	for (i = 0; i < (int)v.n; ++i) {
		for (r = 0; r < 2; ++r) {
			int dir = r << 1 | (int)(v.a[i].y >> 1 & 1), which;
			if (pes[dir].failed) continue;
			which = r << 1 | (int)((v.a[i].y & 1) ^ 1);
			if (y[which] < 0) continue;
			for (k = y[which]; k >= 0; --k) {
				int64_t dist;
				int q;
				double ns;
				pair64_t p;
				if ((int)(v.a[k].y & 3) != which) continue;
				dist = (int64_t)v.a[i].x - (int64_t)v.a[k].x;
				if (dist > pes[dir].high) break;
				if (dist < pes[dir].low) continue;
				ns = (dist - pes[dir].avg) / pes[dir].std;
				q = (int)((v.a[i].y >> 32) + (v.a[k].y >> 32) + .721 * log(2. * erfc(fabs(ns) * M_SQRT1_2)) * opt->a + .499);
				if (q < 0) q = 0;
				p.y = (uint64_t)k << 32 | (uint32_t)i;
				p.x = (uint64_t)q << 32 | (orc_hash64(p.y ^ (uint64_t)(int64_t)(id << 8)) & 0xffffffffU);
				vec_push(u, p);
			}
		}
		y[v.a[i].y & 3] = i;
	}
	if (u.n) {
		int tmp = opt->a + opt->b;
		tmp = tmp > opt->o_del + opt->e_del ? tmp : opt->o_del + opt->e_del;
		tmp = tmp > opt->o_ins + opt->e_ins ? tmp : opt->o_ins + opt->e_ins;
		orc_sort_p128(u.n, u.a);
		i = (int)(u.a[u.n - 1].y >> 32); k = (int)(u.a[u.n - 1].y << 32 >> 32);
		z[v.a[i].y & 1] = (int)(v.a[i].y << 32 >> 34);
		z[v.a[k].y & 1] = (int)(v.a[k].y << 32 >> 34);
		ret = (int)(u.a[u.n - 1].x >> 32);
		*sub = u.n > 1 ? (int)(u.a[u.n - 2].x >> 32) : 0;
		for (i = (int)u.n - 2, *n_sub = 0; i >= 0; --i)
			if (*sub - (int)(u.a[i].x >> 32) <= tmp) ++*n_sub;
	} else { ret = 0; *sub = 0; *n_sub = 0; }
	vec_free(u); vec_free(v);
	return ret;
}

#define RAW_MAPQ(diff, a) ((int)(6.02 * (diff) / (a) + .499))

int orc_mem_sam_pe(const orc_index_t *x, const orc_opt_t *opt, const orc_pestat_t pes[4], uint64_t id,
                   const int l_seq[2], uint8_t *const seq[2], reg_v a[2], blob_t out[2])
{
	int n = 0, i, j, z[2], o, subo, n_sub, extra_flag = 1, n_pri[2];
	aln_t h[2];
	memset(h, 0, sizeof(h));
	z[0] = z[1] = 0;
	if (!(opt->flag & ORC_F_NO_RESCUE)) {
		reg_v b[2];
		vec_init(b[0]); vec_init(b[1]);
		for (i = 0; i < 2; ++i)
			for (j = 0; j < (int)a[i].n; ++j)
				if (a[i].a[j].score >= a[i].a[0].score - opt->pen_unpaired)
					vec_push(b[i], a[i].a[j]);
		for (i = 0; i < 2; ++i)
			for (j = 0; j < (int)b[i].n && j < opt->max_matesw; ++j)
				n += matesw(x, opt, pes, &b[i].a[j], l_seq[!i], seq[!i], &a[!i]);
		vec_free(b[0]); vec_free(b[1]);
	}
	n_pri[0] = orc_mem_mark_primary_se(opt, (int)a[0].n, a[0].a, (int64_t)(id << 1 | 0));
	n_pri[1] = orc_mem_mark_primary_se(opt, (int)a[1].n, a[1].a, (int64_t)(id << 1 | 1));
	if (opt->flag & ORC_F_PRIMARY5) {
		orc_mem_reorder_primary5(opt->T, &a[0]);
		orc_mem_reorder_primary5(opt->T, &a[1]);
	}
	if (opt->flag & ORC_F_NOPAIRING) goto no_pairing;
	if (n_pri[0] && n_pri[1] && (o = mem_pair(x, opt, pes, a, (int)id, &subo, &n_sub, z, n_pri)) > 0) {
		int is_multi[2], q_pe, score_un, q_se[2];
		char **XA[2];
		for (i = 0; i < 2; ++i) {
			for (j = 1; j < n_pri[i]; ++j)
				if (a[i].a[j].secondary < 0 && a[i].a[j].score >= opt->T) break;
			is_multi[i] = j < n_pri[i] ? 1 : 0;
		}
		if (is_multi[0] || is_multi[1]) goto no_pairing;
		score_un = a[0].a[0].score + a[1].a[0].score - opt->pen_unpaired;
		subo = subo > score_un ? subo : score_un;
		q_pe = RAW_MAPQ(o - subo, opt->a);
		if (n_sub > 0) q_pe -= (int)(4.343 * log(n_sub + 1) + .499);
		if (q_pe < 0) q_pe = 0;
		if (q_pe > 60) q_pe = 60;
		q_pe = (int)(q_pe * (1. - .5 * (a[0].a[0].frac_rep + a[1].a[0].frac_rep)) + .499);
		if (o > score_un) {
			reg_t *c[2];
			c[0] = &a[0].a[z[0]]; c[1] = &a[1].a[z[1]];
			for (i = 0; i < 2; ++i) {
				if (c[i]->secondary >= 0) { c[i]->sub = a[i].a[c[i]->secondary].score; c[i]->secondary = -2; }
				q_se[i] = orc_mem_approx_mapq_se(opt, c[i]);
			}
			q_se[0] = q_se[0] > q_pe ? q_se[0] : q_pe < q_se[0] + 40 ? q_pe : q_se[0] + 40;
			q_se[1] = q_se[1] > q_pe ? q_se[1] : q_pe < q_se[1] + 40 ? q_pe : q_se[1] + 40;
			extra_flag |= 2;
			q_se[0] = q_se[0] < RAW_MAPQ(c[0]->score - c[0]->csub, opt->a) ? q_se[0] : RAW_MAPQ(c[0]->score - c[0]->csub, opt->a);
			q_se[1] = q_se[1] < RAW_MAPQ(c[1]->score - c[1]->csub, opt->a) ? q_se[1] : RAW_MAPQ(c[1]->score - c[1]->csub, opt->a);
		} else {
			z[0] = z[1] = 0;
			q_se[0] = orc_mem_approx_mapq_se(opt, &a[0].a[0]);
			q_se[1] = orc_mem_approx_mapq_se(opt, &a[1].a[0]);
		}
		for (i = 0; i < 2; ++i) {
			int k = a[i].a[z[i]].secondary_all;
			if (k >= 0 && k < n_pri[i]) {
				for (j = 0; j < (int)a[i].n; ++j)
					if (a[i].a[j].secondary_all == k || j == k)
						a[i].a[j].secondary_all = z[i];
				a[i].a[z[i]].secondary_all = -1;
			}
		}
		if (!(opt->flag & ORC_F_ALL)) {
			for (i = 0; i < 2; ++i) XA[i] = orc_mem_gen_alt(x, opt, &a[i], l_seq[i], seq[i]);
		} else XA[0] = XA[1] = 0;
		{
			aln_t aa[2][2]; int n_aa[2] = { 0, 0 };
			for (i = 0; i < 2; ++i) {
				h[i] = orc_mem_reg2aln(x, opt, l_seq[i], seq[i], &a[i].a[z[i]]);
				h[i].mapq = (uint32_t)q_se[i];
				h[i].flag |= 0x40 << i | extra_flag;
				h[i].XA = XA[i] ? XA[i][z[i]] : 0;
				aa[i][n_aa[i]++] = h[i];
				if (n_pri[i] < (int)a[i].n) { /* the read has ALT hits: its best ALT hit goes out as a supplementary record */
					reg_t *p = &a[i].a[n_pri[i]];
					aln_t g;
					if (p->score < opt->T || p->secondary >= 0 || !p->is_alt) continue;
					g = orc_mem_reg2aln(x, opt, l_seq[i], seq[i], p);
					g.flag |= 0x800 | 0x40 << i | extra_flag;
					g.XA = XA[i] ? XA[i][n_pri[i]] : 0;
					aa[i][n_aa[i]++] = g;
				}
			}
			for (i = 0; i < n_aa[0]; ++i) orc_aln2blob(&out[0], n_aa[0], aa[0], i, &h[1]);
			for (i = 0; i < n_aa[1]; ++i) orc_aln2blob(&out[1], n_aa[1], aa[1], i, &h[0]);
			for (i = 0; i < 2; ++i) if (n_aa[i] > 1) free(aa[i][1].cigar);
		}
		for (i = 0; i < 2; ++i) {
			free(h[i].cigar);
			if (XA[i]) { for (j = 0; j < (int)a[i].n; ++j) free(XA[i][j]); free(XA[i]); }
		}
	} else goto no_pairing;
	return n;

no_pairing:
	for (i = 0; i < 2; ++i) {
		int which = -1;
		if (a[i].n) {
			if (a[i].a[0].score >= opt->T) which = 0;
			else if (n_pri[i] < (int)a[i].n && a[i].a[n_pri[i]].score >= opt->T) which = n_pri[i];
		}
		if (which >= 0) h[i] = orc_mem_reg2aln(x, opt, l_seq[i], seq[i], &a[i].a[which]);
		else h[i] = orc_mem_reg2aln(x, opt, l_seq[i], seq[i], 0);
	}
	if (!(opt->flag & ORC_F_NOPAIRING) && h[0].rid == h[1].rid && h[0].rid >= 0) {
		int64_t dist;
		int d;
		d = infer_dir(x->l_pac, a[0].a[0].rb, a[1].a[0].rb, &dist);
		if (!pes[d].failed && dist >= pes[d].low && dist <= pes[d].high) extra_flag |= 2;
	}
	orc_mem_reg2sam(x, opt, l_seq[0], seq[0], &a[0], 0x41 | extra_flag, &h[1], &out[0]);
	orc_mem_reg2sam(x, opt, l_seq[1], seq[1], &a[1], 0x81 | extra_flag, &h[0], &out[1]);
	free(h[0].cigar); free(h[1].cigar);
	return n;
}
