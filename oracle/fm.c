/*
 * fm.c -- CPU ORACLE (test infrastructure only; see orc.h): index image loader,
 * FM-index primitives, SMEM seeding, suffix-array lookup, packed-reference access.
 *
 * Restates (un-vendored lh3/bwa @cb95061, reached from jnibwa.c:214):
 *   bwa_mem2idx (image walk; reference side jnibwa.c:154-165), bwt_occ4/2occ4,
 *   bwt_extend, bwt_smem1a, bwt_seed_strategy1, mem_collect_intv, bwt_sa /
 *   bwt_invPsi, bns_pos2rid / bns_depos / bns_intv2rid / bns_get_seq /
 *   bns_fetch_seq.  Behavioural spec: SURVEY.md App. A.3-A.5, B.2, B.3.
 */
#include <unistd.h>
#include <sys/stat.h>
#include <sys/mman.h>
#include "orc_internal.h"

__thread orc_counters_t orc_tls_cnt;

const uint8_t orc_nt4_table[256] = {
	4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
	4,4,4,4,4,4,4,4,4,4,4,4,4,5,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
	4,0,4,1,4,4,4,2,4,4,4,4,4,4,4,4, 4,4,4,4,3,4,4,4,4,4,4,4,4,4,4,4,
	4,0,4,1,4,4,4,2,4,4,4,4,4,4,4,4, 4,4,4,4,3,4,4,4,4,4,4,4,4,4,4,4,
	4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
	4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
	4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
	4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4
};

/* ------------------------------------------------------------------ image */

/* on-disk struct sizes of the x86-64 dump */
#define IMG_BWT_HDR   1120
#define IMG_BNS_HDR   48
#define IMG_AMB_SZ    16
#define IMG_ANN_SZ    40

static int parse_image(orc_index_t *x)
{
	const uint8_t *m = (const uint8_t *)x->mem;
	size_t k = 0, l = x->l_mem;
	int i;
	if (l < IMG_BWT_HDR) return -1;
	memcpy(&x->primary, m + 0, 8);
	memcpy(x->L2, m + 8, 40);
	memcpy(&x->seq_len, m + 48, 8);
	memcpy(&x->bwt_size, m + 56, 8);
	memcpy(&x->sa_intv, m + 72 + 1024, 4);
	memcpy(&x->n_sa, m + 72 + 1024 + 8, 8);
	k = IMG_BWT_HDR;
	if (k + x->bwt_size * 4 > l) return -1;
	x->bwt = (const uint32_t *)(m + k); k += x->bwt_size * 4;
	if (k + x->n_sa * 8 > l) return -1;
	x->sa = (const uint64_t *)(m + k); k += x->n_sa * 8;
	if (k + IMG_BNS_HDR > l) return -1;
	memcpy(&x->l_pac, m + k, 8);
	memcpy(&x->n_seqs, m + k + 8, 4);
	memcpy(&x->n_holes, m + k + 24, 4);
	k += IMG_BNS_HDR;
	k += (size_t)x->n_holes * IMG_AMB_SZ;
	if (k + (size_t)x->n_seqs * IMG_ANN_SZ > l) return -1;
	x->anns = calloc(x->n_seqs, sizeof(orc_contig_t));
	for (i = 0; i < x->n_seqs; ++i) {
		const uint8_t *p = m + k + (size_t)i * IMG_ANN_SZ;
		memcpy(&x->anns[i].offset, p, 8);
		memcpy(&x->anns[i].len, p + 8, 4);
		memcpy(&x->anns[i].n_ambs, p + 12, 4);
		memcpy(&x->anns[i].gi, p + 16, 4);
		memcpy(&x->anns[i].is_alt, p + 20, 4);
	}
	k += (size_t)x->n_seqs * IMG_ANN_SZ;
	for (i = 0; i < x->n_seqs; ++i) { /* strings are parsed by NUL, never by fixed offsets */
		x->anns[i].name = (const char *)(m + k); k += strlen((const char *)(m + k)) + 1;
		x->anns[i].anno = (const char *)(m + k); k += strlen((const char *)(m + k)) + 1;
		if (k > l) return -1;
	}
	x->pac = m + k; k += (size_t)(x->l_pac / 4 + 1);
	if (k != l) return -1; /* the reference asserts the walk ends at the file size */
	return 0;
}

orc_index_t *orc_openIndex(int fd)
{
	struct stat st;
	orc_index_t *x;
	void *mem;
	if (fstat(fd, &st) == -1) return 0;
	mem = mmap(0, st.st_size, PROT_READ, MAP_SHARED, fd, 0);
	close(fd);
	if (mem == MAP_FAILED) return 0;
	x = calloc(1, sizeof(*x));
	x->mem = mem; x->l_mem = st.st_size; x->mem_is_mmap = 1;
	if (parse_image(x) != 0) { munmap(mem, st.st_size); free(x->anns); free(x); return 0; }
	return x;
}

orc_index_t *orc_openIndexFromMemory(const void *mem, size_t len)
{
	orc_index_t *x = calloc(1, sizeof(*x));
	x->mem = malloc(len); memcpy(x->mem, mem, len); x->l_mem = len;
	if (parse_image(x) != 0) { free(x->mem); free(x->anns); free(x); return 0; }
	return x;
}

int orc_destroyIndex(orc_index_t *x)
{
	int r = 0;
	if (!x) return 0;
	if (x->mem_is_mmap) r = munmap(x->mem, x->l_mem); else free(x->mem);
	free(x->anns); free(x);
	return r;
}

/* ------------------------------------------------------------------ occ */

static inline uint64_t cnt2(uint64_t y, int c)
{ /* number of 2-bit fields of y equal to c */
	y = ((c & 2) ? y : ~y) >> 1 & ((c & 1) ? y : ~y) & 0x5555555555555555ull;
	return (uint64_t)__builtin_popcountll(y);
}

/* occurrences of A,C,G,T in BWT rows [0,k] (k is a row number of the SA incl. the '$' row; App. A.4) */
void orc_bwt_occ4(const orc_index_t *x, uint64_t k, uint64_t cnt[4])
{
	const uint32_t *p;
	uint64_t n_sym, full, rem, i, y;
	if (k == (uint64_t)-1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
	k -= (k >= x->primary);
	p = x->bwt + ((k >> 7) << 4);
	memcpy(cnt, p, 32);
	++orc_tls_cnt.n_occ;
	p += 8;
	n_sym = (k & 127) + 1; /* symbols of this block to count; symbol t of a word sits at bit (15-t)*2 */
	full = n_sym >> 5; rem = n_sym & 31;
	for (i = 0; i < full; ++i) {
		y = (uint64_t)p[2 * i] << 32 | p[2 * i + 1];
		cnt[0] += cnt2(y, 0); cnt[1] += cnt2(y, 1); cnt[2] += cnt2(y, 2); cnt[3] += cnt2(y, 3);
	}
	if (rem) {
		y = ((uint64_t)p[2 * full] << 32 | p[2 * full + 1]) & ~((1ull << ((32 - rem) << 1)) - 1);
		cnt[0] += cnt2(y, 0) - (32 - rem); /* masked fields read as A */
		cnt[1] += cnt2(y, 1); cnt[2] += cnt2(y, 2); cnt[3] += cnt2(y, 3);
	}
}

static inline uint64_t bwt_occ1(const orc_index_t *x, uint64_t k, int c)
{
	uint64_t cnt[4];
	if (k == x->seq_len) return x->L2[c + 1] - x->L2[c];
	orc_bwt_occ4(x, k, cnt);
	return cnt[c];
}

static inline int bwt_char(const orc_index_t *x, uint64_t k) /* k: index into the packed BWT ('$' removed) */
{
	const uint32_t *p = x->bwt + ((k >> 7) << 4) + 8;
	return p[(k & 127) >> 4] >> ((15 - (k & 15)) << 1) & 3;
}

void orc_bwt_extend(const orc_index_t *x, const orc_intv_t *ik, orc_intv_t ok[4], int is_back)
{
	uint64_t tk[4], tl[4], a0, s = ik->size, oth;
	int c;
	uint64_t n_before = orc_tls_cnt.n_occ;
	a0 = is_back ? ik->x0 : ik->x1;   /* the side that is walked through the BWT */
	oth = is_back ? ik->x1 : ik->x0;
	orc_bwt_occ4(x, a0 - 1, tk);
	orc_bwt_occ4(x, a0 - 1 + s, tl);
	{ /* bwt_2occ4 touches one block when both bounds fall in the same 128-symbol block */
		uint64_t k = a0 - 1, l = a0 - 1 + s;
		if (k != (uint64_t)-1 && l != (uint64_t)-1) {
			uint64_t k_ = k - (k >= x->primary), l_ = l - (l >= x->primary);
			if ((k_ >> 7) == (l_ >> 7)) orc_tls_cnt.n_occ = n_before + 1;
		}
	}
	++orc_tls_cnt.n_extend;
	for (c = 0; c < 4; ++c) {
		uint64_t v = x->L2[c] + 1 + tk[c];
		if (is_back) ok[c].x0 = v; else ok[c].x1 = v;
		ok[c].size = tl[c] - tk[c];
	}
	{
		uint64_t o3 = oth + (a0 <= x->primary && a0 + s - 1 >= x->primary);
		uint64_t o2 = o3 + ok[3].size, o1 = o2 + ok[2].size, o0 = o1 + ok[1].size;
		if (is_back) { ok[3].x1 = o3; ok[2].x1 = o2; ok[1].x1 = o1; ok[0].x1 = o0; }
		else { ok[3].x0 = o3; ok[2].x0 = o2; ok[1].x0 = o1; ok[0].x0 = o0; }
	}
}

static inline void set_intv(const orc_index_t *x, int c, orc_intv_t *ik)
{
	ik->x0 = x->L2[c] + 1; ik->size = x->L2[c + 1] - x->L2[c]; ik->x1 = x->L2[3 - c] + 1; ik->info = 0;
}

static void reverse_intvs(intv_v *p)
{
	size_t j;
	if (p->n > 1)
		for (j = 0; j < p->n >> 1; ++j) { orc_intv_t t = p->a[p->n - 1 - j]; p->a[p->n - 1 - j] = p->a[j]; p->a[j] = t; }
}

/* all SMEMs through query position `start` with >= min_intv occurrences; returns the next start (App. B.3) */
int orc_bwt_smem1(const orc_index_t *x, int len, const uint8_t *q, int start, uint64_t min_intv, intv_v *mem, intv_v tmp[2])
{
	int i, c, ret;
	size_t j;
	orc_intv_t ik, ok[4];
	intv_v *prev = &tmp[0], *curr = &tmp[1], *swap;
	mem->n = 0;
	if (q[start] > 3) return start + 1;
	if (min_intv < 1) min_intv = 1;
	set_intv(x, q[start], &ik);
	ik.info = (uint64_t)(start + 1);
	curr->n = 0;
	for (i = start + 1; i < len; ++i) { /* forward phase */
		if (q[i] < 4) {
			c = 3 - q[i];
			orc_bwt_extend(x, &ik, ok, 0);
			if (ok[c].size != ik.size) {
				vec_push(*curr, ik);
				if (ok[c].size < min_intv) break;
			}
			ik = ok[c]; ik.info = (uint64_t)(i + 1);
		} else {
			vec_push(*curr, ik);
			break;
		}
	}
	if (i == len) vec_push(*curr, ik);
	reverse_intvs(curr);
	ret = (int)curr->a[0].info;
	swap = curr; curr = prev; prev = swap;
	for (i = start - 1; i >= -1; --i) { /* backward phase */
		c = i < 0 ? -1 : q[i] < 4 ? q[i] : -1;
		for (j = 0, curr->n = 0; j < prev->n; ++j) {
			orc_intv_t *p = &prev->a[j];
			if (c >= 0) orc_bwt_extend(x, p, ok, 1);
			if (c < 0 || ok[c].size < min_intv) {
				if (curr->n == 0) {
					if (mem->n == 0 || (uint64_t)(i + 1) < mem->a[mem->n - 1].info >> 32) {
						ik = *p; ik.info |= (uint64_t)(i + 1) << 32;
						vec_push(*mem, ik);
					}
				}
			} else if (curr->n == 0 || ok[c].size != curr->a[curr->n - 1].size) {
				ok[c].info = p->info;
				vec_push(*curr, ok[c]);
			}
		}
		if (curr->n == 0) break;
		swap = curr; curr = prev; prev = swap;
	}
	reverse_intvs(mem);
	return ret;
}

int orc_bwt_seed_strategy1(const orc_index_t *x, int len, const uint8_t *q, int start, int min_len, uint64_t max_intv, orc_intv_t *mem)
{
	int i, c;
	orc_intv_t ik, ok[4];
	memset(mem, 0, sizeof(*mem));
	if (q[start] > 3) return start + 1;
	set_intv(x, q[start], &ik);
	for (i = start + 1; i < len; ++i) {
		if (q[i] < 4) {
			c = 3 - q[i];
			orc_bwt_extend(x, &ik, ok, 0);
			if (ok[c].size < max_intv && i - start >= min_len) {
				*mem = ok[c];
				mem->info = (uint64_t)start << 32 | (uint64_t)(i + 1);
				return i + 1;
			}
			ik = ok[c];
		} else return i + 1;
	}
	return len;
}

#define INTV_LT(a, b) ((a).info < (b).info)
ORC_SORT_DECL(intv, orc_intv_t, INTV_LT)

void orc_collect_intv_core(const orc_index_t *x, const orc_opt_t *opt, int len, const uint8_t *seq, intv_v *out)
{
	int i, start = 0;
	size_t k, old_n;
	int split_len = (int)(opt->min_seed_len * opt->split_factor + .499);
	intv_v mem1, tmp[2];
	vec_init(mem1); vec_init(tmp[0]); vec_init(tmp[1]);
	out->n = 0;
	while (start < len) { /* round 1: all SMEMs */
		if (seq[start] < 4) {
			start = orc_bwt_smem1(x, len, seq, start, 1, &mem1, tmp);
			for (k = 0; k < mem1.n; ++k) {
				orc_intv_t *p = &mem1.a[k];
				int slen = (int)((uint32_t)p->info - (uint32_t)(p->info >> 32));
				if (slen >= opt->min_seed_len) vec_push(*out, *p);
			}
		} else ++start;
	}
	old_n = out->n;
	for (k = 0; k < old_n; ++k) { /* round 2: re-seed inside long, nearly unique SMEMs */
		orc_intv_t p = out->a[k];
		int beg = (int)(p.info >> 32), end = (int)(int32_t)p.info;
		size_t t;
		if (end - beg < split_len || p.size > (uint64_t)opt->split_width) continue;
		orc_bwt_smem1(x, len, seq, (beg + end) >> 1, p.size + 1, &mem1, tmp);
		for (t = 0; t < mem1.n; ++t)
			if ((int)((uint32_t)mem1.a[t].info - (uint32_t)(mem1.a[t].info >> 32)) >= opt->min_seed_len)
				vec_push(*out, mem1.a[t]);
	}
	if (opt->max_mem_intv > 0) { /* round 3: LAST-like seeds */
		start = 0;
		while (start < len) {
			if (seq[start] < 4) {
				orc_intv_t m;
				start = orc_bwt_seed_strategy1(x, len, seq, start, opt->min_seed_len, opt->max_mem_intv, &m);
				if (m.size > 0) vec_push(*out, m);
			} else ++start;
		}
	}
	orc_sort_intv(out->n, out->a);
	vec_free(mem1); vec_free(tmp[0]); vec_free(tmp[1]);
	(void)i;
}

/* ------------------------------------------------------------------ SA */

uint64_t orc_bwt_sa(const orc_index_t *x, uint64_t k)
{
	uint64_t sa = 0, mask = (uint64_t)x->sa_intv - 1, occ0 = orc_tls_cnt.n_occ;
	++orc_tls_cnt.n_sa_lookup;
	while (k & mask) {
		++sa;
		if (k == x->primary) k = 0;
		else {
			uint64_t kk = k - (k > x->primary);
			int c = bwt_char(x, kk);
			k = x->L2[c] + bwt_occ1(x, k, c);
		}
	}
	++orc_tls_cnt.n_sa;
	orc_tls_cnt.n_occ_sa += orc_tls_cnt.n_occ - occ0;
	return sa + x->sa[k / x->sa_intv];
}

/* ------------------------------------------------------------------ reference access */

int64_t orc_depos(const orc_index_t *x, int64_t pos, int *is_rev)
{
	return (*is_rev = (pos >= x->l_pac)) ? (x->l_pac << 1) - 1 - pos : pos;
}

int orc_pos2rid(const orc_index_t *x, int64_t pos_f)
{
	int left, mid, right;
	if (pos_f >= x->l_pac) return -1;
	left = 0; mid = 0; right = x->n_seqs;
	while (left < right) {
		mid = (left + right) >> 1;
		if (pos_f >= x->anns[mid].offset) {
			if (mid == x->n_seqs - 1) break;
			if (pos_f < x->anns[mid + 1].offset) break;
			left = mid + 1;
		} else right = mid;
	}
	return mid;
}

int orc_intv2rid(const orc_index_t *x, int64_t rb, int64_t re)
{
	int is_rev, rid_b, rid_e;
	if (rb < x->l_pac && re > x->l_pac) return -2;
	rid_b = orc_pos2rid(x, orc_depos(x, rb, &is_rev));
	rid_e = rb < re ? orc_pos2rid(x, orc_depos(x, re - 1, &is_rev)) : rid_b;
	return rid_b == rid_e ? rid_b : -1;
}

#define GET_PAC(pac, l) ((pac)[(l) >> 2] >> ((~(l) & 3) << 1) & 3)

uint8_t *orc_get_seq(int64_t l_pac, const uint8_t *pac, int64_t beg, int64_t end, int64_t *len)
{
	uint8_t *seq = 0;
	if (end < beg) { int64_t t = beg; beg = end; end = t; }
	if (end > l_pac << 1) end = l_pac << 1;
	if (beg < 0) beg = 0;
	if (beg >= l_pac || end <= l_pac) {
		int64_t k, l = 0;
		*len = end - beg;
		seq = malloc(end - beg + 1);
		if (beg >= l_pac) {
			int64_t beg_f = (l_pac << 1) - 1 - end, end_f = (l_pac << 1) - 1 - beg;
			for (k = end_f; k > beg_f; --k) seq[l++] = 3 - GET_PAC(pac, k);
		} else {
			for (k = beg; k < end; ++k) seq[l++] = GET_PAC(pac, k);
		}
	} else *len = 0;
	return seq;
}

uint8_t *orc_fetch_seq(const orc_index_t *x, int64_t *beg, int64_t mid, int64_t *end, int *rid)
{
	int64_t far_beg, far_end, len;
	int is_rev;
	uint8_t *seq;
	if (*end < *beg) { int64_t t = *beg; *beg = *end; *end = t; }
	*rid = orc_pos2rid(x, orc_depos(x, mid, &is_rev));
	far_beg = x->anns[*rid].offset;
	far_end = far_beg + x->anns[*rid].len;
	if (is_rev) {
		int64_t t = far_beg;
		far_beg = (x->l_pac << 1) - far_end;
		far_end = (x->l_pac << 1) - t;
	}
	*beg = *beg > far_beg ? *beg : far_beg;
	*end = *end < far_end ? *end : far_end;
	seq = orc_get_seq(x->l_pac, x->pac, *beg, *end, &len);
	if (seq == 0 || *end - *beg != len) {
		fprintf(stderr, "[orc] fetch_seq: inconsistent window [%lld,%lld) len=%lld\n", (long long)*beg, (long long)*end, (long long)len);
		abort();
	}
	return seq;
}
