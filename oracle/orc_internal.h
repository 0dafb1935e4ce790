/* orc_internal.h -- internals of the CPU oracle (TEST INFRASTRUCTURE ONLY; see orc.h). */
#ifndef ORC_INTERNAL_H_
#define ORC_INTERNAL_H_

#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <limits.h>
#include "orc.h"

/* ---- growable vector ---- */
#define VEC(type) struct { size_t n, m; type *a; }
#define vec_init(v) ((v).n = (v).m = 0, (v).a = 0)
#define vec_free(v) (free((v).a), (v).n = (v).m = 0, (v).a = 0)
#define vec_reserve(v, cap) do { if ((v).m < (size_t)(cap)) { (v).m = (cap); (v).a = realloc((v).a, sizeof(*(v).a) * (v).m); } } while (0)
#define vec_push(v, x) do { if ((v).n == (v).m) { (v).m = (v).m ? (v).m << 1 : 4; (v).a = realloc((v).a, sizeof(*(v).a) * (v).m); } (v).a[(v).n++] = (x); } while (0)

typedef VEC(orc_intv_t) intv_v;
typedef VEC(int) int_v;
typedef VEC(uint64_t) u64_v;

/*
 * klib-style introsort.  The permutation it leaves for
 * EQUAL keys is observable in bwa-mem's results (hazard H4), so the exact
 * control flow is restated: median-of-three with the particular comparison
 * order, Hoare scan that never looks at *s, ranges <=16 left unsorted, depth
 * overflow -> comb sort, one final insertion sort over the whole array.
 */
#define ORC_SORT_DECL(name, type_t, LT) \
static inline void orc_isort_##name(type_t *s, type_t *t) { \
	type_t *i, *j, tmp; \
	for (i = s + 1; i < t; ++i) \
		for (j = i; j > s && LT(*j, *(j - 1)); --j) { tmp = *j; *j = *(j - 1); *(j - 1) = tmp; } \
} \
static inline void orc_combsort_##name(size_t n, type_t *a) { \
	const double shrink = 1.2473309501039786540366528676643; \
	int swapped; size_t gap = n; type_t tmp, *i, *j; \
	do { \
		if (gap > 2) { gap = (size_t)(gap / shrink); if (gap == 9 || gap == 10) gap = 11; } \
		swapped = 0; \
		for (i = a; i < a + n - gap; ++i) { j = i + gap; if (LT(*j, *i)) { tmp = *i; *i = *j; *j = tmp; swapped = 1; } } \
	} while (swapped || gap > 2); \
	if (gap != 1) orc_isort_##name(a, a + n); \
} \
static inline void orc_sort_##name(size_t n, type_t *a) { \
	struct frame { type_t *l, *r; int depth; } stack[8 * 64 + 2], *top = stack; \
	int d; type_t rp, tmp, *s, *t, *i, *j, *k; \
	if (n < 1) return; \
	if (n == 2) { if (LT(a[1], a[0])) { tmp = a[0]; a[0] = a[1]; a[1] = tmp; } return; } \
	for (d = 2; (1ul << d) < n; ++d) {} \
	s = a; t = a + (n - 1); d <<= 1; \
	for (;;) { \
		if (s < t) { \
			if (--d == 0) { orc_combsort_##name((size_t)(t - s) + 1, s); t = s; continue; } \
			i = s; j = t; k = i + ((j - i) >> 1) + 1; \
			if (LT(*k, *i)) { if (LT(*k, *j)) k = j; } \
			else k = LT(*j, *i) ? i : j; \
			rp = *k; \
			if (k != t) { tmp = *k; *k = *t; *t = tmp; } \
			for (;;) { \
				do ++i; while (LT(*i, rp)); \
				do --j; while (i <= j && LT(rp, *j)); \
				if (j <= i) break; \
				tmp = *i; *i = *j; *j = tmp; \
			} \
			tmp = *i; *i = *t; *t = tmp; \
			if (i - s > t - i) { \
				if (i - s > 16) { top->l = s; top->r = i - 1; top->depth = d; ++top; } \
				s = t - i > 16 ? i + 1 : t; \
			} else { \
				if (t - i > 16) { top->l = i + 1; top->r = t; top->depth = d; ++top; } \
				t = i - s > 16 ? i - 1 : s; \
			} \
		} else { \
			if (top == stack) { orc_isort_##name(a, a + n); return; } \
			--top; s = top->l; t = top->r; d = top->depth; \
		} \
	} \
}

/* ---- alignment region (mem_alnreg_t) ---- */
typedef orc_reg_t reg_t;
typedef VEC(reg_t) reg_v;

typedef orc_seed_t seed_t;

typedef struct {
	int n, m, first, rid;
	uint32_t w;      /* weight (29 bits in bwa) */
	int kept, is_alt;
	float frac_rep;
	int64_t pos;
	seed_t *seeds;
	uint64_t ins_seq; /* insertion serial: orders equal-pos chains like the reference's tree (see mem.c) */
} chain_t;
typedef VEC(chain_t) chain_v;

/* ---- alignment record (mem_aln_t) ---- */
typedef struct {
	int64_t pos;
	int rid, flag;
	uint32_t is_rev, is_alt, mapq, NM;
	int n_cigar;
	uint32_t *cigar;   /* n_cigar words followed by the NUL-terminated MD string */
	char *XA;
	int score, sub, alt_sc;
} aln_t;

/* per-thread counters */
extern __thread orc_counters_t orc_tls_cnt;

/* fm.c */
void orc_bwt_occ4(const orc_index_t *x, uint64_t k, uint64_t cnt[4]);
void orc_bwt_extend(const orc_index_t *x, const orc_intv_t *ik, orc_intv_t ok[4], int is_back);
int orc_bwt_smem1(const orc_index_t *x, int len, const uint8_t *q, int start, uint64_t min_intv, intv_v *mem, intv_v tmp[2]);
int orc_bwt_seed_strategy1(const orc_index_t *x, int len, const uint8_t *q, int start, int min_len, uint64_t max_intv, orc_intv_t *mem);
void orc_collect_intv_core(const orc_index_t *x, const orc_opt_t *opt, int len, const uint8_t *seq, intv_v *out);

/* ref.c (in fm.c) */
int orc_pos2rid(const orc_index_t *x, int64_t pos_f);
int64_t orc_depos(const orc_index_t *x, int64_t pos, int *is_rev);
int orc_intv2rid(const orc_index_t *x, int64_t rb, int64_t re);
uint8_t *orc_get_seq(int64_t l_pac, const uint8_t *pac, int64_t beg, int64_t end, int64_t *len);
uint8_t *orc_fetch_seq(const orc_index_t *x, int64_t *beg, int64_t mid, int64_t *end, int *rid);
extern const uint8_t orc_nt4_table[256];

/* mem.c */
void orc_mem_chain(const orc_index_t *x, const orc_opt_t *opt, int len, const uint8_t *seq, chain_v *out);
int orc_mem_chain_flt(const orc_opt_t *opt, int n_chn, chain_t *a);
void orc_mem_flt_chained_seeds(const orc_index_t *x, const orc_opt_t *opt, int l_query, const uint8_t *query, int n_chn, chain_t *a);
void orc_mem_chain2aln(const orc_index_t *x, const orc_opt_t *opt, int l_query, const uint8_t *query, const chain_t *c, reg_v *av);
int orc_mem_sort_dedup_patch(const orc_index_t *x, const orc_opt_t *opt, uint8_t *query, int n, reg_t *a);
void orc_mem_align1_core(const orc_index_t *x, const orc_opt_t *opt, int l_seq, uint8_t *seq, reg_v *regs);
int orc_mem_mark_primary_se(const orc_opt_t *opt, int n, reg_t *a, int64_t id);
void orc_mem_reorder_primary5(int T, reg_v *a);
int orc_mem_approx_mapq_se(const orc_opt_t *opt, const reg_t *a);
aln_t orc_mem_reg2aln(const orc_index_t *x, const orc_opt_t *opt, int l_query, const uint8_t *query, const reg_t *ar);
char **orc_mem_gen_alt(const orc_index_t *x, const orc_opt_t *opt, const reg_v *a, int l_query, const uint8_t *query);
uint32_t *orc_gen_cigar2(const int8_t mat[25], int o_del, int e_del, int o_ins, int e_ins, int w_, int64_t l_pac, const uint8_t *pac,
                         int l_query, uint8_t *query, int64_t rb, int64_t re, int *score, int *n_cigar, int *NM);
uint64_t orc_hash64(uint64_t key);

/* output blob (fmt_BAMish stream; jnibwa.c:43-97) */
typedef struct { size_t l, m; char *s; } blob_t;
void orc_aln2blob(blob_t *str, int n, const aln_t *list, int which, const aln_t *mate);
void orc_mem_reg2sam(const orc_index_t *x, const orc_opt_t *opt, int l_seq, const uint8_t *seq, reg_v *a, int extra_flag, const aln_t *m, blob_t *out);

/* pair.c */
void orc_mem_pestat(const orc_opt_t *opt, int64_t l_pac, int n, const reg_v *regs, orc_pestat_t pes[4]);
int orc_mem_sam_pe(const orc_index_t *x, const orc_opt_t *opt, const orc_pestat_t pes[4], uint64_t id,
                   const int l_seq[2], uint8_t *const seq[2], reg_v a[2], blob_t out[2]);

#endif
