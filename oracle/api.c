/*
 * api.c -- CPU ORACLE (test infrastructure only; see orc.h): the C-ABI of the
 * oracle.  Mirrors /root/reference/src/main/c/jnibwa.c:174-235
 * (jnibwa_getRefContigNames, jnibwa_createAlignments) and the batch driver
 * mem_process_seqs / worker1 / worker2 of un-vendored lh3/bwa.
 */
#include <math.h>
#include <pthread.h>
#include "orc_internal.h"

static orc_counters_t g_cnt;
static pthread_mutex_t g_cnt_lock = PTHREAD_MUTEX_INITIALIZER;

static void flush_counters(void)
{
	uint64_t *g = (uint64_t *)&g_cnt, *t = (uint64_t *)&orc_tls_cnt;
	size_t i, n = sizeof(orc_counters_t) / 8;
	pthread_mutex_lock(&g_cnt_lock);
	for (i = 0; i < n; ++i) { g[i] += t[i]; t[i] = 0; }
	pthread_mutex_unlock(&g_cnt_lock);
}

void orc_get_counters(orc_counters_t *out, int reset)
{
	flush_counters();
	pthread_mutex_lock(&g_cnt_lock);
	if (out) *out = g_cnt;
	if (reset) memset(&g_cnt, 0, sizeof(g_cnt));
	pthread_mutex_unlock(&g_cnt_lock);
}

void orc_free(void *p) { free(p); }

orc_opt_t *orc_opt_init(void)
{
	orc_opt_t *o = calloc(1, sizeof(orc_opt_t));
	int i, j, k;
	o->flag = 0;
	o->a = 1; o->b = 4;
	o->o_del = o->o_ins = 6;
	o->e_del = o->e_ins = 1;
	o->w = 100;
	o->T = 30;
	o->zdrop = 100;
	o->pen_unpaired = 17;
	o->pen_clip5 = o->pen_clip3 = 5;
	o->max_mem_intv = 20;
	o->min_seed_len = 19;
	o->split_width = 10;
	o->max_occ = 500;
	o->max_chain_gap = 10000;
	o->max_ins = 10000;
	o->mask_level = 0.50f;
	o->drop_ratio = 0.50f;
	o->XA_drop_ratio = 0.80f;
	o->split_factor = 1.5f;
	o->chunk_size = 10000000;
	o->n_threads = 1;
	o->max_XA_hits = 5;
	o->max_XA_hits_alt = 200;
	o->max_matesw = 50;
	o->mask_level_redun = 0.95f;
	o->min_chain_weight = 0;
	o->max_chain_extend = 1 << 30;
	o->mapQ_coef_len = 50; o->mapQ_coef_fac = (int)log(o->mapQ_coef_len);
	for (i = k = 0; i < 4; ++i) {
		for (j = 0; j < 4; ++j) o->mat[k++] = i == j ? (int8_t)o->a : (int8_t)-o->b;
		o->mat[k++] = -1;
	}
	for (j = 0; j < 5; ++j) o->mat[k++] = -1;
	return o;
}

void *orc_getRefContigNames(orc_index_t *x, size_t *size)
{
	int i, n = x->n_seqs;
	size_t sz = 4 + 4 * (size_t)n;
	char *buf, *p;
	for (i = 0; i < n; ++i) sz += strlen(x->anns[i].name);
	buf = malloc(sz);
	*(int32_t *)buf = n;
	p = buf + 4;
	for (i = 0; i < n; ++i) {
		int32_t len = (int32_t)strlen(x->anns[i].name);
		memcpy(p, &len, 4); p += 4;
		memcpy(p, x->anns[i].name, len); p += len;
	}
	*size = sz;
	return buf;
}

/* ---- batch driver ---- */
typedef struct {
	const orc_index_t *x;
	const orc_opt_t *opt;
	int n;               /* number of reads */
	int *l_seq;
	uint8_t **seq;
	reg_v *regs;
	blob_t *sam;
	orc_pestat_t pes[4];
	int n_units, phase, n_threads;
	int next; /* next unit to hand out (work stealing over reads, like kt_for) */
} batch_t;

typedef struct { batch_t *b; int tid; } targ_t;

static void do_unit(batch_t *b, int i)
{
	const orc_opt_t *opt = b->opt;
	if (b->phase == 1) {
		if (!(opt->flag & ORC_F_PE)) orc_mem_align1_core(b->x, opt, b->l_seq[i], b->seq[i], &b->regs[i]);
		else {
			orc_mem_align1_core(b->x, opt, b->l_seq[i << 1 | 0], b->seq[i << 1 | 0], &b->regs[i << 1 | 0]);
			orc_mem_align1_core(b->x, opt, b->l_seq[i << 1 | 1], b->seq[i << 1 | 1], &b->regs[i << 1 | 1]);
		}
	} else {
		if (!(opt->flag & ORC_F_PE)) {
			orc_mem_mark_primary_se(opt, (int)b->regs[i].n, b->regs[i].a, i); /* n_processed is always 0 (jnibwa.c:214) */
			if (opt->flag & ORC_F_PRIMARY5) orc_mem_reorder_primary5(opt->T, &b->regs[i]);
			orc_mem_reg2sam(b->x, opt, b->l_seq[i], b->seq[i], &b->regs[i], 0, 0, &b->sam[i]);
			free(b->regs[i].a);
		} else {
			orc_mem_sam_pe(b->x, opt, b->pes, (uint64_t)i, &b->l_seq[i << 1], &b->seq[i << 1], &b->regs[i << 1], &b->sam[i << 1]);
			free(b->regs[i << 1 | 0].a); free(b->regs[i << 1 | 1].a);
		}
	}
}

static void *thread_main(void *p_)
{
	targ_t *t = (targ_t *)p_;
	batch_t *b = t->b;
	int i;
	(void)t;
	for (;;) { /* units are handed out in small runs from a shared cursor: no thread waits for a slow neighbour's static share */
		int lo = __atomic_fetch_add(&b->next, 16, __ATOMIC_RELAXED), hi = lo + 16;
		if (lo >= b->n_units) break;
		if (hi > b->n_units) hi = b->n_units;
		for (i = lo; i < hi; ++i) do_unit(b, i); /* results do not depend on the schedule */
	}
	flush_counters();
	return 0;
}

static void run_phase(batch_t *b, int phase)
{
	int t, nt = b->n_threads;
	b->phase = phase; b->next = 0;
	if (nt <= 1) { int i; for (i = 0; i < b->n_units; ++i) do_unit(b, i); flush_counters(); return; }
	{
		pthread_t *tid = malloc(sizeof(pthread_t) * nt);
		targ_t *ta = malloc(sizeof(targ_t) * nt);
		for (t = 0; t < nt; ++t) { ta[t].b = b; ta[t].tid = t; pthread_create(&tid[t], 0, thread_main, &ta[t]); }
		for (t = 0; t < nt; ++t) pthread_join(tid[t], 0);
		free(tid); free(ta);
	}
}

void *orc_createAlignments(orc_index_t *x, orc_opt_t *opt, orc_pestat_t *pes0, char *pSeq, size_t *pBufSize)
{
	batch_t b;
	uint32_t nSeqs, i;
	size_t total = 0;
	char *res, *p;
	memset(&b, 0, sizeof(b));
	memcpy(&nSeqs, pSeq, 4); pSeq += 4;
	b.x = x; b.opt = opt; b.n = (int)nSeqs;
	b.n_threads = opt->n_threads > 1 ? opt->n_threads : 1;
	b.l_seq = malloc(sizeof(int) * (nSeqs + 1));
	b.seq = malloc(sizeof(uint8_t *) * (nSeqs + 1));
	b.regs = calloc(nSeqs + 1, sizeof(reg_v));
	b.sam = calloc(nSeqs + 1, sizeof(blob_t));
	for (i = 0; i < nSeqs; ++i) { /* the read bytes are encoded in place, like the reference does (hazard H12) */
		size_t l = strlen(pSeq);
		b.l_seq[i] = (int)l; b.seq[i] = (uint8_t *)pSeq;
		pSeq += l + 1;
	}
	b.n_units = (opt->flag & ORC_F_PE) ? (int)(nSeqs >> 1) : (int)nSeqs;
	run_phase(&b, 1);
	if (opt->flag & ORC_F_PE) {
		if (pes0) memcpy(b.pes, pes0, 4 * sizeof(orc_pestat_t));
		else orc_mem_pestat(opt, x->l_pac, (int)nSeqs, b.regs, b.pes);
	}
	run_phase(&b, 2);
	for (i = 0; i < nSeqs; ++i) total += b.sam[i].l; /* an odd trailing read in PE mode yields nothing (hazard H11) */
	res = malloc(total ? total : 1);
	for (i = 0, p = res; i < nSeqs; ++i) {
		if (b.sam[i].s) { memcpy(p, b.sam[i].s, b.sam[i].l); p += b.sam[i].l; free(b.sam[i].s); }
	}
	if ((opt->flag & ORC_F_PE) && (nSeqs & 1)) free(b.regs[nSeqs - 1].a);
	free(b.l_seq); free(b.seq); free(b.regs); free(b.sam);
	*pBufSize = total;
	return res;
}

/* ---- intermediates ---- */
static uint8_t *encode_copy(int len, const char *seq)
{
	uint8_t *s = malloc(len + 1);
	int i;
	for (i = 0; i < len; ++i) s[i] = (uint8_t)seq[i] < 4 ? (uint8_t)seq[i] : orc_nt4_table[(uint8_t)seq[i]];
	return s;
}

orc_intv_t *orc_collect_intv(const orc_index_t *x, const orc_opt_t *opt, int len, const char *seq, int *n)
{
	uint8_t *s = encode_copy(len, seq);
	intv_v v;
	vec_init(v);
	orc_collect_intv_core(x, opt, len, s, &v);
	free(s);
	*n = (int)v.n;
	return v.a;
}

int64_t *orc_chains(const orc_index_t *x, const orc_opt_t *opt, int len, const char *seq, int *n_words)
{
	uint8_t *s = encode_copy(len, seq);
	chain_v chn;
	VEC(int64_t) o;
	size_t i; int j;
	vec_init(o);
	orc_mem_chain(x, opt, len, s, &chn);
	chn.n = (size_t)orc_mem_chain_flt(opt, (int)chn.n, chn.a);
	orc_mem_flt_chained_seeds(x, opt, len, s, (int)chn.n, chn.a);
	vec_push(o, (int64_t)chn.n);
	for (i = 0; i < chn.n; ++i) {
		chain_t *c = &chn.a[i];
		vec_push(o, (int64_t)c->n); vec_push(o, (int64_t)c->rid); vec_push(o, (int64_t)c->w); vec_push(o, (int64_t)c->kept); vec_push(o, c->pos);
		for (j = 0; j < c->n; ++j) { vec_push(o, c->seeds[j].rbeg); vec_push(o, (int64_t)c->seeds[j].qbeg); vec_push(o, (int64_t)c->seeds[j].len); vec_push(o, (int64_t)c->seeds[j].score); }
		free(c->seeds);
	}
	free(chn.a); free(s);
	*n_words = (int)o.n;
	return o.a;
}

orc_reg_t *orc_align1(const orc_index_t *x, const orc_opt_t *opt, int len, const char *seq, int *n)
{
	uint8_t *s = encode_copy(len, seq);
	reg_v regs;
	orc_mem_align1_core(x, opt, len, s, &regs);
	free(s);
	*n = (int)regs.n;
	return regs.a;
}
