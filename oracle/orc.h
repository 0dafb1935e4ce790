/*
 * orc.h -- CPU ORACLE for the BWA-MEM hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This directory is a plain-C, scalar restatement of the algorithm that the
 * reference (broadinstitute/gatk-bwamem-jni) runs behind
 *   jnibwa_createAlignments()            /root/reference/src/main/c/jnibwa.c:197-235
 * i.e. lh3/bwa `mem_process_seqs` @ cb950614ce7217788780b9a8d445c64cd4d8f62e
 * (pinned by /root/reference/src/main/c/Makefile:17,26; NOT vendored, not
 * buildable offline).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product (the CUDA
 * engine under gatk-bwamem-jni_b200/) never links or calls anything here.
 *
 * PARITY STATUS: pinned against the reference's six known-answer tests
 * (BwaMemIndexTest.java:45-127) and the five checked-in index files; every
 * field/heuristic those tests do not assert (mapQ, AS, XS, MD, XA, secondary /
 * supplementary logic, rescue, pestat numerics) is "parity unpinned": it
 * follows the published bwa-mem algorithm as restated in SURVEY.md App. B.
 */
#ifndef ORC_H_
#define ORC_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- option block: 168 bytes, offsets pinned by BwaMemAligner.java:46-137 ---- */
typedef struct {
	int32_t a, b;                 /* 0,4   match score, mismatch penalty */
	int32_t o_del, e_del;         /* 8,12  */
	int32_t o_ins, e_ins;         /* 16,20 */
	int32_t pen_unpaired;         /* 24 */
	int32_t pen_clip5, pen_clip3; /* 28,32 */
	int32_t w;                    /* 36 band width */
	int32_t zdrop;                /* 40 */
	int32_t pad0_;                /* 44 */
	uint64_t max_mem_intv;        /* 48 */
	int32_t T;                    /* 56 */
	int32_t flag;                 /* 60 */
	int32_t min_seed_len;         /* 64 */
	int32_t min_chain_weight;     /* 68 */
	int32_t max_chain_extend;     /* 72 */
	float split_factor;           /* 76 */
	int32_t split_width;          /* 80 */
	int32_t max_occ;              /* 84 */
	int32_t max_chain_gap;        /* 88 */
	int32_t n_threads;            /* 92 */
	int32_t chunk_size;           /* 96 */
	float mask_level;             /* 100 */
	float drop_ratio;             /* 104 */
	float XA_drop_ratio;          /* 108 */
	float mask_level_redun;       /* 112 */
	float mapQ_coef_len;          /* 116 */
	int32_t mapQ_coef_fac;        /* 120 */
	int32_t max_ins;              /* 124 */
	int32_t max_matesw;           /* 128 */
	int32_t max_XA_hits;          /* 132 */
	int32_t max_XA_hits_alt;      /* 136 */
	int8_t mat[25];               /* 140 */
	int8_t pad1_[3];              /* 165 */
} orc_opt_t;

typedef struct { /* mem_pestat_t: 32 bytes (…BwaMemIndex.c:21-40 fills slot 1) */
	int32_t low, high;
	int32_t failed;
	int32_t pad_;
	double avg, std;
} orc_pestat_t;

#define ORC_F_PE        0x2
#define ORC_F_NOPAIRING 0x4
#define ORC_F_ALL       0x8
#define ORC_F_NO_MULTI  0x10
#define ORC_F_NO_RESCUE 0x20
#define ORC_F_PRIMARY5  0x800

/* ---- index image (layout: SURVEY.md App. A.5; written by bwa_idx2mem, jnibwa.c:126-152) ---- */
typedef struct {
	int64_t offset;
	int32_t len, n_ambs;
	uint32_t gi;
	int32_t is_alt;
	const char *name, *anno;
} orc_contig_t;

typedef struct {
	/* FM index */
	uint64_t primary, L2[5], seq_len, bwt_size;
	const uint32_t *bwt;
	int32_t sa_intv;
	uint64_t n_sa;
	const uint64_t *sa;
	/* reference */
	int64_t l_pac;
	int32_t n_seqs, n_holes;
	orc_contig_t *anns;
	const uint8_t *pac;
	/* backing */
	void *mem;
	size_t l_mem;
	int mem_is_mmap;
} orc_index_t;

/* work counters (algorithmic work) */
typedef struct {
	uint64_t n_occ;      /* occ-block touches (64 B each) */
	uint64_t n_sa;       /* sampled-SA reads (8 B each) */
	uint64_t n_extend;   /* bwt_extend calls */
	uint64_t n_sa_lookup;/* bwt_sa calls */
	uint64_t cells_ext;  /* ksw_extend2 cells */
	uint64_t cells_msw;  /* ksw_align2 cells */
	uint64_t cells_glb;  /* ksw_global2 cells */
	uint64_t n_ext_calls, n_msw_calls, n_glb_calls;
	uint64_t n_seeds, n_chains, n_regs;
	uint64_t n_occ_sa;   /* the part of n_occ touched by bwt_sa walks (the rest belongs to SMEM seeding) */
	uint64_t n_ext_retry;/* ksw_extend2 calls that were a second band try (w doubled) */
	uint64_t n_ext_zdrop;/* ksw_extend2 calls ended by the z-drop test */
} orc_counters_t;

/* ---- C-ABI of the oracle (mirrors jnibwa.h:11-16) ---- */
orc_index_t *orc_openIndex(int fd);
orc_index_t *orc_openIndexFromMemory(const void *mem, size_t len); /* copies */
int orc_destroyIndex(orc_index_t *idx);
void *orc_getRefContigNames(orc_index_t *idx, size_t *size);
orc_opt_t *orc_opt_init(void);
void *orc_createAlignments(orc_index_t *idx, orc_opt_t *opt, orc_pestat_t *pes, char *seqs, size_t *size);
void orc_free(void *p);
void orc_get_counters(orc_counters_t *out, int reset);

/* ---- intermediate dumps for stage-by-stage parity tests ---- */
typedef struct { uint64_t x0, x1, size, info; } orc_intv_t;             /* bwtintv_t */
typedef struct { int64_t rbeg; int32_t qbeg, len, score; } orc_seed_t;  /* mem_seed_t */
typedef struct {
	int64_t rb, re;
	int32_t qb, qe;
	int32_t rid;
	int32_t score, truesc, sub, alt_sc, csub, sub_n, w, seedcov, secondary, secondary_all, seedlen0;
	int32_t n_comp, is_alt;
	float frac_rep;
	uint64_t hash;
} orc_reg_t;                                                            /* mem_alnreg_t */

/* seeds of one read (nt4-encodes a private copy): returns malloc'd array */
orc_intv_t *orc_collect_intv(const orc_index_t *idx, const orc_opt_t *opt, int len, const char *seq, int *n);
/* chains after mem_chain + mem_chain_flt: flat dump: for each chain {n_seeds, rid, w, kept, pos(lo,hi)} then seeds */
int64_t *orc_chains(const orc_index_t *idx, const orc_opt_t *opt, int len, const char *seq, int *n_words);
/* regs after mem_align1_core */
orc_reg_t *orc_align1(const orc_index_t *idx, const orc_opt_t *opt, int len, const char *seq, int *n);

/* DP primitives exposed for kernel unit tests */
int orc_ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                    int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                    int *qle, int *tle, int *gtle, int *gscore, int *max_off);
int orc_ksw_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                    int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar, uint32_t **cigar);
typedef struct { int score, te, qe, score2, te2, tb, qb; } orc_kswr_t;
orc_kswr_t orc_ksw_align2(int qlen, uint8_t *query, int tlen, uint8_t *target, const int8_t *mat,
                          int o_del, int e_del, int o_ins, int e_ins, int xtra);
uint64_t orc_bwt_sa(const orc_index_t *idx, uint64_t k);
/* ksw_align2's striped kernel: 1 = SSE2 intrinsics (ksw_sse2.c; the default on x86-64), 0 = the scalar emulation */
void orc_set_simd(int on);
int orc_get_simd(void);
int orc_have_sse2(void);
void orc_sw_striped_sse2(int is_u8, int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                         int o_del, int e_del, int o_ins, int e_ins, int xtra, int out[5]);

#define ORC_KSW_XBYTE  0x10000
#define ORC_KSW_XSTOP  0x20000
#define ORC_KSW_XSUBO  0x40000
#define ORC_KSW_XSTART 0x80000

#ifdef __cplusplus
}
#endif
#endif
