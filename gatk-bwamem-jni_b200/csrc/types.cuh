// types.cuh -- device-side data layout of the B200 BWA-MEM engine.
//
// Option block: the 168-byte mem_opt_t whose offsets the Java side hard-codes
// (/root/reference/src/main/java/org/broadinstitute/hellbender/utils/bwa/BwaMemAligner.java:46-137).
// Index: the BwaMemIndex image resident in HBM.
#pragma once
#include "simt.cuh"

struct MemOpt { // == mem_opt_t, 168 bytes
	i32 a, b, o_del, e_del, o_ins, e_ins, pen_unpaired, pen_clip5, pen_clip3, w, zdrop, pad0_;
	u64 max_mem_intv;
	i32 T, flag, min_seed_len, min_chain_weight, max_chain_extend;
	float split_factor;
	i32 split_width, max_occ, max_chain_gap, n_threads, chunk_size;
	float mask_level, drop_ratio, XA_drop_ratio, mask_level_redun, mapQ_coef_len;
	i32 mapQ_coef_fac, max_ins, max_matesw, max_XA_hits, max_XA_hits_alt;
	i8 mat[25];
	i8 pad1_[3];
};
static_assert(sizeof(MemOpt) == 168, "mem_opt_t layout");

struct PeStat { i32 low, high, failed, pad_; double avg, std; }; // mem_pestat_t, 32 bytes
static_assert(sizeof(PeStat) == 32, "mem_pestat_t layout");

#define MEM_F_PE        0x2
#define MEM_F_NOPAIRING 0x4
#define MEM_F_ALL       0x8
#define MEM_F_NO_MULTI  0x10
#define MEM_F_NO_RESCUE 0x20
#define MEM_F_PRIMARY5  0x800

// ---- capacity / limit flags raised by device code (any flag fails the batch, or re-routes a thread-level item to the warp path)
#define ERR_Z_OVERFLOW     1u
#define ERR_CIGAR_OVERFLOW 2u
#define ERR_MD_OVERFLOW    4u
#define ERR_XA_OVERFLOW    8u
#define ERR_BROW_OVERFLOW  16u
#define ERR_VAR_OVERFLOW   32u
#define ERR_REG_OVERFLOW   64u
#define ERR_INTV_OVERFLOW  128u
#define ERR_QLEN_OVERFLOW  256u
#define ERR_NEED_WARP      512u  // not an error: a thread-level item asks for the warp-level kernel

// ---- index resident in HBM --------------------------------------------------------------------
struct DevIndex {
	u64 primary, L2[5], seq_len;
	const uint4* bwt;   // occurrence array: one 32-byte sector (4 x u32 counts + 64 symbols as two bit-planes) per 64 rows (fm.cuh)
	const u64* occ_sb; i32 occ_sb_shift; // super-block bases of the u32 counts (4 x u64 per 2^shift sectors); nullptr: counts are absolute
	const u64* sa;      // sampled suffix array (row k*sa_intv -> text position)
	u32 sa_intv;        // power of two (32 in the image; smaller when densified at load)
	i64 l_pac;
	i32 n_seqs;
	const i64* ctg_off; // n_seqs+1 cumulative contig offsets on the forward strand
	const u8* ctg_alt;  // n_seqs: 1 for an ALT contig (bntann1_t.is_alt)
	const u8* pac;      // 2 bits per base
	const double* log_tab;  // glibc log(i), i in [0, LOG_TAB_N)  (hazard H5: libm results are tabulated on the host)
	i32 log_tab_n;
	// Bidirectional intervals of every string of 2..kmer_k bases (forward extension from its first base), 16 bytes each,
	// level L at offset (4^L - 16) / 3: a forward scan reads its first kmer_k - 1 extensions here (one 16-byte read,
	// mostly from L2) instead of walking two 64-byte occurrence blocks each.  nullptr / 0: no tables.
	const uint4* kmer_tab; i32 kmer_k;
};

// ---- per-batch stage records ------------------------------------------------------------------
struct Intv { u64 x0, x1, size; u32 beg, end; };            // bwtintv_t; info = beg<<32|end
struct Seed { i64 rbeg; i32 qbeg, len, score, next; };       // mem_seed_t (+ chain link)
struct ChainK { // a kept chain handed to the extension stage
	u32 seed_start, n_seeds; // into the kept-seed pool
	i32 rid; float frac_rep;
	u32 w; i32 kept;
	i64 pos;
};
struct Reg { // mem_alnreg_t
	i64 rb, re;
	i32 qb, qe;
	i32 rid;
	i32 score, truesc, sub, alt_sc, csub, sub_n, w, seedcov, secondary, secondary_all, seedlen0;
	i32 n_comp, is_alt;
	float frac_rep;
	u64 hash;
};

struct AlnRec { // one output alignment, ready for the record encoder (fmt_BAMish, jnibwa.c:43-97)
	i32 flag;             // final SAM flag incl. 0x1/0x4/0x8/0x10/0x20 (bwa sets these on copies before the formatter hook)
	u32 mapq;
	i32 rid, pos, NM, score, sub;
	i32 n_cigar;
	u32 var_off;          // word offset into the var pool: n_cigar words (MIDSH numbering), then the MD bytes
	i32 md_len;
	u32 xa_off; i32 xa_len; // XA text in the var pool (word offset, byte length)
	i32 m_rid, m_pos, tlen; // mate triplet, written iff (flag & 9) == 1
};

struct MateDesc { i32 rid; i64 pos; u32 is_rev; i32 ref_len; }; // what the formatter reads of the mate's mem_aln_t
