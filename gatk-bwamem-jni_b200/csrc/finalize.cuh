// finalize.cuh -- phase 2 of a batch, one warp per read (single-end) or per pair (paired-end):
// primary marking, MAPQ, mate rescue, pairing, XA, CIGAR/NM/MD and the output records.
// Replaces mem_mark_primary_se / mem_approx_mapq_se / mem_reg2aln / mem_gen_alt / mem_reg2sam /
// mem_matesw / mem_pair / mem_sam_pe of un-vendored lh3/bwa
// and the flag finalisation bwa applies before the reference's formatter hook fmt_BAMish
// (/root/reference/src/main/c/jnibwa.c:43-97).
#pragma once
#include "extend.cuh"

// ---- alignment tasks of the two-pass finalize (galign_thread.cuh) -------------------------------------------------
#define GA_CIG_CAP 46
#define GA_MD_CAP 176
#define GA_ITEM_MAX 16   // alignments one item (read / pair) may request; more -> the item is finalized by a warp directly
struct GaTask { u64 qoff; i64 rb, re; i32 qb, qe, w2, truesc; };   // query = codes at qoff + [qb, qe); band start w2
struct GaRes { i32 score, NM, nc, ml, valid; u32 cigar[GA_CIG_CAP]; char md[GA_MD_CAP]; };
enum { FIN_DIRECT = 0, FIN_COLLECT = 1, FIN_CONSUME = 2 };

struct FinalCtx {
	int mode;                         // FIN_DIRECT: alignments computed in place; FIN_COLLECT / FIN_CONSUME: see galign_thread.cuh
	const u8* seq_base;               // P.seq (task descriptors hold offsets)
	GaTask* lt; int lt_n;             // collect: this item's requests (GA_ITEM_MAX entries); lt_n counts all of them
	const GaRes* res; int res_n, res_k; // consume: this item's results and the cursor
	const DevIndex* x; const MemOpt* opt;
	WarpMem* wm; WarpScratch* ws; Counters* cnt;
	u32* var; u32 var_cap; u32* var_cursor; // bump-allocated pool for cigar / MD / XA payloads (words)
	const char* ctg_names; const u32* ctg_name_off; // contig names blob + offsets (n_seqs+1)
	const PeStat* pes;                // 4 orientations
	const double* pair_tab; const i64* pair_tab_off; // per-orientation pairing term for dist in [low,high] (hazard H5)
	bool rescued;                     // mate rescue already performed by the task pipeline (rescue.cuh)
};

template <int W>
DEV u32 var_alloc(FinalCtx& f, u32 n_words)
{
	LANE_DECL;
	WV(u32, o);
	FOR_LANES { L_(o) = 0; }
#ifndef B200_SIM
	if (lane == 0) o = atomic_add_u32(f.var_cursor, n_words);
#else
	o[0] = atomic_add_u32(f.var_cursor, n_words);
#endif
	u32 off = wv_get(o, 0);
	if (off + n_words > f.var_cap) { ONE_LANE { atomic_or_u32(f.ws->err, ERR_VAR_OVERFLOW); } return 0xffffffffu; }
	return off;
}

// ---- primary marking (B.11) --------------------------------------------------------------------
DEV void mark_primary_core(const MemOpt& opt, int n, Reg* a, u32* z)
{ // lane-0-only helper
	int tmp = opt.a + opt.b, nz = 0;
	tmp = opt.o_del + opt.e_del > tmp ? opt.o_del + opt.e_del : tmp;
	tmp = opt.o_ins + opt.e_ins > tmp ? opt.o_ins + opt.e_ins : tmp;
	z[nz++] = 0;
	for (int i = 1; i < n; ++i) {
		int k;
		for (k = 0; k < nz; ++k) {
			int j = (int)z[k];
			int b_max = a[j].qb > a[i].qb ? a[j].qb : a[i].qb;
			int e_min = a[j].qe < a[i].qe ? a[j].qe : a[i].qe;
			if (e_min > b_max) {
				int min_l = a[i].qe - a[i].qb < a[j].qe - a[j].qb ? a[i].qe - a[i].qb : a[j].qe - a[j].qb;
				if ((float)(e_min - b_max) >= (float)min_l * opt.mask_level) {
					if (a[j].sub == 0) a[j].sub = a[i].score;
					if (a[j].score - a[i].score <= tmp && (a[j].is_alt || !a[i].is_alt)) ++a[j].sub_n;
					break;
				}
			}
		}
		if (k == nz) z[nz++] = (u32)i;
		else a[i].secondary = (int)z[k];
	}
}

// mem_mark_primary_se; returns n_pri, the number of hits on the primary assembly.  With hits on ALT contigs (n_pri < n)
// the marking runs twice: over all hits (kept in secondary_all; a primary-assembly hit shadowed by an ALT hit remembers
// that hit's score in alt_sc), then -- hits re-ordered with the primary-assembly ones first -- over those alone.
// ztmp: n words of scratch (the z list of the core, then the rank table of the re-ordering).
struct RegLtHash2 { DEVM bool operator()(const Reg& a, const Reg& b) const { return a.is_alt < b.is_alt || (a.is_alt == b.is_alt && (a.score > b.score || (a.score == b.score && a.hash < b.hash))); } };
template <int W>
DEVNI int warp_mark_primary_se(const MemOpt& opt, int n, Reg* a, i64 id, u32* ztmp)
{
	LANE_DECL;
	if (n == 0) return 0;
	int n_pri = 0;
	for (int i = 0; i < n; ++i) if (!a[i].is_alt) ++n_pri;
	LANE0_ONLY {
		for (int i = 0; i < n; ++i) { a[i].sub = a[i].alt_sc = 0; a[i].secondary = a[i].secondary_all = -1; a[i].hash = hash64((u64)(id + i)); }
		ks_introsort<Reg, RegLtHash>((size_t)n, a, RegLtHash());
		mark_primary_core(opt, n, a, ztmp);
		if (n_pri == n) { for (int i = 0; i < n; ++i) a[i].secondary_all = a[i].secondary; }
		else {
			for (int i = 0; i < n; ++i) {
				a[i].secondary_all = i; // keep the rank of the first round
				if (!a[i].is_alt && a[i].secondary >= 0 && a[a[i].secondary].is_alt) a[i].alt_sc = a[a[i].secondary].score;
			}
			if (n_pri > 0) ks_introsort<Reg, RegLtHash2>((size_t)n, a, RegLtHash2());
			for (int i = 0; i < n; ++i) ztmp[a[i].secondary_all] = (u32)i;
			for (int i = 0; i < n; ++i) {
				if (a[i].secondary >= 0) { a[i].secondary_all = (int)ztmp[a[i].secondary]; if (a[i].is_alt) a[i].secondary = 0x7fffffff; }
				else a[i].secondary_all = -1;
			}
			if (n_pri > 0) { // mark primary for hits to the primary assembly only
				for (int i = 0; i < n_pri; ++i) { a[i].sub = 0; a[i].secondary = -1; }
				mark_primary_core(opt, n_pri, a, ztmp);
			}
		}
	}
	WARP_SYNC();
	return n_pri;
}

template <int W>
DEVNI void warp_reorder_primary5(int T, int n, Reg* a)
{
	LANE_DECL;
	LANE0_ONLY {
		int k, n_pri = 0, left_st = 0x7fffffff, left_k = -1;
		for (k = 0; k < n; ++k) if (a[k].secondary < 0 && !a[k].is_alt && a[k].score >= T) ++n_pri;
		if (n_pri > 1) {
			for (k = 0; k < n; ++k) {
				if (a[k].secondary >= 0 || a[k].is_alt || a[k].score < T) continue;
				if (a[k].qb < left_st) { left_st = a[k].qb; left_k = k; }
			}
			if (left_k != 0) {
				Reg t = a[0]; a[0] = a[left_k]; a[left_k] = t;
				for (k = 1; k < n; ++k) {
					if (a[k].secondary == 0) a[k].secondary = left_k; else if (a[k].secondary == left_k) a[k].secondary = 0;
					if (a[k].secondary_all == 0) a[k].secondary_all = left_k; else if (a[k].secondary_all == left_k) a[k].secondary_all = 0;
				}
			}
		}
	}
	WARP_SYNC();
}

DEV double tab_log(const DevIndex& x, int v, u32* err)
{
	if (v >= 0 && v < x.log_tab_n) return x.log_tab[v];
	atomic_or_u32(err, ERR_QLEN_OVERFLOW);
	return 0.;
}

DEV int approx_mapq_se(const DevIndex& x, const MemOpt& opt, const Reg& a, u32* err)
{
	int mapq, l, sub = a.sub ? a.sub : opt.min_seed_len * opt.a;
	double identity;
	sub = a.csub > sub ? a.csub : sub;
	if (sub >= a.score) return 0;
	l = a.qe - a.qb > a.re - a.rb ? a.qe - a.qb : (int)(a.re - a.rb);
	identity = 1. - (double)(l * opt.a - a.score) / (double)(opt.a + opt.b) / (double)l;
	if (a.score == 0) mapq = 0;
	else if (opt.mapQ_coef_len > 0) {
		double tmp = (float)l < opt.mapQ_coef_len ? 1. : (double)opt.mapQ_coef_fac / tab_log(x, l, err);
		tmp *= identity * identity;
		mapq = (int)(6.02 * (double)(a.score - sub) / (double)opt.a * tmp * tmp + .499);
	} else {
		mapq = (int)(30.0 * (1. - (double)sub / (double)a.score) * tab_log(x, a.seedcov, err) + .499);
		mapq = identity < 0.95 ? (int)((double)mapq * identity * identity + .499) : mapq;
	}
	if (a.sub_n > 0) mapq -= (int)(4.343 * tab_log(x, a.sub_n + 1, err) + .499);
	if (mapq > 60) mapq = 60;
	if (mapq < 0) mapq = 0;
	mapq = (int)((double)mapq * (1. - (double)a.frac_rep) + .499);
	return mapq;
}

// ---- region -> alignment record (B.9) ----------------------------------------------------------
struct Aln { // mem_aln_t with the payload still in warp scratch (ws.cigar / ws.md)
	i64 pos; i32 rid, flag; u32 is_rev, mapq; i32 NM, n_cigar, md_len, score, sub;
};

DEV int infer_bw(int l1, int l2, int score, int a, int q, int r)
{
	if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
	int w = (int)((double)((l1 < l2 ? l1 : l2) * a - score - q) / (double)r + 2.);
	int d = iabs(l1 - l2);
	if (w < d) w = d;
	return w;
}

// mem_reg2aln; ar == nullptr gives the unmapped record.  CIGAR (clips added, edge deletion squeezed) is left in
// ws.cigar[0..n_cigar) and the MD text in ws.md[0..md_len).
template <int W>
DEVNI Aln warp_reg2aln(FinalCtx& f, int l_query, const u8* query, const Reg* ar)
{
	LANE_DECL;
	const DevIndex& x = *f.x; const MemOpt& opt = *f.opt;
	Aln a; a.pos = -1; a.rid = -1; a.flag = 0; a.is_rev = 0; a.mapq = 0; a.NM = 0; a.n_cigar = 0; a.md_len = 0; a.score = 0; a.sub = 0;
	if (ar == nullptr || ar->rb < 0 || ar->re < 0) { a.flag |= 0x4; return a; }
	const Reg r = *ar;
	int qb = r.qb, qe = r.qe; i64 rb = r.rb, re = r.re;
	a.mapq = r.secondary < 0 ? (u32)approx_mapq_se(x, opt, r, f.ws->err) : 0;
	if (r.secondary >= 0) a.flag |= 0x100;
	int tmp = infer_bw(qe - qb, (int)(re - rb), r.truesc, opt.a, opt.o_del, opt.e_del);
	int w2 = infer_bw(qe - qb, (int)(re - rb), r.truesc, opt.a, opt.o_ins, opt.e_ins);
	w2 = w2 > tmp ? w2 : tmp;
	if (w2 > opt.w) w2 = w2 < r.w ? w2 : r.w;
	int i = 0, score = 0, last_sc = -(1 << 30), NM = -1, nc = 0, ml = 0;
	if (f.mode == FIN_COLLECT) { // record the request; the caller only needs a well-formed record to go on deciding
		if (f.lt_n < GA_ITEM_MAX) { GaTask t; t.qoff = (u64)(query - f.seq_base); t.rb = rb; t.re = re; t.qb = qb; t.qe = qe; t.w2 = w2; t.truesc = r.truesc; f.lt[f.lt_n] = t; }
		++f.lt_n;
		NM = 0;
	} else if (f.mode == FIN_CONSUME) {
		const GaRes* g = f.res_k < f.res_n ? &f.res[f.res_k] : nullptr;
		++f.res_k;
		if (g == nullptr || g->valid != 1 || g->nc + 2 > f.ws->cigar_cap || g->ml > f.ws->md_cap - 16) { atomic_or_u32(f.ws->err, ERR_NEED_WARP); NM = 0; } // a warp finalizes the item
		else {
			score = g->score; NM = g->NM; nc = g->nc; ml = g->ml;
			for (int k = 0; k < nc; ++k) f.ws->cigar[k] = g->cigar[k];
			for (int k = 0; k < ml; ++k) f.ws->md[k] = g->md[k];
			f.ws->md[ml] = 0;
		}
	} else
	do {
		w2 = w2 < opt.w << 2 ? w2 : opt.w << 2;
		DBGP(300, w2);
		warp_gen_cigar2<W>(x, opt, *f.wm, *f.ws, f.cnt, w2, qe - qb, query + qb, rb, re, &score, true, &nc, &NM, &ml);
		DBGP(301, score);
		if (score == last_sc || w2 == opt.w << 2) break;
		last_sc = score;
		w2 <<= 1;
	} while (++i < 3 && score < r.truesc - opt.a);
	a.NM = NM;
	int is_rev;
	i64 pos = ref_depos(x, rb < x.l_pac ? rb : re - 1, is_rev);
	a.is_rev = (u32)is_rev;
	u32* cg = f.ws->cigar;
	if (nc > 0) { // squeeze out a leading or (else) a trailing deletion
		if ((cg[0] & 0xf) == 2) {
			pos += cg[0] >> 4;
			--nc;
			LANE0_ONLY { for (int k = 0; k < nc; ++k) cg[k] = cg[k + 1]; }
			WARP_SYNC();
		} else if ((cg[nc - 1] & 0xf) == 2) --nc;
	}
	if (qb != 0 || qe != l_query) {
		int clip5 = is_rev ? l_query - qe : qb, clip3 = is_rev ? qb : l_query - qe;
		if (nc + 2 > f.ws->cigar_cap) { ONE_LANE { atomic_or_u32(f.ws->err, ERR_CIGAR_OVERFLOW); } }
		else {
			if (clip5) {
				LANE0_ONLY { for (int k = nc; k > 0; --k) cg[k] = cg[k - 1]; cg[0] = (u32)clip5 << 4 | 3; }
				++nc; WARP_SYNC();
			}
			if (clip3) { ONE_LANE { cg[nc] = (u32)clip3 << 4 | 3; } ++nc; WARP_SYNC(); }
		}
	}
	a.n_cigar = nc; a.md_len = ml;
	a.rid = ref_pos2rid(x, pos);
	a.pos = pos - x.ctg_off[a.rid];
	a.score = r.score; a.sub = r.sub > r.csub ? r.sub : r.csub;
	return a;
}

DEV int cigar_ref_len(const u32* cg, int n) { int l = 0; for (int k = 0; k < n; ++k) { int op = (int)(cg[k] & 0xf); if (op == 0 || op == 2) l += (int)(cg[k] >> 4); } return l; }

DEV MateDesc mate_of(const Aln& a, const u32* cg)
{
	MateDesc m; m.rid = a.rid; m.pos = a.pos; m.is_rev = a.is_rev; m.ref_len = a.rid >= 0 ? cigar_ref_len(cg, a.n_cigar) : 0;
	return m;
}

// Copies the payload of a mapped alignment (ws.cigar / ws.md) into the var pool so that the warp scratch can be reused
// before the record is written (the paired path needs both ends' alignments before it can emit either).
// Returns the word offset, or 0xffffffff when the pool is exhausted (the error flag is raised).
template <int W>
DEVNI u32 warp_stash(FinalCtx& f, const Aln& p)
{
	LANE_DECL;
	const int n_cigar = p.n_cigar, md_len = n_cigar ? p.md_len : 0;
	const u32 n_words = (u32)n_cigar + (u32)((md_len + 3) >> 2);
	const u32 off = n_words ? var_alloc<W>(f, n_words) : 0;
	if (off == 0xffffffffu) return off;
	FOR_LANES {
		for (int k = lane; k < n_cigar; k += W) f.var[off + k] = f.ws->cigar[k];
		int mdw = (md_len + 3) >> 2;
		for (int k = lane; k < mdw; k += W) {
			u32 wd = 0;
			for (int b = 0; b < 4; ++b) { int ci = k * 4 + b; if (ci < md_len) wd |= (u32)(u8)f.ws->md[ci] << (8 * b); }
			f.var[off + n_cigar + k] = wd;
		}
	}
	WARP_SYNC();
	return off;
}

// Copies the payload of `a` (ws.cigar / ws.md) into the var pool -- unless it was stashed there already (stashed:
// its offset; stashed_ref_len: reference length of its CIGAR) --, applies the flag finalisation bwa performs on
// copies before the formatter hook (B.14) and writes the output record.
template <int W>
DEVNI void warp_emit(FinalCtx& f, const Aln& a_in, const MateDesc* mate, u32 xa_off, int xa_len, AlnRec* out, const u32* stashed = nullptr, int stashed_ref_len = 0)
{
	LANE_DECL;
	Aln p = a_in;
	AlnRec r;
	memset(&r, 0, sizeof(r));
	int ref_len = stashed ? stashed_ref_len : (p.rid >= 0 ? cigar_ref_len(f.ws->cigar, p.n_cigar) : 0);
	MateDesc m; bool has_m = mate != nullptr;
	if (has_m) m = *mate; else { m.rid = -1; m.pos = -1; m.is_rev = 0; m.ref_len = 0; }
	int n_cigar = p.n_cigar;
	p.flag |= has_m ? 0x1 : 0;
	p.flag |= p.rid < 0 ? 0x4 : 0;
	p.flag |= (has_m && m.rid < 0) ? 0x8 : 0;
	if (p.rid < 0 && has_m && m.rid >= 0) { p.rid = m.rid; p.pos = m.pos; p.is_rev = m.is_rev; n_cigar = 0; ref_len = 0; }
	if (has_m && m.rid < 0 && p.rid >= 0) { m.rid = p.rid; m.pos = p.pos; m.is_rev = p.is_rev; m.ref_len = 0; }
	p.flag |= p.is_rev ? 0x10 : 0;
	p.flag |= (has_m && m.is_rev) ? 0x20 : 0;
	r.flag = p.flag; r.mapq = p.mapq;
	if (!(p.flag & 0x4)) {
		r.rid = p.rid; r.pos = (i32)p.pos; r.NM = p.NM; r.score = p.score; r.sub = p.sub; r.n_cigar = n_cigar;
		r.md_len = n_cigar ? p.md_len : 0;
		u32 n_words = (u32)n_cigar + (u32)((r.md_len + 3) >> 2);
		u32 off = stashed ? *stashed : (n_words ? var_alloc<W>(f, n_words) : 0);
		if (off == 0xffffffffu) { off = 0; r.n_cigar = 0; r.md_len = 0; n_cigar = 0; }
		r.var_off = off;
		if (!stashed) FOR_LANES {
			for (int k = lane; k < n_cigar; k += W) f.var[off + k] = f.ws->cigar[k];
			int mdw = (r.md_len + 3) >> 2;
			for (int k = lane; k < mdw; k += W) {
				u32 wd = 0;
				for (int b = 0; b < 4; ++b) { int ci = k * 4 + b; if (ci < r.md_len) wd |= (u32)(u8)f.ws->md[ci] << (8 * b); }
				f.var[off + n_cigar + k] = wd;
			}
		}
		r.xa_off = xa_off; r.xa_len = xa_len;
	}
	if ((p.flag & 0x9) == 1) {
		r.m_rid = m.rid; r.m_pos = (i32)m.pos;
		if ((p.flag & 0x4) || p.rid != m.rid) r.tlen = 0;
		else {
			i64 p0 = p.pos, m0 = m.pos;
			if (p.is_rev) p0 += ref_len - 1;
			if (m.is_rev) m0 += m.ref_len - 1;
			r.tlen = (i32)(m0 - p0 + (p0 > m0 ? -1 : p0 < m0 ? 1 : 0));
		}
	}
	ONE_LANE { *out = r; }
	WARP_SYNC();
}

// ---- XA (B.14) -----------------------------------------------------------------------------------
DEV int get_pri_idx(double XA_drop_ratio, const Reg* a, int i)
{
	int k = a[i].secondary_all;
	if (k >= 0 && (double)a[i].score >= (double)a[k].score * XA_drop_ratio) return k;
	return -1;
}

// XA text of primary index `pri` (mem_gen_alt restricted to one primary); returns the byte length, payload copied to
// the var pool at *xa_off.  Uses ws.cigar / ws.md as scratch (call before the record's own warp_reg2aln).
template <int W>
DEVNI int warp_gen_xa(FinalCtx& f, int l_query, const u8* query, int n, const Reg* a, int pri, u32* xa_off)
{
	LANE_DECL;
	const MemOpt& opt = *f.opt;
	*xa_off = 0;
	int cnt = 0; bool has_alt = false;
	double ratio = (double)opt.XA_drop_ratio;
	for (int i = 0; i < n; ++i) if (get_pri_idx(ratio, a, i) == pri) { ++cnt; if (a[i].is_alt) has_alt = true; }
	if (cnt == 0 || cnt > opt.max_XA_hits_alt || (!has_alt && cnt > opt.max_XA_hits)) return 0;
	char* xa = f.ws->xa; const int cap = f.ws->xa_cap - 32;
	int xl = 0; bool ovf = false;
#define XA_PUTC(ch) do { const char c_ = (char)(ch); if (xl < cap) { ONE_LANE_NOSYNC { xa[xl] = c_; } ++xl; } else ovf = true; } while (0)
#define XA_PUTNUM(v) do { long v_ = (long)(v); char b_[24]; int n_ = 0; if (v_ < 0) { XA_PUTC('-'); v_ = -v_; } do { b_[n_++] = (char)('0' + v_ % 10); v_ /= 10; } while (v_); while (n_) XA_PUTC(b_[--n_]); } while (0)
	for (int i = 0; i < n; ++i) {
		if (get_pri_idx(ratio, a, i) != pri) continue;
		Aln t = warp_reg2aln<W>(f, l_query, query, &a[i]);
		const char* nm = f.ctg_names + f.ctg_name_off[t.rid];
		for (int k = 0; nm[k]; ++k) XA_PUTC(nm[k]);
		XA_PUTC(','); XA_PUTC("+-"[t.is_rev]); XA_PUTNUM(t.pos + 1); XA_PUTC(',');
		for (int k = 0; k < t.n_cigar; ++k) { u32 cg = f.ws->cigar[k]; XA_PUTNUM(cg >> 4); XA_PUTC("MIDSHN"[cg & 0xf]); }
		XA_PUTC(','); XA_PUTNUM(t.NM); XA_PUTC(';');
	}
	WARP_SYNC();
	if (ovf) { ONE_LANE { atomic_or_u32(f.ws->err, ERR_XA_OVERFLOW); } return 0; }
	u32 n_words = (u32)((xl + 3) >> 2);
	u32 off = var_alloc<W>(f, n_words);
	if (off == 0xffffffffu) return 0;
	FOR_LANES {
		for (int k = lane; k < (int)n_words; k += W) {
			u32 wd = 0;
			for (int b = 0; b < 4; ++b) { int ci = k * 4 + b; if (ci < xl) wd |= (u32)(u8)xa[ci] << (8 * b); }
			f.var[off + k] = wd;
		}
	}
	*xa_off = off;
	return xl;
}

// ---- SE emission (mem_reg2sam, B.14).  Returns the number of records written to out[]. ----------
template <int W>
DEVNI int warp_reg2sam(FinalCtx& f, int l_seq, const u8* seq, int n, const Reg* a, int extra_flag, const MateDesc* m, AlnRec* out, int out_cap)
{
	LANE_DECL;
	const MemOpt& opt = *f.opt;
	int l = 0; u32 mapq0 = 0;
	for (int k = 0; k < n; ++k) {
		const Reg p = a[k];
		if (p.score < opt.T) continue;
		if (p.secondary >= 0 && (p.is_alt || !(opt.flag & MEM_F_ALL))) continue;
		if (p.secondary >= 0 && p.secondary < 0x7fffffff && (float)p.score < (float)a[p.secondary].score * opt.drop_ratio) continue;
		if (l >= out_cap) { ONE_LANE { atomic_or_u32(f.ws->err, ERR_REG_OVERFLOW); } break; }
		u32 xa_off = 0; int xa_len = 0;
		DBGP(200, k);
		if (!(opt.flag & MEM_F_ALL)) xa_len = warp_gen_xa<W>(f, l_seq, seq, n, a, k, &xa_off);
		DBGP(201, k);
		Aln q = warp_reg2aln<W>(f, l_seq, seq, &a[k]);
		DBGP(202, k);
		q.flag |= extra_flag;
		if (p.secondary >= 0) q.sub = -1;
		if (l && p.secondary < 0) q.flag |= (opt.flag & MEM_F_NO_MULTI) ? 0x10000 : 0x800;
		if (l && !p.is_alt && q.mapq > mapq0) q.mapq = mapq0;
		if (l == 0) mapq0 = q.mapq;
		warp_emit<W>(f, q, m, xa_off, xa_len, &out[l]);
		DBGP(203, k);
		++l;
	}
	if (l == 0) {
		Aln t = warp_reg2aln<W>(f, l_seq, seq, nullptr);
		t.flag |= extra_flag;
		warp_emit<W>(f, t, m, 0, 0, &out[0]);
		l = 1;
	}
	return l;
}

// ---- pairing ---------------------------------------------------------------------------------------
DEV int infer_dir(i64 l_pac, i64 b1, i64 b2, i64* dist)
{
	int r1 = b1 >= l_pac, r2 = b2 >= l_pac;
	i64 p2 = r1 == r2 ? b2 : (l_pac << 1) - 1 - b2;
	*dist = p2 > b1 ? p2 - b1 : b1 - p2;
	return (r1 == r2 ? 0 : 1) ^ (p2 > b1 ? 0 : 3);
}

DEV int pe_cal_sub(const MemOpt& opt, int n, const Reg* a)
{
	int j;
	for (j = 1; j < n; ++j) {
		int b_max = a[j].qb > a[0].qb ? a[j].qb : a[0].qb;
		int e_min = a[j].qe < a[0].qe ? a[j].qe : a[0].qe;
		if (e_min > b_max) {
			int min_l = a[j].qe - a[j].qb < a[0].qe - a[0].qb ? a[j].qe - a[j].qb : a[0].qe - a[0].qb;
			if ((float)(e_min - b_max) >= (float)min_l * opt.mask_level) break;
		}
	}
	return j < n ? a[j].score : opt.min_seed_len * opt.a;
}

// candidate insert size of one pair for mem_pestat: returns dir<<48 | isize, or -1 (thread per pair)
DEV i64 pestat_candidate(const DevIndex& x, const MemOpt& opt, int n0, const Reg* r0, int n1, const Reg* r1)
{
	if (n0 == 0 || n1 == 0) return -1;
	if ((double)pe_cal_sub(opt, n0, r0) > 0.8 * (double)r0[0].score) return -1;
	if ((double)pe_cal_sub(opt, n1, r1) > 0.8 * (double)r1[0].score) return -1;
	if (r0[0].rid != r1[0].rid) return -1;
	i64 is;
	int dir = infer_dir(x.l_pac, r0[0].rb, r1[0].rb, &is);
	if (is && is <= opt.max_ins) return (i64)dir << 48 | is;
	return -1;
}

// mate rescue (B.8).  a: snapshot hit of one end; ms: mate read codes; ma[0..*n_ma): the mate's live region list.
template <int W>
DEVNI int warp_matesw(FinalCtx& f, const Reg& a, int l_ms, const u8* ms, Reg* ma, int* n_ma, int ma_cap)
{
	LANE_DECL;
	const DevIndex& x = *f.x; const MemOpt& opt = *f.opt; const PeStat* pes = f.pes;
	const i64 l_pac = x.l_pac;
	int skip[4], n = 0, nm = *n_ma;
	for (int r = 0; r < 4; ++r) skip[r] = pes[r].failed ? 1 : 0;
	for (int i = 0; i < nm; ++i) {
		i64 dist;
		int r = infer_dir(l_pac, a.rb, ma[i].rb, &dist);
		if (dist >= pes[r].low && dist <= pes[r].high) skip[r] = 1;
	}
	if (skip[0] + skip[1] + skip[2] + skip[3] == 4) return 0;
	WarpMem& wm = *f.wm;
	for (int r = 0; r < 4; ++r) {
		if (skip[r]) continue;
		int is_rev = (r >> 1) != (r & 1), is_larger = !(r >> 1);
		i64 rb, re;
		if (!is_rev) {
			rb = is_larger ? a.rb + pes[r].low : a.rb - pes[r].high;
			re = (is_larger ? a.rb + pes[r].high : a.rb - pes[r].low) + l_ms;
		} else {
			rb = (is_larger ? a.rb + pes[r].low : a.rb - pes[r].high) - l_ms;
			re = is_larger ? a.rb + pes[r].high : a.rb - pes[r].low;
		}
		if (rb < 0) rb = 0;
		if (re > l_pac << 1) re = l_pac << 1;
		int rid = -1;
		if (rb < re) rid = ref_clip_window(x, rb, (rb + re) >> 1, re);
		if (a.rid == rid && re - rb >= opt.min_seed_len) {
			int xtra = KSW_XSUBO | KSW_XSTART | (l_ms * opt.a < 250 ? KSW_XBYTE : 0) | (opt.min_seed_len * opt.a);
			bool is_u8 = (xtra & KSW_XBYTE) != 0; int lanes_simd = is_u8 ? 16 : 8;
			int tlen = (int)(re - rb);
			int slen = (l_ms + lanes_simd - 1) / lanes_simd, plen = slen * lanes_simd;
			if (plen > wm.cap) { ONE_LANE { atomic_or_u32(f.ws->err, ERR_QLEN_OVERFLOW); } return n; }
			if (tlen > f.ws->brow_cap) { ONE_LANE { atomic_or_u32(f.ws->err, ERR_BROW_OVERFLOW); } return n; }
			// forward pass: query = mate (reverse-complemented when is_rev), padded with code 5
			FOR_LANES {
				for (int j = lane; j < plen; j += W) {
					u8 c = 5;
					if (j < l_ms) { if (is_rev) { u8 b = ms[l_ms - 1 - j]; c = b < 4 ? (u8)(3 - b) : (u8)4; } else c = ms[j]; }
					WQ(wm, j) = c;
				}
			}
			WARP_SYNC();
			TAcc ta; ta.base = rb; ta.stride = 1; ta.piv = -1;
			SwFwd fw = warp_sw_striped<W>(x, opt, wm, lanes_simd, is_u8, l_ms, tlen, ta, xtra, f.ws->brow, f.cnt ? &f.cnt->cells_msw : nullptr);
			if (f.cnt) { ONE_LANE { atomic_add_u64(&f.cnt->n_msw, 1); } }
			int qb_ = -1, tb_ = -1;
			if (!(fw.score < (xtra & 0xffff))) { // XSUBO: the start is only searched for scores above the threshold
				int ql2 = fw.qe + 1;
				int slen2 = (ql2 + lanes_simd - 1) / lanes_simd, plen2 = slen2 * lanes_simd;
				FOR_LANES {
					for (int j = lane; j < plen2; j += W) {
						u8 c = 5;
						if (j < ql2) { int jj = fw.qe - j; if (is_rev) { u8 b = ms[l_ms - 1 - jj]; c = b < 4 ? (u8)(3 - b) : (u8)4; } else c = ms[jj]; }
						WQ(wm, j) = c;
					}
				}
				WARP_SYNC();
				TAcc tr; tr.base = rb; tr.stride = 1; tr.piv = fw.te; // target prefix [0,te] reversed, the rest untouched
				SwFwd rv = warp_sw_striped<W>(x, opt, wm, lanes_simd, is_u8, ql2, tlen, tr, KSW_XSTOP | fw.score, f.ws->brow, f.cnt ? &f.cnt->cells_msw : nullptr);
				if (fw.score == rv.score) { tb_ = fw.te - rv.te; qb_ = fw.qe - rv.qe; }
			}
			if (fw.score >= opt.min_seed_len && qb_ >= 0) {
				Reg b;
				memset(&b, 0, sizeof(Reg));
				b.rid = a.rid; b.is_alt = a.is_alt;
				b.qb = is_rev ? l_ms - (fw.qe + 1) : qb_;
				b.qe = is_rev ? l_ms - qb_ : fw.qe + 1;
				b.rb = is_rev ? (l_pac << 1) - (rb + fw.te + 1) : rb + tb_;
				b.re = is_rev ? (l_pac << 1) - (rb + tb_) : rb + fw.te + 1;
				b.score = fw.score; b.csub = fw.score2; b.secondary = -1;
				b.seedcov = (int)((b.re - b.rb < b.qe - b.qb ? b.re - b.rb : b.qe - b.qb) >> 1);
				if (nm >= ma_cap) { ONE_LANE { atomic_or_u32(f.ws->err, ERR_REG_OVERFLOW); } return n; }
				LANE0_ONLY {
					int i, tmp;
					for (i = 0; i < nm; ++i) if (ma[i].score < b.score) break;
					tmp = i;
					for (i = nm; i > tmp; --i) ma[i] = ma[i - 1];
					ma[tmp] = b;
				}
				++nm;
				WARP_SYNC();
			}
			++n;
		}
		if (n) nm = warp_sort_dedup_patch<W>(x, opt, wm, *f.ws, f.cnt, nullptr, nm, ma);
	}
	*n_ma = nm;
	return n;
}

struct Pair64 { u64 x, y; };
struct Pair64Lt { DEVM bool operator()(const Pair64& a, const Pair64& b) const { return a.x < b.x || (a.x == b.x && a.y < b.y); } };

// mem_pair (B.13).  v: scratch for n_pri[0]+n_pri[1] keys.  The candidate list u is never materialised: the best,
// the runner-up and the n_sub count are obtained by enumerating the candidates twice.
template <int W>
DEVNI int warp_mem_pair(FinalCtx& f, const Reg* a0, const Reg* a1, const int n_pri[2], int id, int* sub, int* n_sub, int z[2], Pair64* v)
{
	LANE_DECL;
	const DevIndex& x = *f.x; const MemOpt& opt = *f.opt; const PeStat* pes = f.pes;
	const i64 l_pac = x.l_pac;
	int nv = n_pri[0] + n_pri[1];
	LANE0_ONLY {
		int t = 0;
		for (int r = 0; r < 2; ++r) {
			const Reg* a = r ? a1 : a0;
			for (int i = 0; i < n_pri[r]; ++i) {
				Pair64 key;
				const Reg& e = a[i];
				key.x = (u64)(e.rb < l_pac ? e.rb : (l_pac << 1) - 1 - e.rb);
				key.x = (u64)e.rid << 32 | (key.x - (u64)x.ctg_off[e.rid]);
				key.y = (u64)e.score << 32 | (u64)(i << 2) | (u64)((e.rb >= l_pac) << 1) | (u64)r;
				v[t++] = key;
			}
		}
		ks_introsort<Pair64, Pair64Lt>((size_t)nv, v, Pair64Lt());
	}
	WARP_SYNC();
	Pair64 best, second; bool have_best = false, have_second = false;
	best.x = best.y = second.x = second.y = 0;
	u64 n_u = 0;
	int ret = 0;
	for (int pass = 0; pass < 2; ++pass) {
		int y[4] = { -1, -1, -1, -1 };
		int cnt_sub = 0;
		int tmp = opt.a + opt.b;
		tmp = tmp > opt.o_del + opt.e_del ? tmp : opt.o_del + opt.e_del;
		tmp = tmp > opt.o_ins + opt.e_ins ? tmp : opt.o_ins + opt.e_ins;
		for (int i = 0; i < nv; ++i) {
			for (int r = 0; r < 2; ++r) {
				int dir = r << 1 | (int)(v[i].y >> 1 & 1);
				if (pes[dir].failed) continue;
				int which = r << 1 | (int)((v[i].y & 1) ^ 1);
				if (y[which] < 0) continue;
				for (int k = y[which]; k >= 0; --k) {
					if ((int)(v[k].y & 3) != which) continue;
					i64 dist = (i64)v[i].x - (i64)v[k].x;
					if (dist > pes[dir].high) break;
					if (dist < pes[dir].low) continue;
					double term = f.pair_tab[f.pair_tab_off[dir] + (dist - pes[dir].low)];
					int q = (int)((double)((v[i].y >> 32) + (v[k].y >> 32)) + term + .499);
					if (q < 0) q = 0;
					Pair64 p;
					p.y = (u64)k << 32 | (u32)i;
					p.x = (u64)q << 32 | (hash64(p.y ^ (u64)(i64)(id << 8)) & 0xffffffffU);
					if (pass == 0) {
						++n_u;
						Pair64Lt lt;
						if (!have_best || lt(best, p)) { second = best; have_second = have_best; best = p; have_best = true; }
						else if (!have_second || lt(second, p)) { second = p; have_second = true; }
					} else {
						bool is_best = p.x == best.x && p.y == best.y;
						if (!is_best && *sub - (int)(p.x >> 32) <= tmp) ++cnt_sub;
					}
				}
			}
			y[v[i].y & 3] = i;
		}
		if (pass == 0) {
			if (n_u == 0) { *sub = 0; *n_sub = 0; return 0; }
			int bi = (int)(best.y >> 32), bk = (int)(best.y << 32 >> 32);
			z[v[bi].y & 1] = (int)(v[bi].y << 32 >> 34);
			z[v[bk].y & 1] = (int)(v[bk].y << 32 >> 34);
			ret = (int)(best.x >> 32);
			*sub = n_u > 1 ? (int)(second.x >> 32) : 0;
		} else *n_sub = cnt_sub;
	}
	return ret;
}

#define RAW_MAPQ(diff, a) ((int)(6.02 * (double)(diff) / (double)(a) + .499))

struct PairIO {
	int l_seq[2]; const u8* seq[2];
	Reg* a[2]; int n[2]; int cap[2];
	Reg* snap[2];           // snapshot scratch (capacity >= initial n of each end)
	u32* ztmp[2];           // u32 scratch per end (capacity cap)
	Pair64* v;              // cap[0]+cap[1]
	AlnRec* out[2]; int out_cap[2]; int n_out[2];
};

// mem_sam_pe (B.13)
template <int W>
DEVNI void warp_sam_pe(FinalCtx& f, u64 id, PairIO& io)
{
	LANE_DECL;
	const DevIndex& x = *f.x; const MemOpt& opt = *f.opt; const PeStat* pes = f.pes;
	int z[2] = { 0, 0 }, n_pri[2], extra_flag = 1;
	if (!(opt.flag & MEM_F_NO_RESCUE) && !f.rescued) {
		int nb[2];
		for (int i = 0; i < 2; ++i) {
			int m = 0;
			for (int j = 0; j < io.n[i]; ++j)
				if (io.a[i][j].score >= io.a[i][0].score - opt.pen_unpaired) { ONE_LANE { io.snap[i][m] = io.a[i][j]; } ++m; }
			nb[i] = m;
		}
		WARP_SYNC();
		for (int i = 0; i < 2; ++i)
			for (int j = 0; j < nb[i] && j < opt.max_matesw; ++j) {
				const Reg s = io.snap[i][j];
				warp_matesw<W>(f, s, io.l_seq[!i], io.seq[!i], io.a[!i], &io.n[!i], io.cap[!i]);
			}
	}
	n_pri[0] = warp_mark_primary_se<W>(opt, io.n[0], io.a[0], (i64)(id << 1 | 0), io.ztmp[0]);
	n_pri[1] = warp_mark_primary_se<W>(opt, io.n[1], io.a[1], (i64)(id << 1 | 1), io.ztmp[1]);
	if (opt.flag & MEM_F_PRIMARY5) { warp_reorder_primary5<W>(opt.T, io.n[0], io.a[0]); warp_reorder_primary5<W>(opt.T, io.n[1], io.a[1]); }
	bool paired_path = false;
	int o = 0, subo = 0, n_sub = 0;
	if (!(opt.flag & MEM_F_NOPAIRING) && n_pri[0] && n_pri[1]) {
		o = warp_mem_pair<W>(f, io.a[0], io.a[1], n_pri, (int)id, &subo, &n_sub, z, io.v);
		if (o > 0) {
			bool is_multi = false;
			for (int i = 0; i < 2; ++i) {
				int j;
				for (j = 1; j < n_pri[i]; ++j) if (io.a[i][j].secondary < 0 && io.a[i][j].score >= opt.T) break;
				if (j < n_pri[i]) is_multi = true;
			}
			paired_path = !is_multi;
		}
	}
	if (paired_path) {
		int q_pe, score_un, q_se[2];
		Reg* a0 = io.a[0]; Reg* a1 = io.a[1];
		score_un = a0[0].score + a1[0].score - opt.pen_unpaired;
		subo = subo > score_un ? subo : score_un;
		q_pe = RAW_MAPQ(o - subo, opt.a);
		if (n_sub > 0) q_pe -= (int)(4.343 * tab_log(x, n_sub + 1, f.ws->err) + .499);
		if (q_pe < 0) q_pe = 0;
		if (q_pe > 60) q_pe = 60;
		q_pe = (int)((double)q_pe * (1. - .5 * (double)(a0[0].frac_rep + a1[0].frac_rep)) + .499);
		if (o > score_un) {
			for (int i = 0; i < 2; ++i) {
				Reg* c = &io.a[i][z[i]];
				if (c->secondary >= 0) { int ps = io.a[i][c->secondary].score; WARP_SYNC(); ONE_LANE { c->sub = ps; c->secondary = -2; } WARP_SYNC(); }
				q_se[i] = approx_mapq_se(x, opt, *c, f.ws->err);
			}
			q_se[0] = q_se[0] > q_pe ? q_se[0] : (q_pe < q_se[0] + 40 ? q_pe : q_se[0] + 40);
			q_se[1] = q_se[1] > q_pe ? q_se[1] : (q_pe < q_se[1] + 40 ? q_pe : q_se[1] + 40);
			extra_flag |= 2;
			for (int i = 0; i < 2; ++i) {
				const Reg& c = io.a[i][z[i]];
				int cap = RAW_MAPQ(c.score - c.csub, opt.a);
				q_se[i] = q_se[i] < cap ? q_se[i] : cap;
			}
		} else {
			z[0] = z[1] = 0;
			q_se[0] = approx_mapq_se(x, opt, a0[0], f.ws->err);
			q_se[1] = approx_mapq_se(x, opt, a1[0], f.ws->err);
		}
		for (int i = 0; i < 2; ++i) {
			int k = io.a[i][z[i]].secondary_all;
			if (k >= 0 && k < n_pri[i]) {
				WARP_SYNC();
				LANE0_ONLY {
					for (int j = 0; j < io.n[i]; ++j) if (io.a[i][j].secondary_all == k || j == k) io.a[i][j].secondary_all = z[i];
					io.a[i][z[i]].secondary_all = -1;
				}
				WARP_SYNC();
			}
		}
		// records: XA first (it reuses the warp scratch), then the chosen hit of each end; the mate descriptor of end i is
		// the other end's record before flag finalisation
		u32 xa_off[2] = { 0, 0 }; int xa_len[2] = { 0, 0 };
		MateDesc md[2]; Aln hs[2]; u32 st_off[2];
		for (int i = 0; i < 2; ++i) { // mate descriptors need pos/is_rev/ref_len of both ends first; each alignment is
			// computed once and its CIGAR / MD parked in the var pool (the scratch is reused by the other end and by XA)
			hs[i] = warp_reg2aln<W>(f, io.l_seq[i], io.seq[i], &io.a[i][z[i]]);
			md[i] = mate_of(hs[i], f.ws->cigar);
			st_off[i] = warp_stash<W>(f, hs[i]);
		}
		for (int i = 0; i < 2; ++i) {
			if (!(opt.flag & MEM_F_ALL)) xa_len[i] = warp_gen_xa<W>(f, io.l_seq[i], io.seq[i], io.n[i], io.a[i], z[i], &xa_off[i]);
			Aln h = hs[i];
			h.mapq = (u32)q_se[i];
			h.flag |= 0x40 << i | extra_flag;
			warp_emit<W>(f, h, &md[!i], xa_off[i], xa_len[i], &io.out[i][0], &st_off[i], md[i].ref_len);
			io.n_out[i] = 1;
			if (n_pri[i] < io.n[i]) { // the read has ALT hits: its best ALT hit goes out as a supplementary record
				const Reg& pa = io.a[i][n_pri[i]];
				if (!(pa.score < opt.T || pa.secondary >= 0 || !pa.is_alt) && io.out_cap[i] >= 2) {
					u32 xo = 0; int xl = 0;
					if (!(opt.flag & MEM_F_ALL)) xl = warp_gen_xa<W>(f, io.l_seq[i], io.seq[i], io.n[i], io.a[i], n_pri[i], &xo);
					Aln g = warp_reg2aln<W>(f, io.l_seq[i], io.seq[i], &pa);
					g.flag |= 0x800 | 0x40 << i | extra_flag;
					warp_emit<W>(f, g, &md[!i], xo, xl, &io.out[i][1]);
					io.n_out[i] = 2;
				}
			}
		}
		return;
	}
	// ---- no_pairing ----
	MateDesc h[2];
	for (int i = 0; i < 2; ++i) {
		int which = -1;
		if (io.n[i]) { if (io.a[i][0].score >= opt.T) which = 0; else if (n_pri[i] < io.n[i] && io.a[i][n_pri[i]].score >= opt.T) which = n_pri[i]; }
		Aln t = warp_reg2aln<W>(f, io.l_seq[i], io.seq[i], which >= 0 ? &io.a[i][which] : nullptr);
		h[i] = mate_of(t, f.ws->cigar);
	}
	if (!(opt.flag & MEM_F_NOPAIRING) && h[0].rid == h[1].rid && h[0].rid >= 0) {
		i64 dist;
		int d = infer_dir(x.l_pac, io.a[0][0].rb, io.a[1][0].rb, &dist);
		if (!pes[d].failed && dist >= pes[d].low && dist <= pes[d].high) extra_flag |= 2;
	}
	io.n_out[0] = warp_reg2sam<W>(f, io.l_seq[0], io.seq[0], io.n[0], io.a[0], 0x41 | extra_flag, &h[1], io.out[0], io.out_cap[0]);
	io.n_out[1] = warp_reg2sam<W>(f, io.l_seq[1], io.seq[1], io.n[1], io.a[1], 0x81 | extra_flag, &h[0], io.out[1], io.out_cap[1]);
}
