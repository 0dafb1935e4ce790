// chain.cuh -- seed chaining, chain weights and the chain filter, one thread per read.
// Replaces mem_chain / test_and_merge / mem_chain_weight / mem_chain_flt of un-vendored lh3/bwa
//  behind jnibwa_createAlignments (/root/reference/src/main/c/jnibwa.c:214).
//
// Layout: every pool below has one slot per suffix-array lookup of the batch; read r owns the slot range
// [seed_off[r], seed_off[r+1]).  Chains own their seeds as linked lists through Seed::next, so no per-chain
// allocation is needed; the kept chains and their seeds are compacted in chain order for the extension stage.
#pragma once
#include "util.cuh"

struct ChainW { // working chain
	i64 pos;
	i32 head, tail, n, rid;
	u32 w; i32 kept, first;
};

struct ChainStageIO {
	const Intv* intv; int n_intv;
	const i64* sd_rbeg;       // SA values, in (interval, occurrence) order
	Seed* S; ChainW* C; u32* ord; u32* tmp; // scratch, one slot per lookup
	u64* kq;                  // scratch, one slot per lookup: (ALT << 63 | weight << 32 | query end << 16 | query begin) of the chains kept so far
	ChainK* ck; Seed* ks;     // outputs (same slot range)
};

DEV bool chain_test_and_merge(const MemOpt& opt, i64 l_pac, ChainW& c, Seed* S, int sidx, int seed_rid)
{
	const Seed& p = S[sidx];
	const Seed& last = S[c.tail];
	const Seed& first = S[c.head];
	i64 qend = last.qbeg + last.len, rend = last.rbeg + last.len;
	if (seed_rid != c.rid) return false;
	if (p.qbeg >= first.qbeg && p.qbeg + p.len <= qend && p.rbeg >= first.rbeg && p.rbeg + p.len <= rend) return true; // contained
	if ((last.rbeg < l_pac || first.rbeg < l_pac) && p.rbeg >= l_pac) return false;
	i64 xq = p.qbeg - last.qbeg, y = p.rbeg - last.rbeg;
	if (y >= 0 && xq - y <= opt.w && y - xq <= opt.w && xq - last.len < opt.max_chain_gap && y - last.len < opt.max_chain_gap) {
		S[c.tail].next = sidx; c.tail = sidx; ++c.n;
		return true;
	}
	return false;
}

DEV int chain_weight(const ChainW& c, const Seed* S)
{
	i64 end = 0; int w = 0, tmp;
	for (int j = c.head; j >= 0; j = S[j].next) {
		const Seed& s = S[j];
		if (s.qbeg >= end) w += s.len;
		else if (s.qbeg + s.len > end) w += (int)(s.qbeg + s.len - end);
		end = end > s.qbeg + s.len ? end : (i64)(s.qbeg + s.len);
	}
	tmp = w; w = 0; end = 0;
	for (int j = c.head; j >= 0; j = S[j].next) {
		const Seed& s = S[j];
		if (s.rbeg >= end) w += s.len;
		else if (s.rbeg + s.len > end) w += (int)(s.rbeg + s.len - end);
		end = end > s.rbeg + s.len ? end : s.rbeg + s.len;
	}
	w = w < tmp ? w : tmp;
	return w < (1 << 30) ? w : (1 << 30) - 1;
}

struct ChainLtW { const ChainW* C; DEVM bool operator()(u32 a, u32 b) const { return C[a].w > C[b].w; } };
struct ChainKeyLtW { DEVM bool operator()(u64 a, u64 b) const { return (u32)(a >> 32) > (u32)(b >> 32); } }; // keys (weight << 32 | chain)

// returns the number of kept chains; *n_kept_seeds gets the number of seeds they hold
DEV int chain_read(const DevIndex& x, const MemOpt& opt, int len, ChainStageIO& io, int* n_kept_seeds)
{
	*n_kept_seeds = 0;
	if (len < opt.min_seed_len) return 0;
	Seed* S = io.S; ChainW* C = io.C; u32* ord = io.ord;
	int nch = 0, si = 0, rid_hint = -1;
	// l_rep: query bases covered by intervals that are too repetitive to be enumerated fully
	int b = 0, e = 0, l_rep = 0;
	for (int i = 0; i < io.n_intv; ++i) {
		const Intv& p = io.intv[i];
		if (p.size <= (u64)opt.max_occ) continue;
		int sb = (int)p.beg, se = (int)p.end;
		if (sb > e) { l_rep += e - b; b = sb; e = se; }
		else e = e > se ? e : se;
	}
	l_rep += e - b;
	for (int i = 0; i < io.n_intv; ++i) {
		const Intv& p = io.intv[i];
		int slen = (int)(p.end - p.beg);
		i64 step = p.size > (u64)opt.max_occ ? (i64)(p.size / (u64)opt.max_occ) : 1;
		int count = 0;
		for (i64 k = 0; (u64)k < p.size && count < opt.max_occ; k += step, ++count) {
			int sidx = si++;
			Seed s; s.rbeg = io.sd_rbeg[sidx]; s.qbeg = (i32)p.beg; s.len = s.score = slen; s.next = -1;
			S[sidx] = s;
			int rid = ref_intv2rid(x, s.rbeg, s.rbeg + s.len, rid_hint);
			if (rid < 0) continue; // spans two contigs or the forward/reverse junction
			rid_hint = rid;
			bool to_add = true; int at = 0;
			if (nch) {
				int lo = 0, hi = nch;
				while (lo < hi) { int mid = (lo + hi) >> 1; if (C[ord[mid]].pos < s.rbeg) lo = mid + 1; else hi = mid; }
				int lower = -1;
				if (lo < nch && C[ord[lo]].pos == s.rbeg) { lower = (int)ord[lo]; at = lo + 1; }
				else if (lo > 0) { lower = (int)ord[lo - 1]; at = lo; }
				else at = 0;
				if (lower >= 0 && chain_test_and_merge(opt, x.l_pac, C[lower], S, sidx, rid)) to_add = false;
			}
			if (to_add) {
				ChainW c; c.pos = s.rbeg; c.head = c.tail = sidx; c.n = 1; c.rid = rid; c.w = 0; c.kept = 0; c.first = -1;
				C[nch] = c;
				for (int j = nch; j > at; --j) ord[j] = ord[j - 1];
				ord[at] = (u32)nch;
				++nch;
			}
		}
	}
	if (nch == 0) return 0;
	float frac_rep = (float)l_rep / len;
	// ---- filter (B.5): weights, weight-descending introsort (tie order observable), overlap rule
	int n = 0;
	for (int i = 0; i < nch; ++i) {
		ChainW& c = C[ord[i]];
		c.first = -1; c.kept = 0;
		c.w = (u32)chain_weight(c, S);
		if ((int)c.w >= opt.min_chain_weight) ord[n++] = ord[i];
	}
	if (n == 0) return 0;
	ChainLtW lt; lt.C = C;
	ks_introsort<u32, ChainLtW>((size_t)n, ord, lt);
	// The scan of a chain against all chains kept so far is quadratic; what it needs of a kept chain -- query span, weight,
	// ALT flag -- sits in one word of a compact list (kq), so a comparison is one sequential load instead of three dependent ones.
	u32* kl = io.tmp; u64* kq = io.kq; int nk = 0; // positions (in ord) of chains kept so far
#define CHAIN_PACK(c_) ((u64)(x.ctg_alt[(c_).rid] ? 1 : 0) << 63 | (u64)(c_).w << 32 | (u64)((u32)(S[(c_).tail].qbeg + S[(c_).tail].len) & 0xffffu) << 16 | (u64)((u32)S[(c_).head].qbeg & 0xffffu))
	C[ord[0]].kept = 3; kl[0] = 0; kq[0] = CHAIN_PACK(C[ord[0]]); nk = 1;
	for (int i = 1; i < n; ++i) {
		ChainW& ci = C[ord[i]];
		const u64 vi = CHAIN_PACK(ci);
		const int ib = (int)(vi & 0xffffu), ie = (int)(vi >> 16 & 0xffffu), wi = (int)(vi >> 32 & 0x7fffffffu); const bool alt_i = (vi >> 63) != 0;
		bool large_ovlp = false; int k;
		for (k = 0; k < nk; ++k) {
			const u64 vj = kq[k];
			const int jb = (int)(vj & 0xffffu), je = (int)(vj >> 16 & 0xffffu), wj = (int)(vj >> 32 & 0x7fffffffu);
			int b_max = jb > ib ? jb : ib, e_min = je < ie ? je : ie;
			if (e_min > b_max && (!(vj >> 63) || alt_i)) { // (a chain on an ALT contig never masks one on the primary assembly)
				int li = ie - ib, lj = je - jb, min_l = li < lj ? li : lj;
				if ((float)(e_min - b_max) >= (float)min_l * opt.mask_level && min_l < opt.max_chain_gap) {
					large_ovlp = true;
					ChainW& cj = C[ord[kl[k]]];
					if (cj.first < 0) cj.first = i;
					if ((float)wi < (float)wj * opt.drop_ratio && wj - wi >= opt.min_seed_len << 1) break;
				}
			}
		}
		if (k == nk) { kl[nk] = (u32)i; kq[nk] = vi; ++nk; ci.kept = large_ovlp ? 2 : 3; }
	}
#undef CHAIN_PACK
	for (int i = 0; i < nk; ++i) { ChainW& c = C[ord[kl[i]]]; if (c.first >= 0) C[ord[c.first]].kept = 1; }
	{
		int i, k;
		for (i = k = 0; i < n; ++i) {
			int kept = C[ord[i]].kept;
			if (kept == 0 || kept == 3) continue;
			if (++k >= opt.max_chain_extend) break;
		}
		for (; i < n; ++i) if (C[ord[i]].kept < 3) C[ord[i]].kept = 0;
	}
	// ---- compact kept chains, seeds in chain order
	int n_out = 0, n_seed_out = 0;
	for (int i = 0; i < n; ++i) {
		ChainW& c = C[ord[i]];
		if (c.kept == 0) continue;
		ChainK o; o.seed_start = (u32)n_seed_out; o.n_seeds = (u32)c.n; o.rid = c.rid; o.frac_rep = frac_rep; o.w = c.w; o.kept = c.kept; o.pos = c.pos;
		for (int j = c.head; j >= 0; j = S[j].next) { Seed s = S[j]; s.next = -1; io.ks[n_seed_out++] = s; }
		io.ck[n_out++] = o;
	}
	*n_kept_seeds = n_seed_out;
	return n_out;
}

// ---- the same procedure for reads with very many seeds, one warp per read -------------------------------------------
// A read from a repeat family brings ~10^3 seeds that form ~10^3 single-seed chains, all overlapping on the query.
// Two steps of the scalar formulation are then quadratic for a single thread: keeping the chain starts sorted
// (insertion = a shift of the order array) and the filter's scan of every chain against all chains kept so far.
// Here the order of evaluation -- and therefore every result -- is unchanged, but the shift and the scan are spread
// over the 32 lanes: a scan step evaluates 32 kept chains at once, the first lane (in list order) that demands a
// break wins, and only chains at or before it receive the side effects the scalar loop would have applied.
struct ChainWarpScratch { i64* pos_sorted; u64* qbw; }; // per seed slot: start of the chain at each sorted position; (w<<32 | qe<<16 | qb) per chain in weight order

template <int W>
DEVNI int chain_read_warp(const DevIndex& x, const MemOpt& opt, int len, ChainStageIO& io, ChainWarpScratch& cs, int* n_kept_seeds)
{
	LANE_DECL;
	*n_kept_seeds = 0;
	if (len < opt.min_seed_len) return 0;
	Seed* S = io.S; ChainW* C = io.C; u32* ord = io.ord; i64* ps = cs.pos_sorted;
	int nch = 0, si = 0;
	int b = 0, e = 0, l_rep = 0;
	for (int i = 0; i < io.n_intv; ++i) {
		const Intv& p = io.intv[i];
		if (p.size <= (u64)opt.max_occ) continue;
		int sb = (int)p.beg, se = (int)p.end;
		if (sb > e) { l_rep += e - b; b = sb; e = se; }
		else e = e > se ? e : se;
	}
	l_rep += e - b;
	for (int i = 0; i < io.n_intv; ++i) {
		const Intv p = io.intv[i];
		const int slen = (int)(p.end - p.beg);
		const i64 step = p.size > (u64)opt.max_occ ? (i64)(p.size / (u64)opt.max_occ) : 1;
		int m = 0; // seeds of this interval (the scalar loop: k = 0, step, ... while k < size and fewer than max_occ taken)
		{ const u64 by_size = (p.size + (u64)step - 1) / (u64)step; m = (int)(by_size < (u64)opt.max_occ ? by_size : (u64)opt.max_occ); }
		for (int c0 = 0; c0 < m; c0 += W) {
			// a lane per seed: its record and its contig (nothing here depends on the chains built so far) ...
			WV(i64, rbv); WV(int, ridv);
			FOR_LANES {
				const int j = c0 + lane;
				L_(rbv) = 0; L_(ridv) = -1;
				if (j < m) {
					Seed s; s.rbeg = io.sd_rbeg[si + j]; s.qbeg = (i32)p.beg; s.len = s.score = slen; s.next = -1;
					S[si + j] = s;
					L_(rbv) = s.rbeg; L_(ridv) = ref_intv2rid(x, s.rbeg, s.rbeg + s.len);
				}
			}
			WARP_SYNC();
			// ... then the seeds one after the other, in order, as the scalar loop inserts them
			const int mt = m - c0 < W ? m - c0 : W;
			for (int t = 0; t < mt; ++t) {
				const int sidx = si + c0 + t;
				const i64 rbeg = wv_get(rbv, t); const int rid = wv_get(ridv, t);
				if (rid < 0) continue;
				bool to_add = true; int at = 0;
				if (nch) {
					// lower bound of rbeg among the chain starts: W probes per round instead of one
					int lo = 0, hi = nch;
					while (hi - lo > W) {
						const int stride = (hi - lo + W - 1) / W;
						WV(bool, below);
						FOR_LANES { const int pos = lo + (lane + 1) * stride - 1; L_(below) = pos < hi ? ps[pos] < rbeg : false; }
						const int cnt = __popc(wv_ballot(below)); // (sorted: the lanes below rbeg are a prefix)
						const int nlo = lo + cnt * stride, nhi = lo + (cnt + 1) * stride - 1;
						lo = nlo < hi ? nlo : hi; hi = nhi < hi ? nhi : hi;
					}
					{
						WV(bool, below);
						FOR_LANES { const int pos = lo + lane; L_(below) = pos < hi ? ps[pos] < rbeg : false; }
						lo += __popc(wv_ballot(below));
					}
					int lower = -1;
					if (lo < nch && ps[lo] == rbeg) { lower = (int)ord[lo]; at = lo + 1; }
					else if (lo > 0) { lower = (int)ord[lo - 1]; at = lo; }
					else at = 0;
					if (lower >= 0) {
						int merged = 0;
						LANE0_ONLY { merged = chain_test_and_merge(opt, x.l_pac, C[lower], S, sidx, rid) ? 1 : 0; }
						WARP_SYNC();
						{ WV(int, mv); FOR_LANES { L_(mv) = merged; } merged = wv_get(mv, 0); }
						if (merged) to_add = false;
					}
				}
				if (to_add) {
					ONE_LANE { ChainW c; c.pos = rbeg; c.head = c.tail = sidx; c.n = 1; c.rid = rid; c.w = 0; c.kept = 0; c.first = -1; C[nch] = c; }
					// shift [at, nch) up by one, highest chunk first
					for (int hi2 = nch; hi2 > at; hi2 -= W) {
						WV(u32, to); WV(i64, tp);
						FOR_LANES { int j = hi2 - 1 - lane; if (j >= at) { L_(to) = ord[j]; L_(tp) = ps[j]; } }
						WARP_SYNC();
						FOR_LANES { int j = hi2 - 1 - lane; if (j >= at) { ord[j + 1] = L_(to); ps[j + 1] = L_(tp); } }
						WARP_SYNC();
					}
					ONE_LANE { ord[at] = (u32)nch; ps[at] = rbeg; }
					WARP_SYNC();
					++nch;
				}
			}
		}
		si += m;
	}
	if (nch == 0) return 0;
	float frac_rep = (float)l_rep / len;
	// ---- weights (a lane per chain), then the chains that pass min_chain_weight, in start order
	FOR_LANES { for (int i = lane; i < nch; i += W) { ChainW& c = C[i]; c.first = -1; c.kept = 0; c.w = (u32)chain_weight(c, S); } }
	WARP_SYNC();
	int n = 0;
	LANE0_ONLY { for (int i = 0; i < nch; ++i) if ((int)C[ord[i]].w >= opt.min_chain_weight) ord[n++] = ord[i]; }
	WARP_SYNC();
	{ WV(int, nv); FOR_LANES { L_(nv) = n; } n = wv_get(nv, 0); }
	if (n == 0) return 0;
	// The weight-descending introsort (tie order observable: one lane, the algorithm itself) over (weight, chain) pairs in
	// one word each: the same comparisons in the same order as on the chain indices, without two dependent loads per compare.
	u64* qbw = cs.qbw;
	FOR_LANES { for (int i = lane; i < n; i += W) qbw[i] = (u64)C[ord[i]].w << 32 | ord[i]; }
	LANE0_ONLY { ChainKeyLtW lt; ks_introsort<u64, ChainKeyLtW>((size_t)n, qbw, lt); }
	WARP_SYNC();
	FOR_LANES {
		for (int i = lane; i < n; i += W) {
			ord[i] = (u32)qbw[i];
			const ChainW& c = C[ord[i]];
			const u32 qb = (u32)S[c.head].qbeg, qe = (u32)(S[c.tail].qbeg + S[c.tail].len);
			qbw[i] = (u64)(x.ctg_alt[c.rid] ? 1 : 0) << 63 | (u64)c.w << 32 | (u64)(qe & 0xffffu) << 16 | (u64)(qb & 0xffffu);
		}
	}
	WARP_SYNC();
	u32* kl = io.tmp; int nk = 0;
	ONE_LANE { C[ord[0]].kept = 3; kl[0] = 0; }
	WARP_SYNC();
	nk = 1;
	for (int i = 1; i < n; ++i) {
		const u64 vi = qbw[i];
		const int ib = (int)(vi & 0xffffu), ie = (int)(vi >> 16 & 0xffffu), wi = (int)(vi >> 32 & 0x7fffffffu); const bool alt_i = (vi >> 63) != 0;
		bool large_ovlp = false, broke = false;
		for (int k0 = 0; k0 < nk && !broke; k0 += W) {
			WV(bool, ov); WV(bool, dr);
			FOR_LANES {
				const int k = k0 + lane;
				bool o = false, d = false;
				if (k < nk) {
					const u64 vj = qbw[kl[k]];
					const int jb = (int)(vj & 0xffffu), je = (int)(vj >> 16 & 0xffffu), wj = (int)(vj >> 32 & 0x7fffffffu);
					const int b_max = jb > ib ? jb : ib, e_min = je < ie ? je : ie;
					if (e_min > b_max && (!(vj >> 63) || alt_i)) {
						const int li = ie - ib, lj = je - jb, min_l = li < lj ? li : lj;
						if ((float)(e_min - b_max) >= (float)min_l * opt.mask_level && min_l < opt.max_chain_gap) {
							o = true;
							d = (float)wi < (float)wj * opt.drop_ratio && wj - wi >= opt.min_seed_len << 1;
						}
					}
				}
				L_(ov) = o; L_(dr) = d;
			}
			const u32 om = wv_ballot(ov), dm = wv_ballot(dr);
			u32 eff = om; // lanes whose side effects the scalar loop would have applied
			if (dm) { const int fb = __ffs(dm) - 1; eff = om & (fb == 31 ? 0xffffffffu : ((2u << fb) - 1u)); broke = true; }
			if (eff) large_ovlp = true;
			FOR_LANES { if ((eff >> lane) & 1u) { ChainW& cj = C[ord[kl[k0 + lane]]]; if (cj.first < 0) cj.first = i; } }
			WARP_SYNC();
		}
		if (!broke) { ONE_LANE { kl[nk] = (u32)i; C[ord[i]].kept = large_ovlp ? 2 : 3; } WARP_SYNC(); ++nk; }
	}
	int n_out = 0, n_seed_out = 0;
	LANE0_ONLY {
		for (int i = 0; i < nk; ++i) { ChainW& c = C[ord[kl[i]]]; if (c.first >= 0) C[ord[c.first]].kept = 1; }
		int i, k;
		for (i = k = 0; i < n; ++i) {
			int kept = C[ord[i]].kept;
			if (kept == 0 || kept == 3) continue;
			if (++k >= opt.max_chain_extend) break;
		}
		for (; i < n; ++i) if (C[ord[i]].kept < 3) C[ord[i]].kept = 0;
		for (i = 0; i < n; ++i) {
			ChainW& c = C[ord[i]];
			if (c.kept == 0) continue;
			ChainK o; o.seed_start = (u32)n_seed_out; o.n_seeds = (u32)c.n; o.rid = c.rid; o.frac_rep = frac_rep; o.w = c.w; o.kept = c.kept; o.pos = c.pos;
			for (int j = c.head; j >= 0; j = S[j].next) { Seed sd = S[j]; sd.next = -1; io.ks[n_seed_out++] = sd; }
			io.ck[n_out++] = o;
		}
	}
	WARP_SYNC();
	{ WV(int, a1); WV(int, a2); FOR_LANES { L_(a1) = n_out; L_(a2) = n_seed_out; } n_out = wv_get(a1, 0); n_seed_out = wv_get(a2, 0); }
	*n_kept_seeds = n_seed_out;
	return n_out;
}
