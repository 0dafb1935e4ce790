// rescue.cuh -- mate rescue (mem_matesw of un-vendored lh3/bwa bwamem_pair.c) as a
// task pipeline instead of a per-pair loop:
//
//   snap   thread per read : b[i] = hits of the read within pen_unpaired of its best hit (taken before any rescue)
//   round k = 0 .. min(max nb, max_matesw)-1, for every read r with k < nb[r] (its hit k rescues into the mate's list):
//     gen    thread per read : the skip[] test against the mate's CURRENT list, windows of the un-skipped
//                              orientations -> up to 4 independent local-alignment tasks
//     sw     warp per task   : ksw_align2 (forward pass, then the reverse pass for the start) -- register-resident
//     apply  thread per read : results inserted in orientation order, each followed by the region de-duplication,
//                              exactly the sequence the per-pair loop performs
//   The two rescue chains of a pair (hits of end 0 into the list of end 1 and vice versa) never read each other's
//   output, and the <= 4 alignments of one mem_matesw call are fixed before the first of them runs, so this order
//   of evaluation yields the same lists as the sequential loop.
//
// Speculative form (the default): the alignment of (hit k, orientation d) depends only on the hit, the insert-size
// statistics and the mate's bases -- not on the mate's list -- so ONE launch aligns the tasks of ALL hits k, chosen by
// the skip test against the mate's ORIGINAL list (gen_all), and one thread per read then replays the sequential
// decisions against the live list (apply_all), consuming the precomputed results.  Rescued hits only ever add
// suppressors, so the speculated set is a superset of the needed one -- except when region de-duplication removed the
// hit that had suppressed an orientation; replay detects that (a needed orientation with no speculated result) and
// hands the read to the round-by-round kernels above, which recompute its chain from the restored list.
#pragma once
#include "finalize.cuh"

struct SwTask {
	i64 rb, re;          // window in doubled-reference coordinates (already clipped to the contig)
	u32 read;            // read whose sequence is aligned (the mate of the rescuing hit)
	i32 is_rev;
	i32 score, te, qe, score2, te2, tb, qb; // results
	i32 pad_;
};
#define SW_SKIP 0
#define SW_NORUN 1   // orientation not skipped but no alignment was run (window on another contig / too short)
#define SW_RUN 2

struct QAcc { const u8* ms; int l_ms; int is_rev; int piv; }; // query accessor: forward code j, or reversed prefix when piv >= 0
DEV u8 qacc_code(const QAcc& a, int j)
{
	int jj = a.piv >= 0 ? a.piv - j : j;
	if (a.is_rev) { u8 b = a.ms[a.l_ms - 1 - jj]; return b < 4 ? (u8)(3 - b) : (u8)4; }
	return a.ms[jj];
}

#define SWF_CMAX 8
struct SwLane { int Hp[SWF_CMAX], E[SWF_CMAX], Hm[SWF_CMAX], H[SWF_CMAX], fin[SWF_CMAX]; u8 qc[SWF_CMAX]; int qs, qn, sub; };

// can the register-resident kernel take this query length?  (every lane must own 1..SWF_CMAX columns of its segment)
DEV bool swf_fits(int qlen, int lanes_simd)
{
	if (qlen <= 0) return false;
	int slen = (qlen + lanes_simd - 1) / lanes_simd, k = 32 / lanes_simd, C = (slen + k - 1) / k;
	return C <= SWF_CMAX && (k - 1) * C < slen;
}

// Local alignment with the observable semantics of the striped kernels (oracle: sw_striped in ksw.c): the same values
// as warp_sw_striped, computed with the whole DP row in registers.  Lane l owns up to C consecutive columns of
// segment l / k (k = 32 / lanes_simd lanes per segment); per row the diagonal enters by one shuffle, the
// segment-local and the full horizontal-gap states come from two max-plus scans over the lanes.
template <int W>
DEVNI SwFwd warp_sw_fast(const DevIndex& x, const MemOpt& opt, int lanes_simd, bool is_u8, int qlen, QAcc qa, int tlen, TAcc ta,
                         int xtra, u64* brow, int brow_cap, u32* err, u64* cell_counter)
{
	LANE_DECL;
	SwFwd r; r.score = 0; r.te = -1; r.qe = -1; r.score2 = -1; r.te2 = -1;
	const int slen = (qlen + lanes_simd - 1) / lanes_simd, plen = slen * lanes_simd;
	if (slen == 0) return r;
	const int k = 32 / lanes_simd, C = (slen + k - 1) / k;
	const int e_del = opt.e_del, e_ins = opt.e_ins, oe_del = opt.o_del + opt.e_del, oe_ins = opt.o_ins + opt.e_ins;
	int shift = 127, mmax = 0;
	for (int t = 0; t < 25; ++t) { if (opt.mat[t] < shift) shift = opt.mat[t]; if (opt.mat[t] > mmax) mmax = opt.mat[t]; }
	shift = (256 - (shift & 0xff)) & 0xff;
	const int minsc = (xtra & KSW_XSUBO) ? (xtra & 0xffff) : 0x10000;
	const int endsc = (xtra & KSW_XSTOP) ? (xtra & 0xffff) : 0x10000;
	WV(SwLane, st);
	FOR_LANES {
		SwLane& s = L_(st);
		int seg = lane / k; s.sub = lane % k;
		s.qs = seg * slen + s.sub * C;
		int qn = slen - s.sub * C; qn = qn < 0 ? 0 : (qn > C ? C : qn);
		s.qn = qn;
#pragma unroll
		for (int c = 0; c < SWF_CMAX; ++c) {
			s.Hp[c] = 0; s.E[c] = 0; s.Hm[c] = 0; s.H[c] = 0; s.fin[c] = 0;
			int q = s.qs + c;
			s.qc[c] = (c < qn && q < qlen) ? qacc_code(qa, q) : (u8)5; // 5 = zero-score padding column
		}
	}
	int gmax = 0, te = -1, n_b = 0;
	u64 cells = 0;
	WV(int, tch);
	FOR_LANES { L_(tch) = 0; }
	for (int i = 0; i < tlen; ++i) {
		if ((i & (W - 1)) == 0) { FOR_LANES { int ii = i + lane; L_(tch) = ii < tlen ? ref_base2(x, tacc_pos(ta, ii)) : 0; } }
		const int tb = wv_get(tch, i & (W - 1));
		// scores of the five query codes against this target base, one byte each; byte 5 (padding) = 0
		u64 rowpack = 0;
		for (int c = 0; c < 5; ++c) rowpack |= (u64)(u8)opt.mat[tb * 5 + c] << (8 * c);
		cells += (u64)qlen;
		WV(int, hlast); WV(int, hdiag0); WV(int, lout); WV(int, useg); WV(int, ufull);
		FOR_LANES { SwLane& s = L_(st); int v = 0;
#pragma unroll
			for (int c = 0; c < SWF_CMAX; ++c) if (c == s.qn - 1) v = s.H[c];
			L_(hlast) = v; }
		wv_shfl_up(hdiag0, hlast, 1);
		FOR_LANES {
			SwLane& s = L_(st);
			int hd = lane == 0 ? 0 : L_(hdiag0);
			int fl = 0;
#pragma unroll
			for (int c = 0; c < SWF_CMAX; ++c) {
				if (c < s.qn) {
					int sc = (int)(i8)(rowpack >> (8 * s.qc[c]));
					int h = hd;
					hd = s.H[c]; // becomes the diagonal of the next column
					if (is_u8) { h = h + sc + shift; if (h > 255) h = 255; h -= shift; if (h < 0) h = 0; }
					else { h = h + sc; if (h > 32767) h = 32767; }
					int e = s.E[c];
					h = h > e ? h : e;          // Hd: best score without a horizontal gap
					s.Hp[c] = h;
					s.fin[c] = fl;              // lane-local F entering this column (carry-in 0)
					int t = h - oe_ins; t = t > 0 ? t : 0;
					fl -= e_ins; fl = fl > 0 ? fl : 0;
					fl = fl > t ? fl : t;
				}
			}
			L_(lout) = fl;
			int u = fl + (s.qs + s.qn) * e_ins;
			L_(useg) = u; L_(ufull) = u;
		}
		// A lane whose outgoing F is 0 cannot raise any later column (carries are floored at 0), so the two scans
		// are skipped on rows where no lane has one -- the common case away from the true alignment.
		WV(bool, has_out);
		FOR_LANES { L_(has_out) = L_(lout) > 0; }
		WV(int, cfull); WV(int, cseg);
		if (wv_ballot(has_out) != 0u) {
			// full scan: F leaving every earlier lane, as seen at a common origin
			WV_SCAN_MAX(ufull);
			wv_shfl_up(cfull, ufull, 1);
			FOR_LANES { L_(cseg) = KSW_NEG_INF; }
			for (int d = 1; d < k; ++d) {
				WV(int, t_); wv_shfl_up(t_, useg, d);
				FOR_LANES { if (L_(st).sub >= d && lane >= d) L_(cseg) = imax(L_(cseg), L_(t_)); }
			}
		} else { FOR_LANES { L_(cfull) = KSW_NEG_INF; L_(cseg) = KSW_NEG_INF; } }
		WV(int, lmax);
		FOR_LANES {
			SwLane& s = L_(st);
			int cf = lane == 0 ? 0 : L_(cfull) - s.qs * e_ins; cf = cf > 0 ? cf : 0;
			int cs = s.sub == 0 ? 0 : L_(cseg) - s.qs * e_ins; cs = cs > 0 ? cs : 0;
			int mx = 0;
#pragma unroll
			for (int c = 0; c < SWF_CMAX; ++c) {
				if (c < s.qn) {
					int hd = s.Hp[c];
					int fs = cs - c * e_ins; fs = fs > s.fin[c] ? fs : s.fin[c]; // >= 0 because fin >= 0
					int ff = cf - c * e_ins; ff = ff > s.fin[c] ? ff : s.fin[c];
					int hp = hd > fs ? hd : fs;   // pass-1 value: segment-local F only
					int hh = hd > ff ? hd : ff;   // pass-2 value: F carried across segments as well
					mx = mx > hp ? mx : hp;
					int en = s.E[c] - e_del; en = en > 0 ? en : 0;
					int t = hp - oe_del; t = t > 0 ? t : 0;
					s.E[c] = en > t ? en : t;
					s.Hp[c] = hh;                // staged: becomes H below
				}
			}
			L_(lmax) = mx;
		}
		const int imx = wv_reduce_max(lmax);
		if (imx >= minsc) {
			if (n_b == 0 || (i32)(u32)brow[n_b - 1] + 1 != i) {
				if (n_b >= brow_cap) { ONE_LANE { atomic_or_u32(err, ERR_BROW_OVERFLOW); } }
				else { ONE_LANE { brow[n_b] = (u64)imx << 32 | (u32)i; } ++n_b; WARP_SYNC(); }
			} else if ((int)(brow[n_b - 1] >> 32) < imx) { ONE_LANE { brow[n_b - 1] = (u64)imx << 32 | (u32)i; } WARP_SYNC(); }
		}
		bool stop = false;
		const bool newmax = imx > gmax;
		if (newmax) {
			gmax = imx; te = i;
			if (is_u8) { if (gmax + shift >= 255 || gmax >= endsc) stop = true; }
			else if (gmax >= endsc) stop = true;
		}
		FOR_LANES {
			SwLane& s = L_(st);
#pragma unroll
			for (int c = 0; c < SWF_CMAX; ++c) { s.H[c] = s.Hp[c]; if (newmax) s.Hm[c] = s.Hp[c]; }
		}
		if (stop) break;
	}
	r.score = is_u8 ? (gmax + shift < 255 ? gmax : 255) : gmax;
	r.te = te;
	if (!is_u8 || r.score != 255) {
		WV(int, best); WV(int, qmin);
		FOR_LANES { SwLane& s = L_(st); int b = -1;
#pragma unroll
			for (int c = 0; c < SWF_CMAX; ++c) if (c < s.qn && s.Hm[c] > b) b = s.Hm[c];
			L_(best) = b; }
		const int mxv = wv_reduce_max(best);
		FOR_LANES { SwLane& s = L_(st); int qq = 0x7fffffff;
#pragma unroll
			for (int c = SWF_CMAX - 1; c >= 0; --c) if (c < s.qn && s.Hm[c] == mxv) qq = s.qs + c;
			L_(qmin) = qq; }
		r.qe = wv_reduce_min(qmin);
		if (n_b > 0) {
			int d = (r.score + mmax - 1) / mmax, low = te - d, high = te + d;
			for (int t = 0; t < n_b; ++t) {
				int e = (i32)(u32)brow[t];
				if ((e < low || e > high) && (int)(brow[t] >> 32) > r.score2) { r.score2 = (int)(brow[t] >> 32); r.te2 = e; }
			}
		}
	}
	(void)plen;
	if (cell_counter) { ONE_LANE { atomic_add_u64(cell_counter, cells); } }
	return r;
}

// ---- snap: b[i] of mem_sam_pe ------------------------------------------------------------------------
struct RescueIO {
	const u64* seq_off; const u8* seq; const u64* reg_off; const u32* reg_cap;
	Reg* regs; u32* n_reg; Reg* snap; u32* nb;
	SwTask* tasks; u8* state; u32* list; u32* n_list; u32* max_nb;
	u32* multi; u32* n_multi; // the reads that take part in the round-by-round kernels (compact list)
	int n_reads; int round;
	// speculative pipeline: per snapshot slot (reg_off[r] + k) and orientation
	u8* spec_state; u32* spec_task; u32* n_reg0; u32* fb_list; u32* n_fb; int force_fb;
	u32* hv_list; u32* n_hv; // reads whose replay is done by a warp (long lists / several rescuing hits)
};
// replay by a warp (rescue_apply_all<32>, warp-cooperative de-duplication) instead of a thread?
DEV bool rescue_is_heavy(const RescueIO& io, i64 r) { return io.nb[r] >= 2u || io.n_reg0[r ^ 1] >= (u32)SDP_FAST_MIN; }

DEV void rescue_snap(const MemOpt& opt, const RescueIO& io, i64 r)
{
	io.n_reg0[r] = io.n_reg[r];
	if ((r | 1) >= io.n_reads) { io.nb[r] = 0; return; } // odd trailing read
	const Reg* a = io.regs + io.reg_off[r];
	Reg* b = io.snap + io.reg_off[r];
	int n = (int)io.n_reg[r], m = 0;
	for (int j = 0; j < n; ++j) if (a[j].score >= a[0].score - opt.pen_unpaired) b[m++] = a[j];
	if (m > opt.max_matesw) m = opt.max_matesw;
	io.nb[r] = (u32)m;
	if (m) atomic_max_u32(io.max_nb, (u32)m);
}
// second pass of the snapshot (needs nb and n_reg0 of both ends): queue the reads whose replay goes to a warp
DEV void rescue_classify(const RescueIO& io, i64 r)
{
	if ((r | 1) >= io.n_reads || io.nb[r] == 0) return;
	if (rescue_is_heavy(io, r)) { u32 slot = atomic_add_u32(io.n_hv, 1u); io.hv_list[slot] = (u32)r; }
}

// ---- gen: the skip test and the windows of mem_matesw for hit `round` of read r -----------------------
DEV void rescue_gen(const DevIndex& x, const MemOpt& opt, const PeStat* pes, const RescueIO& io, i64 r)
{
	if ((u32)io.round >= io.nb[r]) return;
	const i64 m = r ^ 1;
	const Reg a = io.snap[io.reg_off[r] + io.round];
	const Reg* ma = io.regs + io.reg_off[m];
	const int nm = (int)io.n_reg[m];
	const int l_ms = (int)(io.seq_off[m + 1] - io.seq_off[m] - 1);
	const i64 l_pac = x.l_pac;
	int skip[4];
	for (int d = 0; d < 4; ++d) skip[d] = pes[d].failed ? 1 : 0;
	for (int i = 0; i < nm; ++i) {
		i64 dist;
		int d = infer_dir(l_pac, a.rb, ma[i].rb, &dist);
		if (dist >= pes[d].low && dist <= pes[d].high) skip[d] = 1;
	}
	u8* state = io.state + (size_t)r * 4;
	if (skip[0] + skip[1] + skip[2] + skip[3] == 4) { state[0] = state[1] = state[2] = state[3] = SW_SKIP; return; }
	for (int d = 0; d < 4; ++d) {
		if (skip[d]) { state[d] = SW_SKIP; continue; }
		int is_rev = (d >> 1) != (d & 1), is_larger = !(d >> 1);
		i64 rb, re;
		if (!is_rev) {
			rb = is_larger ? a.rb + pes[d].low : a.rb - pes[d].high;
			re = (is_larger ? a.rb + pes[d].high : a.rb - pes[d].low) + l_ms;
		} else {
			rb = (is_larger ? a.rb + pes[d].low : a.rb - pes[d].high) - l_ms;
			re = is_larger ? a.rb + pes[d].high : a.rb - pes[d].low;
		}
		if (rb < 0) rb = 0;
		if (re > l_pac << 1) re = l_pac << 1;
		int rid = -1;
		if (rb < re) rid = ref_clip_window(x, rb, (rb + re) >> 1, re);
		if (a.rid == rid && re - rb >= opt.min_seed_len) {
			SwTask t;
			t.rb = rb; t.re = re; t.read = (u32)m; t.is_rev = is_rev;
			t.score = 0; t.te = t.qe = t.score2 = t.te2 = t.tb = t.qb = -1; t.pad_ = 0;
			u32 slot = atomic_add_u32(io.n_list, 1u);
			io.tasks[slot] = t;
			io.list[(size_t)r * 4 + d] = slot;
			state[d] = SW_RUN;
		} else state[d] = SW_NORUN;
	}
}

// ---- gen_all: the tasks of every rescuing hit of read r, chosen against the mate's original list -------------
DEV void rescue_gen_all(const DevIndex& x, const MemOpt& opt, const PeStat* pes, const RescueIO& io, i64 r)
{
	const int nb = (int)io.nb[r];
	if (nb == 0) return;
	const i64 m = r ^ 1;
	const Reg* ma = io.regs + io.reg_off[m];
	const int nm = (int)io.n_reg[m];
	const int l_ms = (int)(io.seq_off[m + 1] - io.seq_off[m] - 1);
	const i64 l_pac = x.l_pac;
	for (int k = 0; k < nb; ++k) {
		const Reg a = io.snap[io.reg_off[r] + k];
		int skip[4];
		for (int d = 0; d < 4; ++d) skip[d] = pes[d].failed ? 1 : 0;
		for (int i = 0; i < nm; ++i) {
			i64 dist;
			int d = infer_dir(l_pac, a.rb, ma[i].rb, &dist);
			if (dist >= pes[d].low && dist <= pes[d].high) skip[d] = 1;
		}
		u8* state = io.spec_state + (io.reg_off[r] + k) * 4;
		u32* tix = io.spec_task + (io.reg_off[r] + k) * 4;
		for (int d = 0; d < 4; ++d) {
			if (skip[d]) { state[d] = SW_SKIP; continue; }
			int is_rev = (d >> 1) != (d & 1), is_larger = !(d >> 1);
			i64 rb, re;
			if (!is_rev) {
				rb = is_larger ? a.rb + pes[d].low : a.rb - pes[d].high;
				re = (is_larger ? a.rb + pes[d].high : a.rb - pes[d].low) + l_ms;
			} else {
				rb = (is_larger ? a.rb + pes[d].low : a.rb - pes[d].high) - l_ms;
				re = is_larger ? a.rb + pes[d].high : a.rb - pes[d].low;
			}
			if (rb < 0) rb = 0;
			if (re > l_pac << 1) re = l_pac << 1;
			int rid = -1;
			if (rb < re) rid = ref_clip_window(x, rb, (rb + re) >> 1, re);
			if (a.rid == rid && re - rb >= opt.min_seed_len) {
				SwTask t;
				t.rb = rb; t.re = re; t.read = (u32)m; t.is_rev = is_rev;
				t.score = 0; t.te = t.qe = t.score2 = t.te2 = t.tb = t.qb = -1; t.pad_ = 0;
				u32 slot = atomic_add_u32(io.n_list, 1u);
				io.tasks[slot] = t;
				tix[d] = slot;
				state[d] = SW_RUN;
			} else state[d] = SW_NORUN;
		}
	}
}

// ---- sw: one ksw_align2 of mem_matesw --------------------------------------------------------------------
template <int W>
DEVNI void rescue_sw(const DevIndex& x, const MemOpt& opt, WarpMem& wm, WarpScratch& ws, Counters* cnt, const RescueIO& io, u32 task_id)
{
	LANE_DECL;
	SwTask t = io.tasks[task_id];
	const u8* ms = io.seq + io.seq_off[t.read];
	const int l_ms = (int)(io.seq_off[t.read + 1] - io.seq_off[t.read] - 1);
	const int xtra = KSW_XSUBO | KSW_XSTART | (l_ms * opt.a < 250 ? KSW_XBYTE : 0) | (opt.min_seed_len * opt.a);
	const bool is_u8 = (xtra & KSW_XBYTE) != 0; const int lanes_simd = is_u8 ? 16 : 8;
	const int tlen = (int)(t.re - t.rb);
	SwFwd fw, rv;
	int qb_ = -1, tb_ = -1;
	TAcc ta; ta.base = t.rb; ta.stride = 1; ta.piv = -1;
	if (W == 32 && swf_fits(l_ms, lanes_simd)) {
		QAcc qa; qa.ms = ms; qa.l_ms = l_ms; qa.is_rev = t.is_rev; qa.piv = -1;
		fw = warp_sw_fast<W>(x, opt, lanes_simd, is_u8, l_ms, qa, tlen, ta, xtra, ws.brow, ws.brow_cap, ws.err, cnt ? &cnt->cells_msw : nullptr);
	} else {
		int slen = (l_ms + lanes_simd - 1) / lanes_simd, plen = slen * lanes_simd;
		if (plen > wm.cap) { ONE_LANE { atomic_or_u32(ws.err, ERR_QLEN_OVERFLOW); } return; }
		if (tlen > ws.brow_cap) { ONE_LANE { atomic_or_u32(ws.err, ERR_BROW_OVERFLOW); } return; }
		FOR_LANES { for (int j = lane; j < plen; j += W) { u8 c = 5; if (j < l_ms) { if (t.is_rev) { u8 b = ms[l_ms - 1 - j]; c = b < 4 ? (u8)(3 - b) : (u8)4; } else c = ms[j]; } WQ(wm, j) = c; } }
		WARP_SYNC();
		fw = warp_sw_striped<W>(x, opt, wm, lanes_simd, is_u8, l_ms, tlen, ta, xtra, ws.brow, cnt ? &cnt->cells_msw : nullptr);
	}
	if (cnt) { ONE_LANE { atomic_add_u64(&cnt->n_msw, 1); } }
	if (!(fw.score < (xtra & 0xffff))) { // XSUBO: the start is searched only for scores at or above the threshold
		const int ql2 = fw.qe + 1;
		TAcc tr; tr.base = t.rb; tr.stride = 1; tr.piv = fw.te; // target prefix [0,te] reversed, the rest untouched
		if (W == 32 && swf_fits(ql2, lanes_simd)) {
			QAcc qa; qa.ms = ms; qa.l_ms = l_ms; qa.is_rev = t.is_rev; qa.piv = fw.qe;
			rv = warp_sw_fast<W>(x, opt, lanes_simd, is_u8, ql2, qa, tlen, tr, KSW_XSTOP | fw.score, ws.brow, ws.brow_cap, ws.err, cnt ? &cnt->cells_msw : nullptr);
		} else {
			int slen2 = (ql2 + lanes_simd - 1) / lanes_simd, plen2 = slen2 * lanes_simd;
			if (plen2 > wm.cap) { ONE_LANE { atomic_or_u32(ws.err, ERR_QLEN_OVERFLOW); } return; }
			FOR_LANES { for (int j = lane; j < plen2; j += W) { u8 c = 5; if (j < ql2) { int jj = fw.qe - j; if (t.is_rev) { u8 b = ms[l_ms - 1 - jj]; c = b < 4 ? (u8)(3 - b) : (u8)4; } else c = ms[jj]; } WQ(wm, j) = c; } }
			WARP_SYNC();
			rv = warp_sw_striped<W>(x, opt, wm, lanes_simd, is_u8, ql2, tlen, tr, KSW_XSTOP | fw.score, ws.brow, cnt ? &cnt->cells_msw : nullptr);
		}
		if (fw.score == rv.score) { tb_ = fw.te - rv.te; qb_ = fw.qe - rv.qe; }
	}
	ONE_LANE {
		SwTask* o = &io.tasks[task_id];
		o->score = fw.score; o->te = fw.te; o->qe = fw.qe; o->score2 = fw.score2; o->te2 = fw.te2; o->tb = tb_; o->qb = qb_;
	}
	WARP_SYNC();
}

// the region a finished alignment task contributes (false: the task found nothing usable)
DEV bool rescue_make_reg(const MemOpt& opt, i64 l_pac, const Reg& a, const SwTask& t, int l_ms, Reg& b)
{
	if (!(t.score >= opt.min_seed_len && t.qb >= 0)) return false;
	const int is_rev = t.is_rev;
	memset(&b, 0, sizeof(Reg));
	b.rid = a.rid; b.is_alt = a.is_alt;
	b.qb = is_rev ? l_ms - (t.qe + 1) : t.qb;
	b.qe = is_rev ? l_ms - t.qb : t.qe + 1;
	b.rb = is_rev ? (l_pac << 1) - (t.rb + t.te + 1) : t.rb + t.tb;
	b.re = is_rev ? (l_pac << 1) - (t.rb + t.tb) : t.rb + t.te + 1;
	b.score = t.score; b.csub = t.score2; b.secondary = -1;
	b.seedcov = (int)((b.re - b.rb < b.qe - b.qb ? b.re - b.rb : b.qe - b.qb) >> 1);
	return true;
}

DEV bool sdp_redundant(const MemOpt& opt, const Reg& p, const Reg& q) // the test of mem_sort_dedup_patch; p ends after q
{
	i64 or_ = q.re - p.rb;
	i64 oq = q.qb < p.qb ? q.qe - p.qb : p.qe - q.qb;
	i64 mr = q.re - q.rb < p.re - p.rb ? q.re - q.rb : p.re - p.rb;
	i64 mq = q.qe - q.qb < p.qe - p.qb ? q.qe - q.qb : p.qe - p.qb;
	return (float)or_ > opt.mask_level_redun * (float)mr && (float)oq > opt.mask_level_redun * (float)mq;
}

// mem_matesw's "insert the hit, then mem_sort_dedup_patch" for ONE new hit b on a SETTLED list (in final order, all
// end positions and all (score, rb, qb) keys distinct, every pair within scan range already found non-redundant):
// O(n / 32) per lane instead of two sorts.  Why this is exact: with all end positions distinct the order of the scan is
// fixed, a new element can only shorten the scans of the others (never extend them), and pairs of old elements were
// examined in the same orientation before -- so only pairs that involve b can have an effect:
//   * b's own scan walks down the elements that end before it, on its contig, within max_chain_gap, until the
//     nearest element of another contig; the first redundant one with a higher score removes b, redundant ones met
//     before that are removed by b;
//   * the elements that end after b and reach it (same contig, start within max_chain_gap of b's end, no element of
//     another contig in between) meet b in ascending order of their ends while it is alive: a redundant one with a
//     lower score is removed, the first redundant one with a score >= b's removes b.
// Returns the new length, or -1 when an equal end position / equal key makes the order the reference's unstable sort
// would produce observable (the caller then takes the general path).  *settled stays true only if nothing was removed.
template <int W>
DEVNI int rescue_insert_settled(const MemOpt& opt, WarpScratch& ws, Reg* ma, int nm, const Reg& b, bool* settled)
{
	LANE_DECL;
	const i64 BIG = (i64)1 << 62;
	// pass 1: ties, the nearest elements of other contigs below / above b's end
	WV(bool, tie); WV(i64, blo); WV(i64, bhi);
	FOR_LANES {
		bool tv = false; i64 lo = -BIG, hi = BIG;
		for (int i = lane; i < nm; i += W) {
			const Reg& e = ma[i];
			tv |= e.re == b.re || (e.score == b.score && e.rb == b.rb && e.qb == b.qb);
			if (e.rid != b.rid) { if (e.re < b.re) lo = e.re > lo ? e.re : lo; else hi = e.re < hi ? e.re : hi; }
		}
		L_(tie) = tv; L_(blo) = lo; L_(bhi) = hi;
	}
	SDP_STAT(5);
	if (wv_ballot(tie) != 0u) { SDP_STAT(6); return -1; }
	const i64 B_lo = wv_reduce_max64(blo), B_hi = wv_reduce_min64(bhi);
	// pass 2: b's own scan (downwards): the redundant element with a higher score that ends last stops it
	WV(i64, rs);
	FOR_LANES {
		i64 v = -BIG;
		for (int i = lane; i < nm; i += W) {
			const Reg& e = ma[i];
			if (e.rid == b.rid && e.re < b.re && e.re > B_lo && b.rb < e.re + opt.max_chain_gap && sdp_redundant(opt, b, e) && b.score < e.score) v = e.re > v ? e.re : v;
		}
		L_(rs) = v;
	}
	const i64 R_star = wv_reduce_max64(rs);
	bool b_alive = R_star == -BIG;
	// pass 3: the elements above that reach b; the first (lowest end) redundant one with score >= b's removes b
	i64 R_hat = BIG;
	if (b_alive) {
		WV(i64, rh);
		FOR_LANES {
			i64 v = BIG;
			for (int i = lane; i < nm; i += W) {
				const Reg& e = ma[i];
				if (e.rid == b.rid && e.re > b.re && e.re < B_hi && e.rb < b.re + opt.max_chain_gap && sdp_redundant(opt, e, b) && !(e.score < b.score)) v = e.re < v ? e.re : v;
			}
			L_(rh) = v;
		}
		R_hat = wv_reduce_min64(rh);
		if (R_hat != BIG) b_alive = false;
	}
	// pass 4: survivors in order, b at its rank; built in the scratch, then copied back
	Reg* t = ws.rt;
	int n_out = 0, n_del = 0;
	bool placed = !b_alive;
	for (int i0 = 0; i0 < nm; i0 += W) {
		WV(bool, keep); WV(bool, before);
		FOR_LANES {
			const int i = i0 + lane;
			bool kp = false, bf = false;
			if (i < nm) {
				const Reg& e = ma[i];
				bool del = false;
				if (e.rid == b.rid) {
					if (e.re < b.re) del = e.re > B_lo && e.re > R_star && b.rb < e.re + opt.max_chain_gap && !(b.score < e.score) && sdp_redundant(opt, b, e);
					else if (R_star == -BIG) del = e.re < B_hi && e.re < R_hat && e.rb < b.re + opt.max_chain_gap && e.score < b.score && sdp_redundant(opt, e, b);
				}
				kp = !del;
				bf = e.score > b.score || (e.score == b.score && (e.rb < b.rb || (e.rb == b.rb && e.qb < b.qb))); // e sorts before b
			}
			L_(keep) = kp; L_(before) = bf;
		}
		const u32 km = wv_ballot(keep), bm = wv_ballot(before);
		const int cnt = i0 + W <= nm ? W : nm - i0;
		const u32 valid = cnt == 32 ? 0xffffffffu : ((1u << cnt) - 1u);
		// the list is ordered, so the elements that sort before b form a prefix: b goes in front of the first one that does not
		u32 after = ~bm & valid;
		int split = (!placed && after) ? __ffs(after) - 1 : -1; // lane index in this chunk in front of which b is placed
		FOR_LANES {
			const int i = i0 + lane;
			if (i < nm && ((km >> lane) & 1u)) {
				int pos = n_out + __popc(km & ((1u << lane) - 1u));
				if (split >= 0 && lane >= split) ++pos;
				t[pos] = ma[i];
			}
		}
		if (split >= 0) { ONE_LANE { Reg bb = b; bb.n_comp = 1; t[n_out + __popc(km & ((1u << split) - 1u))] = bb; } placed = true; ++n_out; }
		n_out += __popc(km);
		n_del += cnt - __popc(km);
	}
	if (!placed) { ONE_LANE { Reg bb = b; bb.n_comp = 1; t[n_out] = bb; } ++n_out; }
	WARP_SYNC();
	FOR_LANES { for (int i = lane; i < n_out; i += W) ma[i] = t[i]; }
	WARP_SYNC();
	*settled = n_del == 0 && b_alive;
	if (!*settled) SDP_STAT(7);
	return n_out;
}

// the hit of one finished alignment task, inserted by score into the mate's list (false: no room)
DEV bool rescue_insert(const MemOpt& opt, i64 l_pac, const Reg& a, const SwTask& t, int l_ms, Reg* ma, int* n_ma, int ma_cap)
{
	if (!(t.score >= opt.min_seed_len && t.qb >= 0)) return true;
	const int is_rev = t.is_rev;
	int nm = *n_ma;
	Reg b;
	memset(&b, 0, sizeof(Reg));
	b.rid = a.rid; b.is_alt = a.is_alt;
	b.qb = is_rev ? l_ms - (t.qe + 1) : t.qb;
	b.qe = is_rev ? l_ms - t.qb : t.qe + 1;
	b.rb = is_rev ? (l_pac << 1) - (t.rb + t.te + 1) : t.rb + t.tb;
	b.re = is_rev ? (l_pac << 1) - (t.rb + t.tb) : t.rb + t.te + 1;
	b.score = t.score; b.csub = t.score2; b.secondary = -1;
	b.seedcov = (int)((b.re - b.rb < b.qe - b.qb ? b.re - b.rb : b.qe - b.qb) >> 1);
	if (nm >= ma_cap) return false;
	int i, tmp;
	for (i = 0; i < nm; ++i) if (ma[i].score < b.score) break;
	tmp = i;
	for (i = nm; i > tmp; --i) ma[i] = ma[i - 1];
	ma[tmp] = b;
	*n_ma = nm + 1;
	return true;
}

// ---- apply: insert the hits of one mem_matesw call into the mate's list, de-duplicating as the loop does ----
template <int W>
DEV void rescue_apply(const DevIndex& x, const MemOpt& opt, WarpMem& wm, WarpScratch& ws, const RescueIO& io, i64 r)
{
	if ((u32)io.round >= io.nb[r]) return;
	const u8* state = io.state + (size_t)r * 4;
	if (state[0] == SW_SKIP && state[1] == SW_SKIP && state[2] == SW_SKIP && state[3] == SW_SKIP) return;
	const i64 m = r ^ 1;
	const Reg a = io.snap[io.reg_off[r] + io.round];
	Reg* ma = io.regs + io.reg_off[m];
	int nm = (int)io.n_reg[m], n = 0;
	const int ma_cap = (int)io.reg_cap[m];
	const int l_ms = (int)(io.seq_off[m + 1] - io.seq_off[m] - 1);
	const i64 l_pac = x.l_pac;
	for (int d = 0; d < 4; ++d) {
		if (state[d] == SW_SKIP) continue;
		if (state[d] == SW_RUN) {
			const SwTask t = io.tasks[io.list[(size_t)r * 4 + d]];
			if (!rescue_insert(opt, l_pac, a, t, l_ms, ma, &nm, ma_cap)) { atomic_or_u32(ws.err, ERR_REG_OVERFLOW); return; }
			++n;
		}
		if (n) nm = warp_sort_dedup_patch<W>(x, opt, wm, ws, nullptr, nullptr, nm, ma);
	}
	io.n_reg[m] = (u32)nm;
}

// ---- apply_all: replay of the whole rescue chain of read r (hits 0 .. nb-1 into the mate's list) from the
// speculated results.  Returns false when the read has to take the round-by-round path instead.
template <int W>
DEV bool rescue_apply_all(const DevIndex& x, const MemOpt& opt, const PeStat* pes, WarpMem& wm, WarpScratch& ws, const RescueIO& io, i64 r)
{
	LANE_DECL;
	const int nb = (int)io.nb[r];
	if (nb == 0) return true;
	const i64 m = r ^ 1;
	Reg* ma = io.regs + io.reg_off[m];
	int nm = (int)io.n_reg[m];
	const int ma_cap = (int)io.reg_cap[m];
	const int l_ms = (int)(io.seq_off[m + 1] - io.seq_off[m] - 1);
	const i64 l_pac = x.l_pac;
	bool ok = true;
	bool settled = false; // the list is a fixed point of the de-duplication (nothing inserted since a clean pass)
	for (int k = 0; k < nb && ok; ++k) {
		const Reg a = io.snap[io.reg_off[r] + k];
		int skip[4];
		for (int d = 0; d < 4; ++d) skip[d] = pes[d].failed ? 1 : 0;
		{ // which orientations already have a hit at a proper distance (lanes share the list)
			WV(int, sk);
			FOR_LANES {
				int bits = 0;
				for (int i = lane; i < nm; i += W) {
					i64 dist;
					int d = infer_dir(l_pac, a.rb, ma[i].rb, &dist);
					if (dist >= pes[d].low && dist <= pes[d].high) bits |= 1 << d;
				}
				L_(sk) = bits;
			}
			int all = 0;
			for (int d = 0; d < 4; ++d) { WV(bool, bt); FOR_LANES { L_(bt) = (L_(sk) >> d & 1) != 0; } if (wv_ballot(bt) != 0u) all |= 1 << d; }
			for (int d = 0; d < 4; ++d) if (all >> d & 1) skip[d] = 1;
		}
		if (skip[0] + skip[1] + skip[2] + skip[3] == 4) continue;
		const u8* state = io.spec_state + (io.reg_off[r] + k) * 4;
		const u32* tix = io.spec_task + (io.reg_off[r] + k) * 4;
		for (int d = 0; d < 4; ++d) if (!skip[d] && state[d] == SW_SKIP) ok = false; // needed, not speculated
		if (!ok) break;
		int n = 0;
		for (int d = 0; d < 4; ++d) {
			if (skip[d]) continue;
			bool deduped = false; // this orientation's de-duplication has been done (fused with the insertion)
			if (state[d] == SW_RUN) {
				const SwTask t = io.tasks[tix[d]];
				Reg b;
				if (rescue_make_reg(opt, l_pac, a, t, l_ms, b)) {
					if (nm >= ma_cap) { ONE_LANE { atomic_or_u32(ws.err, ERR_REG_OVERFLOW); io.n_reg[m] = (u32)nm; } return true; }
					int nm2 = -1;
					if (W == 32 && settled && nm >= SDP_FAST_MIN && nm + 1 <= ws.rt_cap && ws.rt != nullptr) nm2 = rescue_insert_settled<W>(opt, ws, ma, nm, b, &settled);
					if (nm2 >= 0) { nm = nm2; deduped = true; }
					else {
						LANE0_ONLY { int i, tmp; for (i = 0; i < nm; ++i) if (ma[i].score < b.score) break; tmp = i; for (i = nm; i > tmp; --i) ma[i] = ma[i - 1]; ma[tmp] = b; }
						WARP_SYNC();
						++nm; settled = false;
					}
				}
				++n;
			}
			if (n && !deduped && !settled) nm = warp_sort_dedup_fast<W>(x, opt, wm, ws, nullptr, nullptr, nm, ma, &settled);
		}
	}
	ONE_LANE { io.n_reg[m] = (u32)nm; }
	return ok && !io.force_fb;
}
