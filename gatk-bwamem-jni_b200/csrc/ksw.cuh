// ksw.cuh -- the three dynamic-programming primitives as warp-cooperative device functions.
//
// Replaces ksw_extend2, ksw_global2 and ksw_align2 (-> ksw_u8 / ksw_i16) of un-vendored lh3/bwa
// (SURVEY.md App. B.7, B.9, B.8), reached from /root/reference/src/main/c/jnibwa.c:214.
//
// Scheme (one warp per task, rows sequential, 32 query columns per step):
//   * In ksw_extend2 / ksw_global2 the gap states are fed by M (the diagonal move), not by H, so
//     within a row M and E are lane-local and F is a max-plus prefix scan over lanes
//     (v_k = g_k + k*e_ins, inclusive max-scan with 5 shuffles, F_j = scan_j - j*e_ins).  Each row
//     completes before the next starts, so band clamping, the per-row [beg,end) shrink, z-drop and the
//     tie rules of the scalar formulation are reproduced exactly.
//   * ksw_align2 reproduces the observable semantics of the striped SIMD kernels: zero-score padding
//     columns, segment-local F in the first pass (segmented max-plus scan), E fed by the
//     pre-correction H, then the carried F of earlier segments applied on top.
//   * Integer arithmetic only (IMNMX / VIMNMX3 / VIADDMNMX on sm_100a); no tensor cores -- this is
//     dynamic programming, not a dense contraction.
// The target base of row i is a warp-uniform value: lanes prefetch 32 consecutive reference bases from
// the 2-bit packed reference in HBM/L2 every 32 rows and broadcast one per row with a shuffle.
#pragma once
#include "util.cuh"

#define KSW_NEG_INF (-0x40000000)
#define KSW_XBYTE  0x10000
#define KSW_XSTOP  0x20000
#define KSW_XSUBO  0x40000
#define KSW_XSTART 0x80000

// per-warp shared-memory working set
struct WarpMem {
	i32* H;    // [cap+2]
	i32* E;    // [cap+2]
	i32* H2;   // [cap+2]   (align2: second H row)
	i32* Hm;   // [cap+2]   (align2: snapshot of the best row)
	u8* q;     // [cap+2]   query codes in DP orientation (5 = padding column); element j is q[j * q_stride] (WQ)
	int q_stride; // 1 for a warp's private row; the block size when threads of a block interleave their rows in shared memory
	i32* fend; // [16]
	int cap;
	// thread-level kernels (W = 1): ring buffers for the banded global alignment, in shared memory, one slot column per
	// thread (element k of this thread's ring is ring?[k * ring_stride]; stride = block size keeps every access on bank tid % 32)
	i32* ringH; i32* ringE; int ring_n; int ring_stride;
};
#define THREAD_RING_N 64
// element stride of a thread's shared-memory arrays in the extension-side / rescue task kernels: the block size, a
// compile-time constant there, so that the unrolled cells of a row are addressed with immediate offsets
#ifdef B200_SIM
#define TX_STRIDE 1
#else
#define TX_STRIDE THREADS_PER_BLOCK
#endif
#define WQ(wm_, j_) (wm_).q[(size_t)(j_) * (wm_).q_stride]

// target accessor over the doubled reference text
struct TAcc { i64 base; int stride; int piv; };
DEV i64 tacc_pos(const TAcc& a, int i) { return a.stride < 0 ? a.base - i : (i <= a.piv ? a.base + a.piv - i : a.base + i); }

struct ExtResult { int score, qle, tle, gtle, gscore, max_off; };

// inclusive max-scan over the 32 lanes
#define WV_SCAN_MAX(pm) do { \
	for (int d_ = 1; d_ < W; d_ <<= 1) { \
		WV(int, t_); wv_shfl_up(t_, pm, d_); \
		FOR_LANES { if (lane >= d_) L_(pm) = imax(L_(pm), L_(t_)); } \
	} } while (0)


// ---- thread-per-task formulations (W = 1): the scalar recurrences, one cell at a time ------------------------
DEVNI void thread_extend2(const DevIndex& x, const MemOpt& opt, WarpMem& wm, int qlen, int tlen, TAcc ta,
                          int w, int end_bonus, int zdrop, int h0, ExtResult& res, u64* cell_counter)
{
	const int o_del = opt.o_del, e_del = opt.e_del, o_ins = opt.o_ins, e_ins = opt.e_ins;
	const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	i32* H = wm.H; i32* E = wm.E; const u8* q = wm.q; const size_t qs = (size_t)wm.q_stride;
	H[0] = h0; E[0] = 0;
	{
		int v = h0 > oe_ins ? h0 - oe_ins : 0;
		if (qlen >= 1) { H[1] = v; E[1] = 0; }
		int j;
		for (j = 2; j <= qlen && H[j - 1] > e_ins; ++j) { H[j] = H[j - 1] - e_ins; E[j] = 0; }
		for (; j <= qlen; ++j) { H[j] = 0; E[j] = 0; }
	}
	int mx = 0;
	for (int k = 0; k < 25; ++k) mx = mx > opt.mat[k] ? mx : (int)opt.mat[k];
	int max_ins = (int)((double)(qlen * mx + end_bonus - o_ins) / e_ins + 1.); max_ins = max_ins > 1 ? max_ins : 1;
	w = w < max_ins ? w : max_ins;
	int max_del = (int)((double)(qlen * mx + end_bonus - o_del) / e_del + 1.); max_del = max_del > 1 ? max_del : 1;
	w = w < max_del ? w : max_del;
	int max = h0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, max_off = 0, beg = 0, end = qlen;
	u64 cells = 0;
	for (int i = 0; i < tlen; ++i) {
		int f = 0, h1, m = 0, mj = -1, j;
		const int tb = ref_base2(x, tacc_pos(ta, i));
		u64 rowpack = 0;
		for (int c = 0; c < 5; ++c) rowpack |= (u64)(u8)opt.mat[tb * 5 + c] << (8 * c);
		if (beg < i - w) beg = i - w;
		if (end > i + w + 1) end = i + w + 1;
		if (end > qlen) end = qlen;
		if (beg == 0) { h1 = h0 - (o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; } else h1 = 0;
		if (end > beg) cells += (u64)(end - beg);
		for (j = beg; j < end; ++j) {
			int h, M = H[j], e = E[j], t;
			H[j] = h1;
			M = M ? M + (int)(i8)(rowpack >> (8 * q[j * qs])) : 0;
			h = M > e ? M : e;
			h = h > f ? h : f;
			h1 = h;
			mj = m > h ? mj : j;
			m = m > h ? m : h;
			t = M - oe_del; t = t > 0 ? t : 0;
			e -= e_del; e = e > t ? e : t;
			E[j] = e;
			t = M - oe_ins; t = t > 0 ? t : 0;
			f -= e_ins; f = f > t ? f : t;
		}
		H[end] = h1; E[end] = 0;
		if (j == qlen) { max_ie = gscore > h1 ? max_ie : i; gscore = gscore > h1 ? gscore : h1; }
		if (m == 0) break;
		if (m > max) { max = m; max_i = i; max_j = mj; int off = iabs(mj - i); max_off = max_off > off ? max_off : off; }
		else if (zdrop > 0) {
			if (i - max_i > mj - max_j) { if (max - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop) break; }
			else { if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) break; }
		}
		for (j = beg; j < end && H[j] == 0 && E[j] == 0; ++j) {}
		beg = j;
		for (j = end; j >= beg && H[j] == 0 && E[j] == 0; --j) {}
		end = j + 2 < qlen ? j + 2 : qlen;
	}
	res.score = max; res.qle = max_j + 1; res.tle = max_i + 1; res.gtle = max_ie + 1; res.gscore = gscore; res.max_off = max_off;
	if (cell_counter) atomic_add_u64(cell_counter, cells);
}

DEVNI int thread_global2(const DevIndex& x, const MemOpt& opt, WarpMem& wm, int qlen, int tlen, TAcc ta, int w,
                         u8* z, u32* cigar, int cigar_cap, u32* err, int* n_cigar_out, u64* cell_counter)
{
	const int o_del = opt.o_del, e_del = opt.e_del, o_ins = opt.o_ins, e_ins = opt.e_ins;
	const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	const u8* q = wm.q; const size_t qs = (size_t)wm.q_stride;
	// H/E rows: a ring of ring_n >= 2w+2 column slots when the caller provides one (a row touches columns
	// [i-w, i+w+1] only and every column entering the band was written as H[end] by the row before), else the full rows
	const bool ring = wm.ringH != nullptr && 2 * w + 2 <= wm.ring_n;
	i32* H = ring ? wm.ringH : wm.H; i32* E = ring ? wm.ringE : wm.E;
	const int rmask = ring ? wm.ring_n - 1 : -1, rstride = ring ? wm.ring_stride : 1;
#define HX(j) (((j) & rmask) * rstride)
	const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
	const int zs = (n_col + 3) & ~3; // traceback row stride: whole 32-bit words, so four cells go out per store
	const bool bt = cigar != nullptr;
	if (n_cigar_out) *n_cigar_out = 0;
	{
		int j;
		H[HX(0)] = 0; E[HX(0)] = KSW_NEG_INF;
		for (j = 1; j <= qlen && j <= w; ++j) { H[HX(j)] = -(o_ins + e_ins * j); E[HX(j)] = KSW_NEG_INF; }
		if (!ring) for (; j <= qlen; ++j) { H[j] = KSW_NEG_INF; E[j] = KSW_NEG_INF; }
		else if (j <= qlen) { H[HX(j)] = KSW_NEG_INF; E[HX(j)] = KSW_NEG_INF; } // column w+1: only ever read after row 0 overwrote it
	}
	u64 cells = 0;
	for (int i = 0; i < tlen; ++i) {
		const int tb = ref_base2(x, tacc_pos(ta, i));
		u64 rowpack = 0;
		for (int c = 0; c < 5; ++c) rowpack |= (u64)(u8)opt.mat[tb * 5 + c] << (8 * c);
		int f = KSW_NEG_INF, h1, t;
		const int beg = i > w ? i - w : 0;
		const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
		h1 = beg == 0 ? -(o_del + e_del * (i + 1)) : KSW_NEG_INF;
		if (end > beg) cells += (u64)(end - beg);
		u32* zi = bt ? (u32*)(z + (size_t)i * zs) : nullptr;
		u32 zw = 0;
		for (int j = beg; j < end; ++j) {
			const int jx = HX(j);
			int h, m = H[jx], e = E[jx];
			u8 d;
			H[jx] = h1;
			m += (int)(i8)(rowpack >> (8 * q[j * qs]));
			d = m >= e ? 0 : 1;
			h = m >= e ? m : e;
			d = h >= f ? d : 2;
			h = h >= f ? h : f;
			h1 = h;
			t = m - oe_del;
			e -= e_del;
			d |= e > t ? 1 << 2 : 0;
			e = e > t ? e : t;
			E[jx] = e;
			t = m - oe_ins;
			f -= e_ins;
			d |= f > t ? 2 << 4 : 0;
			f = f > t ? f : t;
			if (zi) {
				const int k = j - beg;
				zw |= (u32)d << ((k & 3) << 3);
				if ((k & 3) == 3 || j == end - 1) { zi[k >> 2] = zw; zw = 0; }
			}
		}
		H[HX(end)] = h1; E[HX(end)] = KSW_NEG_INF;
	}
	const int score = H[HX(qlen)];
#undef HX
	if (bt) {
		bool cg_ovf = false;
		int n_cigar = 0, which = 0, i = tlen - 1, k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
		u32 cur_op = 0xffffffffu, cur_len = 0;
#define CG_PUT1() do { if (n_cigar < cigar_cap) cigar[n_cigar] = cur_len << 4 | cur_op; else cg_ovf = true; ++n_cigar; } while (0)
		while (i >= 0 && k >= 0) {
			which = z[(size_t)i * zs + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
			u32 op;
			if (which == 0) { op = 0; --i; --k; }
			else if (which == 1) { op = 2; --i; }
			else { op = 1; --k; }
			if (op == cur_op) ++cur_len;
			else { if (cur_op != 0xffffffffu) { CG_PUT1(); } cur_op = op; cur_len = 1; }
		}
		if (i >= 0) { if (cur_op == 2) cur_len += (u32)(i + 1); else { if (cur_op != 0xffffffffu) { CG_PUT1(); } cur_op = 2; cur_len = (u32)(i + 1); } }
		if (k >= 0) { if (cur_op == 1) cur_len += (u32)(k + 1); else { if (cur_op != 0xffffffffu) { CG_PUT1(); } cur_op = 1; cur_len = (u32)(k + 1); } }
		if (cur_op != 0xffffffffu) { CG_PUT1(); }
#undef CG_PUT1
		if (cg_ovf) { atomic_or_u32(err, ERR_CIGAR_OVERFLOW); n_cigar = 0; }
		else for (int a = 0; a < n_cigar >> 1; ++a) { u32 t = cigar[a]; cigar[a] = cigar[n_cigar - 1 - a]; cigar[n_cigar - 1 - a] = t; }
		*n_cigar_out = n_cigar;
	}
	if (cell_counter) atomic_add_u64(cell_counter, cells);
	return score;
}

template <int W>
DEVNI void warp_extend2(const DevIndex& x, const MemOpt& opt, WarpMem& wm, int qlen, int tlen, TAcc ta,
                        int w, int end_bonus, int zdrop, int h0, ExtResult& res, u64* cell_counter)
{
	if (W == 1) { thread_extend2(x, opt, wm, qlen, tlen, ta, w, end_bonus, zdrop, h0, res, cell_counter); return; }
	LANE_DECL;
	const int o_del = opt.o_del, e_del = opt.e_del, o_ins = opt.o_ins, e_ins = opt.e_ins;
	const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	i32* H = wm.H; i32* E = wm.E;
	// first row: H[0]=h0, H[1]=h0-oe_ins, then decay by e_ins while positive
	FOR_LANES {
		for (int j = lane; j <= qlen; j += W) {
			int v;
			if (j == 0) v = h0;
			else { v = h0 - oe_ins - (j - 1) * e_ins; if (v < 0) v = 0; }
			H[j] = v; E[j] = 0;
		}
	}
	// NB: the scalar fill stops at the first value <= e_ins ("while H[j-1] > e_ins"), i.e. H[j] = H[j-1]-e_ins is only
	// written while the predecessor exceeds e_ins; a predecessor in (0, e_ins] leaves the rest at 0.  h0-oe_ins-(j-1)e
	// clamped at 0 differs only if some positive value <= e_ins were followed by a positive one, which cannot happen.
	int mx = 0;
	for (int k = 0; k < 25; ++k) mx = mx > opt.mat[k] ? mx : (int)opt.mat[k];
	int max_ins = (int)((double)(qlen * mx + end_bonus - o_ins) / e_ins + 1.); max_ins = max_ins > 1 ? max_ins : 1;
	w = w < max_ins ? w : max_ins;
	int max_del = (int)((double)(qlen * mx + end_bonus - o_del) / e_del + 1.); max_del = max_del > 1 ? max_del : 1;
	w = w < max_del ? w : max_del;
	int max = h0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, max_off = 0, beg = 0, end = qlen;
	u64 cells = 0;
	WARP_SYNC();
	WV(int, tch);
	FOR_LANES { L_(tch) = 0; }
	for (int i = 0; i < tlen; ++i) {
		if ((i & (W - 1)) == 0) { FOR_LANES { int ii = i + lane; L_(tch) = ii < tlen ? ref_base2(x, tacc_pos(ta, ii)) : 0; } }
		const int tb = wv_get(tch, i & (W - 1));
		if (beg < i - w) beg = i - w;
		if (end > i + w + 1) end = i + w + 1;
		if (end > qlen) end = qlen;
		int h1c, fcar = 0, m = 0, mj = -1, nzfirst = -1, nzlast = -1;
		if (beg == 0) { h1c = h0 - (o_del + e_del * (i + 1)); if (h1c < 0) h1c = 0; } else h1c = 0;
		if (end > beg) cells += (u64)(end - beg);
		for (int j0 = beg; j0 < end; j0 += W) {
			WV(int, M); WV(int, e); WV(int, pm); WV(bool, act);
			FOR_LANES {
				int j = j0 + lane;
				L_(act) = j < end;
				int hold = L_(act) ? H[j] : 0;
				L_(e) = L_(act) ? E[j] : 0;
				int mm = hold ? hold + (int)opt.mat[tb * 5 + WQ(wm, L_(act) ? j : 0)] : 0;
				L_(M) = mm;
				int g = mm - oe_ins; g = g > 0 ? g : 0;
				L_(pm) = L_(act) ? g + lane * e_ins : KSW_NEG_INF;
			}
			WV_SCAN_MAX(pm);
			WV(int, fout); WV(int, fin); WV(int, h); WV(int, hp);
			FOR_LANES { L_(fout) = imax(L_(pm) - lane * e_ins, fcar - (lane + 1) * e_ins); }
			wv_shfl_up(fin, fout, 1);
			FOR_LANES {
				if (lane == 0) L_(fin) = fcar;
				int hh = imax(L_(M), L_(e));
				L_(h) = imax(hh, L_(fin));
			}
			wv_shfl_up(hp, h, 1);
			WV(bool, nz); WV(int, hm);
			FOR_LANES {
				if (lane == 0) L_(hp) = h1c;
				bool nzv = false;
				if (L_(act)) {
					int j = j0 + lane;
					int t = L_(M) - oe_del; t = t > 0 ? t : 0;
					int en = imax(L_(e) - e_del, t);
					H[j] = L_(hp); E[j] = en;
					nzv = (L_(hp) != 0) || (en != 0);
				}
				L_(nz) = nzv;
				L_(hm) = L_(act) ? L_(h) : -1;
			}
			int cmx = wv_reduce_max(hm);
			WV(bool, ismax);
			FOR_LANES { L_(ismax) = L_(hm) == cmx; }
			u32 bmax = wv_ballot(ismax), bnz = wv_ballot(nz);
			if (cmx >= m) { m = cmx; mj = j0 + 31 - __clz(bmax); }
			if (bnz) { if (nzfirst < 0) nzfirst = j0 + __ffs(bnz) - 1; nzlast = j0 + 31 - __clz(bnz); }
			int la = end - 1 - j0; la = la < W - 1 ? la : W - 1;
			fcar = wv_get(fout, la); h1c = wv_get(h, la);
		}
		ONE_LANE { H[end] = h1c; E[end] = 0; }
		WARP_SYNC();
		int jfin = beg < end ? end : beg;
		if (jfin == qlen) { max_ie = gscore > h1c ? max_ie : i; gscore = gscore > h1c ? gscore : h1c; }
		if (m == 0) break;
		if (m > max) { max = m; max_i = i; max_j = mj; int off = iabs(mj - i); max_off = max_off > off ? max_off : off; }
		else if (zdrop > 0) {
			if (i - max_i > mj - max_j) { if (max - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop) break; }
			else { if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) break; }
		}
		int beg_new = nzfirst >= 0 ? nzfirst : end;
		int last = h1c != 0 ? end : (nzlast >= 0 ? nzlast : beg_new - 1);
		end = last + 2 < qlen ? last + 2 : qlen;
		beg = beg_new;
	}
	res.score = max; res.qle = max_j + 1; res.tle = max_i + 1; res.gtle = max_ie + 1; res.gscore = gscore; res.max_off = max_off;
	if (cell_counter) { ONE_LANE { atomic_add_u64(cell_counter, cells); } }
}

// Banded global alignment.  z (traceback bytes, n_col per target row) and cigar live in per-warp global scratch.
// Returns the score; *n_cigar_out = 0 when no traceback was requested (cigar == nullptr).
template <int W>
DEVNI int warp_global2(const DevIndex& x, const MemOpt& opt, WarpMem& wm, int qlen, int tlen, TAcc ta, int w,
                       u8* z, u32* cigar, int cigar_cap, u32* err, int* n_cigar_out, u64* cell_counter)
{
	if (W == 1) return thread_global2(x, opt, wm, qlen, tlen, ta, w, z, cigar, cigar_cap, err, n_cigar_out, cell_counter);
	LANE_DECL;
	const int o_del = opt.o_del, e_del = opt.e_del, o_ins = opt.o_ins, e_ins = opt.e_ins;
	const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	i32* H = wm.H; i32* E = wm.E;
	const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
	const bool bt = cigar != nullptr;
	if (n_cigar_out) *n_cigar_out = 0;
	FOR_LANES {
		for (int j = lane; j <= qlen; j += W) {
			int v = j == 0 ? 0 : (j <= w ? -(o_ins + e_ins * j) : KSW_NEG_INF);
			H[j] = v; E[j] = KSW_NEG_INF;
		}
	}
	WARP_SYNC();
	u64 cells = 0;
	WV(int, tch);
	FOR_LANES { L_(tch) = 0; }
	for (int i = 0; i < tlen; ++i) {
		if ((i & (W - 1)) == 0) { FOR_LANES { int ii = i + lane; L_(tch) = ii < tlen ? ref_base2(x, tacc_pos(ta, ii)) : 0; } }
		const int tb = wv_get(tch, i & (W - 1));
		int beg = i > w ? i - w : 0;
		int end = i + w + 1 < qlen ? i + w + 1 : qlen;
		int h1c = beg == 0 ? -(o_del + e_del * (i + 1)) : KSW_NEG_INF;
		int fcar = KSW_NEG_INF;
		if (end > beg) cells += (u64)(end - beg);
		for (int j0 = beg; j0 < end; j0 += W) {
			WV(int, M); WV(int, e); WV(int, pm); WV(bool, act);
			FOR_LANES {
				int j = j0 + lane;
				L_(act) = j < end;
				int hold = L_(act) ? H[j] : KSW_NEG_INF;
				L_(e) = L_(act) ? E[j] : KSW_NEG_INF;
				int mm = hold + (int)opt.mat[tb * 5 + WQ(wm, L_(act) ? j : 0)];
				L_(M) = mm;
				L_(pm) = L_(act) ? (mm - oe_ins) + lane * e_ins : 2 * KSW_NEG_INF + 1;
			}
			WV_SCAN_MAX(pm);
			WV(int, fout); WV(int, fin); WV(int, h); WV(int, hp);
			FOR_LANES { L_(fout) = imax(L_(pm) - lane * e_ins, fcar - (lane + 1) * e_ins); }
			wv_shfl_up(fin, fout, 1);
			WV(int, dd);
			FOR_LANES {
				if (lane == 0) L_(fin) = fcar;
				int mm = L_(M), ee = L_(e), ff = L_(fin);
				int d = mm >= ee ? 0 : 1;
				int hh = mm >= ee ? mm : ee;
				d = hh >= ff ? d : 2;
				hh = hh >= ff ? hh : ff;
				L_(h) = hh; L_(dd) = d;
			}
			wv_shfl_up(hp, h, 1);
			FOR_LANES {
				if (lane == 0) L_(hp) = h1c;
				if (L_(act)) {
					int j = j0 + lane;
					int d = L_(dd);
					int t = L_(M) - oe_del, en = L_(e) - e_del;
					if (en > t) d |= 1 << 2; else en = t;
					int t2 = L_(M) - oe_ins, fn = L_(fin) - e_ins;
					if (fn > t2) d |= 2 << 4;
					H[j] = L_(hp); E[j] = en;
					if (bt) z[(size_t)i * n_col + (j - beg)] = (u8)d;
				}
			}
			int la = end - 1 - j0; la = la < W - 1 ? la : W - 1;
			fcar = wv_get(fout, la); h1c = wv_get(h, la);
		}
		ONE_LANE { H[end] = h1c; E[end] = KSW_NEG_INF; }
		WARP_SYNC();
	}
	int score = H[qlen];
	if (bt) {
		// traceback: sequential, warp-uniform; the op list is bounded by cigar_cap (overflow is reported, never written)
#define CG_PUT() do { if (n_cigar < cigar_cap) { ONE_LANE { cigar[n_cigar] = cur_len << 4 | cur_op; } } else cg_ovf = true; ++n_cigar; } while (0)
		bool cg_ovf = false;
		int n_cigar = 0, which = 0, i = tlen - 1, k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
		u32 cur_op = 0xffffffffu, cur_len = 0;
		// ops are produced back to front; emitted reversed at the end
		while (i >= 0 && k >= 0) {
			which = z[(size_t)i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
			u32 op;
			if (which == 0) { op = 0; --i; --k; }
			else if (which == 1) { op = 2; --i; }
			else { op = 1; --k; }
			if (op == cur_op) ++cur_len;
			else { if (cur_op != 0xffffffffu) { CG_PUT(); } cur_op = op; cur_len = 1; }
		}
		if (i >= 0) { if (cur_op == 2) cur_len += (u32)(i + 1); else { if (cur_op != 0xffffffffu) { CG_PUT(); } cur_op = 2; cur_len = (u32)(i + 1); } }
		if (k >= 0) { if (cur_op == 1) cur_len += (u32)(k + 1); else { if (cur_op != 0xffffffffu) { CG_PUT(); } cur_op = 1; cur_len = (u32)(k + 1); } }
		if (cur_op != 0xffffffffu) { CG_PUT(); }
		WARP_SYNC();
		if (!cg_ovf) { ONE_LANE { for (int a = 0; a < n_cigar >> 1; ++a) { u32 t = cigar[a]; cigar[a] = cigar[n_cigar - 1 - a]; cigar[n_cigar - 1 - a] = t; } } }
		WARP_SYNC();
		if (cg_ovf) { ONE_LANE { atomic_or_u32(err, ERR_CIGAR_OVERFLOW); } n_cigar = 0; }
		*n_cigar_out = n_cigar;
#undef CG_PUT
	}
	if (cell_counter) { ONE_LANE { atomic_add_u64(cell_counter, cells); } }
	return score;
}

// ---- local alignment with striped-kernel semantics ------------------------------------------------
struct SwFwd { int score, te, qe, score2, te2; };

// wm.q must hold the query in DP orientation for columns [0,qlen) ; columns [qlen,plen) are padding.
// brow: per-warp global scratch for the (row maximum,row) list, capacity >= tlen.
template <int W>
DEVNI SwFwd warp_sw_striped(const DevIndex& x, const MemOpt& opt, WarpMem& wm, int lanes_simd, bool is_u8, int qlen, int tlen, TAcc ta,
                            int xtra, u64* brow, u64* cell_counter)
{
	LANE_DECL;
	SwFwd r; r.score = 0; r.te = -1; r.qe = -1; r.score2 = -1; r.te2 = -1;
	const int slen = (qlen + lanes_simd - 1) / lanes_simd, plen = slen * lanes_simd;
	if (slen == 0) return r;
	const int e_del = opt.e_del, e_ins = opt.e_ins, oe_del = opt.o_del + opt.e_del, oe_ins = opt.o_ins + opt.e_ins;
	int shift = 127, mmax = 0;
	for (int k = 0; k < 25; ++k) { if (opt.mat[k] < shift) shift = opt.mat[k]; if (opt.mat[k] > mmax) mmax = opt.mat[k]; }
	shift = (256 - (shift & 0xff)) & 0xff;
	const int minsc = (xtra & KSW_XSUBO) ? (xtra & 0xffff) : 0x10000;
	const int endsc = (xtra & KSW_XSTOP) ? (xtra & 0xffff) : 0x10000;
	i32* H0 = wm.H; i32* H1 = wm.H2; i32* E = wm.E; i32* Hm = wm.Hm;
	FOR_LANES { for (int q = lane; q < plen; q += W) { H0[q] = 0; H1[q] = 0; E[q] = 0; Hm[q] = 0; } }
	WARP_SYNC();
	int gmax = 0, te = -1, n_b = 0;
	u64 cells = 0;
	WV(int, tch);
	FOR_LANES { L_(tch) = 0; }
	for (int i = 0; i < tlen; ++i) {
		if ((i & (W - 1)) == 0) { FOR_LANES { int ii = i + lane; L_(tch) = ii < tlen ? ref_base2(x, tacc_pos(ta, ii)) : 0; } }
		const int tb = wv_get(tch, i & (W - 1));
		int imx = 0, fcar = 0;
		cells += (u64)qlen;
		for (int q0 = 0; q0 < plen; q0 += W) { // pass 1
			WV(int, hd); WV(int, e); WV(int, pm); WV(bool, act); WV(int, head);
			FOR_LANES {
				int q = q0 + lane;
				L_(act) = q < plen;
				int qq = L_(act) ? q : 0;
				int code = WQ(wm, qq);
				int s = code < 5 ? (int)opt.mat[tb * 5 + code] : 0;
				int hh = qq > 0 ? H0[qq - 1] : 0;
				if (is_u8) { hh = hh + s + shift; if (hh > 255) hh = 255; hh -= shift; if (hh < 0) hh = 0; }
				else { hh = hh + s; if (hh > 32767) hh = 32767; }
				int ee = E[qq];
				L_(e) = ee;
				hh = hh > ee ? hh : ee; // max(diag, E); F joins below
				L_(hd) = hh;
				int g = hh - oe_ins; g = g > 0 ? g : 0;
				L_(pm) = L_(act) ? g + lane * e_ins : KSW_NEG_INF;
				L_(head) = (qq - qq % slen) - q0; // lane index of this column's segment head (may be < 0)
			}
			// segmented inclusive max-scan: a lane only takes values from lanes of its own segment
			for (int d_ = 1; d_ < W; d_ <<= 1) {
				WV(int, t_); wv_shfl_up(t_, pm, d_);
				FOR_LANES { if (lane >= d_ && lane - d_ >= L_(head)) L_(pm) = imax(L_(pm), L_(t_)); }
			}
			WV(int, fout); WV(int, fin); WV(int, h);
			FOR_LANES {
				int o = L_(pm) - lane * e_ins;
				if (L_(head) < 0) o = imax(o, fcar - (lane + 1) * e_ins);
				L_(fout) = o;
			}
			wv_shfl_up(fin, fout, 1);
			WV(int, hm);
			FOR_LANES {
				int fi = lane == 0 ? fcar : L_(fin);
				if (lane == L_(head)) fi = 0;
				if (fi < 0) fi = 0;
				int hh = imax(L_(hd), fi);
				L_(h) = hh;
				if (L_(act)) {
					int q = q0 + lane;
					H1[q] = hh;
					int en = L_(e) - e_del; if (en < 0) en = 0;
					int t = hh - oe_del; if (t < 0) t = 0;
					E[q] = en > t ? en : t;
					if (q % slen == slen - 1) { // F leaving this segment = max(F_in - e_ins, h - oe_ins), >= 0
						int fo = fi - e_ins; if (fo < 0) fo = 0;
						int t2 = hh - oe_ins; if (t2 < 0) t2 = 0;
						wm.fend[q / slen] = fo > t2 ? fo : t2;
					}
				}
				L_(hm) = L_(act) ? hh : 0;
			}
			int cmx = wv_reduce_max(hm);
			imx = imx > cmx ? imx : cmx;
			// F leaving the last column of this chunk (for columns of the same segment in the next chunk)
			{
				WV(int, fl);
				FOR_LANES { int fi = lane == 0 ? fcar : L_(fin); if (lane == L_(head)) fi = 0; if (fi < 0) fi = 0;
					int fo = fi - e_ins; if (fo < 0) fo = 0; int t2 = L_(h) - oe_ins; if (t2 < 0) t2 = 0; L_(fl) = fo > t2 ? fo : t2; }
				int la = plen - 1 - q0; la = la < W - 1 ? la : W - 1;
				fcar = wv_get(fl, la);
			}
		}
		WARP_SYNC();
		// pass 2: F carried out of earlier segments
		for (int q0 = 0; q0 < plen; q0 += W) {
			FOR_LANES {
				int q = q0 + lane;
				if (q < plen) {
					int seg = q / slen, carry = 0;
					for (int s2 = 0; s2 < seg; ++s2) { int c = wm.fend[s2] - (q - (s2 + 1) * slen) * e_ins; carry = carry > c ? carry : c; }
					if (carry > H1[q]) H1[q] = carry;
				}
			}
		}
		WARP_SYNC();
		if (imx >= minsc) {
			if (n_b == 0 || (i32)(u32)brow[n_b - 1] + 1 != i) { ONE_LANE { brow[n_b] = (u64)imx << 32 | (u32)i; } ++n_b; WARP_SYNC(); }
			else if ((int)(brow[n_b - 1] >> 32) < imx) { ONE_LANE { brow[n_b - 1] = (u64)imx << 32 | (u32)i; } WARP_SYNC(); }
		}
		bool stop = false;
		if (imx > gmax) {
			gmax = imx; te = i;
			FOR_LANES { for (int q = lane; q < plen; q += W) Hm[q] = H1[q]; }
			if (is_u8) { if (gmax + shift >= 255 || gmax >= endsc) stop = true; }
			else if (gmax >= endsc) stop = true;
		}
		if (stop) break;
		{ i32* t = H1; H1 = H0; H0 = t; }
	}
	WARP_SYNC();
	r.score = is_u8 ? (gmax + shift < 255 ? gmax : 255) : gmax;
	r.te = te;
	if (!is_u8 || r.score != 255) {
		// smallest query column holding the maximum of the snapshot row
		WV(int, best);
		FOR_LANES { int b = -1; for (int q = lane; q < plen; q += W) b = Hm[q] > b ? Hm[q] : b; L_(best) = b; }
		int mxv = wv_reduce_max(best);
		WV(int, qmin);
		FOR_LANES { int qq = 0x7fffffff; for (int q = lane; q < plen; q += W) if (Hm[q] == mxv) { qq = q; break; } L_(qmin) = qq; }
		r.qe = wv_reduce_min(qmin);
		if (n_b > 0) {
			int d = (r.score + mmax - 1) / mmax, low = te - d, high = te + d;
			for (int k = 0; k < n_b; ++k) {
				int e = (i32)(u32)brow[k];
				if ((e < low || e > high) && (int)(brow[k] >> 32) > r.score2) { r.score2 = (int)(brow[k] >> 32); r.te2 = e; }
			}
		}
	}
	if (cell_counter) { ONE_LANE { atomic_add_u64(cell_counter, cells); } }
	return r;
}
