// extend.cuh -- seed extension into alignment regions, CIGAR generation and region de-duplication,
// one warp per read.  Replaces mem_chain2aln / cal_max_gap / mem_sort_dedup_patch / mem_patch_reg /
// bwa_gen_cigar2 of un-vendored lh3/bwa behind
// jnibwa_createAlignments (/root/reference/src/main/c/jnibwa.c:214).
//
// The per-read decision logic (containment tests, band retries, clip-vs-extend) is a sequential state
// machine executed warp-uniformly; every DP it triggers runs across the 32 lanes (ksw.cuh).
#pragma once
#include "ksw.cuh"

struct Reg;
struct WarpScratch { // per-warp global scratch
	u8* z; size_t z_cap;
	u32* cigar; int cigar_cap;
	char* md; int md_cap;
	char* xa; int xa_cap;
	u64* brow; int brow_cap;
	u32* err; // engine-wide error flags
	// region-list scratch of the warp-cooperative de-duplication (warp_sort_dedup_fast): rt_cap entries each
	Reg* rt; i64* rkey; u32* rrank; int rt_cap;
};


// Result of one seed of a heavy read extended ahead of time (extend_thread.cuh, task mode; consumed by warp_chain2aln)
struct ExtRes { i64 rb, re; i32 qb, qe, score, truesc, w, valid; };
struct ExtTask { u32 r, c, t, pad_; }; // read, chain index within the read, seed index within the chain

struct Counters { u64 cells_ext, cells_glb, cells_msw, n_ext, n_glb, n_msw; };

// cal_max_gap: l = max(int(double(qlen*a - o_del) / e_del + 1.), int(double(qlen*a - o_ins) / e_ins + 1.)), clamped to
// [1, 2w].  Integer restatement: for x >= 0 and e > 0, int(double(x) / e + 1.) == x / e + 1 (the quotient of two ints
// below 2^31 is off by far less than 1/e from the true value, so truncation lands on the same integer); for x < 0 the
// value is < 1 and the clamp to >= 1 hides it.  Avoids fp64 divisions (slow on this part) in per-seed loops.
DEV int cal_max_gap(const MemOpt& opt, int qlen)
{
	const int xd = qlen * opt.a - opt.o_del, xi = qlen * opt.a - opt.o_ins;
	const int l_del = xd >= 0 ? (opt.e_del == 1 ? xd + 1 : xd / opt.e_del + 1) : 0;
	const int l_ins = xi >= 0 ? (opt.e_ins == 1 ? xi + 1 : xi / opt.e_ins + 1) : 0;
	int l = l_del > l_ins ? l_del : l_ins;
	l = l > 1 ? l : 1;
	return l < opt.w << 1 ? l : opt.w << 1;
}

// clip a window to the contig holding `mid` (bns_fetch_seq); returns the contig id
DEV int ref_clip_window(const DevIndex& x, i64& beg, i64 mid, i64& end)
{
	int is_rev;
	if (end < beg) { i64 t = beg; beg = end; end = t; }
	int rid = ref_pos2rid(x, ref_depos(x, mid, is_rev));
	i64 far_beg = x.ctg_off[rid], far_end = x.ctg_off[rid + 1];
	if (is_rev) { i64 t = far_beg; far_beg = (x.l_pac << 1) - far_end; far_end = (x.l_pac << 1) - t; }
	beg = beg > far_beg ? beg : far_beg;
	end = end < far_end ? end : far_end;
	return rid;
}

// bwa_gen_cigar2.  query points at the nt4 codes of the read segment [0,l_query) in global memory.
// want_cigar: write cigar words to ws.cigar and the MD text to ws.md (NUL-terminated), set *n_cigar, *NM, *md_len.
// Returns false when the request is rejected (empty / straddling the strand junction).
template <int W>
DEVNI bool warp_gen_cigar2(const DevIndex& x, const MemOpt& opt, WarpMem& wm, WarpScratch& ws, Counters* cnt, int w_,
                           int l_query, const u8* query, i64 rb, i64 re, int* score, bool want_cigar, int* n_cigar, int* NM, int* md_len)
{
	LANE_DECL;
	const i64 l_pac = x.l_pac;
	if (n_cigar) *n_cigar = 0;
	if (NM) *NM = -1;
	if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return false;
	if (l_query > wm.cap) { ONE_LANE { atomic_or_u32(ws.err, ERR_QLEN_OVERFLOW); } return false; }
	const i64 rlen = re - rb;
	const bool rev = rb >= l_pac; // reverse (not complement) both sequences so that indels left-align on the forward strand
	TAcc ta; ta.piv = -1;
	if (rev) { ta.base = re - 1; ta.stride = -1; } else { ta.base = rb; ta.stride = 1; }
	FOR_LANES { for (int j = lane; j < l_query; j += W) WQ(wm, j) = rev ? query[l_query - 1 - j] : query[j]; }
	WARP_SYNC();
	int nc = 0;
	if (l_query == rlen && w_ == 0) { // ungapped: one M run, score by direct summation
		WV(int, part);
		FOR_LANES { int s = 0; for (int j = lane; j < l_query; j += W) s += (int)opt.mat[ref_base2(x, tacc_pos(ta, j)) * 5 + WQ(wm, j)]; L_(part) = s; }
		int tot = 0;
		for (int l = 0; l < W; ++l) tot += wv_get(part, l);
		*score = tot;
		if (want_cigar) { ONE_LANE { ws.cigar[0] = (u32)l_query << 4; } nc = 1; WARP_SYNC(); }
	} else {
		int max_ins = (int)((double)(((l_query + 1) >> 1) * opt.mat[0] - opt.o_ins) / opt.e_ins + 1.);
		int max_del = (int)((double)(((l_query + 1) >> 1) * opt.mat[0] - opt.o_del) / opt.e_del + 1.);
		int max_gap = max_ins > max_del ? max_ins : max_del;
		max_gap = max_gap > 1 ? max_gap : 1;
		int dl = iabs((int)rlen - l_query);
		int w = (max_gap + dl + 1) >> 1;
		w = w < w_ ? w : w_;
		int min_w = dl + 3;
		w = w > min_w ? w : min_w;
		if (want_cigar) {
			int n_col = l_query < 2 * w + 1 ? l_query : 2 * w + 1;
			if ((size_t)(W == 1 ? ((n_col + 3) & ~3) : n_col) * (size_t)rlen > ws.z_cap) { ONE_LANE { atomic_or_u32(ws.err, ERR_Z_OVERFLOW); } return false; }
		}
		if (W == 1) { // thread-level rows: a shared-memory ring when the band fits it; the full rows of the thread's global scratch in kernels without rings
			const bool ring_ok = wm.ringH != nullptr && 2 * w + 2 <= wm.ring_n;
			const bool rows_ok = wm.ringH == nullptr && wm.H != nullptr && l_query <= wm.cap;
			if (!ring_ok && !rows_ok) { atomic_or_u32(ws.err, ERR_NEED_WARP); return false; } // a warp takes the item
		}
		*score = warp_global2<W>(x, opt, wm, l_query, (int)rlen, ta, w, ws.z, want_cigar ? ws.cigar : nullptr, ws.cigar_cap, ws.err, &nc, cnt ? &cnt->cells_glb : nullptr);
		if (cnt) { ONE_LANE { atomic_add_u64(&cnt->n_glb, 1); } }
	}
	DBGP(400, nc);
	if (want_cigar) {
		// MD / NM: sequential walk (warp-uniform), text written by one lane
		int xq = 0, y = 0, u = 0, n_mm = 0, n_gap = 0, ml = 0;
		const char* int2base = rb < l_pac ? "ACGTN" : "TGCAN";
		char* md = ws.md;
		const int cap = ws.md_cap - 16;
		bool ovf = false;
#define MD_PUTC(ch) do { const char c_ = (char)(ch); /* evaluated by every lane: the argument may have side effects */ if (ml < cap) { ONE_LANE_NOSYNC { md[ml] = c_; } ++ml; } else ovf = true; } while (0)
#define MD_PUTNUM(v) do { int v_ = (v); char b_[12]; int n_ = 0; do { b_[n_++] = (char)('0' + v_ % 10); v_ /= 10; } while (v_); while (n_) MD_PUTC(b_[--n_]); } while (0)
		for (int k = 0; k < nc; ++k) {
			u32 cg = ws.cigar[k];
			int op = (int)(cg & 0xf), len = (int)(cg >> 4);
			if (op == 0) {
				for (int i = 0; i < len; ++i) {
					int rbs = ref_base2(x, tacc_pos(ta, y + i));
					if ((int)WQ(wm, xq + i) != rbs) { MD_PUTNUM(u); MD_PUTC(int2base[rbs]); ++n_mm; u = 0; }
					else ++u;
				}
				xq += len; y += len;
			} else if (op == 2) {
				if (k > 0 && k < nc - 1) {
					MD_PUTNUM(u); MD_PUTC('^');
					for (int i = 0; i < len; ++i) MD_PUTC(int2base[ref_base2(x, tacc_pos(ta, y + i))]);
					u = 0; n_gap += len;
				}
				y += len;
			} else if (op == 1) { xq += len; n_gap += len; }
		}
		MD_PUTNUM(u);
		ONE_LANE { md[ml] = 0; }
		WARP_SYNC();
		if (ovf) { ONE_LANE { atomic_or_u32(ws.err, ERR_MD_OVERFLOW); } }
		*NM = n_mm + n_gap; *n_cigar = nc; *md_len = ml;
	}
	return true;
}

template <int W>
DEV int patch_reg(const DevIndex& x, const MemOpt& opt, WarpMem& wm, WarpScratch& ws, Counters* cnt, const u8* query, const Reg& a, const Reg& b, int* w_out)
{
	if (query == nullptr) return 0;
	if (a.rb < x.l_pac && b.rb >= x.l_pac) return 0;
	if (a.qb >= b.qb || a.qe >= b.qe || a.re >= b.re) return 0;
	int w = (int)((a.re - b.rb) - (a.qe - b.qb));
	w = w > 0 ? w : -w;
	double r = (double)(a.re - b.rb) / (double)(b.re - a.rb) - (double)(a.qe - b.qb) / (double)(b.qe - a.qb);
	r = r > 0. ? r : -r;
	if (a.re < b.rb || a.qe < b.qb) { if (w > opt.w << 1 || r >= (double)0.05f) return 0; }
	else if (w > opt.w << 2 || r >= (double)(0.05f * 2)) return 0;
	w += a.w + b.w;
	w = w < opt.w << 2 ? w : opt.w << 2;
	int score = 0;
	warp_gen_cigar2<W>(x, opt, wm, ws, cnt, w, b.qe - a.qb, query + a.qb, a.rb, b.re, &score, false, nullptr, nullptr, nullptr);
	int q_s = (int)((double)(b.qe - a.qb) / (double)((b.qe - b.qb) + (a.qe - a.qb)) * (double)(b.score + a.score) + .499);
	int r_s = (int)((double)(b.re - a.rb) / (double)((b.re - b.rb) + (a.re - a.rb)) * (double)(b.score + a.score) + .499);
	if ((double)score / (double)(q_s > r_s ? q_s : r_s) < (double)0.90f) return 0;
	*w_out = w;
	return score;
}

struct U64Lt { DEVM bool operator()(u64 a, u64 b) const { return a < b; } };
struct RegLtEnd { DEVM bool operator()(const Reg& a, const Reg& b) const { return a.re < b.re; } };
struct RegLtScore { DEVM bool operator()(const Reg& a, const Reg& b) const { return a.score > b.score || (a.score == b.score && (a.rb < b.rb || (a.rb == b.rb && a.qb < b.qb))); } };
struct RegLtHash { DEVM bool operator()(const Reg& a, const Reg& b) const { return a.score > b.score || (a.score == b.score && (a.is_alt < b.is_alt || (a.is_alt == b.is_alt && a.hash < b.hash))); } };


// mem_sort_dedup_patch; query == nullptr disables patching (the mate-rescue call sites)
// Sort keys of a region with the record's position: the scalar sorts below run the reference's (unstable) introsort on
// these 16/24-byte pairs instead of the 96-byte records when per-warp scratch is available.  The comparator reads the
// key alone, so the sequence of comparisons and exchanges -- hence the resulting permutation, ties included -- is the
// one the sort of the records themselves would produce; the records are then moved once.
struct KeyEnd { i64 re; u32 idx; u32 pad_; };
struct KeyEndLt { DEVM bool operator()(const KeyEnd& a, const KeyEnd& b) const { return a.re < b.re; } };
struct KeyScore { i64 rb; i32 score, qb; u32 idx; u32 pad_; };
struct KeyScoreLt { DEVM bool operator()(const KeyScore& a, const KeyScore& b) const { return a.score > b.score || (a.score == b.score && (a.rb < b.rb || (a.rb == b.rb && a.qb < b.qb))); } };

// stage: 0 = the whole procedure; 1 = the list is already sorted by end position (n_comp set); 2 = only the tail
// (compaction of the survivors, order by (score desc, rb, qb), removal of identical hits)
template <int W>
DEVNI int warp_sort_dedup_patch(const DevIndex& x, const MemOpt& opt, WarpMem& wm, WarpScratch& ws, Counters* cnt, const u8* query, int n, Reg* a, int stage = 0)
{
	LANE_DECL;
	if (n <= 1) return n;
	// per-warp scratch ws.rt: first half = record staging, second half = the key array (n <= rt_cap / 2)
	const bool by_key = W == 32 && ws.rt != nullptr && n >= 8 && n <= (ws.rt_cap >> 1);
	if (stage == 0) {
		if (by_key) {
			KeyEnd* ke = (KeyEnd*)(ws.rt + ((size_t)ws.rt_cap >> 1));
			{
				FOR_LANES { for (int i = lane; i < n; i += W) { KeyEnd k; k.re = a[i].re; k.idx = (u32)i; k.pad_ = 0; ke[i] = k; } }
				WARP_SYNC();
				LANE0_ONLY { ks_introsort<KeyEnd, KeyEndLt>((size_t)n, ke, KeyEndLt()); }
				WARP_SYNC();
				FOR_LANES { for (int i = lane; i < n; i += W) { Reg r = a[ke[i].idx]; r.n_comp = 1; ws.rt[i] = r; } }
				WARP_SYNC();
				FOR_LANES { for (int i = lane; i < n; i += W) a[i] = ws.rt[i]; }
			}
		} else { LANE0_ONLY { ks_introsort<Reg, RegLtEnd>((size_t)n, a, RegLtEnd()); for (int i = 0; i < n; ++i) a[i].n_comp = 1; } }
	}
	WARP_SYNC();
	for (int i = 1; i < n && stage < 2; ++i) {
		Reg p = a[i];
		if (p.rid != a[i - 1].rid || p.rb >= a[i - 1].re + opt.max_chain_gap) continue;
		bool p_dirty = false;
		for (int j = i - 1; j >= 0 && p.rid == a[j].rid && p.rb < a[j].re + opt.max_chain_gap; --j) {
			Reg q = a[j];
			if (q.qe == q.qb) continue;
			i64 or_ = q.re - p.rb;
			i64 oq = q.qb < p.qb ? q.qe - p.qb : p.qe - q.qb;
			i64 mr = q.re - q.rb < p.re - p.rb ? q.re - q.rb : p.re - p.rb;
			i64 mq = q.qe - q.qb < p.qe - p.qb ? q.qe - q.qb : p.qe - p.qb;
			int score, w;
			if ((float)or_ > opt.mask_level_redun * (float)mr && (float)oq > opt.mask_level_redun * (float)mq) {
				if (p.score < q.score) { p.qe = p.qb; p_dirty = true; break; }
				else { ONE_LANE { a[j].qe = a[j].qb; } WARP_SYNC(); }
			} else if (q.rb < p.rb && (score = patch_reg<W>(x, opt, wm, ws, cnt, query, q, p, &w)) > 0) {
				p.n_comp += q.n_comp + 1;
				p.seedcov = p.seedcov > q.seedcov ? p.seedcov : q.seedcov;
				p.sub = p.sub > q.sub ? p.sub : q.sub;
				p.csub = p.csub > q.csub ? p.csub : q.csub;
				p.qb = q.qb; p.rb = q.rb;
				p.truesc = p.score = score;
				p.w = w;
				p_dirty = true;
				ONE_LANE { a[j].qb = a[j].qe; } WARP_SYNC();
			}
		}
		if (p_dirty) { ONE_LANE { a[i] = p; } WARP_SYNC(); }
	}
	int m = 0;
	LANE0_ONLY {
		int i, mm;
		for (i = 0, mm = 0; i < n; ++i) if (a[i].qe > a[i].qb) { if (mm != i) a[mm++] = a[i]; else ++mm; }
		int n2 = mm;
		if (by_key) { // same permutation as sorting the records (see KeyScore)
			KeyScore* ks = (KeyScore*)(ws.rt + ((size_t)ws.rt_cap >> 1));
			for (i = 0; i < n2; ++i) { KeyScore k; k.rb = a[i].rb; k.score = a[i].score; k.qb = a[i].qb; k.idx = (u32)i; k.pad_ = 0; ks[i] = k; }
			ks_introsort<KeyScore, KeyScoreLt>((size_t)n2, ks, KeyScoreLt());
			for (i = 0; i < n2; ++i) ws.rt[i] = a[ks[i].idx];
			for (i = 0; i < n2; ++i) a[i] = ws.rt[i];
		} else ks_introsort<Reg, RegLtScore>((size_t)n2, a, RegLtScore());
		for (i = 1; i < n2; ++i) if (a[i].score == a[i - 1].score && a[i].rb == a[i - 1].rb && a[i].qb == a[i - 1].qb) a[i].qe = a[i].qb;
		for (i = 1, mm = 1; i < n2; ++i) if (a[i].qe > a[i].qb) { if (mm != i) a[mm++] = a[i]; else ++mm; }
		m = n2 < 1 ? n2 : mm;
	}
	WARP_SYNC();
	{ WV(int, mv); FOR_LANES { L_(mv) = m; } m = wv_get(mv, 0); }
	return m;
}

// mem_chain2aln for one chain; av[0..*n_av) accumulates the regions of the read
template <int W>
DEVNI void warp_chain2aln(const DevIndex& x, const MemOpt& opt, WarpMem& wm, WarpScratch& ws, Counters* cnt, int l_query, const u8* query,
                          const ChainK& c, const Seed* seeds, u64* srt, Reg* av, int* n_av, int av_cap, const ExtRes* spec = nullptr, bool spec_only = false)
{
	LANE_DECL;
	const i64 l_pac = x.l_pac;
	const int cn = (int)c.n_seeds;
	if (cn == 0) return;
	i64 rmax0 = l_pac << 1, rmax1 = 0;
	for (int i = 0; i < cn; ++i) {
		const Seed& t = seeds[i];
		i64 b = t.rbeg - (t.qbeg + cal_max_gap(opt, t.qbeg));
		i64 e = t.rbeg + t.len + ((l_query - t.qbeg - t.len) + cal_max_gap(opt, l_query - t.qbeg - t.len));
		rmax0 = rmax0 < b ? rmax0 : b;
		rmax1 = rmax1 > e ? rmax1 : e;
	}
	rmax0 = rmax0 > 0 ? rmax0 : 0;
	rmax1 = rmax1 < l_pac << 1 ? rmax1 : l_pac << 1;
	if (rmax0 < l_pac && l_pac < rmax1) { if (seeds[0].rbeg < l_pac) rmax1 = l_pac; else rmax0 = l_pac; }
	ref_clip_window(x, rmax0, seeds[0].rbeg, rmax1);
	// seeds by (score, index) ascending, visited from the top; keys are unique so the order is canonical
	LANE0_ONLY {
		for (int i = 0; i < cn; ++i) srt[i] = (u64)(u32)seeds[i].score << 32 | (u32)i;
		ks_introsort<u64, U64Lt>((size_t)cn, srt, U64Lt());
	}
	WARP_SYNC();
	for (int k = cn - 1; k >= 0; --k) {
		const Seed s = seeds[(u32)srt[k]];
		int n = *n_av, i;
		// inside an earlier region and close to its diagonal?  (only WHETHER some region qualifies matters: lanes share the list)
		bool contained = false;
		{
			WV(bool, hit);
			FOR_LANES {
				bool h = false;
				for (int t = lane; t < n && !h; t += W) {
					const Reg& p = av[t];
					if (s.rbeg < p.rb || s.rbeg + s.len > p.re || s.qbeg < p.qb || s.qbeg + s.len > p.qe) continue;
					if ((double)(s.len - p.seedlen0) > .1 * (double)l_query) continue;
					int qd = s.qbeg - p.qb; i64 rd = s.rbeg - p.rb;
					int max_gap = cal_max_gap(opt, qd < rd ? qd : (int)rd);
					int w = max_gap < p.w ? max_gap : p.w;
					if (qd - rd < w && rd - qd < w) { h = true; break; }
					qd = p.qe - (s.qbeg + s.len); rd = p.re - (s.rbeg + s.len);
					max_gap = cal_max_gap(opt, qd < rd ? qd : (int)rd);
					w = max_gap < p.w ? max_gap : p.w;
					if (qd - rd < w && rd - qd < w) { h = true; break; }
				}
				L_(hit) = h;
			}
			contained = wv_ballot(hit) != 0u;
		}
		if (contained) { // extend anyway only if an overlapping seed lies on another diagonal
			WV(bool, other);
			FOR_LANES {
				bool h = false;
				for (int t = k + 1 + lane; t < cn && !h; t += W) {
					if (srt[t] == 0) continue;
					const Seed& u = seeds[(u32)srt[t]];
					if ((double)u.len < (double)s.len * .95) continue;
					if (s.qbeg <= u.qbeg && s.qbeg + s.len - u.qbeg >= s.len >> 2 && u.qbeg - s.qbeg != u.rbeg - s.rbeg) h = true;
					else if (u.qbeg <= s.qbeg && u.qbeg + u.len - s.qbeg >= s.len >> 2 && s.qbeg - u.qbeg != s.rbeg - u.rbeg) h = true;
				}
				L_(other) = h;
			}
			if (wv_ballot(other) == 0u) { ONE_LANE { srt[k] = 0; } WARP_SYNC(); continue; }
		}
		if (n >= av_cap) { ONE_LANE { atomic_or_u32(ws.err, ERR_REG_OVERFLOW); } return; }
		Reg a;
		memset(&a, 0, sizeof(Reg));
		int aw0 = opt.w, aw1 = opt.w;
		a.w = opt.w; a.score = a.truesc = -1; a.rid = c.rid;
		const ExtRes* pre = spec ? spec + (u32)srt[k] : nullptr;
		if (pre && pre->valid == 1) { // extended ahead of time by the task kernel (extend_thread.cuh)
			a.rb = pre->rb; a.re = pre->re; a.qb = pre->qb; a.qe = pre->qe; a.score = pre->score; a.truesc = pre->truesc;
			aw0 = aw1 = pre->w;
		} else if (spec_only) { // the caller cannot extend (replay by a thread): the read goes to the warp kernel
			ONE_LANE { atomic_or_u32(ws.err, ERR_NEED_WARP); }
			return;
		} else {
		if (s.qbeg) { // left extension: both sequences reversed
			FOR_LANES { for (int j = lane; j < s.qbeg; j += W) WQ(wm, j) = query[s.qbeg - 1 - j]; }
			WARP_SYNC();
			TAcc ta; ta.base = s.rbeg - 1; ta.stride = -1; ta.piv = -1;
			int tlen = (int)(s.rbeg - rmax0);
			ExtResult er; er.score = -1; er.qle = er.tle = er.gtle = 0; er.gscore = -1; er.max_off = 0;
			for (int t = 0; t < 2; ++t) {
				int prev = a.score;
				aw0 = opt.w << t;
				warp_extend2<W>(x, opt, wm, s.qbeg, tlen, ta, aw0, opt.pen_clip5, opt.zdrop, s.len * opt.a, er, cnt ? &cnt->cells_ext : nullptr);
				if (cnt) { ONE_LANE { atomic_add_u64(&cnt->n_ext, 1); } }
				a.score = er.score;
				if (a.score == prev || er.max_off < (aw0 >> 1) + (aw0 >> 2)) break;
			}
			if (er.gscore <= 0 || er.gscore <= a.score - opt.pen_clip5) { a.qb = s.qbeg - er.qle; a.rb = s.rbeg - er.tle; a.truesc = a.score; }
			else { a.qb = 0; a.rb = s.rbeg - er.gtle; a.truesc = er.gscore; }
		} else { a.score = a.truesc = s.len * opt.a; a.qb = 0; a.rb = s.rbeg; }
		if (s.qbeg + s.len != l_query) { // right extension
			int qe = s.qbeg + s.len, sc0 = a.score;
			i64 re = s.rbeg + s.len;
			FOR_LANES { for (int j = lane; j < l_query - qe; j += W) WQ(wm, j) = query[qe + j]; }
			WARP_SYNC();
			TAcc ta; ta.base = re; ta.stride = 1; ta.piv = -1;
			int tlen = (int)(rmax1 - re);
			ExtResult er; er.score = -1; er.qle = er.tle = er.gtle = 0; er.gscore = -1; er.max_off = 0;
			for (int t = 0; t < 2; ++t) {
				int prev = a.score;
				aw1 = opt.w << t;
				warp_extend2<W>(x, opt, wm, l_query - qe, tlen, ta, aw1, opt.pen_clip3, opt.zdrop, sc0, er, cnt ? &cnt->cells_ext : nullptr);
				if (cnt) { ONE_LANE { atomic_add_u64(&cnt->n_ext, 1); } }
				a.score = er.score;
				if (a.score == prev || er.max_off < (aw1 >> 1) + (aw1 >> 2)) break;
			}
			if (er.gscore <= 0 || er.gscore <= a.score - opt.pen_clip3) { a.qe = qe + er.qle; a.re = re + er.tle; a.truesc += a.score - sc0; }
			else { a.qe = l_query; a.re = re + er.gtle; a.truesc += er.gscore - sc0; }
		} else { a.qe = l_query; a.re = s.rbeg + s.len; }
		}
		a.seedcov = 0;
		for (i = 0; i < cn; ++i) {
			const Seed& t = seeds[i];
			if (t.qbeg >= a.qb && t.qbeg + t.len <= a.qe && t.rbeg >= a.rb && t.rbeg + t.len <= a.re) a.seedcov += t.len;
		}
		a.w = aw0 > aw1 ? aw0 : aw1;
		a.seedlen0 = s.len;
		a.frac_rep = c.frac_rep;
		ONE_LANE { av[n] = a; }
		*n_av = n + 1;
		WARP_SYNC();
	}
}

// ---- warp-cooperative mem_sort_dedup_patch for long region lists (reads from repeats: hundreds of hits) -------------
// The scalar procedure is two sorts around a sequential scan; run by one lane it costs ~n log n dependent steps on
// 96-byte records, and repeat-derived reads call it with n ~ 500 once per extension and once per rescued hit.  Here:
//   * both orders are computed by rank (n^2 / 32 comparisons per lane on a compact key array) and applied with one
//     scatter.  When all keys are distinct every correct sort yields the same array, so the unstable introsort of
//     the reference need not be imitated; EQUAL keys are detected while ranking and send the list to the scalar
//     path (at that stage only), which reproduces the reference's tie order;
//   * the redundancy scan only relates hits that lie within max_chain_gap of each other on one contig: the list,
//     sorted by end position, falls into independent segments (repeat copies sit at unrelated loci, so segments are
//     short) that the lanes scan concurrently, each with the scalar logic;
//   * region patching (a banded global alignment between two colinear hits) stays on the scalar path: a dry pass over
//     all pairs that could be examined detects candidates first.
#define SDP_FAST_MIN 24
#ifdef B200_SIM
static long g_sim_sdp[8]; // kernel-debug harness only: calls, tie fallbacks (end), patch-candidate fallbacks, tie fallbacks (score), completed
#define SDP_STAT(k) (++g_sim_sdp[k])
#else
#define SDP_STAT(k) ((void)0)
#endif
DEV bool patch_precheck(const DevIndex& x, const MemOpt& opt, const Reg& a, const Reg& b)
{ // the tests of patch_reg that precede its alignment
	if (a.rb < x.l_pac && b.rb >= x.l_pac) return false;
	if (a.qb >= b.qb || a.qe >= b.qe || a.re >= b.re) return false;
	int w = (int)((a.re - b.rb) - (a.qe - b.qb));
	w = w > 0 ? w : -w;
	double r = (double)(a.re - b.rb) / (double)(b.re - a.rb) - (double)(a.qe - b.qb) / (double)(b.qe - a.qb);
	r = r > 0. ? r : -r;
	if (a.re < b.rb || a.qe < b.qb) { if (w > opt.w << 1 || r >= (double)0.05f) return false; }
	else if (w > opt.w << 2 || r >= (double)(0.05f * 2)) return false;
	return true;
}

// *fixed_point (optional) is set when the call went through without any tie or patch candidate AND removed nothing:
// the list is then in its final order with all sort keys distinct and no redundant pair in scan range, so a further
// call on the unchanged list would return it unchanged (same end-position order => every pair is examined in the
// same orientation and found non-redundant again; same score order).  Callers use it to skip such calls.
template <int W>
DEVNI int warp_sort_dedup_fast(const DevIndex& x, const MemOpt& opt, WarpMem& wm, WarpScratch& ws, Counters* cnt, const u8* query, int n, Reg* a, bool* fixed_point = nullptr)
{
	if (fixed_point) *fixed_point = false;
	if (W != 32 || n < SDP_FAST_MIN || n > ws.rt_cap || ws.rt == nullptr) return warp_sort_dedup_patch<W>(x, opt, wm, ws, cnt, query, n, a);
	LANE_DECL;
	SDP_STAT(0);
	i64* key = ws.rkey; u32* rk = ws.rrank; Reg* t = ws.rt;
	// ---- order by end position
	FOR_LANES { for (int i = lane; i < n; i += W) key[i] = a[i].re; }
	WARP_SYNC();
	WV(bool, tie);
	FOR_LANES {
		bool tv = false;
		for (int i = lane; i < n; i += W) {
			const i64 ki = key[i];
			u32 less = 0;
			for (int j = 0; j < n; ++j) { const i64 kj = key[j]; less += kj < ki ? 1u : 0u; tv |= (kj == ki) & (j != i); }
			rk[i] = less;
		}
		L_(tie) = tv;
	}
	if (wv_ballot(tie) != 0u) { SDP_STAT(1); return warp_sort_dedup_patch<W>(x, opt, wm, ws, cnt, query, n, a); } // equal keys: the scalar sort decides
	FOR_LANES { for (int i = lane; i < n; i += W) { Reg r = a[i]; r.n_comp = 1; t[rk[i]] = r; } }
	WARP_SYNC();
	// ---- reach of every hit (lowest index its scan can touch) and patch candidates
	WV(bool, cand);
	FOR_LANES {
		bool cv = false;
		for (int i = lane; i < n; i += W) {
			const Reg p = t[i];
			int lo = i;
			for (int j = i - 1; j >= 0 && p.rid == t[j].rid && p.rb < t[j].re + opt.max_chain_gap; --j) {
				lo = j;
				if (query != nullptr && t[j].rb < p.rb && patch_precheck(x, opt, t[j], p)) cv = true;
			}
			rk[i] = (u32)lo;
		}
		L_(cand) = cv;
	}
	const bool any_cand = wv_ballot(cand) != 0u;
	FOR_LANES { for (int i = lane; i < n; i += W) a[i] = t[i]; } // a: sorted by end position, n_comp = 1
	WARP_SYNC();
	if (any_cand) { SDP_STAT(2); return warp_sort_dedup_patch<W>(x, opt, wm, ws, cnt, query, n, a, 1); }
	// ---- segments: a boundary lies before i when no hit at or after i reaches below i (suffix minimum of the reaches)
	int n_seg = 0;
	u32* seg = (u32*)key; // segment start indices, ascending after the reversal below
	LANE0_ONLY {
		u32 mn = (u32)n;
		int m = 0;
		for (int i = n - 1; i >= 0; --i) { mn = rk[i] < mn ? rk[i] : mn; if (mn >= (u32)i) seg[m++] = (u32)i; }
		for (int b = 0, e = m - 1; b < e; ++b, --e) { u32 tt = seg[b]; seg[b] = seg[e]; seg[e] = tt; }
		seg[m] = (u32)n;
		rk[0] = (u32)m; // (reaches are no longer needed)
	}
	WARP_SYNC();
	{ WV(int, mv); FOR_LANES { L_(mv) = (int)rk[0]; } n_seg = wv_get(mv, 0); }
	// ---- the redundancy scan, one lane per segment (scalar logic, no patching: no candidate exists)
	FOR_LANES {
		for (int sg = lane; sg < n_seg; sg += W) {
			const int b = (int)seg[sg], e = (int)seg[sg + 1];
			for (int i = b + 1; i < e; ++i) {
				Reg p = a[i];
				if (p.rid != a[i - 1].rid || p.rb >= a[i - 1].re + opt.max_chain_gap) continue;
				for (int j = i - 1; j >= b && p.rid == a[j].rid && p.rb < a[j].re + opt.max_chain_gap; --j) {
					const Reg q = a[j];
					if (q.qe == q.qb) continue;
					i64 or_ = q.re - p.rb;
					i64 oq = q.qb < p.qb ? q.qe - p.qb : p.qe - q.qb;
					i64 mr = q.re - q.rb < p.re - p.rb ? q.re - q.rb : p.re - p.rb;
					i64 mq = q.qe - q.qb < p.qe - p.qb ? q.qe - q.qb : p.qe - p.qb;
					if ((float)or_ > opt.mask_level_redun * (float)mr && (float)oq > opt.mask_level_redun * (float)mq) {
						if (p.score < q.score) { a[i].qe = a[i].qb; break; }
						else a[j].qe = a[j].qb;
					}
				}
			}
		}
	}
	WARP_SYNC();
	// ---- survivors, compacted in order into t
	int n2 = 0;
	for (int i0 = 0; i0 < n; i0 += W) {
		WV(bool, keep);
		FOR_LANES { const int i = i0 + lane; L_(keep) = i < n && a[i].qe > a[i].qb; }
		const u32 bm = wv_ballot(keep);
		FOR_LANES { const int i = i0 + lane; if (i < n && ((bm >> lane) & 1u)) t[n2 + __popc(bm & ((1u << lane) - 1u))] = a[i]; }
		n2 += __popc(bm);
	}
	WARP_SYNC();
	// ---- order by (score desc, rb, qb); identical triples are the reference's duplicates -> scalar tail decides
	FOR_LANES {
		bool tv = false;
		for (int i = lane; i < n2; i += W) {
			const Reg ri = t[i];
			u32 less = 0;
			for (int j = 0; j < n2; ++j) {
				const Reg& rj = t[j];
				const bool lt = rj.score > ri.score || (rj.score == ri.score && (rj.rb < ri.rb || (rj.rb == ri.rb && rj.qb < ri.qb)));
				less += lt ? 1u : 0u;
				tv |= (rj.score == ri.score) & (rj.rb == ri.rb) & (rj.qb == ri.qb) & (j != i);
			}
			rk[i] = less;
		}
		L_(tie) = tv;
	}
	if (wv_ballot(tie) != 0u) { // (a still holds the scanned list: the scalar tail redoes the compaction itself)
		SDP_STAT(3);
		return warp_sort_dedup_patch<W>(x, opt, wm, ws, cnt, query, n, a, 2);
	}
	FOR_LANES { for (int i = lane; i < n2; i += W) a[rk[i]] = t[i]; }
	WARP_SYNC();
	SDP_STAT(4);
	if (fixed_point) *fixed_point = n2 == n; // (a removed hit may have ended scans that would now run on: not a fixed point)
	return n2;
}
