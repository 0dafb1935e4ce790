// extend_thread.cuh -- seed extension of one read by ONE thread (mem_chain2aln + ksw_extend2 of un-vendored lh3/bwa,
// SURVEY.md App. B.6 / B.7, behind /root/reference/src/main/c/jnibwa.c:214).
//
// Most reads have a handful of kept seeds and two short extensions (~5-7 k DP cells), far too little to keep the 32
// lanes of a warp busy row by row.  Here every lane owns a read and the warp runs 32 independent scalar DPs:
//   * the DP row lives in shared memory, one u32 per query column (H in the low half-word, E in the high one; both are
//     non-negative in ksw_extend2), interleaved over the threads of the block so that column j of every thread sits in
//     bank tid % 32 -- one conflict-free LDS + STS per cell;
//   * the read's base codes live in shared memory too (same interleaving); the left extension walks them backwards, the
//     right extension forwards, so no reversed copy is made;
//   * the per-read control flow (chains, seeds by score, containment test, left / right extension, band retry) is a
//     state machine around ONE DP-row site: each trip of the outer loop advances the machine (lane-divergent, cheap)
//     to the next row that has to be computed, then all lanes compute up to P.tx_chunk cells of their row together (a
//     longer row continues on the next trip).  Lanes therefore stay in step at chunk granularity no matter how their
//     extensions differ in length and width;
//   * reads are not assigned statically: a lane that finishes its read takes the next one from a batch-wide counter
//     (fetching is one more state of the machine), so no lane idles while a neighbour works through a long read; a
//     warp vote at the top of every trip is both the exit test and the reconvergence point.
// Exactness: the row code is the scalar recurrence of ksw_extend2 verbatim (band clamp, [beg,end) shrink, z-drop,
// tie rules); the only representational change is 16-bit storage, guarded by a bound on the largest reachable score
// (reads that could exceed it, or do not fit the shared rows, are handed to the warp-per-read kernel).
#pragma once
#include "extend.cuh"

struct ThreadExtMem {
	void* row; int row_stride; int row_cap;  // DP row: cell j at ((CELL*)row)[j * row_stride], j in [0, row_cap)
	u8* rd; int rd_stride; int rd_cap;       // read codes: element j at rd[j * rd_stride]
	const u64* rowpack;                      // [5]: byte c of rowpack[t] = mat[t*5 + c]
};

enum { TX_FETCH, TX_NEXT_CHAIN, TX_NEXT_SEED, TX_EXT_START, TX_EXT_DONE, TX_RIGHT_CHECK, TX_FINISH_SEED, TX_READ_DONE };

DEV bool tx_all_done(bool done) { return lanes_all(done); }

// Extends reads taken from `counter` until the batch is exhausted.  Every lane of the warp must call this (the loop
// votes with the full mask).  Reads with many kept seeds, and reads that do not fit the thread-level rows, are queued
// on the heavy list for the warp-per-read kernel.
//
// TASKS = true: the work items are single seeds of heavy reads (ExtTask list written by wbody_heavy_prep): a read from
// a repeat family has hundreds of chains, and extending them one after the other in one thread or warp leaves the
// rest of the GPU waiting for it.  Every seed is extended unconditionally (its result does not depend on the hits
// found so far; only the decision to USE it does -- the containment test is replayed by wbody_extend_heavy).
template <bool TASKS, class CELL>
DEV void thread_extend_queue(const BatchParams& P, i64 n_reads, u32* counter, const ThreadExtMem& tm, WarpMem& wm, WarpScratch& ws, u32* lerr)
{
	const DevIndex& x = P.x; const MemOpt& opt = P.opt; Counters* cnt = P.cnt; u32* err = P.err;
	const i64 l_pac = x.l_pac;
	const int o_del = opt.o_del, e_del = opt.e_del, o_ins = opt.o_ins, e_ins = opt.e_ins;
	const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	int mx = 0;
	for (int k = 0; k < 25; ++k) mx = mx > opt.mat[k] ? mx : (int)opt.mat[k];
	const size_t RS = (size_t)tm.row_stride, QS = (size_t)tm.rd_stride;
	CELL* const row = (CELL*)tm.row;
	constexpr int HB = (int)sizeof(CELL) * 4;          // bits of H and of E in one cell
	constexpr u32 HMASK = (1u << HB) - 1u;
	// largest score a cell may have to hold: h0 + qlen * max match <= l_query * max(match, a) -- an extension starts from at
	// most (bases before it) * a and gains at most max match per base; the clipping bonus is only ever compared, never stored
	constexpr i64 SCORE_CAP = sizeof(CELL) == 2 ? 255 : 60000;
	const int chunk = P.tx_chunk > 0 ? P.tx_chunk : 1 << 30;
	// per-read state
	i64 r = -1; int l_query = 0, nck = 0, av_cap = 0;
	const u8* query = nullptr; const ChainK* cks = nullptr; const Seed* ks_base = nullptr; u64* srt = nullptr; Reg* av = nullptr;

	int phase = TX_FETCH, n_av = 0;
	bool in_ext = false, done = false;
	u64 res_slot = 0; // task mode: where the result goes
#define TX_TO_WARP() do { if (TASKS) P.ext_res[res_slot].valid = 0; /* the replay extends this seed itself */ \
		else { u32 slot_ = atomic_add_u32(P.n_heavy, 1u); P.heavy_list[slot_] = (u32)r; } phase = TX_FETCH; } while (0)
	int c = -1, k = -1, cn = 0;
	const Seed* seeds = nullptr; ChainK ck; ck.n_seeds = 0; ck.seed_start = 0; ck.rid = 0; ck.frac_rep = 0.f;
	i64 rmax0 = 0, rmax1 = 0;
	Seed s; s.rbeg = 0; s.qbeg = s.len = s.score = s.next = 0;
	Reg a; memset(&a, 0, sizeof(Reg));
	int aw0 = 0, aw1 = 0, side = 0, tr = 0, prev = 0, sc0 = 0;
	// extension in flight
	int qlen = 0, tlen = 0, w = 0, end_bonus = 0, h0 = 0, qbase = 0, qdir = 1;
	i64 tbase = 0; int tdir = 1;
	int i = 0, beg = 0, end = 0, max = 0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, max_off = 0;
	u64 cells = 0, n_ext = 0;
	u32 pac_word = 0; i64 pac_widx = -1;
	// row in flight (a row may span several visits of the DP site)
	int jc = -1, h1 = 0, f = 0, m = 0, mj = -1; u64 rp = 0;

	for (;;) {
		// ---- advance to the next DP row (or to the end of the read)
		while (!in_ext && !done) {
			switch (phase) {
			case TX_FETCH: {
				if (TASKS) {
					const i64 tid = (i64)atomic_add_u32(counter, 1u);
					if (tid >= n_reads) { done = true; break; }
					const ExtTask tk = P.ext_tasks[tid];
					r = (i64)tk.r;
					const u64 t0 = P.seed_off[r];
					ck = P.ck[t0 + tk.c];
					res_slot = t0 + ck.seed_start + tk.t;
					l_query = (int)(P.seq_off[r + 1] - P.seq_off[r] - 1);
					query = P.seq + P.seq_off[r];
					if (l_query > tm.rd_cap || (i64)(sizeof(CELL) == 2 ? 1 : 2) * l_query * (mx > opt.a ? mx : opt.a) > SCORE_CAP) { TX_TO_WARP(); break; }
					for (int j = 0; j < l_query; ++j) tm.rd[(size_t)j * QS] = query[j];
					seeds = P.ks + t0 + ck.seed_start; cn = 0; // (seed coverage is computed by the replay)
					s = seeds[tk.t];
					rmax0 = P.ck_rmax[2 * (t0 + tk.c)]; rmax1 = P.ck_rmax[2 * (t0 + tk.c) + 1];
					memset(&a, 0, sizeof(Reg));
					aw0 = aw1 = opt.w;
					a.w = opt.w; a.score = a.truesc = -1; a.rid = ck.rid;
					if (s.qbeg) {
						side = 0; tr = 0;
						qlen = s.qbeg; qbase = s.qbeg - 1; qdir = -1;
						tbase = s.rbeg - 1; tdir = -1; tlen = (int)(s.rbeg - rmax0);
						end_bonus = opt.pen_clip5; h0 = s.len * opt.a;
						phase = TX_EXT_START;
					} else { a.score = a.truesc = s.len * opt.a; a.qb = 0; a.rb = s.rbeg; phase = TX_RIGHT_CHECK; }
					break;
				}
				r = (i64)atomic_add_u32(counter, 1u);
				if (r >= n_reads) { done = true; break; }
				nck = (int)P.n_ck[r];
				const int nks = (int)P.n_ks[r];
				if (nck == 0) { P.n_reg[r] = 0; break; }
				if (nks > P.light_max_seeds || nck > P.light_max_chains) { TX_TO_WARP(); break; }
				const u64 t0 = P.seed_off[r];
				l_query = (int)(P.seq_off[r + 1] - P.seq_off[r] - 1);
				query = P.seq + P.seq_off[r];
				// storage bound: no H/E value exceeds (seed + left + right) scores <= l_query * max(match, a)
				if (l_query > tm.rd_cap || (i64)(sizeof(CELL) == 2 ? 1 : 2) * l_query * (mx > opt.a ? mx : opt.a) > SCORE_CAP) { TX_TO_WARP(); break; }
				cks = P.ck + t0; ks_base = P.ks + t0; srt = P.srt + t0; av = P.regs_tmp + t0; av_cap = nks;
				for (int j = 0; j < l_query; ++j) tm.rd[(size_t)j * QS] = query[j];
				c = -1; n_av = 0; *lerr = 0;
				phase = TX_NEXT_CHAIN;
				break; }
			case TX_READ_DONE: { // region de-duplication / patching of the read, then the next read
				int n2 = warp_sort_dedup_patch<1>(x, opt, wm, ws, cnt, query, n_av, av);
				if (err_peek(lerr) & ERR_NEED_WARP) { TX_TO_WARP(); break; } // recomputed from scratch by a warp
				if (err_peek(lerr)) atomic_or_u32(err, err_peek(lerr));
				P.n_reg[r] = (u32)n2;
				phase = TX_FETCH;
				break; }
			case TX_NEXT_CHAIN: {
				++c;
				if (c >= nck) { phase = TX_READ_DONE; break; }
				ck = cks[c];
				cn = (int)ck.n_seeds;
				if (cn == 0) break;
				seeds = ks_base + ck.seed_start;
				rmax0 = l_pac << 1; rmax1 = 0;
				for (int t = 0; t < cn; ++t) {
					const Seed& sd = seeds[t];
					i64 b = sd.rbeg - (sd.qbeg + cal_max_gap(opt, sd.qbeg));
					i64 e = sd.rbeg + sd.len + ((l_query - sd.qbeg - sd.len) + cal_max_gap(opt, l_query - sd.qbeg - sd.len));
					rmax0 = rmax0 < b ? rmax0 : b;
					rmax1 = rmax1 > e ? rmax1 : e;
				}
				rmax0 = rmax0 > 0 ? rmax0 : 0;
				rmax1 = rmax1 < l_pac << 1 ? rmax1 : l_pac << 1;
				if (rmax0 < l_pac && l_pac < rmax1) { if (seeds[0].rbeg < l_pac) rmax1 = l_pac; else rmax0 = l_pac; }
				ref_clip_window(x, rmax0, seeds[0].rbeg, rmax1);
				for (int t = 0; t < cn; ++t) srt[t] = (u64)(u32)seeds[t].score << 32 | (u32)t;
				ks_introsort<u64, U64Lt>((size_t)cn, srt, U64Lt());
				k = cn;
				phase = TX_NEXT_SEED;
				break; }
			case TX_NEXT_SEED: {
				--k;
				if (k < 0) { phase = TX_NEXT_CHAIN; break; }
				s = seeds[(u32)srt[k]];
				int t;
				for (t = 0; t < n_av; ++t) { // inside an earlier region and close to its diagonal?
					const Reg& p = av[t];
					if (s.rbeg < p.rb || s.rbeg + s.len > p.re || s.qbeg < p.qb || s.qbeg + s.len > p.qe) continue;
					if ((double)(s.len - p.seedlen0) > .1 * (double)l_query) continue;
					int qd = s.qbeg - p.qb; i64 rd = s.rbeg - p.rb;
					int max_gap = cal_max_gap(opt, qd < rd ? qd : (int)rd);
					int ww = max_gap < p.w ? max_gap : p.w;
					if (qd - rd < ww && rd - qd < ww) break;
					qd = p.qe - (s.qbeg + s.len); rd = p.re - (s.rbeg + s.len);
					max_gap = cal_max_gap(opt, qd < rd ? qd : (int)rd);
					ww = max_gap < p.w ? max_gap : p.w;
					if (qd - rd < ww && rd - qd < ww) break;
				}
				if (t < n_av) { // contained: extend anyway only if an overlapping seed lies on another diagonal
					for (t = k + 1; t < cn; ++t) {
						if (srt[t] == 0) continue;
						const Seed& u = seeds[(u32)srt[t]];
						if ((double)u.len < (double)s.len * .95) continue;
						if (s.qbeg <= u.qbeg && s.qbeg + s.len - u.qbeg >= s.len >> 2 && u.qbeg - s.qbeg != u.rbeg - s.rbeg) break;
						if (u.qbeg <= s.qbeg && u.qbeg + u.len - s.qbeg >= s.len >> 2 && s.qbeg - u.qbeg != s.rbeg - u.rbeg) break;
					}
					if (t == cn) { srt[k] = 0; break; }
				}
				if (n_av >= av_cap) { atomic_or_u32(err, ERR_REG_OVERFLOW); P.n_reg[r] = 0; phase = TX_FETCH; break; }
				memset(&a, 0, sizeof(Reg));
				aw0 = aw1 = opt.w;
				a.w = opt.w; a.score = a.truesc = -1; a.rid = ck.rid;
				if (s.qbeg) { // left extension: both sequences walked backwards
					side = 0; tr = 0;
					qlen = s.qbeg; qbase = s.qbeg - 1; qdir = -1;
					tbase = s.rbeg - 1; tdir = -1; tlen = (int)(s.rbeg - rmax0);
					end_bonus = opt.pen_clip5; h0 = s.len * opt.a;
					phase = TX_EXT_START;
				} else { a.score = a.truesc = s.len * opt.a; a.qb = 0; a.rb = s.rbeg; phase = TX_RIGHT_CHECK; }
				break; }
			case TX_RIGHT_CHECK:
				if (s.qbeg + s.len != l_query) {
					side = 1; tr = 0; sc0 = a.score;
					const int qe = s.qbeg + s.len;
					qlen = l_query - qe; qbase = qe; qdir = 1;
					tbase = s.rbeg + s.len; tdir = 1; tlen = (int)(rmax1 - tbase);
					end_bonus = opt.pen_clip3; h0 = sc0;
					phase = TX_EXT_START;
				} else { a.qe = l_query; a.re = s.rbeg + s.len; phase = TX_FINISH_SEED; }
				break;
			case TX_EXT_START: { // one band try of ksw_extend2: first row of H/E, band clamp, running maxima
				prev = a.score;
				w = opt.w << tr;
				if (side == 0) aw0 = w; else aw1 = w;
				if (qlen + 1 > tm.row_cap) { TX_TO_WARP(); break; }
				{
					CELL* rw = row;
					rw[0] = (CELL)h0;
					int v = h0 > oe_ins ? h0 - oe_ins : 0;
					for (int j = 1; j <= qlen; ++j) { rw[(size_t)j * RS] = (CELL)v; v = v > e_ins ? v - e_ins : 0; }
				}
				// band clamp of ksw_extend2, int(double(x) / e + 1.) clamped to >= 1, in integers (see cal_max_gap)
				const int xi_ = qlen * mx + end_bonus - o_ins, xd_ = qlen * mx + end_bonus - o_del;
				int max_ins = xi_ >= 0 ? (e_ins == 1 ? xi_ + 1 : xi_ / e_ins + 1) : 0; max_ins = max_ins > 1 ? max_ins : 1;
				w = w < max_ins ? w : max_ins;
				int max_del = xd_ >= 0 ? (e_del == 1 ? xd_ + 1 : xd_ / e_del + 1) : 0; max_del = max_del > 1 ? max_del : 1;
				w = w < max_del ? w : max_del;
				max = h0; max_i = -1; max_j = -1; max_ie = -1; gscore = -1; max_off = 0; beg = 0; end = qlen; i = 0;
				++n_ext;
				if (tlen > 0) in_ext = true; else phase = TX_EXT_DONE;
				break; }
			case TX_EXT_DONE: { // consume the result of the try
				const int score = max, qle = max_j + 1, tle = max_i + 1, gtle = max_ie + 1;
				const int wt = opt.w << tr;
				a.score = score;
				const bool more = !(a.score == prev || max_off < (wt >> 1) + (wt >> 2)) && tr == 0;
				if (more) { tr = 1; phase = TX_EXT_START; break; }
				if (side == 0) {
					if (gscore <= 0 || gscore <= a.score - opt.pen_clip5) { a.qb = s.qbeg - qle; a.rb = s.rbeg - tle; a.truesc = a.score; }
					else { a.qb = 0; a.rb = s.rbeg - gtle; a.truesc = gscore; }
					phase = TX_RIGHT_CHECK;
				} else {
					const int qe = s.qbeg + s.len; const i64 re = s.rbeg + s.len;
					if (gscore <= 0 || gscore <= a.score - opt.pen_clip3) { a.qe = qe + qle; a.re = re + tle; a.truesc += a.score - sc0; }
					else { a.qe = l_query; a.re = re + gtle; a.truesc += gscore - sc0; }
					phase = TX_FINISH_SEED;
				}
				break; }
			case TX_FINISH_SEED: {
				if (TASKS) {
					ExtRes o; o.rb = a.rb; o.re = a.re; o.qb = a.qb; o.qe = a.qe; o.score = a.score; o.truesc = a.truesc; o.w = aw0 > aw1 ? aw0 : aw1; o.valid = 1;
					P.ext_res[res_slot] = o;
					phase = TX_FETCH;
					break;
				}
				a.seedcov = 0;
				for (int t = 0; t < cn; ++t) {
					const Seed& u = seeds[t];
					if (u.qbeg >= a.qb && u.qbeg + u.len <= a.qe && u.rbeg >= a.rb && u.rbeg + u.len <= a.re) a.seedcov += u.len;
				}
				a.w = aw0 > aw1 ? aw0 : aw1;
				a.seedlen0 = s.len;
				a.frac_rep = ck.frac_rep;
				av[n_av++] = a;
				phase = TX_NEXT_SEED;
				break; }
			default: done = true; break;
			}
		}
		if (tx_all_done(done)) break; // full-mask vote: exit test and reconvergence point
		// ---- one visit of the DP site: up to P.tx_chunk cells of the current row (the scalar recurrence of ksw_extend2,
		// verbatim).  A row longer than the chunk is continued on the next trip, so a lane with a short row waits for at
		// most one chunk of its neighbours' long rows, not for the whole row.
		if (in_ext) {
			if (jc < 0) { // row start
				// target base of row i from the 2-bit packed reference (one 32-bit word holds 16 consecutive forward bases)
				const i64 p = tbase + (i64)tdir * i;
				const bool rv = p >= l_pac;
				const i64 fp = rv ? (l_pac << 1) - 1 - p : p;
				const i64 widx = fp >> 4;
				if (widx != pac_widx) { pac_word = ldg((const u32*)x.pac + widx); pac_widx = widx; }
				int tb = (int)((pac_word >> (8 * (int)((fp >> 2) & 3))) >> ((~fp & 3) << 1) & 3);
				if (rv) tb = 3 - tb;
				rp = tm.rowpack[tb];
				if (beg < i - w) beg = i - w;
				if (end > i + w + 1) end = i + w + 1;
				if (end > qlen) end = qlen;
				f = 0; m = 0; mj = -1;
				if (beg == 0) { h1 = h0 - (o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; } else h1 = 0;
				if (end > beg) cells += (u64)(end - beg);
				jc = beg;
			}
			{
				const int jstop = end - jc > chunk ? jc + chunk : end;
				CELL* rw = row + (size_t)jc * RS;
				const u8* qp = tm.rd + (size_t)(qbase + qdir * jc) * QS;
				const ptrdiff_t qstep = (ptrdiff_t)qdir * (ptrdiff_t)QS;
#pragma unroll 4
				for (int j = jc; j < jstop; ++j) {
					const u32 he = (u32)*rw;
					int M = (int)(he & HMASK), e = (int)(he >> HB);
					const int sc = (int)(i8)(rp >> (8 * (int)*qp));
					M = M ? M + sc : 0;
					int h = M > e ? M : e;
					h = h > f ? h : f;
					mj = m > h ? mj : j;
					m = m > h ? m : h;
					int t = M - oe_del; t = t > 0 ? t : 0;
					e -= e_del; e = e > t ? e : t;
					*rw = (CELL)((u32)h1 | (u32)e << HB);
					h1 = h;
					t = M - oe_ins; t = t > 0 ? t : 0;
					f -= e_ins; f = f > t ? f : t;
					rw += RS; qp += qstep;
				}
				jc = jstop > jc ? jstop : jc;
			}
			if (jc >= end) { // row end
				jc = -1;
				row[(size_t)end * RS] = (CELL)h1; // H[end] = h1, E[end] = 0
				bool fin = false;
				const int jfin = beg < end ? end : beg;
				if (jfin == qlen) { max_ie = gscore > h1 ? max_ie : i; gscore = gscore > h1 ? gscore : h1; }
				if (m == 0) fin = true;
				else {
					if (m > max) { max = m; max_i = i; max_j = mj; int off = iabs(mj - i); max_off = max_off > off ? max_off : off; }
					else if (opt.zdrop > 0) {
						if (i - max_i > mj - max_j) { if (max - m - ((i - max_i) - (mj - max_j)) * e_del > opt.zdrop) fin = true; }
						else { if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > opt.zdrop) fin = true; }
					}
					if (!fin) {
						// shrink [beg, end) to the cells that can still score (leading / trailing zeros of the row)
						int jb = beg; const CELL* pb = row + (size_t)beg * RS;
						while (jb < end && *pb == 0) { ++jb; pb += RS; }
						beg = jb;
						int je = end; const CELL* pe = row + (size_t)end * RS;
						while (je >= jb && *pe == 0) { --je; pe -= RS; }
						end = je + 2 < qlen ? je + 2 : qlen;
						++i;
						if (i >= tlen) fin = true;
					}
				}
				if (fin) { in_ext = false; phase = TX_EXT_DONE; }
			}
		}
	}
#undef TX_TO_WARP
	if (cnt) { if (cells) atomic_add_u64(&cnt->cells_ext, cells); if (n_ext) atomic_add_u64(&cnt->n_ext, n_ext); }
}

// =====================================================================================================================
// Extension as UNIFORM SIDE TASKS (the production route; the read-level machine above is kept as the comparison route
// behind B200_EXT_LEGACY).
//
// The read-level machine spends ~45 % of its instructions in per-read control flow executed by 1-3 lanes of a warp.
// Here the stage is three kinds of launches with no state machine in any of them:
//   prep    (thread per light read, warp per heavy read)  reference windows of the chains, seeds ordered by score, and
//           one extension TASK per seed that bwa will certainly or probably extend: the best seed of every chain of a
//           light read, every kept seed of a heavy read;
//   sides   two launches, LEFT then RIGHT.  A task's left and right extensions are separate work items (the right one
//           starts from the left one's score), each list ordered by its query length, and a warp takes 32 neighbours
//           of that order at a time: its lanes run the plain nested loops of ksw_extend2 over rows of (nearly) the
//           same width, so they stay converged through the inner loop and meet again at every row end;
//   replay  (thread per light read, warp per heavy read)  mem_chain2aln's decision sequence -- containment tests, the
//           "seed on another diagonal" exception, seed coverage -- consuming the finished extensions.  An extension the
//           replay needs but nobody computed (a second seed of a chain that turns out not to be contained) sends the
//           read to the warp kernel, which computes it itself.
// Exactness: an extension's result does not depend on the hits found so far; only the decision to USE it does, and
// that is replayed in the reference's order.

template <class CELL>
struct TxCell {
	static constexpr int HB = (int)sizeof(CELL) * 4;
	static constexpr u32 HMASK = (1u << HB) - 1u;
	static constexpr i64 SCORE_CAP = sizeof(CELL) == 2 ? 255 : 60000;
};

// ksw_extend2 by one thread over its shared-memory row: the scalar recurrence verbatim, H and E of a
// column in one CELL.  rd[j * QS] = query code of DP column j (already in DP orientation).
template <class CELL>
DEV void tx_extend2(const DevIndex& x, const MemOpt& opt, const ThreadExtMem& tm, int mx, int qlen, int tlen, i64 tbase, int tdir,
                    int w, int end_bonus, int h0, ExtResult& res, u64& cells)
{
	constexpr int HB = TxCell<CELL>::HB; constexpr u32 HMASK = TxCell<CELL>::HMASK;
	const int o_del = opt.o_del, e_del = opt.e_del, o_ins = opt.o_ins, e_ins = opt.e_ins;
	const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins, zdrop = opt.zdrop;
	constexpr size_t RS = TX_STRIDE, QS = TX_STRIDE;
	CELL* const row = (CELL*)tm.row;
	const i64 l_pac = x.l_pac;
	{
		CELL* rw = row;
		rw[0] = (CELL)h0;
		int v = h0 > oe_ins ? h0 - oe_ins : 0;
		for (int j = 1; j <= qlen; ++j) { rw[(size_t)j * RS] = (CELL)v; v = v > e_ins ? v - e_ins : 0; }
	}
	// band clamp of ksw_extend2, int(double(x) / e + 1.) clamped to >= 1, in integers (see cal_max_gap)
	const int xi_ = qlen * mx + end_bonus - o_ins, xd_ = qlen * mx + end_bonus - o_del;
	int max_ins = xi_ >= 0 ? (e_ins == 1 ? xi_ + 1 : xi_ / e_ins + 1) : 0; max_ins = max_ins > 1 ? max_ins : 1;
	w = w < max_ins ? w : max_ins;
	int max_del = xd_ >= 0 ? (e_del == 1 ? xd_ + 1 : xd_ / e_del + 1) : 0; max_del = max_del > 1 ? max_del : 1;
	w = w < max_del ? w : max_del;
	int max = h0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, max_off = 0, beg = 0, end = qlen;
	u32 pac_word = 0; i64 pac_widx = -1;
	for (int i = 0; i < tlen; ++i) {
		// target base of row i from the 2-bit packed reference (one 32-bit word holds 16 consecutive forward bases)
		const i64 p = tbase + (i64)tdir * i;
		const bool rv = p >= l_pac;
		const i64 fp = rv ? (l_pac << 1) - 1 - p : p;
		const i64 widx = fp >> 4;
		if (widx != pac_widx) { pac_word = ldg((const u32*)x.pac + widx); pac_widx = widx; }
		int tb = (int)((pac_word >> (8 * (int)((fp >> 2) & 3))) >> ((~fp & 3) << 1) & 3);
		if (rv) tb = 3 - tb;
		const u64 rp = tm.rowpack[tb];
		if (beg < i - w) beg = i - w;
		if (end > i + w + 1) end = i + w + 1;
		if (end > qlen) end = qlen;
		int f = 0, m = 0, mj = -1, h1;
		if (beg == 0) { h1 = h0 - (o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; } else h1 = 0;
		if (end > beg) cells += (u64)(end - beg);
		{
			CELL* rw = row + (size_t)beg * RS;
			const u8* qp = tm.rd + (size_t)beg * QS;
			// the row maximum and the LAST column that attains it in one running maximum of (h << 16 | j) -- `mj = m > h ? mj : j`
			// moves to j whenever h >= m (h < 2^15, j < 2^16)
			u32 mk = 0u;
#pragma unroll 4
			for (int j = beg; j < end; ++j) {
				const u32 he = (u32)*rw;
				int M = (int)(he & HMASK), e = (int)(he >> HB);
				const int sc = (int)(i8)(rp >> (int)*qp); // (the codes are stored times 8: the shift count)
				M = M ? M + sc : 0;
				int h = M > e ? M : e;
				h = h > f ? h : f;
				const u32 key = (u32)h << 16 | (u32)j;
				mk = mk > key ? mk : key;
				int t = M - oe_del; t = t > 0 ? t : 0;
				e -= e_del; e = e > t ? e : t;
				*rw = (CELL)((u32)h1 | (u32)e << HB);
				h1 = h;
				t = M - oe_ins; t = t > 0 ? t : 0;
				f -= e_ins; f = f > t ? f : t;
				rw += RS; qp += QS;
			}
			if (end > beg) { m = (int)(mk >> 16); mj = (int)(mk & 0xffffu); }
		}
		row[(size_t)end * RS] = (CELL)h1; // H[end] = h1, E[end] = 0
		const int jfin = beg < end ? end : beg;
		if (jfin == qlen) { max_ie = gscore > h1 ? max_ie : i; gscore = gscore > h1 ? gscore : h1; }
		if (m == 0) break;
		if (m > max) { max = m; max_i = i; max_j = mj; int off = iabs(mj - i); max_off = max_off > off ? max_off : off; }
		else if (zdrop > 0) {
			if (i - max_i > mj - max_j) { if (max - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop) break; }
			else { if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) break; }
		}
		// shrink [beg, end) to the cells that can still score (leading / trailing zeros of the row)
		int jb = beg; const CELL* pb = row + (size_t)beg * RS;
		while (jb < end && *pb == 0) { ++jb; pb += RS; }
		beg = jb;
		int je = end; const CELL* pe = row + (size_t)end * RS;
		while (je >= jb && *pe == 0) { --je; pe -= RS; }
		end = je + 2 < qlen ? je + 2 : qlen;
	}
	res.score = max; res.qle = max_j + 1; res.tle = max_i + 1; res.gtle = max_ie + 1; res.gscore = gscore; res.max_off = max_off;
}

// One launch = one side (0 left, 1 right) of every task.  `order` lists task indices by ascending query length of that
// side; warps take 32 neighbours at a time from the long end (longest jobs first).
template <int SIDE, class CELL>
DEV void thread_extend_sides(const BatchParams& P, u32* counter, const ThreadExtMem& tm)
{
	const DevIndex& x = P.x; const MemOpt& opt = P.opt;
	const i64 lo = (i64)P.ext_lo, n = (i64)P.ext_hi - lo; // this launch: positions [ext_lo, ext_hi) of the side's order
	const u32* order = SIDE == 0 ? P.ext_order_l : P.ext_order_r;
	int mx = 0;
	for (int k = 0; k < 25; ++k) mx = mx > opt.mat[k] ? mx : (int)opt.mat[k];
	const size_t QS = (size_t)tm.rd_stride;
	const int lane = simt_lane();
	u64 cells = 0, n_ext = 0;
	for (;;) {
		u32 base = 0;
		if (lane == 0) base = atomic_add_u32(counter, (u32)SIMT_GROUP);
		base = lane_bcast0(base);
		if ((i64)base >= n) break;
		const i64 pos = lo + n - 1 - ((i64)base + lane);
		if (pos >= lo) {
			const u32 ti = order ? order[pos] : (u32)pos;
			const ExtTask tk = P.ext_tasks[ti];
			const i64 r = (i64)tk.r;
			const u64 t0 = P.seed_off[r];
			const ChainK ck = P.ck[t0 + tk.c];
			const u64 slot = t0 + ck.seed_start + tk.t;
			const Seed s = P.ks[slot];
			const int l_query = (int)(P.seq_off[r + 1] - P.seq_off[r] - 1);
			const u8* query = P.seq + P.seq_off[r];
			ExtRes o = P.ext_res[slot];
			const int qlen = SIDE == 0 ? s.qbeg : l_query - (s.qbeg + s.len);
			// 8-/16-bit storage bound: no H/E value exceeds (seed + left + right) scores <= l_query * max(match, a)
			const bool fits = qlen + 1 <= tm.row_cap && qlen <= tm.rd_cap && (i64)(sizeof(CELL) == 2 ? 1 : 2) * l_query * (mx > opt.a ? mx : opt.a) <= TxCell<CELL>::SCORE_CAP;
			if (SIDE == 0) {
				if (qlen == 0) { o.score = o.truesc = s.len * opt.a; o.qb = 0; o.rb = s.rbeg; o.w = opt.w; o.valid = 2; }
				else if (!fits) o.valid = 0; // the replay extends this seed itself
				else {
					for (int j = 0; j < qlen; ++j) tm.rd[(size_t)j * QS] = (u8)(query[s.qbeg - 1 - j] << 3); // both sequences walked backwards; codes times 8 (tx_extend2)
					const int tlen = (int)(s.rbeg - P.ck_rmax[2 * (t0 + tk.c)]);
					ExtResult er; er.score = -1; er.qle = er.tle = er.gtle = 0; er.gscore = -1; er.max_off = 0;
					int score = -1, aw = opt.w;
					for (int tr = 0; tr < 2; ++tr) {
						const int prev = score;
						aw = opt.w << tr;
						if (tlen > 0) tx_extend2<CELL>(x, opt, tm, mx, qlen, tlen, s.rbeg - 1, -1, aw, opt.pen_clip5, s.len * opt.a, er, cells);
						else { er.score = s.len * opt.a; er.qle = er.tle = er.gtle = 0; er.gscore = -1; er.max_off = 0; }
						++n_ext;
						score = er.score;
						if (score == prev || er.max_off < (aw >> 1) + (aw >> 2)) break;
					}
					o.score = score;
					if (er.gscore <= 0 || er.gscore <= score - opt.pen_clip5) { o.qb = s.qbeg - er.qle; o.rb = s.rbeg - er.tle; o.truesc = score; }
					else { o.qb = 0; o.rb = s.rbeg - er.gtle; o.truesc = er.gscore; }
					o.w = aw; o.valid = 2;
				}
			} else if (o.valid == 2) { // (0: the left side was refused)
				const int qe = s.qbeg + s.len; const i64 re = s.rbeg + s.len;
				if (qlen == 0) { o.qe = l_query; o.re = re; o.valid = 1; }
				else if (!fits) o.valid = 0;
				else {
					for (int j = 0; j < qlen; ++j) tm.rd[(size_t)j * QS] = (u8)(query[qe + j] << 3);
					const int tlen = (int)(P.ck_rmax[2 * (t0 + tk.c) + 1] - re);
					const int sc0 = o.score;
					ExtResult er; er.score = -1; er.qle = er.tle = er.gtle = 0; er.gscore = -1; er.max_off = 0;
					int score = sc0, aw = opt.w;
					for (int tr = 0; tr < 2; ++tr) {
						const int prev = score;
						aw = opt.w << tr;
						if (tlen > 0) tx_extend2<CELL>(x, opt, tm, mx, qlen, tlen, re, 1, aw, opt.pen_clip3, sc0, er, cells);
						else { er.score = sc0; er.qle = er.tle = er.gtle = 0; er.gscore = -1; er.max_off = 0; }
						++n_ext;
						score = er.score;
						if (score == prev || er.max_off < (aw >> 1) + (aw >> 2)) break;
					}
					if (er.gscore <= 0 || er.gscore <= score - opt.pen_clip3) { o.qe = qe + er.qle; o.re = re + er.tle; o.truesc += score - sc0; }
					else { o.qe = l_query; o.re = re + er.gtle; o.truesc += er.gscore - sc0; }
					o.score = score;
					o.w = o.w > aw ? o.w : aw; o.valid = 1;
				}
			}
			P.ext_res[slot] = o;
		}
		warp_converge();
	}
	if (P.cnt) { if (cells) atomic_add_u64(&P.cnt->cells_ext, cells); if (n_ext) atomic_add_u64(&P.cnt->n_ext, n_ext); }
}
