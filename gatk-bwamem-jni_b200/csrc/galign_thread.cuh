// galign_thread.cuh -- the CIGAR-producing global alignments of the finalize stage as a queue of independent tasks
// (mem_reg2aln's band loop around bwa_gen_cigar2 / ksw_global2 of un-vendored lh3/bwa, SURVEY.md App. B.9, behind
// /root/reference/src/main/c/jnibwa.c:214).
//
// In the thread-per-pair finalize kernel only the reads with >= 3 mismatches or an indel need a banded alignment
// (the others take bwa's ungapped shortcut), so at any moment a handful of the 32 lanes of a warp sit in the DP loop
// while the rest wait.  The finalize stage is therefore run in two passes around this kernel:
//   collect  the whole per-pair decision logic runs with warp_reg2aln recording WHAT has to be aligned (no decision
//            depends on an alignment's outcome: records are assembled from them, nothing is chosen by them);
//   tasks    (this file) every recorded alignment is computed by a lane of a persistent queue kernel whose control
//            flow is a state machine around ONE DP-row site, like the seed extension (extend_thread.cuh);
//   consume  the decision logic runs again on the restored region lists, warp_reg2aln now returning the k-th
//            precomputed alignment of the item.
// A task the thread-level formulation cannot take (band wider than the shared-memory ring, scores beyond the 16-bit
// rows, traceback or CIGAR larger than the thread scratch) is marked invalid; its item is then redone by the warp-level kernel, as before.
#pragma once
#include "finalize.cuh"

enum { GA_FETCH, GA_TRY, GA_DP_END, GA_TRY_END, GA_FINISH };

DEV void thread_galign_queue(const BatchParams& P, i64 n_tasks, u32* counter, WarpMem& wm, WarpScratch& ws, u32* lerr)
{
	const DevIndex& x = P.x; const MemOpt& opt = P.opt;
	const i64 l_pac = x.l_pac;
	const int o_del = opt.o_del, e_del = opt.e_del, o_ins = opt.o_ins, e_ins = opt.e_ins;
	const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	const size_t qs = (size_t)wm.q_stride;
	// ring of column slots, H and E of a column packed into one word as 16-bit halves (one LDS + one STS per cell, half
	// the shared memory).  "Minus infinity" shrinks with them: every value of the recurrence is either real or
	// GA_NEG_INF plus a real offset, two infinite values compare by their offsets whatever the constant is, and an
	// infinite one loses against every real one as long as the offsets stay below GA_SPAN -- tasks that could exceed
	// that bound are left to the warp kernel (32-bit rows).
	u32* const R = (u32*)wm.ringH;
	const int rmask = wm.ring_n - 1, rstride = wm.ring_stride;
	enum { GA_NEG_INF = -0x4000, GA_SPAN = 0x1800 };
	int pen_max = 1;
	for (int k = 0; k < 25; ++k) { int v = iabs((int)opt.mat[k]); pen_max = pen_max > v ? pen_max : v; }
	pen_max = pen_max > e_del ? pen_max : e_del; pen_max = pen_max > e_ins ? pen_max : e_ins;
#define HX(j) (((j) & rmask) * rstride)
#define GA_PACK(h_, e_) ((u32)(u16)(i16)(h_) | (u32)(u16)(i16)(e_) << 16)
	int phase = GA_FETCH;
	bool in_dp = false, done = false;
	i64 tid = -1;
	GaTask tk; memset(&tk, 0, sizeof(tk));
	int l_query = 0, rlen = 0, w2 = 0, tries = 0, last_sc = 0, score = 0, nc = 0;
	bool rev = false; TAcc ta; ta.base = 0; ta.stride = 1; ta.piv = -1;
	// DP in flight
	int w = 0, n_col = 0, zs = 0, i = 0;
	u64 cells = 0, n_glb = 0;
	u32 pac_word = 0; i64 pac_widx = -1;
// a task this formulation cannot take goes on the list of the warp-per-task kernel (wbody_galign_fb); valid = 0 stands
// until that kernel has produced the alignment (an item whose task stays invalid is redone by a warp as a whole)
#define GA_INVALID() do { P.ga_res[tid].valid = 0; { const u32 s_ = atomic_add_u32(P.n_ga_fb, 1u); P.ga_fb[s_] = (u32)tid; } phase = GA_FETCH; } while (0)
	for (;;) {
		while (!in_dp && !done) {
			switch (phase) {
			case GA_FETCH: {
				tid = (i64)atomic_add_u32(counter, 1u);
				if (tid >= n_tasks) { done = true; break; }
				if (P.ga_order) tid = (i64)P.ga_order[n_tasks - 1 - tid]; // tasks by descending cost: the lanes of a warp hold similar ones
				tk = P.ga_tasks[tid];
				l_query = tk.qe - tk.qb; rlen = (int)(tk.re - tk.rb);
				if (l_query <= 0 || tk.rb >= tk.re || (tk.rb < l_pac && tk.re > l_pac) || l_query > wm.cap) { GA_INVALID(); break; }
				rev = tk.rb >= l_pac; // reverse (not complement) both sequences so that indels left-align on the forward strand
				ta.piv = -1;
				if (rev) { ta.base = tk.re - 1; ta.stride = -1; } else { ta.base = tk.rb; ta.stride = 1; }
				{
					const u8* query = P.seq + tk.qoff + tk.qb;
					for (int j = 0; j < l_query; ++j) WQ(wm, j) = rev ? query[l_query - 1 - j] : query[j];
				}
				w2 = tk.w2; tries = 0; last_sc = -(1 << 30); score = 0; nc = 0;
				*lerr = 0;
				phase = GA_TRY;
				break; }
			case GA_TRY: { // one call of bwa_gen_cigar2 with band parameter w2
				w2 = w2 < opt.w << 2 ? w2 : opt.w << 2;
				nc = 0;
				if (l_query == rlen && w2 == 0) { // ungapped: one M run, score by direct summation
					int s = 0;
					for (int j = 0; j < l_query; ++j) s += (int)opt.mat[ref_base2(x, tacc_pos(ta, j)) * 5 + WQ(wm, j)];
					score = s;
					ws.cigar[0] = (u32)l_query << 4; nc = 1;
					phase = GA_TRY_END;
					break;
				}
				int max_ins = (int)((double)(((l_query + 1) >> 1) * opt.mat[0] - opt.o_ins) / opt.e_ins + 1.);
				int max_del = (int)((double)(((l_query + 1) >> 1) * opt.mat[0] - opt.o_del) / opt.e_del + 1.);
				int max_gap = max_ins > max_del ? max_ins : max_del;
				max_gap = max_gap > 1 ? max_gap : 1;
				const int dl = iabs(rlen - l_query);
				w = (max_gap + dl + 1) >> 1;
				w = w < w2 ? w : w2;
				const int min_w = dl + 3;
				w = w > min_w ? w : min_w;
				n_col = l_query < 2 * w + 1 ? l_query : 2 * w + 1;
				zs = (n_col + 3) & ~3;
				if ((size_t)zs * (size_t)rlen > ws.z_cap || 2 * w + 2 > wm.ring_n || (i64)(rlen + l_query + 2) * pen_max + oe_del + oe_ins >= GA_SPAN) { GA_INVALID(); break; } // a warp takes the item
				{ // first row of the banded global alignment (ksw_global2)
					int j;
					R[HX(0)] = GA_PACK(0, GA_NEG_INF);
					for (j = 1; j <= l_query && j <= w; ++j) R[HX(j)] = GA_PACK(-(o_ins + e_ins * j), GA_NEG_INF);
					if (j <= l_query) R[HX(j)] = GA_PACK(GA_NEG_INF, GA_NEG_INF); // column w+1: only ever read after row 0 overwrote it
				}
				i = 0; ++n_glb;
				if (rlen > 0) in_dp = true; else phase = GA_DP_END;
				break; }
			case GA_DP_END: { // score and traceback -> CIGAR (ops are produced back to front)
				score = (int)(i16)(u16)(R[HX(l_query)] & 0xffffu);
				const u8* z = ws.z; u32* cigar = ws.cigar; const int cigar_cap = ws.cigar_cap;
				bool cg_ovf = false;
				int n_cigar = 0, which = 0, ti = rlen - 1, k = (ti + w + 1 < l_query ? ti + w + 1 : l_query) - 1;
				u32 cur_op = 0xffffffffu, cur_len = 0;
#define CG_PUT1() do { if (n_cigar < cigar_cap) cigar[n_cigar] = cur_len << 4 | cur_op; else cg_ovf = true; ++n_cigar; } while (0)
				while (ti >= 0 && k >= 0) {
					which = z[(size_t)ti * zs + (k - (ti > w ? ti - w : 0))] >> (which << 1) & 3;
					u32 op;
					if (which == 0) { op = 0; --ti; --k; }
					else if (which == 1) { op = 2; --ti; }
					else { op = 1; --k; }
					if (op == cur_op) ++cur_len;
					else { if (cur_op != 0xffffffffu) { CG_PUT1(); } cur_op = op; cur_len = 1; }
				}
				if (ti >= 0) { if (cur_op == 2) cur_len += (u32)(ti + 1); else { if (cur_op != 0xffffffffu) { CG_PUT1(); } cur_op = 2; cur_len = (u32)(ti + 1); } }
				if (k >= 0) { if (cur_op == 1) cur_len += (u32)(k + 1); else { if (cur_op != 0xffffffffu) { CG_PUT1(); } cur_op = 1; cur_len = (u32)(k + 1); } }
				if (cur_op != 0xffffffffu) { CG_PUT1(); }
#undef CG_PUT1
				if (cg_ovf) { GA_INVALID(); break; }
				for (int a = 0; a < n_cigar >> 1; ++a) { u32 t = cigar[a]; cigar[a] = cigar[n_cigar - 1 - a]; cigar[n_cigar - 1 - a] = t; }
				nc = n_cigar;
				phase = GA_TRY_END;
				break; }
			case GA_TRY_END: // mem_reg2aln's band loop: up to three tries with a doubled band until the score stops improving
				if (score == last_sc || w2 == opt.w << 2) { phase = GA_FINISH; break; }
				last_sc = score;
				w2 <<= 1;
				if (++tries < 3 && score < tk.truesc - opt.a) phase = GA_TRY; else phase = GA_FINISH;
				break;
			case GA_FINISH: { // MD / NM of the final CIGAR (bwa_gen_cigar2), then the result record
				GaRes& o = P.ga_res[tid];
				int xq = 0, y = 0, u = 0, n_mm = 0, n_gap = 0, ml = 0;
				const char* int2base = tk.rb < l_pac ? "ACGTN" : "TGCAN";
				bool ovf = nc > GA_CIG_CAP;
#define MDP(ch) do { const char c_ = (char)(ch); if (ml < GA_MD_CAP) o.md[ml] = c_; else ovf = true; ++ml; } while (0)
#define MDN(v) do { int v_ = (v); char b_[12]; int n_ = 0; do { b_[n_++] = (char)('0' + v_ % 10); v_ /= 10; } while (v_); while (n_) MDP(b_[--n_]); } while (0)
				for (int k = 0; k < nc && !ovf; ++k) {
					const u32 cg = ws.cigar[k];
					const int op = (int)(cg & 0xf), len = (int)(cg >> 4);
					if (op == 0) {
						for (int t = 0; t < len; ++t) {
							const int rbs = ref_base2(x, tacc_pos(ta, y + t));
							if ((int)WQ(wm, xq + t) != rbs) { MDN(u); MDP(int2base[rbs]); ++n_mm; u = 0; }
							else ++u;
						}
						xq += len; y += len;
					} else if (op == 2) {
						if (k > 0 && k < nc - 1) {
							MDN(u); MDP('^');
							for (int t = 0; t < len; ++t) MDP(int2base[ref_base2(x, tacc_pos(ta, y + t))]);
							u = 0; n_gap += len;
						}
						y += len;
					} else if (op == 1) { xq += len; n_gap += len; }
				}
				MDN(u);
#undef MDP
#undef MDN
				if (ovf) { GA_INVALID(); break; }
				for (int k = 0; k < nc; ++k) o.cigar[k] = ws.cigar[k];
				o.score = score; o.NM = n_mm + n_gap; o.nc = nc; o.ml = ml; o.valid = 1;
				phase = GA_FETCH;
				break; }
			default: done = true; break;
			}
		}
		if (lanes_all(done)) break;
		// ---- one row of the banded global alignment (the scalar recurrence of ksw_global2, verbatim)
		if (in_dp) {
			const i64 p = tacc_pos(ta, i);
			const bool rv = p >= l_pac;
			const i64 fp = rv ? (l_pac << 1) - 1 - p : p;
			const i64 widx = fp >> 4;
			if (widx != pac_widx) { pac_word = ldg((const u32*)x.pac + widx); pac_widx = widx; }
			int tb = (int)((pac_word >> (8 * (int)((fp >> 2) & 3))) >> ((~fp & 3) << 1) & 3);
			if (rv) tb = 3 - tb;
			const u64 rowpack = P.rowpack[tb];
			int f = GA_NEG_INF, h1, t;
			const int beg = i > w ? i - w : 0;
			const int end = i + w + 1 < l_query ? i + w + 1 : l_query;
			h1 = beg == 0 ? -(o_del + e_del * (i + 1)) : GA_NEG_INF;
			if (end > beg) cells += (u64)(end - beg);
			u32* zi = (u32*)(ws.z + (size_t)i * zs);
			u32 zw = 0;
			for (int j = beg; j < end; ++j) {
				const int jx = HX(j);
				const u32 he = R[jx];
				int h, m = (int)(i16)(u16)(he & 0xffffu), e = (int)he >> 16;
				u32 d;
				const int hl = h1; // H(i, j-1): the value column j holds for the next row
				m += (int)(i8)(rowpack >> (8 * (int)WQ(wm, j)));
				d = m >= e ? 0u : 1u;
				h = m >= e ? m : e;
				d = h >= f ? d : 2u;
				h = h >= f ? h : f;
				h1 = h;
				t = m - oe_del;
				e -= e_del;
				d |= e > t ? 1u << 2 : 0u;
				e = e > t ? e : t;
				R[jx] = GA_PACK(hl, e);
				t = m - oe_ins;
				f -= e_ins;
				d |= f > t ? 2u << 4 : 0u;
				f = f > t ? f : t;
				const int k = j - beg;
				zw |= d << ((k & 3) << 3);
				if ((k & 3) == 3 || j == end - 1) { zi[k >> 2] = zw; zw = 0; }
			}
			R[HX(end)] = GA_PACK(h1, GA_NEG_INF);
			++i;
			if (i >= rlen) { in_dp = false; phase = GA_DP_END; }
		}
	}
#undef HX
#undef GA_PACK
#undef GA_INVALID
	(void)qs;
	if (P.cnt) { if (cells) atomic_add_u64(&P.cnt->cells_glb, cells); if (n_glb) atomic_add_u64(&P.cnt->n_glb, n_glb); }
}
