// fm.cuh -- FM-index walks on the HBM-resident index: occurrence counting, bidirectional
// interval extension, SMEM discovery (3 seeding rounds) and suffix-array resolution.
//
// Replaces, behind jnibwa_createAlignments (/root/reference/src/main/c/jnibwa.c:197-235, call into
// bwa at :214): bwt_occ4 / bwt_2occ4 / bwt_extend / bwt_smem1a / bwt_seed_strategy1 /
// mem_collect_intv / bwt_sa / bwt_invPsi of un-vendored lh3/bwa.
//
// One occurrence block is 64 bytes = two 32-byte sectors: sector 0 holds the four u64 base counts,
// sector 1 the 128 two-bit symbols.  A lookup is four 128-bit read-only loads of one aligned block
// (algorithmic traffic: 64 B per block touch).
#pragma once
#include "types.cuh"

struct OccBlock { uint4 c01, c23, s0, s1; };

DEV OccBlock load_block(const DevIndex& x, u64 blk)
{
	const uint4* p = x.bwt + (blk << 2);
	OccBlock b;
	b.c01 = ldg(p); b.c23 = ldg(p + 1); b.s0 = ldg(p + 2); b.s1 = ldg(p + 3);
	return b;
}

// counts of C, G, T among the first n (1..32) symbols of a 32-symbol word (first symbol in the top bits)
DEV void count32(u64 y, int n, u32& cC, u32& cG, u32& cT)
{
	if (n < 32) y &= ~((1ull << ((32 - n) << 1)) - 1); // masked fields read as A and are never counted below
	u64 lo = y & 0x5555555555555555ull, hi = (y >> 1) & 0x5555555555555555ull;
	u32 t = (u32)popc64(lo & hi);
	cT += t; cC += (u32)popc64(lo) - t; cG += (u32)popc64(hi) - t;
}

// occ of A,C,G,T in packed-BWT positions [0, kk] given the block that contains kk
DEV void occ4_in_block(const OccBlock& b, u64 kk, u64 cnt[4])
{
	int n = (int)(kk & 127) + 1;
	u32 cC = 0, cG = 0, cT = 0;
	u64 y0 = (u64)b.s0.x << 32 | b.s0.y, y1 = (u64)b.s0.z << 32 | b.s0.w;
	u64 y2 = (u64)b.s1.x << 32 | b.s1.y, y3 = (u64)b.s1.z << 32 | b.s1.w;
	count32(y0, n < 32 ? n : 32, cC, cG, cT);
	if (n > 32) count32(y1, n < 64 ? n - 32 : 32, cC, cG, cT);
	if (n > 64) count32(y2, n < 96 ? n - 64 : 32, cC, cG, cT);
	if (n > 96) count32(y3, n - 96, cC, cG, cT);
	cnt[0] = ((u64)b.c01.x | (u64)b.c01.y << 32) + (u64)((u32)n - cC - cG - cT);
	cnt[1] = ((u64)b.c01.z | (u64)b.c01.w << 32) + cC;
	cnt[2] = ((u64)b.c23.x | (u64)b.c23.y << 32) + cG;
	cnt[3] = ((u64)b.c23.z | (u64)b.c23.w << 32) + cT;
}

// k is an SA row number including the '$' row; (u64)-1 yields zeros
DEV void occ4(const DevIndex& x, u64 k, u64 cnt[4])
{
	if (k == ~0ull) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
	k -= (k >= x.primary);
	OccBlock b = load_block(x, k >> 7);
	occ4_in_block(b, k, cnt);
}

struct BiIntv { u64 x0, x1, size; };

// bwt_extend: ok[c] for the four bases; is_back extends the pattern to the left
DEV void fm_extend(const DevIndex& x, const BiIntv& ik, BiIntv ok[4], int is_back)
{
	u64 tk[4], tl[4];
	u64 a0 = is_back ? ik.x0 : ik.x1, oth = is_back ? ik.x1 : ik.x0, s = ik.size;
	u64 k = a0 - 1, l = a0 - 1 + s;
	if (k != ~0ull && l != ~0ull) {
		u64 k_ = k - (k >= x.primary), l_ = l - (l >= x.primary);
		OccBlock bk = load_block(x, k_ >> 7);
		occ4_in_block(bk, k_, tk);
		if ((l_ >> 7) == (k_ >> 7)) occ4_in_block(bk, l_, tl); // both bounds in one block: one touch (bwt_2occ4)
		else { OccBlock bl = load_block(x, l_ >> 7); occ4_in_block(bl, l_, tl); }
	} else { occ4(x, k, tk); occ4(x, l, tl); }
#pragma unroll
	for (int c = 0; c < 4; ++c) {
		u64 v = x.L2[c] + 1 + tk[c];
		if (is_back) ok[c].x0 = v; else ok[c].x1 = v;
		ok[c].size = tl[c] - tk[c];
	}
	u64 o3 = oth + ((a0 <= x.primary && a0 + s - 1 >= x.primary) ? 1 : 0);
	u64 o2 = o3 + ok[3].size, o1 = o2 + ok[2].size, o0 = o1 + ok[1].size;
	if (is_back) { ok[3].x1 = o3; ok[2].x1 = o2; ok[1].x1 = o1; ok[0].x1 = o0; }
	else { ok[3].x0 = o3; ok[2].x0 = o2; ok[1].x0 = o1; ok[0].x0 = o0; }
}

DEV void fm_set_intv(const DevIndex& x, int c, BiIntv& ik)
{
	ik.x0 = x.L2[c] + 1; ik.size = x.L2[c + 1] - x.L2[c]; ik.x1 = x.L2[3 - c] + 1;
}

// ---- SMEM discovery -----------------------------------------------------------------------------
// Scratch lists of one read (thread-private): candidate intervals carry the query end in `end`.
struct SmemScratch { Intv* a; Intv* b; int cap; };

struct IntvOut { Intv* v; int n, cap; bool overflow; };
DEV void out_push(IntvOut& o, const Intv& t) { if (o.n < o.cap) o.v[o.n++] = t; else o.overflow = true; }

// mem_collect_intv (bwt_smem1a with max_intv = 0 for rounds 1 and 2, bwt_seed_strategy1 for round 3; App. B.3) as ONE
// loop with ONE fm_extend call site.  The three rounds and the forward / backward phases of SMEM discovery are states
// of a small machine: each trip of the loop first advances the machine, without touching the index, to the next
// interval that has to be extended (cheap, lane-divergent bookkeeping), then all lanes of the warp meet at the same
// fm_extend (two occurrence-block fetches, the expensive part) and consume its result.  With nested per-phase loops the
// lanes of a warp sit in different loops and the block fetches serialise (measured: 8 of 32 lanes active); with a
// single call site they issue together.  The interval lists it produces are those of the nested-loop formulation.
enum { SM_NEXT_R1, SM_FWD, SM_BWD_ROW, SM_BWD_ITEM, SM_NEXT_R2, SM_NEXT_R3, SM_S3, SM_DONE };
enum { SMX_NONE, SMX_FWD, SMX_BWD, SMX_S3 };

DEV void intv_set(Intv& t, const BiIntv& b, u32 beg, u32 end) { t.x0 = b.x0; t.x1 = b.x1; t.size = b.size; t.beg = beg; t.end = end; }

DEV void fm_collect_intv(const DevIndex& x, const MemOpt& opt, int len, const u8* q, SmemScratch& sc, IntvOut& o)
{
	const int split_len = (int)(opt.min_seed_len * opt.split_factor + .499);
	const int min_seed_len = opt.min_seed_len;
	o.n = 0; o.overflow = false;
	int st = SM_NEXT_R1, round = 1, pend = SMX_NONE;
	int xx = 0;                 // rounds 1 and 3: next start position
	int k2 = 0, old_n = 0;      // round 2: cursor over the round-1 SMEMs
	// one SMEM search (bwt_smem1a) in flight
	int start = 0, i = 0, ret = 0, c = -1, j = 0, first_out = 0;
	u64 min_intv = 1, last_size = 0;
	BiIntv ik, ext; u32 ik_end = 0;
	Intv* prev = sc.a; Intv* curr = sc.b;
	int n_prev = 0, n_curr = 0;
	u32 last_beg = 0; bool have_out = false;
	Intv pcur; // backward phase: the candidate being extended
	BiIntv ok[4];
	ik.x0 = ik.x1 = ik.size = 0; ext = ik;
	memset(&pcur, 0, sizeof(pcur));
#define SM_PUSH_CURR(bi, e) do { if (n_curr < sc.cap) { intv_set(curr[n_curr], bi, 0, e); ++n_curr; } else o.overflow = true; } while (0)
	for (;;) {
		// ---- advance to the next extension (or to the end of the read)
		while (pend == SMX_NONE && st != SM_DONE) {
			switch (st) {
			case SM_NEXT_R1:
				while (xx < len && q[xx] > 3) ++xx;
				if (xx >= len) { old_n = o.n; k2 = 0; round = 2; st = SM_NEXT_R2; break; }
				start = xx; min_intv = 1;
				fm_set_intv(x, q[start], ik); ik_end = (u32)(start + 1); n_curr = 0; i = start + 1; prev = sc.a; curr = sc.b;
				st = SM_FWD;
				break;
			case SM_NEXT_R2: {
				bool go = false;
				while (k2 < old_n) {
					const Intv p = o.v[k2];
					if ((int)(p.end - p.beg) < split_len || p.size > (u64)opt.split_width) { ++k2; continue; }
					start = (int)((p.beg + p.end) >> 1); min_intv = p.size + 1;
					go = true;
					break;
				}
				if (!go) { xx = 0; round = 3; st = opt.max_mem_intv > 0 ? SM_NEXT_R3 : SM_DONE; break; }
				if (q[start] > 3) { ++k2; break; } // (an SMEM never holds an ambiguous base; kept for fidelity)
				fm_set_intv(x, q[start], ik); ik_end = (u32)(start + 1); n_curr = 0; i = start + 1; prev = sc.a; curr = sc.b;
				st = SM_FWD;
				break; }
			case SM_FWD: // forward phase: record the interval at every change of size
				if (i < len && q[i] < 4) { ext = ik; pend = SMX_FWD; break; }
				SM_PUSH_CURR(ik, ik_end); // i == len, or an ambiguous base ends the match
				goto end_forward;
			end_forward:
				for (int a = 0; a < n_curr >> 1; ++a) { Intv t = curr[a]; curr[a] = curr[n_curr - 1 - a]; curr[n_curr - 1 - a] = t; } // longest match first
				ret = (int)curr[0].end;
				{ Intv* t = curr; curr = prev; prev = t; n_prev = n_curr; }
				first_out = o.n; have_out = false; last_beg = 0;
				i = start - 1;
				st = SM_BWD_ROW;
				break;
			case SM_BWD_ROW:
				if (i < -1) goto end_backward;
				c = i < 0 ? -1 : (q[i] < 4 ? (int)q[i] : -1);
				n_curr = 0; last_size = 0; j = 0;
				st = SM_BWD_ITEM;
				break;
			case SM_BWD_ITEM:
				if (j == n_prev) { // row finished
					if (n_curr == 0) goto end_backward;
					{ Intv* t = curr; curr = prev; prev = t; n_prev = n_curr; }
					--i;
					st = SM_BWD_ROW;
					break;
				}
				pcur = prev[j];
				if (c >= 0) { ext.x0 = pcur.x0; ext.x1 = pcur.x1; ext.size = pcur.size; pend = SMX_BWD; break; }
				// no base to extend with: the candidate cannot grow
				if (n_curr == 0 && (!have_out || (u32)(i + 1) < last_beg)) {
					Intv t = pcur; t.beg = (u32)(i + 1);
					have_out = true; last_beg = t.beg;
					out_push(o, t);
				}
				++j;
				break;
			end_backward: {
				for (int a = first_out, b = o.n - 1; a < b; ++a, --b) { Intv t = o.v[a]; o.v[a] = o.v[b]; o.v[b] = t; } // ascending begin
				int w = first_out;
				for (int a = first_out; a < o.n; ++a) if ((int)(o.v[a].end - o.v[a].beg) >= min_seed_len) o.v[w++] = o.v[a];
				o.n = w;
				if (round == 1) { xx = ret; st = SM_NEXT_R1; } else { ++k2; st = SM_NEXT_R2; }
				break; }
			case SM_NEXT_R3:
				while (xx < len && q[xx] > 3) ++xx;
				if (xx >= len) { st = SM_DONE; break; }
				start = xx;
				fm_set_intv(x, q[start], ik); i = start + 1;
				st = SM_S3;
				break;
			case SM_S3: // bwt_seed_strategy1: shortest forward match of >= min_seed_len+1 bases with < max_mem_intv occurrences
				if (i >= len) { xx = len; st = SM_NEXT_R3; break; }
				if (q[i] > 3) { xx = i + 1; st = SM_NEXT_R3; break; }
				ext = ik; pend = SMX_S3;
				break;
			default: st = SM_DONE; break;
			}
		}
		if (st == SM_DONE && pend == SMX_NONE) break;
		// ---- the one place where the index is read
		fm_extend(x, ext, ok, pend == SMX_BWD ? 1 : 0);
		// ---- consume
		if (pend == SMX_FWD) {
			const int cc = 3 - q[i];
			bool stop = false;
			if (ok[cc].size != ik.size) {
				SM_PUSH_CURR(ik, ik_end);
				if (ok[cc].size < min_intv) stop = true;
			}
			if (!stop) { ik = ok[cc]; ik_end = (u32)(i + 1); ++i; st = SM_FWD; }
			else { // same epilogue as the other forward exits (no second push: the loop ended before i == len)
				for (int a = 0; a < n_curr >> 1; ++a) { Intv t = curr[a]; curr[a] = curr[n_curr - 1 - a]; curr[n_curr - 1 - a] = t; }
				ret = (int)curr[0].end;
				{ Intv* t = curr; curr = prev; prev = t; n_prev = n_curr; }
				first_out = o.n; have_out = false; last_beg = 0;
				i = start - 1;
				st = SM_BWD_ROW;
			}
		} else if (pend == SMX_BWD) {
			if (ok[c].size < min_intv) {
				if (n_curr == 0 && (!have_out || (u32)(i + 1) < last_beg)) {
					Intv t = pcur; t.beg = (u32)(i + 1);
					have_out = true; last_beg = t.beg;
					out_push(o, t);
				}
			} else if (n_curr == 0 || ok[c].size != last_size) {
				intv_set(curr[n_curr], ok[c], 0, pcur.end); ++n_curr; last_size = ok[c].size; // n_curr <= n_prev <= cap
			}
			++j;
			st = SM_BWD_ITEM;
		} else { // SMX_S3
			const int cc = 3 - q[i];
			if (ok[cc].size < opt.max_mem_intv && i - start >= min_seed_len) {
				if (ok[cc].size > 0) { Intv m; intv_set(m, ok[cc], (u32)start, (u32)(i + 1)); out_push(o, m); }
				xx = i + 1; st = SM_NEXT_R3;
			} else { ik = ok[cc]; ++i; st = SM_S3; }
		}
		pend = SMX_NONE;
	}
#undef SM_PUSH_CURR
	// ordered by (beg, end).  Equal keys are identical intervals, so the sort algorithm is unobservable here (unlike
	// the hazard-H4 sorts) and a plain insertion sort is used.
	for (int a = 1; a < o.n; ++a) {
		Intv t = o.v[a];
		u64 key = (u64)t.beg << 32 | t.end;
		int b = a - 1;
		while (b >= 0 && ((u64)o.v[b].beg << 32 | o.v[b].end) > key) { o.v[b + 1] = o.v[b]; --b; }
		o.v[b + 1] = t;
	}
}

// ---- suffix-array resolution ----------------------------------------------------------------------
// bwt_sa: LF-walk from row k to a sampled row.  One block touch per step (the BWT symbol and its rank
// live in the same 64-byte block) + one 8-byte SA read.
DEV u64 fm_sa(const DevIndex& x, u64 k)
{
	u64 sa = 0, mask = (u64)x.sa_intv - 1;
	while (k & mask) {
		++sa;
		if (k == x.primary) { k = 0; continue; }
		u64 kk = k - (k > x.primary); // position of this row's symbol in the packed BWT
		OccBlock b = load_block(x, kk >> 7);
		u32 wsel = (u32)(kk & 127) >> 4;
		u32 wd = wsel < 4 ? (wsel == 0 ? b.s0.x : wsel == 1 ? b.s0.y : wsel == 2 ? b.s0.z : b.s0.w)
		                  : (wsel == 4 ? b.s1.x : wsel == 5 ? b.s1.y : wsel == 6 ? b.s1.z : b.s1.w);
		int c = (int)(wd >> ((15 - (kk & 15)) << 1) & 3);
		// occ(k, c) counts rows [0,k]; k > primary maps to packed position k-1 == kk, k < primary to k == kk
		u64 cnt[4];
		occ4_in_block(b, kk, cnt);
		k = x.L2[c] + cnt[c];
	}
	return sa + ldg(&x.sa[k / x.sa_intv]);
}
