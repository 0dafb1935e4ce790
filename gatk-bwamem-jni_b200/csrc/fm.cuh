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

// bwt_smem1a with max_intv = 0 (App. B.3).  Emits SMEMs (ascending begin) at o.v[o.n..]; returns the next start.
DEV int fm_smem1(const DevIndex& x, int len, const u8* q, int start, u64 min_intv, SmemScratch& sc, IntvOut& o, int min_seed_len)
{
	if (q[start] > 3) return start + 1;
	if (min_intv < 1) min_intv = 1;
	Intv* prev = sc.a; Intv* curr = sc.b;
	int n_prev, n_curr = 0, i, ret;
	BiIntv ik, ok[4];
	u32 ik_end = (u32)(start + 1);
	fm_set_intv(x, q[start], ik);
	for (i = start + 1; i < len; ++i) { // forward phase: record the interval at every change of size
		if (q[i] < 4) {
			int c = 3 - q[i];
			fm_extend(x, ik, ok, 0);
			if (ok[c].size != ik.size) {
				if (n_curr < sc.cap) { Intv t; t.x0 = ik.x0; t.x1 = ik.x1; t.size = ik.size; t.beg = 0; t.end = ik_end; curr[n_curr++] = t; } else o.overflow = true;
				if (ok[c].size < min_intv) break;
			}
			ik = ok[c]; ik_end = (u32)(i + 1);
		} else {
			if (n_curr < sc.cap) { Intv t; t.x0 = ik.x0; t.x1 = ik.x1; t.size = ik.size; t.beg = 0; t.end = ik_end; curr[n_curr++] = t; } else o.overflow = true;
			break;
		}
	}
	if (i == len) { if (n_curr < sc.cap) { Intv t; t.x0 = ik.x0; t.x1 = ik.x1; t.size = ik.size; t.beg = 0; t.end = ik_end; curr[n_curr++] = t; } else o.overflow = true; }
	for (int j = 0; j < n_curr >> 1; ++j) { Intv t = curr[j]; curr[j] = curr[n_curr - 1 - j]; curr[n_curr - 1 - j] = t; } // longest match first
	ret = (int)curr[0].end;
	{ Intv* t = curr; curr = prev; prev = t; n_prev = n_curr; }
	int first_out = o.n;
	u32 last_beg = 0; bool have_out = false;
	for (i = start - 1; i >= -1; --i) { // backward phase
		int c = i < 0 ? -1 : (q[i] < 4 ? (int)q[i] : -1);
		n_curr = 0;
		u64 last_size = 0;
		for (int j = 0; j < n_prev; ++j) {
			Intv p = prev[j];
			bool can = false;
			if (c >= 0) {
				BiIntv pk; pk.x0 = p.x0; pk.x1 = p.x1; pk.size = p.size;
				fm_extend(x, pk, ok, 1);
				can = ok[c].size >= min_intv;
			}
			if (!can) {
				if (n_curr == 0) {
					if (!have_out || (u32)(i + 1) < last_beg) {
						Intv t = p; t.beg = (u32)(i + 1);
						have_out = true; last_beg = t.beg;
						out_push(o, t);
					}
				}
			} else if (n_curr == 0 || ok[c].size != last_size) {
				Intv t; t.x0 = ok[c].x0; t.x1 = ok[c].x1; t.size = ok[c].size; t.beg = 0; t.end = p.end;
				curr[n_curr++] = t; last_size = t.size; // n_curr <= n_prev <= cap
			}
		}
		if (n_curr == 0) break;
		{ Intv* t = curr; curr = prev; prev = t; n_prev = n_curr; }
	}
	// emitted in descending begin -> reverse, then keep only seeds of at least min_seed_len
	for (int a = first_out, b = o.n - 1; a < b; ++a, --b) { Intv t = o.v[a]; o.v[a] = o.v[b]; o.v[b] = t; }
	int w = first_out;
	for (int a = first_out; a < o.n; ++a)
		if ((int)(o.v[a].end - o.v[a].beg) >= min_seed_len) o.v[w++] = o.v[a];
	o.n = w;
	return ret;
}

// bwt_seed_strategy1 (round 3)
DEV int fm_seed_strategy1(const DevIndex& x, int len, const u8* q, int start, int min_len, u64 max_intv, Intv& mem)
{
	BiIntv ik, ok[4];
	mem.size = 0; mem.x0 = mem.x1 = 0; mem.beg = mem.end = 0;
	if (q[start] > 3) return start + 1;
	fm_set_intv(x, q[start], ik);
	for (int i = start + 1; i < len; ++i) {
		if (q[i] < 4) {
			int c = 3 - q[i];
			fm_extend(x, ik, ok, 0);
			if (ok[c].size < max_intv && i - start >= min_len) {
				mem.x0 = ok[c].x0; mem.x1 = ok[c].x1; mem.size = ok[c].size; mem.beg = (u32)start; mem.end = (u32)(i + 1);
				return i + 1;
			}
			ik = ok[c];
		} else return i + 1;
	}
	return len;
}

// mem_collect_intv: three rounds, then ordered by (beg, end).  Equal keys are identical intervals, so the
// sort algorithm is unobservable here (unlike hazard H4 sorts) and a plain insertion sort is used.
DEV void fm_collect_intv(const DevIndex& x, const MemOpt& opt, int len, const u8* seq, SmemScratch& sc, IntvOut& o)
{
	int start = 0;
	int split_len = (int)(opt.min_seed_len * opt.split_factor + .499);
	o.n = 0; o.overflow = false;
	while (start < len) {
		if (seq[start] < 4) start = fm_smem1(x, len, seq, start, 1, sc, o, opt.min_seed_len);
		else ++start;
	}
	int old_n = o.n;
	for (int k = 0; k < old_n; ++k) {
		Intv p = o.v[k];
		int beg = (int)p.beg, end = (int)p.end;
		if (end - beg < split_len || p.size > (u64)opt.split_width) continue;
		fm_smem1(x, len, seq, (beg + end) >> 1, p.size + 1, sc, o, opt.min_seed_len);
	}
	if (opt.max_mem_intv > 0) {
		start = 0;
		while (start < len) {
			if (seq[start] < 4) {
				Intv m;
				start = fm_seed_strategy1(x, len, seq, start, opt.min_seed_len, opt.max_mem_intv, m);
				if (m.size > 0) out_push(o, m);
			} else ++start;
		}
	}
	for (int i = 1; i < o.n; ++i) {
		Intv t = o.v[i];
		u64 key = (u64)t.beg << 32 | t.end;
		int j = i - 1;
		while (j >= 0 && ((u64)o.v[j].beg << 32 | o.v[j].end) > key) { o.v[j + 1] = o.v[j]; --j; }
		o.v[j + 1] = t;
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
