// fm.cuh -- FM-index walks on the HBM-resident index: occurrence counting, bidirectional
// interval extension, SMEM discovery (3 seeding rounds) and suffix-array resolution.
//
// Replaces, behind jnibwa_createAlignments (/root/reference/src/main/c/jnibwa.c:197-235, call into
// bwa at :214): bwt_occ4 / bwt_2occ4 / bwt_extend / bwt_smem1a / bwt_seed_strategy1 /
// mem_collect_intv / bwt_sa / bwt_invPsi of un-vendored lh3/bwa.
//
// DEVICE LAYOUT OF THE OCCURRENCE ARRAY (re-laid out once at upload by fm_relayout_block; every result is unchanged).
// The image interleaves 4 x u64 running counts with 128 two-bit symbols: a 64-byte block.  The device keeps one 32-byte
// SECTOR per 64 rows instead:
//   words 0..3  the counts of A, C, G, T before the sector, as u32 -- relative to the sector's SUPER-BLOCK when a base
//               occurs 2^32 times or more in the text (DevIndex.occ_sb: u64 counts at every 2^occ_sb_shift-th sector;
//               nullptr for every genome whose per-base totals fit 32 bits, e.g. human: the common case reads nothing
//               extra and pays one uniform branch);
//   words 4..5  the low bits of the 64 symbols (symbol j at bit j & 31 of word 4 + (j >> 5)), words 6..7 the high bits.
// A lookup is two 128-bit read-only loads of one aligned sector and six population counts (one position mask per 32
// symbols shared by the low, high and both planes): half the load instructions, L1 requests and population counts of a
// lookup in the image layout, same bytes in HBM.  (The L2 fetches 64 bytes from HBM on a miss whatever the request and
// ignores the fetch-granularity hint, so DRAM traffic is unchanged -- profiles/README.md -- the gain is instruction issue.
// Three 40-bit counts with A derived from the position need no super-blocks, but their decoding made the compiler
// split the batched sector loads of the backward sweeps: 12.8 -> 16.4 ms.)
#pragma once
#include "types.cuh"

struct OccBlock { uint4 cnt, sym; };
#define OCC_SHIFT 6 /* rows per sector = 64 */
#define OCC_MASK 63

DEV OccBlock load_block(const DevIndex& x, u64 blk)
{
	const uint4* p = x.bwt + (blk << 1);
	OccBlock b;
	b.cnt = ldg(p); b.sym = ldg(p + 1);
	return b;
}

// image layout -> device layout for image block b (src: the image's occ-interleaved BWT of bwt_size words; dst: two
// sectors = 16 words per image block).  sb / sb_shift: the super-block table (or nullptr), read here as the base to
// subtract; super-blocks start on image-block boundaries (sb_shift >= 1).
HOSTDEV void fm_image_counts(const u32* src, u64 bwt_size, u64 b, u64 cnt[4])
{
	for (int c = 0; c < 4; ++c) { const u64 i = (b << 4) + 2 * c; cnt[c] = i + 1 < bwt_size ? ((u64)src[i] | (u64)src[i + 1] << 32) : 0; }
}
DEV void fm_relayout_block(const u32* src, u64 bwt_size, u32* dst, u64 b, const u64* sb, int sb_shift)
{
	const u64 base = b << 4;
	u32 w[16];
	for (int t = 0; t < 16; ++t) w[t] = base + t < bwt_size ? src[base + t] : 0u;
	u64 cnt[4];
	for (int c = 0; c < 4; ++c) cnt[c] = (u64)w[2 * c] | (u64)w[2 * c + 1] << 32;
	for (int h = 0; h < 2; ++h) {
		const u64 sec = b * 2 + h;
		u32 p0[2] = { 0, 0 }, p1[2] = { 0, 0 }; u32 n[4] = { 0, 0, 0, 0 };
		for (int j = 0; j < 64; ++j) {
			const int jj = h * 64 + j;
			const u32 sym = w[8 + (jj >> 4)] >> ((15 - (jj & 15)) << 1) & 3u;
			p0[j >> 5] |= (sym & 1u) << (j & 31);
			p1[j >> 5] |= (sym >> 1) << (j & 31);
			++n[sym];
		}
		u32* d = dst + sec * 8;
		for (int c = 0; c < 4; ++c) d[c] = (u32)(cnt[c] - (sb ? sb[(sec >> sb_shift) * 4 + c] : 0));
		d[4] = p0[0]; d[5] = p0[1]; d[6] = p1[0]; d[7] = p1[1];
		for (int c = 0; c < 4; ++c) cnt[c] += n[c]; // (padding symbols past the end of the text count as A: never looked up)
	}
}

// (a call on purpose: inlined, these few loads are if-converted and every lookup of the common case -- no super-blocks --
// issues them predicated off: ~12 % of the instructions of a backward sweep)
DEVNI u64 occ_super_block_base(const u64* sb, u64 idx, int c) { return ldg(sb + idx * 4 + c); }

// Occurrences of A,C,G,T in packed-BWT positions [0, kk] given the sector that holds kk.  Branch-free: every lane runs
// the same six population counts whatever its offset in the sector.  Symbols past the offset are masked out of both
// planes (= A) and never counted; A is what remains of n.
DEV void occ4_in_block(const DevIndex& x, const OccBlock& b, u64 kk, u64 cnt[4])
{
	const int n = (int)(kk & OCC_MASK) + 1; // 1..64 symbols of this sector
	const u32 m0 = n >= 32 ? 0xffffffffu : (1u << n) - 1u, m1 = n >= 64 ? 0xffffffffu : (n > 32 ? (1u << (n - 32)) - 1u : 0u);
	const u32 lo0 = b.sym.x & m0, lo1 = b.sym.y & m1, hi0 = b.sym.z & m0, hi1 = b.sym.w & m1;
	const u32 sLo = (u32)__popc(lo0) + (u32)__popc(lo1), sHi = (u32)__popc(hi0) + (u32)__popc(hi1), sT = (u32)__popc(lo0 & hi0) + (u32)__popc(lo1 & hi1);
	const u32 cC = sLo - sT, cG = sHi - sT, cT = sT;
	cnt[0] = (u64)(b.cnt.x + ((u32)n - cC - cG - cT));
	cnt[1] = (u64)(b.cnt.y + cC);
	cnt[2] = (u64)(b.cnt.z + cG);
	cnt[3] = (u64)(b.cnt.w + cT);
	if (x.occ_sb) { // (uniform over the launch)
		const u64 idx = (kk >> OCC_SHIFT) >> x.occ_sb_shift;
		cnt[0] += occ_super_block_base(x.occ_sb, idx, 0); cnt[1] += occ_super_block_base(x.occ_sb, idx, 1);
		cnt[2] += occ_super_block_base(x.occ_sb, idx, 2); cnt[3] += occ_super_block_base(x.occ_sb, idx, 3);
	}
}

// k is an SA row number including the '$' row; (u64)-1 yields zeros
DEV void occ4(const DevIndex& x, u64 k, u64 cnt[4])
{
	if (k == ~0ull) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
	k -= (k >= x.primary);
	OccBlock b = load_block(x, k >> OCC_SHIFT);
	occ4_in_block(x, b, k, cnt);
}

struct BiIntv { u64 x0, x1, size; };

// bwt_extend for ONE base c (the only one a caller ever consumes): no per-base arrays, so nothing is spilled to local
// memory by lane-varying indexing.  The complementary coordinate needs the sizes of the bases above c only.
// (in two halves, so that a caller can issue the sector loads of several kinds of step together)
struct ExtRows { u64 k_, l_; bool kz, lz; };
DEV ExtRows fm_extend1_rows(const DevIndex& x, const BiIntv& ik, int is_back)
{
	const u64 a0 = is_back ? ik.x0 : ik.x1;
	const u64 k = a0 - 1, l = a0 - 1 + ik.size;
	ExtRows r;
	r.kz = k == ~0ull; r.lz = l == ~0ull;
	r.k_ = r.kz ? 0 : k - (k >= x.primary); r.l_ = r.lz ? 0 : l - (l >= x.primary);
	return r;
}
DEV BiIntv fm_extend1_finish(const DevIndex& x, const BiIntv& ik, int c, int is_back, const ExtRows& er, const OccBlock& bk, const OccBlock& bl)
{
	u64 tk[4], tl[4];
	const u64 a0 = is_back ? ik.x0 : ik.x1, oth = is_back ? ik.x1 : ik.x0, s = ik.size;
	occ4_in_block(x, bk, er.k_, tk);
	occ4_in_block(x, bl, er.l_, tl);
	if (er.kz) { tk[0] = tk[1] = tk[2] = tk[3] = 0; }
	if (er.lz) { tl[0] = tl[1] = tl[2] = tl[3] = 0; }
	const u64 s1 = tl[1] - tk[1], s2 = tl[2] - tk[2], s3 = tl[3] - tk[3];
	const u64 kc = c == 0 ? tk[0] : c == 1 ? tk[1] : c == 2 ? tk[2] : tk[3];
	const u64 lc = c == 0 ? tl[0] : c == 1 ? tl[1] : c == 2 ? tl[2] : tl[3];
	const u64 L2c = c == 0 ? x.L2[0] : c == 1 ? x.L2[1] : c == 2 ? x.L2[2] : x.L2[3];
	const u64 above = (c < 3 ? s3 : 0) + (c < 2 ? s2 : 0) + (c < 1 ? s1 : 0);
	const u64 v = L2c + 1 + kc;
	const u64 o = oth + ((a0 <= x.primary && a0 + s - 1 >= x.primary) ? 1 : 0) + above;
	BiIntv r;
	r.size = lc - kc;
	if (is_back) { r.x0 = v; r.x1 = o; } else { r.x1 = v; r.x0 = o; }
	return r;
}
DEV BiIntv fm_extend1(const DevIndex& x, const BiIntv& ik, int c, int is_back)
{
	const ExtRows er = fm_extend1_rows(x, ik, is_back);
	const OccBlock bk = load_block(x, er.k_ >> OCC_SHIFT);
	const OccBlock bl = load_block(x, er.l_ >> OCC_SHIFT);
	return fm_extend1_finish(x, ik, c, is_back, er, bk, bl);
}

// N independent backward extensions by the same base (candidates of one row of a backward sweep): all 4N block loads
// are issued before any of them is consumed, so one lane keeps N times as many index reads in flight.
#ifndef SEED_BWD_WAYS
#define SEED_BWD_WAYS 2
#endif
template <int N>
DEV void fm_extend1_back_xn(const DevIndex& x, const BiIntv* iv, int c, BiIntv* out)
{
	u64 kk[2 * N]; bool zz[2 * N];
#pragma unroll
	for (int t = 0; t < N; ++t) {
		const u64 k = iv[t].x0 - 1, l = iv[t].x0 - 1 + iv[t].size;
		zz[2 * t] = k == ~0ull; zz[2 * t + 1] = l == ~0ull;
		kk[2 * t] = zz[2 * t] ? 0 : k - (k >= x.primary);
		kk[2 * t + 1] = zz[2 * t + 1] ? 0 : l - (l >= x.primary);
	}
	OccBlock b[2 * N];
#pragma unroll
	for (int t = 0; t < 2 * N; ++t) b[t] = load_block(x, kk[t] >> OCC_SHIFT);
	const u64 L2c = c == 0 ? x.L2[0] : c == 1 ? x.L2[1] : c == 2 ? x.L2[2] : x.L2[3];
#pragma unroll
	for (int t = 0; t < N; ++t) {
		u64 tk[4], tl[4];
		occ4_in_block(x, b[2 * t], kk[2 * t], tk); occ4_in_block(x, b[2 * t + 1], kk[2 * t + 1], tl);
		if (zz[2 * t]) { tk[0] = tk[1] = tk[2] = tk[3] = 0; }
		if (zz[2 * t + 1]) { tl[0] = tl[1] = tl[2] = tl[3] = 0; }
		const u64 s1 = tl[1] - tk[1], s2 = tl[2] - tk[2], s3 = tl[3] - tk[3];
		const u64 kc = c == 0 ? tk[0] : c == 1 ? tk[1] : c == 2 ? tk[2] : tk[3];
		const u64 lc = c == 0 ? tl[0] : c == 1 ? tl[1] : c == 2 ? tl[2] : tl[3];
		const u64 above = (c < 3 ? s3 : 0) + (c < 2 ? s2 : 0) + (c < 1 ? s1 : 0);
		const u64 a0 = iv[t].x0, s = iv[t].size;
		BiIntv r;
		r.size = lc - kc;
		r.x0 = L2c + 1 + kc;
		r.x1 = iv[t].x1 + ((a0 <= x.primary && a0 + s - 1 >= x.primary) ? 1 : 0) + above;
		out[t] = r;
	}
}

DEV void fm_set_intv(const DevIndex& x, int c, BiIntv& ik)
{
	ik.x0 = x.L2[c] + 1; ik.size = x.L2[c + 1] - x.L2[c]; ik.x1 = x.L2[3 - c] + 1;
}

// ---- prefix tables -------------------------------------------------------------------------------
DEV uint4 cand_pack(const BiIntv& b, u32 end);
DEV void cand_unpack(const uint4 e, BiIntv& b, u32& end);
HOSTDEV u64 kmer_level_off(int L) { return (((u64)1 << (2 * L)) - 16) / 3; } // entries of the levels 2 .. L-1
// level L + 1 from level L (thread per entry of level L + 1): the forward extension of string p by base c is entry p * 4 + c
DEV void kmer_build_entry(const DevIndex& x, uint4* tab, int L1, u64 e)
{
	const u64 p = e >> 2; const int c = (int)(e & 3);
	BiIntv ik; u32 dummy;
	if (L1 == 2) fm_set_intv(x, (int)p, ik); else cand_unpack(tab[kmer_level_off(L1 - 1) + p], ik, dummy);
	const BiIntv ok = fm_extend1(x, ik, 3 - c, 0);
	tab[kmer_level_off(L1) + e] = cand_pack(ok, 0u);
}
// One forward extension of a scan that has consumed `plen` bases whose codes form `pidx` (base-4, first base most
// significant), by base code c < 4: from the tables when the extended string still fits them, else from the index.  Both
// sources are read unconditionally (the unused one at a harmless address) so that a warp whose lanes mix the two pays
// one memory latency, not two.
DEV BiIntv fm_extend_fwd(const DevIndex& x, const BiIntv& ik, int c, int plen, u32 pidx)
{
	const bool tab = x.kmer_tab != nullptr && plen >= 1 && plen < x.kmer_k;
	BiIntv in = ik;
	if (tab) { in.x0 = 1; in.x1 = 1; in.size = 1; }
	const uint4 te = ldg(x.kmer_tab ? x.kmer_tab + (tab ? kmer_level_off(plen + 1) + ((u64)pidx << 2 | (u32)c) : 0) : (const uint4*)x.bwt);
	BiIntv ok = fm_extend1(x, in, 3 - c, 0);
	if (tab) { u32 dummy; cand_unpack(te, ok, dummy); }
	return ok;
}

// (The interval of a string does not depend on the order in which it was grown, so the backward sweeps could read their
// short extended strings here as well; measured on the B200 that made them slower -- 14.3 -> 20.0 ms: they are bound by
// instruction issue and latency, not by DRAM traffic, and the second source adds instructions to every trip.)

// ---- SMEM discovery -----------------------------------------------------------------------------
// Working set of one read (one thread): the candidate list of the SMEM search in flight and the read's base codes.
// Both live in shared memory, interleaved over the threads of the block (element k of this thread's row is at
// k * stride, so every access stays on bank tid % 32); candidates beyond `lcap` spill to the thread's slice of a
// global pool.  A candidate is packed in 16 bytes: the low words of x0, x1 and size, then one byte each of their bits
// 32..39 and the query end -- enough for |T| < 2^40 and reads of <= 255 bases (longer reads keep everything global).
struct SeedWork {
	uint4* L; int lstride, lcap;   // shared candidate list (packed); lcap = 0 -> global only
	Intv* G; int gcap;             // global spill of the list, entries [lcap, lcap + gcap)
	u8* qs; int qstride;           // shared copy of the read (nullptr -> read through q)
};
struct IntvOut { Intv* v; int n, cap; bool overflow; };
DEV void out_push(IntvOut& o, const Intv& t) { if (o.n < o.cap) o.v[o.n++] = t; else o.overflow = true; }

DEV void cand_set(const SeedWork& w, int k, const BiIntv& b, u32 end)
{
	if (k < w.lcap) {
		uint4 e;
		e.x = (u32)b.x0; e.y = (u32)b.x1; e.z = (u32)b.size;
		e.w = (u32)(b.x0 >> 32 & 0xff) | (u32)(b.x1 >> 32 & 0xff) << 8 | (u32)(b.size >> 32 & 0xff) << 16 | end << 24;
		w.L[(size_t)k * w.lstride] = e;
	} else { Intv& t = w.G[k - w.lcap]; t.x0 = b.x0; t.x1 = b.x1; t.size = b.size; t.beg = 0; t.end = end; }
}
DEV void cand_get(const SeedWork& w, int k, BiIntv& b, u32& end)
{
	if (k < w.lcap) {
		const uint4 e = w.L[(size_t)k * w.lstride];
		b.x0 = (u64)e.x | (u64)(e.w & 0xff) << 32; b.x1 = (u64)e.y | (u64)(e.w >> 8 & 0xff) << 32; b.size = (u64)e.z | (u64)(e.w >> 16 & 0xff) << 32;
		end = e.w >> 24;
	} else { const Intv& t = w.G[k - w.lcap]; b.x0 = t.x0; b.x1 = t.x1; b.size = t.size; end = t.end; }
}

// mem_collect_intv (bwt_smem1a with max_intv = 0 for rounds 1 and 2, bwt_seed_strategy1 for round 3; App. B.3) as ONE
// loop with ONE fm_extend call site.  The three rounds and the forward / backward phases of SMEM discovery are states
// of a small machine: each trip of the loop first advances the machine, without touching the index, to the next
// interval that has to be extended (lane-divergent bookkeeping on shared memory and registers), then all lanes of the
// warp meet at the same fm_extend (two occurrence-block fetches, the expensive part) and consume its result.
// The backward phase compacts the candidate list in place: row i reads candidate j and writes at most j survivors, so
// the "previous" and "current" lists of the textbook formulation share one array.
enum { SM_NEXT_R1, SM_FWD, SM_BWD_ROW, SM_BWD_ITEM, SM_NEXT_R2, SM_NEXT_R3, SM_S3, SM_DONE, SM_FETCH, SM_ALLDONE };
enum { SMX_NONE, SMX_FWD, SMX_BWD, SMX_S3 };

DEV void intv_set(Intv& t, const BiIntv& b, u32 beg, u32 end) { t.x0 = b.x0; t.x1 = b.x1; t.size = b.size; t.beg = beg; t.end = end; }

// Reads are not assigned to lanes statically: `src.next` hands the lane its next read (a batch-wide counter), fetching is
// one more state of the machine, and a full-mask vote per trip is both the exit test and the point where the lanes
// of the warp meet again -- no lane idles while a neighbour works through a read with more extensions.
//
// PHASES selects what one launch does for a read: 3 = the three rounds in one machine; 1 = rounds 1 and 2 (SMEMs and
// re-seeding, two kinds of extension steps); 2 = round 3 alone (forward-only seeds: one kind of step, so the lanes of a
// warp run in lock-step).  Rounds 1/2 and round 3 are independent of each other and the interval list is ordered by
// (begin, end) at the end, so running them in two launches yields the same list.
template <int PHASES, class Src>
DEV void fm_collect_queue(const DevIndex& x, const MemOpt& opt, const SeedWork& wk, Src& src)
{
	const int split_len = (int)(opt.min_seed_len * opt.split_factor + .499);
	const int min_seed_len = opt.min_seed_len;
	const int lcap_all = wk.lcap + wk.gcap;
	IntvOut o; o.v = nullptr; o.n = 0; o.cap = 0; o.overflow = false;
	int len = 0; const u8* qg = nullptr;
	const u8* q = nullptr; size_t qst = 1;
#define Q(i_) q[(size_t)(i_) * qst]
	int st = SM_FETCH, round = 1, pend = SMX_NONE;
	int xx = 0;                 // rounds 1 and 3: next start position
	int k2 = 0, old_n = 0;      // round 2: cursor over the round-1 SMEMs
	int start = 0, i = 0, ret = 0, c = -1, j = 0, first_out = 0;
	u64 min_intv = 1, last_size = 0;
	BiIntv ik, ext, pcur; u32 ik_end = 0, pcur_end = 0;
	int n_prev = 0, n_curr = 0;
	u32 last_beg = 0; bool have_out = false;
	BiIntv okc; int ext_c = 0;
	int plen = 0; u32 pidx = 0; // bases consumed by the forward scan in flight and their codes (prefix tables: fm_extend_fwd)
	ik.x0 = ik.x1 = ik.size = 0; ext = ik; pcur = ik; okc = ik;
#define SM_PUSH(bi, e) do { if (n_curr < lcap_all) { cand_set(wk, n_curr, bi, e); ++n_curr; } else o.overflow = true; } while (0)
#define SM_END_FORWARD() do { \
		for (int a_ = 0; a_ < n_curr >> 1; ++a_) { /* longest match first */ \
			BiIntv t0_, t1_; u32 e0_, e1_; cand_get(wk, a_, t0_, e0_); cand_get(wk, n_curr - 1 - a_, t1_, e1_); \
			cand_set(wk, a_, t1_, e1_); cand_set(wk, n_curr - 1 - a_, t0_, e0_); } \
		{ BiIntv t0_; u32 e0_ = 0; if (n_curr > 0) cand_get(wk, 0, t0_, e0_); ret = (int)e0_; } \
		n_prev = n_curr; first_out = o.n; have_out = false; last_beg = 0; i = start - 1; st = SM_BWD_ROW; } while (0)
#define SM_EMIT_PCUR() do { if (n_curr == 0 && (!have_out || (u32)(i + 1) < last_beg)) { \
		Intv t_; intv_set(t_, pcur, (u32)(i + 1), pcur_end); have_out = true; last_beg = t_.beg; out_push(o, t_); } } while (0)
	for (;;) {
		// ---- advance to the next extension (or to the end of the batch)
		while (pend == SMX_NONE && st != SM_ALLDONE) {
			switch (st) {
			case SM_FETCH:
				if (!src.next(len, qg, o)) { st = SM_ALLDONE; break; }
				// the read's codes: shared copy when provided
				q = qg; qst = 1;
				if (wk.qs) { for (int t = 0; t < len; ++t) wk.qs[(size_t)t * wk.qstride] = qg[t]; q = wk.qs; qst = (size_t)wk.qstride; }
				round = 1; xx = 0; k2 = 0; old_n = 0; n_prev = n_curr = 0; have_out = false; last_beg = 0;
				if (PHASES == 2) { round = 3; st = opt.max_mem_intv > 0 ? SM_NEXT_R3 : SM_DONE; } // (src.next set o.n to what rounds 1/2 left)
				else st = SM_NEXT_R1;
				break;
			case SM_DONE: { // the read is finished: order its intervals by (beg, end) and hand them over
				// Equal keys are identical intervals, so the sort algorithm is unobservable here (unlike the hazard-H4 sorts).
				for (int a = 1; a < o.n && PHASES != 1; ++a) {
					Intv t = o.v[a];
					u64 key = (u64)t.beg << 32 | t.end;
					int b = a - 1;
					while (b >= 0 && ((u64)o.v[b].beg << 32 | o.v[b].end) > key) { o.v[b + 1] = o.v[b]; --b; }
					o.v[b + 1] = t;
				}
				src.finish(o);
				st = SM_FETCH;
				break; }
			case SM_NEXT_R1:
				while (xx < len && Q(xx) > 3) ++xx;
				if (xx >= len) { old_n = o.n; k2 = 0; round = 2; st = SM_NEXT_R2; break; }
				start = xx; min_intv = 1;
				fm_set_intv(x, Q(start), ik); ik_end = (u32)(start + 1); n_curr = 0; i = start + 1; plen = 1; pidx = (u32)Q(start);
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
				if (!go) { xx = 0; round = 3; st = (PHASES == 3 && opt.max_mem_intv > 0) ? SM_NEXT_R3 : SM_DONE; break; }
				if (Q(start) > 3) { ++k2; break; } // (an SMEM never holds an ambiguous base; kept for fidelity)
				fm_set_intv(x, Q(start), ik); ik_end = (u32)(start + 1); n_curr = 0; i = start + 1; plen = 1; pidx = (u32)Q(start);
				st = SM_FWD;
				break; }
			case SM_FWD: // forward phase: record the interval at every change of size
				if (i < len && Q(i) < 4) { ext = ik; ext_c = 3 - Q(i); pend = SMX_FWD; break; }
				SM_PUSH(ik, ik_end); // i == len, or an ambiguous base ends the match
				SM_END_FORWARD();
				break;
			case SM_BWD_ROW:
				if (i < -1) goto end_backward;
				c = i < 0 ? -1 : (Q(i) < 4 ? (int)Q(i) : -1);
				n_curr = 0; last_size = 0; j = 0;
				st = SM_BWD_ITEM;
				break;
			case SM_BWD_ITEM:
				if (j == n_prev) { // row finished
					if (n_curr == 0) goto end_backward;
					n_prev = n_curr;
					--i;
					st = SM_BWD_ROW;
					break;
				}
				cand_get(wk, j, pcur, pcur_end);
				if (c >= 0) { ext = pcur; ext_c = c; pend = SMX_BWD; break; }
				SM_EMIT_PCUR(); // no base to extend with: the candidate cannot grow
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
				while (xx < len && Q(xx) > 3) ++xx;
				if (xx >= len) { st = SM_DONE; break; }
				start = xx;
				fm_set_intv(x, Q(start), ik); i = start + 1; plen = 1; pidx = (u32)Q(start);
				st = SM_S3;
				break;
			case SM_S3: // bwt_seed_strategy1: shortest forward match of >= min_seed_len+1 bases with < max_mem_intv occurrences
				if (i >= len) { xx = len; st = SM_NEXT_R3; break; }
				if (Q(i) > 3) { xx = i + 1; st = SM_NEXT_R3; break; }
				ext = ik; ext_c = 3 - Q(i); pend = SMX_S3;
				break;
			default: st = SM_DONE; break;
			}
		}
		if (lanes_all(st == SM_ALLDONE)) break;
		if (pend == SMX_NONE) continue; // (this lane has run out of reads; its neighbours have not)
		// ---- the one place where the index is read (forward steps: the prefix tables while the scan is short enough)
		if (pend == SMX_BWD) okc = fm_extend1(x, ext, ext_c, 1);
		else { okc = fm_extend_fwd(x, ext, 3 - ext_c, plen, pidx); pidx = pidx << 2 | (u32)(3 - ext_c); ++plen; }
		// ---- consume
		if (pend == SMX_FWD) {
			bool stop = false;
			if (okc.size != ik.size) {
				SM_PUSH(ik, ik_end);
				if (okc.size < min_intv) stop = true;
			}
			if (!stop) { ik = okc; ik_end = (u32)(i + 1); ++i; st = SM_FWD; }
			else SM_END_FORWARD(); // (no second push: the scan ended before i == len)
		} else if (pend == SMX_BWD) {
			if (okc.size < min_intv) SM_EMIT_PCUR();
			else if (n_curr == 0 || okc.size != last_size) { cand_set(wk, n_curr, okc, pcur_end); ++n_curr; last_size = okc.size; } // n_curr <= j
			++j;
			st = SM_BWD_ITEM;
		} else { // SMX_S3
			if (okc.size < opt.max_mem_intv && i - start >= min_seed_len) {
				if (okc.size > 0) { Intv m; intv_set(m, okc, (u32)start, (u32)(i + 1)); out_push(o, m); }
				xx = i + 1; st = SM_NEXT_R3;
			} else { ik = okc; ++i; st = SM_S3; }
		}
		pend = SMX_NONE;
	}
#undef SM_PUSH
#undef SM_END_FORWARD
#undef SM_EMIT_PCUR
#undef Q
}

// ---- suffix-array resolution ----------------------------------------------------------------------
// bwt_sa: LF-walk from row k to a sampled row.  One sector touch per step (the BWT symbol and its rank
// live in the same 32-byte sector) + one 8-byte SA read.
// (one LF step from row k != primary, given the sector of its packed position kk = k - (k > primary))
DEV u64 fm_lf_finish(const DevIndex& x, const OccBlock& b, u64 kk)
{
	const u32 bit = (u32)(kk & 31);
	const u32 lo = (kk & 32) ? b.sym.y : b.sym.x, hi = (kk & 32) ? b.sym.w : b.sym.z;
	const int c = (int)((lo >> bit & 1u) | (hi >> bit & 1u) << 1);
	// occ(k, c) counts rows [0,k]; k > primary maps to packed position k-1 == kk, k < primary to k == kk
	u64 cnt[4];
	occ4_in_block(x, b, kk, cnt);
	return x.L2[c] + cnt[c];
}
DEV u64 fm_sa(const DevIndex& x, u64 k)
{
	u64 sa = 0, mask = (u64)x.sa_intv - 1;
	while (k & mask) {
		++sa;
		if (k == x.primary) { k = 0; continue; }
		const u64 kk = k - (k > x.primary); // position of this row's symbol in the packed BWT
		const OccBlock b = load_block(x, kk >> OCC_SHIFT);
		k = fm_lf_finish(x, b, kk);
	}
	return sa + ldg(&x.sa[k / x.sa_intv]);
}

// ---- SMEM discovery as separate forward and backward launches -----------------------------------------------------
// bwt_smem1a is a forward scan (extend to the right from x, remember the interval at every change of size) followed
// by a backward sweep over the remembered intervals.  In one per-read state machine the two kinds of steps -- and the
// bookkeeping between them -- make the lanes of a warp diverge.  The forward scans of round 1 do not depend on the
// backward sweeps (the next scan starts where the longest match of the previous one ended, which the forward scan
// itself knows), and the scans of round 2 only need the finished SMEMs of round 1.  So:
//   forward launch  (a lane per read)  every scan of the read back to back: one kind of step, candidates appended to
//                                      the read's slice of a global pool, one SeedTask per scan;
//   backward launch (a lane per task)  the sweep of one scan, compacting its candidates in place in the pool; SMEMs
//                                      are appended to the read's interval list with an atomic cursor (the list is
//                                      ordered by (begin, end) at the very end, so the order of arrival is free).
// Round 1 = forward<1> + backward, round 2 = forward<2> + backward, round 3 = fm_collect_queue<2> (which also orders
// the list).  Candidates are packed in 16 bytes (|T| < 2^40, reads <= 255 bases; other batches use fm_collect_queue).
struct SeedTask { u32 read, c_off, n_c, start; u64 min_intv; };
struct SeedSplitIO {
	const u8* seq; const u64* seq_off; i64 n_reads;
	i64 read_base;                    // the forward launch covers reads [read_base, read_base + n_reads): a batch is seeded in slices when its pool would be too large
	uint4* pool; u32 pool_cap;        // pool_cap candidates per read of the slice (row = read - read_base)
	u32* pool_used;                   // per read: candidates appended so far (round 2 continues behind round 1)
	SeedTask* tasks; u32 task_cap;    // light tasks fill the array from the front, scans with many candidates from the back (swept first: longest jobs first)
	u32* n_tasks;                     // [0],[1]: ONE 64-bit reservation counter (low word = light scans, high word = heavy scans);
	                                  // [2],[3]: the two counts at the moment the list became full (written by the reservation that found it so)
	Intv* intv; int intv_cap; u32* n_intv;
	u32* err; int smem_cap;
	u32* fb_list; u32* n_fb;          // reads whose scans did not fit the pool / task list: redone by fm_collect_queue<1>
};
#ifndef SEED_HEAVY_CANDS
#define SEED_HEAVY_CANDS 24
#endif
#define SEED_POOL_OVERFLOW 0xffffffffu // value of pool_used[r] for such a read (its tasks are skipped)
// Reservation of a task slot.  Both counts live in one 64-bit word, so the atomic add returns a consistent pair: the
// first task_cap reservations (in the order of the atomics) succeed and own distinct slots, every later one fails --
// its read is redone as a whole by the one-machine kernel (fb_list) -- and the one that finds the list exactly full
// records the two counts, which is what the backward launches then run (the counters themselves overshoot).
DEV bool seed_task_reserve(const SeedSplitIO& io, bool heavy, u32& slot)
{
	const u64 old = atomic_add_u64((u64*)io.n_tasks, heavy ? (u64)1 << 32 : (u64)1);
	const u32 nl = (u32)old, nh = (u32)(old >> 32);
	const u64 tot = (u64)nl + nh;
	if (tot >= io.task_cap) { if (tot == io.task_cap) { io.n_tasks[2] = nl; io.n_tasks[3] = nh; } return false; }
	slot = heavy ? io.task_cap - 1 - nh : nl;
	return true;
}
DEV void seed_task_counts(const SeedSplitIO& io, u32& nl, u32& nh)
{
	nl = io.n_tasks[0]; nh = io.n_tasks[1];
	if ((u64)nl + nh > io.task_cap) { nl = io.n_tasks[2]; nh = io.n_tasks[3]; }
}
DEV uint4 cand_pack(const BiIntv& b, u32 end)
{
	uint4 e;
	e.x = (u32)b.x0; e.y = (u32)b.x1; e.z = (u32)b.size;
	e.w = (u32)(b.x0 >> 32 & 0xff) | (u32)(b.x1 >> 32 & 0xff) << 8 | (u32)(b.size >> 32 & 0xff) << 16 | end << 24;
	return e;
}
DEV void cand_unpack(const uint4 e, BiIntv& b, u32& end)
{
	b.x0 = (u64)e.x | (u64)(e.w & 0xff) << 32; b.x1 = (u64)e.y | (u64)(e.w >> 8 & 0xff) << 32; b.size = (u64)e.z | (u64)(e.w >> 16 & 0xff) << 32;
	end = e.w >> 24;
}

// ROUND 1: scans from x = 0, each starting where the longest match of the previous one ended (min_intv 1).
// ROUND 2: one scan from the middle of every round-1 SMEM of >= split_len bases and <= split_width occurrences
//          (min_intv = its occurrences + 1).
template <int ROUND>
DEV void fm_forward_queue(const DevIndex& x, const MemOpt& opt, const SeedSplitIO& io, u32* counter)
{
	const int split_len = (int)(opt.min_seed_len * opt.split_factor + .499);
	enum { F_FETCH, F_NEXT, F_SCAN, F_ALLDONE };
	int st = F_FETCH; bool pend = false;
	i64 r = -1; int len = 0; const u8* q = nullptr;
	int xx = 0, k2 = 0, n1 = 0, start = 0, i = 0, ext_c = 0;
	u32 used = 0, c_off = 0, n_c = 0, ik_end = 0, last_end = 0;
	u64 min_intv = 1;
	int plen = 0; u32 pidx = 0; // bases consumed by the scan in flight and their codes (prefix tables: fm_extend_fwd)
	BiIntv ik; ik.x0 = ik.x1 = ik.size = 0;
	uint4* pool = nullptr;
	bool ovf = false; // the read does not fit the pool / the task list: it is handed to the one-machine formulation
#define F_PUSH() do { if (used < io.pool_cap) { pool[used] = cand_pack(ik, ik_end); ++used; ++n_c; last_end = ik_end; } else ovf = true; } while (0)
#define F_END_SCAN() do { \
		if (n_c && !ovf) { u32 slot_ = 0; \
			if (seed_task_reserve(io, n_c >= SEED_HEAVY_CANDS, slot_)) { SeedTask t_; t_.read = (u32)r; t_.c_off = c_off; t_.n_c = n_c; t_.start = (u32)start; t_.min_intv = min_intv; io.tasks[slot_] = t_; } \
			else ovf = true; } \
		if (ROUND == 1) xx = (int)last_end; else ++k2; \
		st = ovf ? F_FETCH : F_NEXT; } while (0)
	for (;;) {
		while (!pend && st != F_ALLDONE) {
			switch (st) {
			case F_FETCH:
				if (r >= 0) {
					if (ovf) { io.pool_used[r] = SEED_POOL_OVERFLOW; const u32 slot_ = atomic_add_u32(io.n_fb, 1u); io.fb_list[slot_] = (u32)r; ovf = false; }
					else io.pool_used[r] = used;
				}
				for (;;) {
					r = io.read_base + (i64)atomic_add_u32(counter, 1u);
					if (r >= io.read_base + io.n_reads) break;
					len = (int)(io.seq_off[r + 1] - io.seq_off[r] - 1);
					if (len >= opt.min_seed_len && len <= io.smem_cap && (ROUND == 1 || io.pool_used[r] != SEED_POOL_OVERFLOW)) break;
					if (ROUND == 1) io.pool_used[r] = 0;
				}
				if (r >= io.read_base + io.n_reads) { r = -1; st = F_ALLDONE; break; }
				q = io.seq + io.seq_off[r];
				pool = io.pool + (size_t)(r - io.read_base) * io.pool_cap;
				used = 0; // (round 1's candidates are dead once its sweeps have run)
				xx = 0; k2 = 0; n1 = ROUND == 1 ? 0 : (int)(io.n_intv[r] < (u32)io.intv_cap ? io.n_intv[r] : (u32)io.intv_cap);
				st = F_NEXT;
				break;
			case F_NEXT:
				if (ROUND == 1) {
					while (xx < len && q[xx] > 3) ++xx;
					if (xx >= len) { st = F_FETCH; break; }
					start = xx; min_intv = 1;
				} else {
					const Intv* v = io.intv + (size_t)r * io.intv_cap;
					bool go = false;
					while (k2 < n1) {
						const Intv p = v[k2];
						if ((int)(p.end - p.beg) < split_len || p.size > (u64)opt.split_width) { ++k2; continue; }
						start = (int)((p.beg + p.end) >> 1); min_intv = p.size + 1;
						go = true;
						break;
					}
					if (!go) { st = F_FETCH; break; }
					if (q[start] > 3) { ++k2; break; } // (an SMEM never holds an ambiguous base; kept for fidelity)
				}
				fm_set_intv(x, q[start], ik); ik_end = (u32)(start + 1); i = start + 1; plen = 1; pidx = (u32)q[start];
				c_off = used; n_c = 0; last_end = 0;
				st = F_SCAN;
				break;
			case F_SCAN: // record the interval at every change of size
				if (i < len && q[i] < 4) { ext_c = 3 - q[i]; pend = true; break; }
				F_PUSH(); // i == len, or an ambiguous base ends the match
				F_END_SCAN();
				break;
			default: st = F_ALLDONE; break;
			}
		}
		if (lanes_all(st == F_ALLDONE)) break;
		if (!pend) continue;
		const BiIntv okc = fm_extend_fwd(x, ik, 3 - ext_c, plen, pidx);
		pidx = pidx << 2 | (u32)(3 - ext_c); ++plen;
		bool stop = false;
		if (okc.size != ik.size) {
			F_PUSH();
			if (okc.size < min_intv) stop = true;
		}
		if (!stop) { ik = okc; ik_end = (u32)(i + 1); ++i; }
		else F_END_SCAN(); // (no second push: the scan ended before i == len)
		pend = false;
	}
#undef F_PUSH
#undef F_END_SCAN
}

// The backward sweep of one forward scan.  Candidate a of the textbook list (longest match first) is pool slot
// c_off + n_c - 1 - a; row i reads candidate j and writes at most j survivors, so the list is compacted in place.
//
// A lane per scan.  The candidates of a scan (< SEED_HEAVY_CANDS of them) are copied into the lane's column of shared
// memory when the scan is taken, so the bookkeeping between two index reads -- which runs with the lanes of a warp
// diverged -- touches no global memory.  Lanes that have finished their scan wait (state B_IDLE) until SEED_REFILL_MIN
// lanes of the warp are idle and then take new scans TOGETHER: one reservation, one round of task / candidate loads for
// all of them, instead of one full memory latency paid by the whole warp for every single lane's fetch.
#ifndef SEED_REFILL_MIN
#define SEED_REFILL_MIN 8
#endif
HOSTDEV size_t seed_bwd_smem_bytes(int block_threads) { return (size_t)block_threads * SEED_HEAVY_CANDS * 16; }
#ifdef B200_SIM
static uint4 g_sim_bwd_cands[SEED_HEAVY_CANDS];
#endif
DEV void fm_backward_queue(const DevIndex& x, const MemOpt& opt, const SeedSplitIO& io, i64 n_tasks, u32* counter)
{
	enum { B_IDLE, B_ROW, B_ITEM, B_ALLDONE };
#ifdef B200_SIM
	uint4* const sc = g_sim_bwd_cands; const size_t SS = 1;
#else
	extern __shared__ __align__(16) u8 tsm_[];
	uint4* const sc = (uint4*)tsm_ + threadIdx.x; const size_t SS = (size_t)blockDim.x;
#endif
#define CL(a_) sc[(size_t)(a_) * SS] /* candidate a of the current row's list */
	const int lane = simt_lane();
	int st = B_IDLE; int pend = 0; // number of candidates of the current row in flight (0, 1 or 2)
	SeedTask tk; tk.read = 0; tk.c_off = 0; tk.n_c = 0; tk.start = 0; tk.min_intv = 1;
	const u8* q = nullptr;
	int i = 0, c = -1, j = 0, n_prev = 0, n_curr = 0;
	u64 last_size = 0; u32 last_beg = 0, pcur_end = 0; bool have_out = false;
	BiIntv pcur; pcur.x0 = pcur.x1 = pcur.size = 0;
	BiIntv pcs[SEED_BWD_WAYS]; u32 pes[SEED_BWD_WAYS];
	for (int t = 0; t < SEED_BWD_WAYS; ++t) { pcs[t] = pcur; pes[t] = 0; }
#define B_EMIT_PCUR() do { if (n_curr == 0 && (!have_out || (u32)(i + 1) < last_beg)) { \
		have_out = true; last_beg = (u32)(i + 1); \
		if ((int)(pcur_end - (u32)(i + 1)) >= opt.min_seed_len) { /* the length filter of the end of the sweep */ \
			const u32 slot_ = atomic_add_u32(&io.n_intv[tk.read], 1u); \
			if (slot_ < (u32)io.intv_cap) { Intv t_; t_.x0 = pcur.x0; t_.x1 = pcur.x1; t_.size = pcur.size; t_.beg = (u32)(i + 1); t_.end = pcur_end; io.intv[(size_t)tk.read * io.intv_cap + slot_] = t_; } \
			else atomic_or_u32(io.err, ERR_INTV_OVERFLOW); } } } while (0)
	for (;;) {
		// ---- idle lanes take new scans together
		const u32 idle = lanes_ballot(st == B_IDLE), busy = lanes_ballot(st == B_ROW || st == B_ITEM);
		if (idle != 0u && (__popc(idle) >= SEED_REFILL_MIN || busy == 0u)) {
			const int leader = __ffs(idle) - 1;
			u32 base = 0;
			if (lane == leader) base = atomic_add_u32(counter, (u32)__popc(idle));
			base = lanes_shfl(base, leader);
			if (st == B_IDLE) {
				const i64 t = (i64)base + __popc(idle & ((1u << lane) - 1u));
				if (t >= n_tasks) st = B_ALLDONE;
				else {
					tk = io.tasks[t]; // (scans with many candidates sit at the other end of the array: fm_backward_task_warp)
					if (io.pool_used[tk.read] != SEED_POOL_OVERFLOW && tk.n_c <= (u32)SEED_HEAVY_CANDS) { // (else the read is redone as a whole: stay idle)
						q = io.seq + io.seq_off[tk.read];
						const uint4* src = io.pool + (size_t)((i64)tk.read - io.read_base) * io.pool_cap + tk.c_off + (tk.n_c - 1);
						for (int a = 0; a < (int)tk.n_c; ++a) CL(a) = src[-a];
						n_prev = (int)tk.n_c; i = (int)tk.start - 1; have_out = false; last_beg = 0;
						st = B_ROW;
					}
				}
			}
		}
		if (lanes_ballot(st != B_ALLDONE) == 0u) break;
		// ---- bookkeeping up to the next index read (shared memory and registers only, but for the read's bases)
		while (!pend && (st == B_ROW || st == B_ITEM)) {
			if (st == B_ROW) {
				if (i < -1) { st = B_IDLE; break; }
				c = i < 0 ? -1 : (q[i] < 4 ? (int)q[i] : -1);
				n_curr = 0; last_size = 0; j = 0;
				st = B_ITEM;
			} else {
				if (j == n_prev) { // row finished
					if (n_curr == 0) { st = B_IDLE; break; }
					n_prev = n_curr; --i;
					st = B_ROW;
					continue;
				}
				cand_unpack(CL(j), pcur, pcur_end);
				if (c >= 0) { // the candidates of a row are independent: up to SEED_BWD_WAYS of them are extended per trip
					pcs[0] = pcur; pes[0] = pcur_end;
					pend = n_prev - j < SEED_BWD_WAYS ? n_prev - j : SEED_BWD_WAYS;
#pragma unroll
					for (int t = 1; t < SEED_BWD_WAYS; ++t) if (t < pend) cand_unpack(CL(j + t), pcs[t], pes[t]);
					break;
				}
				B_EMIT_PCUR(); // no base to extend with: the candidate cannot grow
				++j;
			}
		}
		if (lanes_ballot(pend != 0) == 0u) continue;
		if (!pend) continue;
		BiIntv oks[SEED_BWD_WAYS];
		// (two code paths on purpose: in the late rows of a sweep no lane of the warp has a second candidate and the warp
		// skips the wide one altogether -- folding the two into one measured 15 % slower)
		if (pend == SEED_BWD_WAYS) fm_extend1_back_xn<SEED_BWD_WAYS>(x, pcs, c, oks);
		else if (SEED_BWD_WAYS > 2 && pend >= 2) { fm_extend1_back_xn<2>(x, pcs, c, oks); for (int t = 2; t < pend; ++t) oks[t] = fm_extend1(x, pcs[t], c, 1); }
		else oks[0] = fm_extend1(x, pcs[0], c, 1);
#pragma unroll
		for (int t = 0; t < SEED_BWD_WAYS; ++t) { // consumed in list order, exactly as one at a time
			if (t >= pend) break;
			pcur = pcs[t]; pcur_end = pes[t];
			const BiIntv okc = oks[t];
			if (okc.size < tk.min_intv) B_EMIT_PCUR();
			else if (n_curr == 0 || okc.size != last_size) { CL(n_curr) = cand_pack(okc, pcur_end); ++n_curr; last_size = okc.size; } // n_curr <= j (in-place compaction only writes slots <= j)
			++j;
		}
		pend = 0;
	}
#undef B_EMIT_PCUR
#undef CL
}

// The backward sweep of a scan with MANY candidates (low-complexity and repeat-derived reads: up to one candidate per
// base), one warp per task.  Swept by a single lane such a task is ~10^4 dependent index lookups -- milliseconds, the
// tail of the whole launch.  The extensions of the candidates of one row are independent of each other, so the lanes
// extend 32 candidates at once; only the bookkeeping that follows (which candidate survives, which one is reported)
// is order-dependent, and it needs nothing but the new interval sizes: every lane replays it over the 32 sizes
// (shuffles) and then carries out the action that fell to its own candidate.
template <int W>
DEVNI void fm_backward_task_warp(const DevIndex& x, const MemOpt& opt, const SeedSplitIO& io, const SeedTask& tk)
{
	LANE_DECL;
	if (io.pool_used[tk.read] == SEED_POOL_OVERFLOW) return; // the read is redone as a whole
	const u8* q = io.seq + io.seq_off[tk.read];
	uint4* cl = io.pool + (size_t)((i64)tk.read - io.read_base) * io.pool_cap + tk.c_off + (tk.n_c - 1); // cl[-a] = candidate a
	int n_prev = (int)tk.n_c; u32 last_beg = 0; bool have_out = false;
	for (int i = (int)tk.start - 1; i >= -1; --i) {
		const int c = i < 0 ? -1 : (q[i] < 4 ? (int)q[i] : -1);
		int n_curr = 0; u64 last_size = 0;
		for (int j0 = 0; j0 < n_prev; j0 += W) {
			WV(BiIntv, pc); WV(u32, pe); WV(BiIntv, ok); WV(u64, sz); WV(int, act); WV(int, idx);
			FOR_LANES {
				const int j = j0 + lane;
				L_(sz) = 0; L_(act) = 0; L_(idx) = 0; L_(pe) = 0;
				L_(pc).x0 = L_(pc).x1 = L_(pc).size = 0; L_(ok) = L_(pc);
				if (j < n_prev) {
					cand_unpack(cl[-j], L_(pc), L_(pe));
					if (c >= 0) { L_(ok) = fm_extend1(x, L_(pc), c, 1); L_(sz) = L_(ok).size; }
				}
			}
			WARP_SYNC(); // every candidate of the chunk is in registers: slots up to j0 + 31 may now be overwritten
			const int cnt = n_prev - j0 < W ? n_prev - j0 : W;
			for (int t = 0; t < cnt; ++t) { // the scalar bookkeeping, replayed by all lanes on the broadcast sizes
				const u64 s_t = wv_get(sz, t);
				int a = 0, at = 0;
				if (c < 0 || s_t < tk.min_intv) { // the candidate cannot grow: report it if it is the longest of its end so far
					if (n_curr == 0 && (!have_out || (u32)(i + 1) < last_beg)) { have_out = true; last_beg = (u32)(i + 1); a = 1; }
				} else if (n_curr == 0 || s_t != last_size) { a = 2; at = n_curr; ++n_curr; last_size = s_t; }
				FOR_LANES { if (lane == t) { L_(act) = a; L_(idx) = at; } }
			}
			FOR_LANES {
				if (L_(act) == 2) cl[-L_(idx)] = cand_pack(L_(ok), L_(pe));
				else if (L_(act) == 1 && (int)(L_(pe) - (u32)(i + 1)) >= opt.min_seed_len) { // the length filter of the end of the sweep
					const u32 slot = atomic_add_u32(&io.n_intv[tk.read], 1u);
					if (slot < (u32)io.intv_cap) { Intv t_; t_.x0 = L_(pc).x0; t_.x1 = L_(pc).x1; t_.size = L_(pc).size; t_.beg = (u32)(i + 1); t_.end = L_(pe); io.intv[(size_t)tk.read * io.intv_cap + slot] = t_; }
					else atomic_or_u32(io.err, ERR_INTV_OVERFLOW);
				}
			}
			WARP_SYNC();
		}
		if (n_curr == 0) break;
		n_prev = n_curr;
	}
}
