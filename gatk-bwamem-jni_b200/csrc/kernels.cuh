// kernels.cuh -- the kernels of one createAlignments batch (bodies; the launch wrappers are in engine.cu).
//
// Stage order (reference entry /root/reference/src/main/c/jnibwa.c:197-235):
//   encode  -> seed (SMEM, 3 rounds) -> [scan] -> sal (SA lookups) -> chain (+filter) -> extend (+dedup)
//   -> [PE: pestat candidates -> host mem_pestat] -> capacity/scan/compact -> finalize (SE | PE) -> size/scan -> write
// Thread kernels take one work item per thread; warp kernels one per warp.
#pragma once
#include "fm.cuh"
#include "chain.cuh"
#include "finalize.cuh"
#include "rescue.cuh"
#include "sw_thread.cuh"

struct BatchParams {
	DevIndex x;
	MemOpt opt;
	// reads
	int n_reads;            // reads in the call
	int n_units;            // reads (SE) or pairs (PE)
	const u8* blob;         // raw bytes of the request (after the u32 count)
	u8* seq;                // nt4 codes, same offsets as blob
	const u64* seq_off;     // n_reads+1 byte offsets (each read followed by its NUL)
	u64 blob_len;
	// seeding
	Intv* intv; int intv_cap; u32* n_intv; u32* n_seed; // per read
	Intv* smem_scratch; int smem_cap;                  // per thread: smem_cap spill entries of the candidate list
	int seed_shared;                                    // candidate list (+ read codes) in shared memory (fm.cuh SeedWork)
	uint4* seed_pool; u32 seed_pool_cap; u32* seed_pool_used; SeedTask* seed_tasks; u32* n_seed_tasks; u32 seed_task_cap; // forward / backward split (fm.cuh)
	u32* seed_fb; u32* n_seed_fb;                        // reads redone by the one-machine formulation (pool / task overflow)
	int seed_lcap, seed_shared_read;                    // shared candidates per thread; whether the read's codes are staged in shared memory too
	u64* seed_off;          // n_reads+1 (exclusive scan of n_seed)
	u64 n_seed_total;
	i64* sd_rbeg;           // [T]
	// chaining pools, one slot per lookup
	Seed* S; ChainW* C; u32* ord; u32* tmp32; ChainK* ck; Seed* ks; u64* srt; Reg* regs_tmp;
	u32* n_ck; u32* n_ks; u32* n_reg;   // per read
	// final regions
	u32* reg_cap; u64* reg_off; Reg* regs; Reg* snap; u32* ztmp; Pair64* pairv; AlnRec* recs; u32* n_rec;
	i64* pe_cand;           // per pair
	// mate rescue (rescue.cuh)
	u32* nb; SwTask* sw_tasks; u8* sw_state; u32* sw_list; u32* sw_n_list; u32* max_nb; u32* multi; u32* n_multi; int round; int rescued;
	u32* rs_heavy; u32* n_rs_heavy; // rescue replay by a warp (rescue.cuh rescue_is_heavy)
	// thread-per-task rescue alignments (sw_thread.cuh): score tables, shared-memory words per array, global scratch, and the
	// lists of tasks that need the reverse pass / stay with the warp-per-task kernel
	u64 swtab[4]; int sw_shift, sw_mmax, tsw_capw, tsw_brow_cap; u32* tsw_hmax; u64* tsw_brow; u32* sw_rev; u32* n_sw_rev; u32* sw_fb; u32* n_sw_fb;
	Reg* ws_rt; i64* ws_rkey; u32* ws_rrank; int ws_rt_cap; // per-warp region-list scratch (warp_sort_dedup_fast)
	u8* spec_state; u32* spec_task; u32* n_reg0; int force_fb; // speculative rescue (rescue.cuh); multi / n_multi double as the fallback list
	PeStat pes[4];
	const double* pair_tab; i64 pair_tab_off[4];
	// output
	u32* var; u32 var_cap; u32* var_cursor;
	u32* out_words; u64* out_off; u32* out;
	const char* ctg_names; const u32* ctg_name_off;
	// warp scratch (per resident warp)
	u8* ws_z; size_t ws_z_cap; u32* ws_cigar; int ws_cigar_cap; char* ws_md; int ws_md_cap; char* ws_xa; int ws_xa_cap; u64* ws_brow; int ws_brow_cap;
	int wm_cap;             // per-warp DP column capacity
	int wm_smem_cap;        // ... as far as the rows live in shared memory: wm_cap, or 0 when
	u8* ws_rows; size_t ws_rows_stride; // the rows of a warp slot live in global memory (batches of long reads: warp_smem_bytes(wm_cap) exceeds an SM)
	// thread scratch (per resident thread of the thread-per-item kernels): DP rows, traceback, cigar, MD, XA
	u8* t_buf; size_t t_stride; int t_z_cap, t_cigar_cap, t_md_cap, t_xa_cap;
	u32* redo_list; u32* n_redo; // items a thread-level kernel handed over to the warp-level kernel
	u32* heavy_list; u32* n_heavy; int light_max_seeds, light_max_chains; // extension: reads with many seeds go to the warp-level kernel
	int tx_chunk;                           // thread-level extension: cells of one DP row per visit of the row site
	u32* flt_tasks; u32* n_flt_tasks; // mem_flt_chained_seeds: (read, seed slot) pairs
	ExtTask* ext_tasks; u32* n_ext_tasks; ExtRes* ext_res; i64* ck_rmax; // per-seed extension tasks, results, per-chain windows
	u32* ext_key_l; u32* ext_key_r; const u32* ext_order_l; const u32* ext_order_r; // query length of a task's left / right side; task indices ordered by it (nullptr: unordered)
	int ext_sides;          // the side-task route (extend_thread.cuh thread_extend_sides) is in use: tasks carry keys
	u32 ext_lo, ext_hi;     // one side launch covers positions [ext_lo, ext_hi) of the order (a class of query lengths)
	int ext_rw;             // ... with shared-memory rows of ext_rw cells per thread (query lengths < ext_rw)
	i64 seed_read_base;     // rounds 1/2 of the seeding run over reads [seed_read_base, ...) of the batch (slices: engine_impl.inl)
	u32* chain_heavy; u32* n_chain_heavy; int chain_heavy_seeds; // reads chained by a warp
	int chain_listed;       // the list of those reads was written beforehand (body_chain_classify): body_chain only skips them
	GaTask* ga_tasks; GaRes* ga_res; u32* n_ga_tasks; u32 ga_cap; u32* ga_item0; u32* ga_itemn; // two-pass finalize (galign_thread.cuh)
	int final_heavy_hits;   // finalize items with more hits than this skip the thread-level passes (warp-level kernel)
	u32* ga_key; const u32* ga_order; u32* ga_fb; u32* n_ga_fb; // cost key per task, tasks ordered by it, tasks left to the warp-per-task kernel
	int use_ext_res;        // wbody_extend consumes ext_res (set for the replay launch of the heavy reads only)
	int t_ext_rw;           // thread-level extension: shared-memory words per thread (DP row, doubling as the two NW rings)
	u64 rowpack[5];         // byte c of rowpack[t] = mat[t*5 + c] (score of query code c against target base t)
	u32* err; Counters* cnt; u32* work_counter;
};

#include "extend_thread.cuh"
#include "galign_thread.cuh"

// ---- encode: ASCII -> nt4 (bytes already < 4 are taken as codes; hazard H12) --------------------
DEV u8 nt4_of(u8 c)
{
	if (c < 4) return c;
	switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; case '-': return 5; default: return 4; }
}
DEV void body_encode(const BatchParams& P, i64 i) { P.seq[i] = nt4_of(P.blob[i]); }

// ---- seed: one thread per read -------------------------------------------------------------------
#define SEED_LCAP 16   // candidates per thread kept in shared memory (16 bytes each); the rest spill to the global pool
HOSTDEV size_t seed_block_smem_bytes(int block_threads, int max_len, int lcap, int with_read) { return (size_t)block_threads * ((size_t)lcap * 16 + (with_read ? (size_t)((max_len + 1 + 3) & ~3) : 0)); }
#ifdef B200_SIM
static uint4 g_sim_seed_l[SEED_LCAP];
#endif
// hands reads to the lanes of a seeding kernel and takes their interval lists back.  PHASES as in fm_collect_queue:
// 1 leaves the unsorted list of rounds 1/2 in place for the round-3 launch (2), which appends, sorts and counts.
template <int PHASES>
struct SeedSrc {
	const BatchParams& P; i64 n; u32* counter; i64 r;
	DEVM SeedSrc(const BatchParams& P_, i64 n_, u32* c_) : P(P_), n(n_), counter(c_), r(-1) {}
	const u32* list = nullptr; // when set, the work items are list[0..n)
	DEVM bool next(int& len, const u8*& seq, IntvOut& o)
	{
		for (;;) {
			r = (i64)atomic_add_u32(counter, 1u);
			if (r >= n) return false;
			if (list) r = (i64)list[r];
			len = (int)(P.seq_off[r + 1] - P.seq_off[r] - 1);
			if (len >= P.opt.min_seed_len && len <= P.smem_cap) break;
			if (len > P.smem_cap && PHASES != 2) atomic_or_u32(P.err, ERR_QLEN_OVERFLOW);
			P.n_intv[r] = 0; P.n_seed[r] = 0; // no seed can exist
		}
		seq = P.seq + P.seq_off[r];
		o.v = P.intv + (size_t)r * P.intv_cap; o.n = PHASES == 2 ? (int)P.n_intv[r] : 0; o.cap = P.intv_cap; o.overflow = false;
		if (o.n > o.cap) { o.n = o.cap; o.overflow = true; } // (the sweeps count what they could not store: the batch is run again with more room)
		return true;
	}
	DEVM void finish(const IntvOut& o)
	{
		if (o.overflow) atomic_or_u32(P.err, ERR_INTV_OVERFLOW);
		P.n_intv[r] = (u32)o.n;
		if (PHASES == 1) return;
		u32 ns = 0;
		for (int i = 0; i < o.n; ++i) ns += (u32)(o.v[i].size < (u64)P.opt.max_occ ? o.v[i].size : (u64)P.opt.max_occ);
		P.n_seed[r] = ns;
	}
};
template <int PHASES>
DEV void seed_queue(const BatchParams& P, i64 n, u32* counter, i64 tid, const u32* list = nullptr)
{
	SeedWork wk;
	wk.G = P.smem_scratch + (size_t)tid * P.smem_cap; wk.gcap = P.smem_cap;
	wk.L = nullptr; wk.lstride = 0; wk.lcap = 0; wk.qs = nullptr; wk.qstride = 0;
	if (P.seed_shared) { // packed shared-memory working set (reads of <= 255 bases, |T| < 2^40)
#ifdef B200_SIM
		wk.L = g_sim_seed_l; wk.lstride = 1; wk.lcap = SEED_LCAP;
#else
		extern __shared__ __align__(16) u8 tsm_[];
		wk.L = (uint4*)tsm_ + threadIdx.x; wk.lstride = (int)blockDim.x; wk.lcap = P.seed_lcap;
		if (P.seed_shared_read) { wk.qs = tsm_ + (size_t)blockDim.x * P.seed_lcap * 16 + threadIdx.x; wk.qstride = (int)blockDim.x; }
#endif
	}
	SeedSrc<PHASES> src(P, n, counter);
	src.list = list;
	fm_collect_queue<PHASES>(P.x, P.opt, wk, src);
}
DEV SeedSplitIO make_seed_split_io(const BatchParams& P, i64 n_reads)
{
	SeedSplitIO io;
	io.seq = P.seq; io.seq_off = P.seq_off; io.n_reads = n_reads; io.read_base = P.seed_read_base;
	io.pool = P.seed_pool; io.pool_cap = P.seed_pool_cap; io.pool_used = P.seed_pool_used;
	io.tasks = P.seed_tasks; io.n_tasks = P.n_seed_tasks; io.task_cap = P.seed_task_cap;
	io.intv = P.intv; io.intv_cap = P.intv_cap; io.n_intv = P.n_intv;
	io.err = P.err; io.smem_cap = P.smem_cap;
	io.fb_list = P.seed_fb; io.n_fb = P.n_seed_fb;
	return io;
}
// rounds 1 and 2 as forward scans (a lane per read) and backward sweeps (a lane per scan): fm.cuh
DEV void body_seedF1_q(const BatchParams& P, i64 n, u32* counter, i64 tid) { (void)tid; SeedSplitIO io = make_seed_split_io(P, n); fm_forward_queue<1>(P.x, P.opt, io, counter); }
DEV void body_seedF2_q(const BatchParams& P, i64 n, u32* counter, i64 tid) { (void)tid; SeedSplitIO io = make_seed_split_io(P, n); fm_forward_queue<2>(P.x, P.opt, io, counter); }
DEV void body_seedB_q(const BatchParams& P, i64 n, u32* counter, i64 tid) { (void)tid; (void)n; SeedSplitIO io = make_seed_split_io(P, P.n_reads); u32 nl, nh; seed_task_counts(io, nl, nh);
	fm_backward_queue(P.x, P.opt, io, (i64)nl, counter); }
template <int W>
DEV void wbody_seedB_heavy(const BatchParams& P, i64 item, WarpMem& wm, WarpScratch& ws)
{
	(void)wm; (void)ws;
	SeedSplitIO io = make_seed_split_io(P, P.n_reads);
	u32 nl, nh; seed_task_counts(io, nl, nh);
	if ((u32)item >= nh) return; // (a reservation past the end of the task list: that read is redone as a whole)
	const SeedTask tk = P.seed_tasks[P.seed_task_cap - 1 - (u32)item];
	fm_backward_task_warp<W>(P.x, P.opt, io, tk);
}
DEV void body_seedB2_q(const BatchParams& P, i64 n, u32* counter, i64 tid) { body_seedB_q(P, n, counter, tid); }
DEV void body_seed_q(const BatchParams& P, i64 n, u32* counter, i64 tid) { seed_queue<3>(P, n, counter, tid); }
// rounds 1/2 of the reads the split launches could not hold (list written by the forward launches)
DEV void body_seed12_fb_q(const BatchParams& P, i64 n, u32* counter, i64 tid) { (void)n; seed_queue<1>(P, (i64)*P.n_seed_fb, counter, tid, P.seed_fb); }    // all rounds in one launch
DEV void body_seed12_q(const BatchParams& P, i64 n, u32* counter, i64 tid) { seed_queue<1>(P, n, counter, tid); }  // SMEMs + re-seeding
DEV void body_seed3_q(const BatchParams& P, i64 n, u32* counter, i64 tid) { seed_queue<2>(P, n, counter, tid); }   // forward-only seeds, sort, count

// ---- sal: one thread per lookup ------------------------------------------------------------------
DEV void body_sal(const BatchParams& P, i64 t)
{
	// read owning slot t: last r with seed_off[r] <= t
	i64 lo = 0, hi = P.n_reads;
	while (hi - lo > 1) { i64 mid = (lo + hi) >> 1; if (P.seed_off[mid] <= (u64)t) lo = mid; else hi = mid; }
	i64 r = lo;
	u64 rel = (u64)t - P.seed_off[r];
	const Intv* iv = P.intv + (size_t)r * P.intv_cap;
	int n = (int)P.n_intv[r];
	for (int i = 0; i < n; ++i) {
		u64 c = iv[i].size < (u64)P.opt.max_occ ? iv[i].size : (u64)P.opt.max_occ;
		if (rel < c) {
			u64 step = iv[i].size > (u64)P.opt.max_occ ? iv[i].size / (u64)P.opt.max_occ : 1;
			P.sd_rbeg[t] = (i64)fm_sa(P.x, iv[i].x0 + rel * step);
			return;
		}
		rel -= c;
	}
}

// ---- chain: one thread per read ------------------------------------------------------------------
#define CHAIN_HEAVY_SEEDS 64 // reads with more suffix-array hits than this are chained by a warp (chain.cuh chain_read_warp)
// reads with many SA hits are chained by a warp each (wbody_chain_heavy); listing them first lets that kernel run beside body_chain
DEV void body_chain_classify(const BatchParams& P, i64 r)
{
	if (P.seed_off[r + 1] - P.seed_off[r] > (u64)P.chain_heavy_seeds) { u32 slot = atomic_add_u32(P.n_chain_heavy, 1u); P.chain_heavy[slot] = (u32)r; }
}
DEV void body_chain(const BatchParams& P, i64 r)
{
	u64 t0 = P.seed_off[r];
	if (P.seed_off[r + 1] - t0 > (u64)P.chain_heavy_seeds) { if (!P.chain_listed) { u32 slot = atomic_add_u32(P.n_chain_heavy, 1u); P.chain_heavy[slot] = (u32)r; } return; }
	int len = (int)(P.seq_off[r + 1] - P.seq_off[r] - 1);
	ChainStageIO io;
	io.intv = P.intv + (size_t)r * P.intv_cap; io.n_intv = (int)P.n_intv[r];
	io.sd_rbeg = P.sd_rbeg + t0;
	io.S = P.S + t0; io.C = P.C + t0; io.ord = P.ord + t0; io.tmp = P.tmp32 + t0; io.ck = P.ck + t0; io.ks = P.ks + t0;
	io.kq = P.srt + t0; // (a pool of the extension stage, free at this point)
	int nks = 0;
	int nck = P.seed_off[r + 1] > t0 ? chain_read(P.x, P.opt, len, io, &nks) : 0;
	P.n_ck[r] = (u32)nck; P.n_ks[r] = (u32)nks;
}

template <int W>
DEV void wbody_chain_heavy(const BatchParams& P, i64 item, WarpMem& wm, WarpScratch& ws)
{
	LANE_DECL;
	(void)wm; (void)ws;
	const i64 r = (i64)P.chain_heavy[item];
	const u64 t0 = P.seed_off[r];
	const int len = (int)(P.seq_off[r + 1] - P.seq_off[r] - 1);
	ChainStageIO io;
	io.intv = P.intv + (size_t)r * P.intv_cap; io.n_intv = (int)P.n_intv[r];
	io.sd_rbeg = P.sd_rbeg + t0;
	io.S = P.S + t0; io.C = P.C + t0; io.ord = P.ord + t0; io.tmp = P.tmp32 + t0; io.ck = P.ck + t0; io.ks = P.ks + t0;
	io.kq = nullptr;
	ChainWarpScratch cs; cs.pos_sorted = (i64*)(P.regs_tmp + t0); cs.qbw = P.srt + t0; // pools of later stages, free at this point
	int nks = 0;
	const int nck = chain_read_warp<W>(P.x, P.opt, len, io, cs, &nks);
	ONE_LANE { P.n_ck[r] = (u32)nck; P.n_ks[r] = (u32)nks; }
}

// ---- mem_flt_chained_seeds / mem_seed_sw (SURVEY.md App. B.5 tail; reads long enough that 5.5 ln l <= 0.05 l, or
// min_chain_weight > 0): every seed of a kept chain shorter than 200 bases is re-scored by a local alignment of the
// seed +-50 bases against the reference +-50 bases; seeds scoring below min_HSP_score are dropped from their chain.
DEV bool flt_applies(const DevIndex& x, const MemOpt& opt, int l_query, int* min_hsp, u32* err)
{
	// log(l_query) from the host's table (glibc values, hazard H5); the arithmetic is bwa's: float constants, double products
	double lg = 0.;
	if (!opt.min_chain_weight) { if (l_query < x.log_tab_n) lg = x.log_tab[l_query]; else { atomic_or_u32(err, ERR_QLEN_OVERFLOW); return false; } }
	const double min_l = opt.min_chain_weight ? (double)(1.1f * (float)opt.min_chain_weight) : (double)5.5f * lg;
	*min_hsp = (int)((double)opt.a * min_l + .499);
	return !(min_l > (double)(0.05f * (float)l_query));
}
// thread per read: one task per seed of its kept chains (task = slot of the seed in the kept-seed pool)
DEV void body_flt_gen(const BatchParams& P, i64 r)
{
	const int l_query = (int)(P.seq_off[r + 1] - P.seq_off[r] - 1);
	int min_hsp;
	if (l_query < 1 || P.n_ck[r] == 0 || !flt_applies(P.x, P.opt, l_query, &min_hsp, P.err)) return;
	const u32 nks = P.n_ks[r];
	if (!nks) return;
	const u32 base = atomic_add_u32(P.n_flt_tasks, nks);
	for (u32 k = 0; k < nks; ++k) { P.flt_tasks[2 * (size_t)(base + k)] = (u32)r; P.flt_tasks[2 * (size_t)(base + k) + 1] = k; }
}
// warp per task: mem_seed_sw.  The score goes to tmp32[slot] (-1: the seed is exempt).
template <int W>
DEV void wbody_flt_sw(const BatchParams& P, i64 item, WarpMem& wm, WarpScratch& ws)
{
	LANE_DECL;
	const i64 r = (i64)P.flt_tasks[2 * (size_t)item];
	const u64 slot = P.seed_off[r] + P.flt_tasks[2 * (size_t)item + 1];
	const Seed s = P.ks[slot];
	const int l_query = (int)(P.seq_off[r + 1] - P.seq_off[r] - 1);
	const u8* query = P.seq + P.seq_off[r];
	const i64 l_pac = P.x.l_pac;
	int score = -1;
	if (s.len < 200) {
		int qb = s.qbeg, qe = s.qbeg + s.len;
		i64 rb = s.rbeg, re = s.rbeg + s.len;
		const i64 mid = (rb + re) >> 1;
		qb -= 50; qb = qb > 0 ? qb : 0;
		qe += 50; qe = qe < l_query ? qe : l_query;
		rb -= 50; rb = rb > 0 ? rb : 0;
		re += 50; re = re < l_pac << 1 ? re : l_pac << 1;
		if (rb < l_pac && l_pac < re) { if (mid < l_pac) re = l_pac; else rb = l_pac; }
		if (!(qe - qb >= 200 || re - rb >= 200)) {
			ref_clip_window(P.x, rb, mid, re); // bns_fetch_seq clips the window to the contig of its middle
			const int qlen = qe - qb, tlen = (int)(re - rb);
			// ksw_align2 without KSW_XBYTE: the 16-bit kernel (8 lanes); only the forward score is consumed
			const int lanes_simd = 8, slen = (qlen + lanes_simd - 1) / lanes_simd, plen = slen * lanes_simd;
			if (plen + 2 > wm.cap) { ONE_LANE { atomic_or_u32(ws.err, ERR_QLEN_OVERFLOW); } }
			else {
				FOR_LANES { for (int j = lane; j < plen; j += W) WQ(wm, j) = j < qlen ? query[qb + j] : (u8)5; }
				WARP_SYNC();
				TAcc ta; ta.base = rb; ta.stride = 1; ta.piv = -1;
				const SwFwd fw = warp_sw_striped<W>(P.x, P.opt, wm, lanes_simd, false, qlen, tlen, ta, KSW_XSTART, ws.brow, P.cnt ? &P.cnt->cells_msw : nullptr);
				score = fw.score;
			}
		}
	}
	ONE_LANE { P.tmp32[slot] = (u32)score; }
}
// thread per read: drop the seeds that scored below min_HSP_score, chain by chain, keeping the order
DEV void body_flt_apply(const BatchParams& P, i64 r)
{
	const int l_query = (int)(P.seq_off[r + 1] - P.seq_off[r] - 1);
	int min_hsp;
	if (l_query < 1 || P.n_ck[r] == 0 || !flt_applies(P.x, P.opt, l_query, &min_hsp, P.err)) return;
	const u64 t0 = P.seed_off[r];
	const int nck = (int)P.n_ck[r];
	for (int c = 0; c < nck; ++c) {
		ChainK& ck = P.ck[t0 + c];
		Seed* seeds = P.ks + t0 + ck.seed_start;
		u32 k = 0;
		for (u32 j = 0; j < ck.n_seeds; ++j) {
			Seed s = seeds[j];
			const int sc = (int)P.tmp32[t0 + ck.seed_start + j];
			if (sc < 0 || sc >= min_hsp) { s.score = sc < 0 ? s.len * P.opt.a : sc; seeds[k++] = s; }
		}
		ck.n_seeds = k;
	}
}

DEV void make_warp_ctx(const BatchParams& P, i64 warp_slot, u8* smem, WarpMem& wm, WarpScratch& ws)
{
	int cap = P.wm_cap;
	i32* p = P.ws_rows ? (i32*)(P.ws_rows + (size_t)warp_slot * P.ws_rows_stride) : (i32*)smem;
	wm.H = p; wm.E = p + (cap + 2); wm.H2 = p + 2 * (cap + 2); wm.Hm = p + 3 * (cap + 2); wm.fend = p + 4 * (cap + 2);
	wm.ringH = nullptr; wm.ringE = nullptr; wm.ring_n = 0; wm.ring_stride = 0;
	wm.q = (u8*)(p + 4 * (cap + 2) + 16); wm.q_stride = 1;
	wm.cap = cap;
	ws.z = P.ws_z + (size_t)warp_slot * P.ws_z_cap; ws.z_cap = P.ws_z_cap;
	ws.cigar = P.ws_cigar + (size_t)warp_slot * P.ws_cigar_cap; ws.cigar_cap = P.ws_cigar_cap;
	ws.md = P.ws_md + (size_t)warp_slot * P.ws_md_cap; ws.md_cap = P.ws_md_cap;
	ws.xa = P.ws_xa + (size_t)warp_slot * P.ws_xa_cap; ws.xa_cap = P.ws_xa_cap;
	ws.brow = P.ws_brow + (size_t)warp_slot * P.ws_brow_cap; ws.brow_cap = P.ws_brow_cap;
	ws.err = P.err;
	ws.rt_cap = P.ws_rt_cap;
	ws.rt = P.ws_rt ? P.ws_rt + (size_t)warp_slot * P.ws_rt_cap : nullptr;
	ws.rkey = P.ws_rkey ? P.ws_rkey + (size_t)warp_slot * (P.ws_rt_cap + 2) : nullptr;
	ws.rrank = P.ws_rrank ? P.ws_rrank + (size_t)warp_slot * (P.ws_rt_cap + 2) : nullptr;
}
HOSTDEV size_t warp_smem_bytes(int cap) { return (size_t)(4 * (cap + 2) + 16) * 4 + (size_t)((cap + 2 + 15) & ~15); }

// ---- extend: one warp per read -------------------------------------------------------------------
template <int W>
DEV void wbody_extend(const BatchParams& P, i64 r, WarpMem& wm, WarpScratch& ws)
{
	LANE_DECL;
	u64 t0 = P.seed_off[r];
	int len = (int)(P.seq_off[r + 1] - P.seq_off[r] - 1);
	const u8* seq = P.seq + P.seq_off[r];
	int nck = (int)P.n_ck[r], n_av = 0, av_cap = (int)P.n_ks[r];
	Reg* av = P.regs_tmp + t0;
	if (len > wm.cap) { if (nck) { ONE_LANE { atomic_or_u32(P.err, ERR_QLEN_OVERFLOW); } } nck = 0; }
	for (int c = 0; c < nck; ++c) {
		const ChainK ck = P.ck[t0 + c];
		warp_chain2aln<W>(P.x, P.opt, wm, ws, P.cnt, len, seq, ck, P.ks + t0 + ck.seed_start, P.srt + t0, av, &n_av, av_cap,
		                  P.use_ext_res ? P.ext_res + t0 + ck.seed_start : nullptr);
	}
	n_av = warp_sort_dedup_fast<W>(P.x, P.opt, wm, ws, P.cnt, seq, n_av, av);
	ONE_LANE { P.n_reg[r] = (u32)n_av; }
}

// ---- PE: insert-size candidate of a pair (thread per pair) and region capacity per read -----------
DEV void body_pe_cand(const BatchParams& P, i64 pr)
{
	i64 r0 = pr << 1, r1 = r0 + 1;
	P.pe_cand[pr] = pestat_candidate(P.x, P.opt, (int)P.n_reg[r0], P.regs_tmp + P.seed_off[r0], (int)P.n_reg[r1], P.regs_tmp + P.seed_off[r1]);
}

DEV void body_reg_cap(const BatchParams& P, i64 r)
{
	u32 cap = P.n_reg[r];
	if ((P.opt.flag & MEM_F_PE) && !(P.opt.flag & MEM_F_NO_RESCUE) && (r | 1) < P.n_reads) {
		i64 m = r ^ 1; // rescues launched from the mate's hits land in this read's list
		int n_ok = 0;
		for (int d = 0; d < 4; ++d) n_ok += P.pes[d].failed ? 0 : 1;
		const Reg* a = P.regs_tmp + P.seed_off[m];
		int nm = (int)P.n_reg[m], nb = 0;
		for (int j = 0; j < nm; ++j) if (a[j].score >= a[0].score - P.opt.pen_unpaired) ++nb;
		nb = nb < P.opt.max_matesw ? nb : P.opt.max_matesw;
		cap += (u32)(nb * n_ok);
	}
	P.reg_cap[r] = cap > 0 ? cap : 1;
}

DEV void body_compact(const BatchParams& P, i64 r)
{
	const Reg* s = P.regs_tmp + P.seed_off[r];
	Reg* d = P.regs + P.reg_off[r];
	int n = (int)P.n_reg[r];
	for (int i = 0; i < n; ++i) { Reg t = s[i]; if (t.rid >= 0 && P.x.ctg_alt[t.rid]) t.is_alt = 1; d[i] = t; } // (mem_align1_core tags the hits on ALT contigs)
}

DEV RescueIO make_rescue_io(const BatchParams& P)
{
	RescueIO io;
	io.seq_off = P.seq_off; io.seq = P.seq; io.reg_off = P.reg_off; io.reg_cap = P.reg_cap; io.regs = P.regs; io.n_reg = P.n_reg;
	io.snap = P.snap; io.nb = P.nb; io.tasks = P.sw_tasks; io.state = P.sw_state; io.list = P.sw_list; io.n_list = P.sw_n_list;
	io.max_nb = P.max_nb; io.multi = P.multi; io.n_multi = P.n_multi; io.n_reads = P.n_units << 1; io.round = P.round;
	io.spec_state = P.spec_state; io.spec_task = P.spec_task; io.n_reg0 = P.n_reg0; io.fb_list = P.multi; io.n_fb = P.n_multi; io.force_fb = P.force_fb;
	io.hv_list = P.rs_heavy; io.n_hv = P.n_rs_heavy;
	return io;
}
DEV void body_rescue_snap(const BatchParams& P, i64 r) { RescueIO io = make_rescue_io(P); rescue_snap(P.opt, io, r); }
DEV void body_rescue_gen(const BatchParams& P, i64 r) { if (r >= ((i64)P.n_units << 1)) return; RescueIO io = make_rescue_io(P); rescue_gen(P.x, P.opt, P.pes, io, r); }
DEV void body_rescue_gen_all(const BatchParams& P, i64 r) { if (r >= ((i64)P.n_units << 1)) return; RescueIO io = make_rescue_io(P); rescue_gen_all(P.x, P.opt, P.pes, io, r); }
// replay of the speculated rescue chain of read r; reads that cannot be replayed are queued for the round-by-round kernels
DEV void body_rescue_apply_all(const BatchParams& P, i64 r)
{
	if (r >= ((i64)P.n_units << 1)) return;
	RescueIO io = make_rescue_io(P);
	WarpMem wm; memset(&wm, 0, sizeof(wm));
	WarpScratch ws; memset(&ws, 0, sizeof(ws)); ws.err = P.err;
	if (io.nb[r] == 0 || rescue_is_heavy(io, r)) return; // nothing to replay / replayed by a warp (wbody_rescue_apply_heavy)
	if (!rescue_apply_all<1>(P.x, P.opt, P.pes, wm, ws, io, r)) { u32 slot = atomic_add_u32(io.n_fb, 1u); io.fb_list[slot] = (u32)r; }
}
DEV void body_rescue_classify(const BatchParams& P, i64 r) { RescueIO io = make_rescue_io(P); rescue_classify(io, r); }
template <int W>
DEV void wbody_rescue_apply_heavy(const BatchParams& P, i64 item, WarpMem& wm, WarpScratch& ws)
{
	LANE_DECL;
	RescueIO io = make_rescue_io(P);
	const i64 r = (i64)P.rs_heavy[item];
	const bool ok = rescue_apply_all<W>(P.x, P.opt, P.pes, wm, ws, io, r);
	if (!ok) { ONE_LANE { u32 slot = atomic_add_u32(io.n_fb, 1u); io.fb_list[slot] = (u32)r; } }
	WARP_SYNC();
}
// round-by-round path: the mate's list of a queued read is put back to its pre-rescue state first
DEV void body_rescue_fb_restore(const BatchParams& P, i64 i)
{
	const i64 r = (i64)P.multi[i], m = r ^ 1;
	const Reg* s = P.regs_tmp + P.seed_off[m]; Reg* d = P.regs + P.reg_off[m];
	const int n = (int)P.n_reg0[m];
	for (int k = 0; k < n; ++k) { Reg t = s[k]; if (t.rid >= 0 && P.x.ctg_alt[t.rid]) t.is_alt = 1; d[k] = t; } // (as body_compact: the temporary list is untagged)
	P.n_reg[m] = (u32)n;
}
// the round-by-round kernels visit only the queued reads (compact list)
DEV void body_rescue_gen_multi(const BatchParams& P, i64 i) { if (i < (i64)*P.n_multi) body_rescue_gen(P, (i64)P.multi[i]); }
DEV void body_rescue_apply(const BatchParams& P, i64 r);
DEV void body_rescue_apply_multi(const BatchParams& P, i64 i) { if (i < (i64)*P.n_multi) body_rescue_apply(P, (i64)P.multi[i]); }
DEV void body_rescue_apply(const BatchParams& P, i64 r)
{
	if (r >= ((i64)P.n_units << 1)) return;
	RescueIO io = make_rescue_io(P);
	WarpMem wm; memset(&wm, 0, sizeof(wm));
	WarpScratch ws; memset(&ws, 0, sizeof(ws)); ws.err = P.err;
	rescue_apply<1>(P.x, P.opt, wm, ws, io, r);
}
// one warp per alignment task; `item` indexes the compact task list of the round
template <int W>
DEV void wbody_rescue_sw(const BatchParams& P, i64 item, WarpMem& wm, WarpScratch& ws)
{
	RescueIO io = make_rescue_io(P);
	rescue_sw<W>(P.x, P.opt, wm, ws, P.cnt, io, (u32)item); // tasks are stored compactly: the item is the task index
}

// the warp-per-task kernel over the tasks the thread-level one left on its fall-back list
template <int W>
DEV void wbody_rescue_sw_fb(const BatchParams& P, i64 item, WarpMem& wm, WarpScratch& ws)
{
	RescueIO io = make_rescue_io(P);
	rescue_sw<W>(P.x, P.opt, wm, ws, P.cnt, io, P.sw_fb[item]);
}
HOSTDEV size_t tsw_block_smem_bytes(int block_threads, int capw) { return (size_t)block_threads * (size_t)capw * 10; }
#ifdef B200_SIM
static u32 g_sim_tsw_h[4096], g_sim_tsw_e[4096]; static u16 g_sim_tsw_s[4096];
#endif
// DIR 0: the forward pass of every task (tasks with a score at or above the threshold are listed for the reverse pass);
// DIR 1: the reverse pass of the listed tasks.  A warp takes 32 tasks at a time.
template <int DIR>
DEV void thread_sw_queue(const BatchParams& P, u32* counter, i64 tid)
{
	TswMem tm;
	tm.capw = P.tsw_capw;
#ifdef B200_SIM
	tm.H = g_sim_tsw_h; tm.E = g_sim_tsw_e; tm.sel = g_sim_tsw_s; tm.stride = 1; if (tm.capw > 4096) tm.capw = 4096;
#else
	extern __shared__ __align__(16) u8 tsm_[];
	const int T = (int)blockDim.x;
	tm.H = (u32*)tsm_ + threadIdx.x; tm.E = (u32*)tsm_ + (size_t)tm.capw * T + threadIdx.x; tm.sel = (u16*)(tsm_ + (size_t)tm.capw * T * 8) + threadIdx.x; tm.stride = T;
#endif
	tm.hmax = P.tsw_hmax + (size_t)(tid >> 5) * (size_t)P.tsw_capw * 32 + (size_t)(tid & 31); tm.hmax_stride = 32;
	tm.brow = P.tsw_brow + (size_t)tid * (size_t)P.tsw_brow_cap; tm.brow_cap = P.tsw_brow_cap;
	const i64 n = DIR == 0 ? (i64)*P.sw_n_list : (i64)*P.n_sw_rev;
	const int lane = simt_lane();
	const int thr = P.opt.min_seed_len * P.opt.a; // KSW_XSUBO threshold of mem_matesw
	u64 cells = 0, n_msw = 0;
	for (;;) {
		u32 base = 0;
		if (lane == 0) base = atomic_add_u32(counter, (u32)SIMT_GROUP);
		base = lane_bcast0(base);
		if ((i64)base >= n) break;
		const i64 idx = (i64)base + lane;
		if (idx < n) {
			const u32 id = DIR == 0 ? (u32)idx : P.sw_rev[idx];
			SwTask t = P.sw_tasks[id];
			const u8* ms = P.seq + P.seq_off[t.read];
			const int l_ms = (int)(P.seq_off[t.read + 1] - P.seq_off[t.read] - 1);
			const int tlen = (int)(t.re - t.rb);
			if (DIR == 0) {
				if (!tsw_takes(P.opt, l_ms, tlen, tm.capw, tm.brow_cap)) { const u32 slot = atomic_add_u32(P.n_sw_fb, 1u); P.sw_fb[slot] = id; }
				else {
					QAcc qa; qa.ms = ms; qa.l_ms = l_ms; qa.is_rev = t.is_rev; qa.piv = -1;
					TAcc ta; ta.base = t.rb; ta.stride = 1; ta.piv = -1;
					const SwFwd fw = tsw_pass_u8(P.x, P.opt, tm, P.swtab, P.sw_shift, P.sw_mmax, l_ms, qa, tlen, ta, KSW_XSUBO | KSW_XSTART | KSW_XBYTE | thr, cells);
					++n_msw;
					SwTask* o = &P.sw_tasks[id];
					o->score = fw.score; o->te = fw.te; o->qe = fw.qe; o->score2 = fw.score2; o->te2 = fw.te2; o->tb = -1; o->qb = -1;
					if (!(fw.score < thr)) { const u32 slot = atomic_add_u32(P.n_sw_rev, 1u); P.sw_rev[slot] = id; } // XSUBO: the start is searched only for scores at or above the threshold
				}
			} else {
				QAcc qa; qa.ms = ms; qa.l_ms = l_ms; qa.is_rev = t.is_rev; qa.piv = t.qe;
				TAcc tr; tr.base = t.rb; tr.stride = 1; tr.piv = t.te; // target prefix [0, te] reversed, the rest untouched
				const SwFwd rv = tsw_pass_u8(P.x, P.opt, tm, P.swtab, P.sw_shift, P.sw_mmax, t.qe + 1, qa, tlen, tr, KSW_XSTOP | t.score, cells);
				if (t.score == rv.score) { SwTask* o = &P.sw_tasks[id]; o->tb = t.te - rv.te; o->qb = t.qe - rv.qe; }
			}
		}
		warp_converge();
	}
	if (P.cnt) { if (cells) atomic_add_u64(&P.cnt->cells_msw, cells); if (n_msw) atomic_add_u64(&P.cnt->n_msw, n_msw); }
}
DEV void body_tsw_fwd_q(const BatchParams& P, i64 n, u32* counter, i64 tid) { (void)n; thread_sw_queue<0>(P, counter, tid); }
DEV void body_tsw_rev_q(const BatchParams& P, i64 n, u32* counter, i64 tid) { (void)n; thread_sw_queue<1>(P, counter, tid); }

struct FinalMode { int mode; GaTask* lt; int lt_n; const GaRes* res; int res_n, res_k; }; // how warp_reg2aln obtains alignments
DEV void make_final_ctx(const BatchParams& P, WarpMem& wm, WarpScratch& ws, FinalCtx& f, const FinalMode* fm = nullptr)
{
	f.mode = FIN_DIRECT; f.seq_base = P.seq; f.lt = nullptr; f.lt_n = 0; f.res = nullptr; f.res_n = 0; f.res_k = 0;
	if (fm) { f.mode = fm->mode; f.lt = fm->lt; f.lt_n = fm->lt_n; f.res = fm->res; f.res_n = fm->res_n; f.res_k = fm->res_k; }
	f.rescued = P.rescued != 0;
	f.x = &P.x; f.opt = &P.opt; f.wm = &wm; f.ws = &ws; f.cnt = P.cnt;
	f.var = P.var; f.var_cap = P.var_cap; f.var_cursor = P.var_cursor;
	f.ctg_names = P.ctg_names; f.ctg_name_off = P.ctg_name_off;
	f.pes = P.pes; f.pair_tab = P.pair_tab; f.pair_tab_off = P.pair_tab_off;
}

// ---- finalize, single-end: one warp per read ------------------------------------------------------
template <int W>
DEV void wbody_final_se(const BatchParams& P, i64 r, WarpMem& wm, WarpScratch& ws, FinalMode* fm = nullptr)
{
	LANE_DECL;
	FinalCtx f; make_final_ctx(P, wm, ws, f, fm);
	int len = (int)(P.seq_off[r + 1] - P.seq_off[r] - 1);
	const u8* seq = P.seq + P.seq_off[r];
	Reg* a = P.regs + P.reg_off[r];
	int n = (int)P.n_reg[r];
	DBGP(100, n);
	warp_mark_primary_se<W>(P.opt, n, a, r, P.ztmp + P.reg_off[r]); // id = index within the call (hazard H2)
	DBGP(101, r);
	if (P.opt.flag & MEM_F_PRIMARY5) warp_reorder_primary5<W>(P.opt.T, n, a);
	int nr = warp_reg2sam<W>(f, len, seq, n, a, 0, nullptr, P.recs + P.reg_off[r], (int)P.reg_cap[r]);
	ONE_LANE { P.n_rec[r] = (u32)nr; }
	if (fm) { fm->lt_n = f.lt_n; fm->res_k = f.res_k; }
}

// ---- finalize, paired-end: one warp per pair ------------------------------------------------------
template <int W>
DEV void wbody_final_pe(const BatchParams& P, i64 pr, WarpMem& wm, WarpScratch& ws, FinalMode* fm = nullptr)
{
	LANE_DECL;
	FinalCtx f; make_final_ctx(P, wm, ws, f, fm);
	PairIO io;
	for (int e = 0; e < 2; ++e) {
		i64 r = (pr << 1) + e;
		io.l_seq[e] = (int)(P.seq_off[r + 1] - P.seq_off[r] - 1);
		io.seq[e] = P.seq + P.seq_off[r];
		io.a[e] = P.regs + P.reg_off[r]; io.n[e] = (int)P.n_reg[r]; io.cap[e] = (int)P.reg_cap[r];
		io.snap[e] = P.snap + P.reg_off[r];
		io.ztmp[e] = P.ztmp + P.reg_off[r];
		io.out[e] = P.recs + P.reg_off[r]; io.out_cap[e] = (int)P.reg_cap[r]; io.n_out[e] = 0;
	}
	io.v = P.pairv + P.reg_off[pr << 1];
	warp_sam_pe<W>(f, (u64)pr, io);
	ONE_LANE { P.n_rec[pr << 1] = (u32)io.n_out[0]; P.n_rec[(pr << 1) + 1] = (u32)io.n_out[1]; }
	if (fm) { fm->lt_n = f.lt_n; fm->res_k = f.res_k; }
}

// ---- thread-per-item variants (W = 1) with hand-over of oversized items to the warp-level kernels -------------
HOSTDEV size_t thread_scratch_bytes(int wm_cap, int z_cap, int cigar_cap, int md_cap, int xa_cap)
{
	size_t n = (size_t)(2 * (wm_cap + 2)) * 4 + (size_t)((wm_cap + 2 + 15) & ~15) + (size_t)((z_cap + 15) & ~15) + (size_t)cigar_cap * 4 + (size_t)((md_cap + 15) & ~15) + (size_t)((xa_cap + 15) & ~15);
	n += 16; // the thread's error word (thread_err_word)
	return (n + 127) & ~(size_t)127;
}
#ifdef B200_SIM
static i32 g_sim_ring[2 * THREAD_RING_N];
#endif
// shared memory of a thread-level block: two rings of THREAD_RING_N i32 per thread
HOSTDEV size_t thread_block_smem_bytes(int block_threads, int wm_cap) { return (size_t)block_threads * (THREAD_RING_N * 2 * 4 + (size_t)((wm_cap + 2 + 3) & ~3)); }
// Limit flags raised by a thread-level item ("redo me with the warp-level scratch") are collected in a word of the
// thread's GLOBAL scratch, not in a local variable: the device functions raise flags with atomicOr through a generic
// pointer, and an atomic on an address in the local window is an illegal access when it is not inlined away.
DEV u32* thread_err_word(const BatchParams& P, i64 tid) { return (u32*)(P.t_buf + (size_t)(tid + 1) * P.t_stride - 16); }
DEV void make_thread_ctx(const BatchParams& P, i64 tid, WarpMem& wm, WarpScratch& ws, u32* lerr, bool shared_rings = true)
{
	const int cap = P.wm_cap;
	u8* b = P.t_buf + (size_t)tid * P.t_stride;
	i32* p = (i32*)b;
	wm.H = p; wm.E = p + (cap + 2); wm.H2 = nullptr; wm.Hm = nullptr; wm.fend = nullptr;
	b += (size_t)(2 * (cap + 2)) * 4;
	wm.q = b; wm.q_stride = 1; b += (cap + 2 + 15) & ~15;
	wm.cap = cap;
	wm.ringH = nullptr; wm.ringE = nullptr; wm.ring_n = 0; wm.ring_stride = 0;
#ifdef B200_SIM
	if (shared_rings) { wm.ringH = g_sim_ring; wm.ringE = g_sim_ring + THREAD_RING_N; wm.ring_n = THREAD_RING_N; wm.ring_stride = 1; }
#else
	if (shared_rings) { // (kernels launched without dynamic shared memory keep the rows and the query in the thread's global scratch)
		extern __shared__ __align__(16) u8 tsm_[];
		i32* r = (i32*)tsm_;
		wm.ringH = r + threadIdx.x; wm.ringE = r + (size_t)blockDim.x * THREAD_RING_N + threadIdx.x;
		wm.ring_n = THREAD_RING_N; wm.ring_stride = (int)blockDim.x;
		wm.q = tsm_ + (size_t)blockDim.x * THREAD_RING_N * 8 + threadIdx.x; wm.q_stride = (int)blockDim.x; // query row, interleaved like the rings
	}
#endif
	ws.z = b; ws.z_cap = (size_t)P.t_z_cap; b += (P.t_z_cap + 15) & ~15;
	ws.cigar = (u32*)b; ws.cigar_cap = P.t_cigar_cap; b += (size_t)P.t_cigar_cap * 4;
	ws.md = (char*)b; ws.md_cap = P.t_md_cap; b += (P.t_md_cap + 15) & ~15;
	ws.xa = (char*)b; ws.xa_cap = P.t_xa_cap;
	ws.brow = nullptr; ws.brow_cap = 0;
	ws.rt = nullptr; ws.rkey = nullptr; ws.rrank = nullptr; ws.rt_cap = 0;
	ws.err = lerr; // limits hit here are not errors: the item is redone by a warp with the large scratch
}
// extension, thread per read (extend_thread.cuh): reads with few kept seeds are extended by one thread with its DP row
// in shared memory; the others, and reads that do not fit the thread-level rows, are queued for the warp-per-read kernel
// cell = bytes of one DP cell: 4 (H and E as 16-bit halves) or 2 (8-bit halves, when no score of the batch can exceed 255).
// The banded global alignment of region patching (rare) needs two i32 rows of its own: with 4-byte cells they are rings
// over the thread's DP row (same interleaving, so the words are the thread's own); 2-byte cells are interleaved
// differently -- a 32-bit word over them would belong to two other threads -- so there the alignment uses the full
// rows of the thread's global scratch instead (shared memory spent on rings would cost the kernel L1 for its stack).
HOSTDEV size_t thread_ext_smem_bytes(int block_threads, int rw, int wm_cap, int cell)
{
	return (size_t)block_threads * ((((size_t)rw * cell + 3) & ~(size_t)3) + (size_t)((wm_cap + 2 + 3) & ~3));
}
// Byte layout of a thread's share of the block's dynamic shared memory (element k of an array sits at
// base + (k * threads + tid) * element size: bank tid % 32 for 4-byte elements whatever k is).  A host function so
// that tests/test_smem_layout.py can check that no two threads of a block ever share a byte (b200_extLayout).
struct ExtLayout { size_t row_off, q_off, ringH_off, ringE_off; int ring_n; };
HOSTDEV ExtLayout thread_ext_layout(int threads, int tid, int rw, int wm_cap, int cell, bool patches)
{
	ExtLayout L;
	const size_t row_bytes = (((size_t)rw * cell) + 3) & ~(size_t)3, qb = (size_t)((wm_cap + 2 + 3) & ~3);
	L.row_off = (size_t)tid * cell;                       // cells: stride threads * cell
	L.q_off = (size_t)threads * row_bytes + tid;          // read codes: stride threads
	L.ringH_off = L.ringE_off = 0; L.ring_n = 0;
	if (!patches) return L;
	if (cell == 4) { // the rings reuse the thread's own DP row (same interleaving)
		int ring_n = THREAD_RING_N;
		while (ring_n > 1 && (size_t)ring_n * 8 > (size_t)rw * cell) ring_n >>= 1;
		L.ring_n = ring_n; L.ringH_off = (size_t)tid * 4; L.ringE_off = ((size_t)ring_n * threads + tid) * 4;
	}
	(void)qb;
	return L; // 2-byte cells: no rings (ring_n = 0), see thread_ext_smem_bytes
}
#ifdef B200_SIM
static u32 g_sim_ext_row[4096]; static u8 g_sim_ext_q[4096];
#endif
template <class CELL>
DEV void make_thread_ext_ctx(const BatchParams& P, i64 tid, ThreadExtMem& tm, WarpMem& wm, WarpScratch& ws, u32* lerr, bool patches)
{
	make_thread_ctx(P, tid, wm, ws, lerr); // global scratch (traceback, cigar, MD) of this thread slot; the rings are replaced below
#ifdef B200_SIM
	const int st = 1, rw = P.t_ext_rw < 4096 ? P.t_ext_rw : 4096;
	u8* base = (u8*)g_sim_ext_row; u8* qrow = g_sim_ext_q;
	ExtLayout L = thread_ext_layout(1, 0, rw, P.wm_cap, (int)sizeof(CELL), patches);
	i32* ringH = (i32*)base; i32* ringE = ringH + L.ring_n;
#else
	extern __shared__ __align__(16) u8 tsm_[];
	const int st = (int)blockDim.x, rw = P.t_ext_rw;
	u8* base = tsm_;
	const ExtLayout L = thread_ext_layout(st, (int)threadIdx.x, rw, P.wm_cap, (int)sizeof(CELL), patches);
	u8* qrow = base + L.q_off;
	i32* ringH = (i32*)(base + L.ringH_off); i32* ringE = (i32*)(base + L.ringE_off);
#endif
	tm.row = (void*)(base + L.row_off); tm.row_stride = st; tm.row_cap = rw;
	tm.rd = qrow; tm.rd_stride = st; tm.rd_cap = P.wm_cap + 2;
	tm.rowpack = P.rowpack;
	// the banded global alignment of region patching runs after the extensions: two rings of ring_n i32 per thread
	// (a power of two) over the DP row, or wm.H / wm.E in global scratch when there are none; the query row is shared
	if (L.ring_n == 0) { wm.ringH = nullptr; wm.ringE = nullptr; wm.ring_n = 0; wm.ring_stride = 0; }
	else { wm.ringH = ringH; wm.ringE = ringE; wm.ring_n = L.ring_n; wm.ring_stride = st; }
	wm.q = qrow; wm.q_stride = st;
}
DEV void body_extend_q(const BatchParams& P, i64 n, u32* counter, i64 tid)
{
	u32* const lerr = thread_err_word(P, tid); *lerr = 0;
	ThreadExtMem tm; WarpMem wm; WarpScratch ws; make_thread_ext_ctx<u32>(P, tid, tm, wm, ws, lerr, true);
	thread_extend_queue<false, u32>(P, n, counter, tm, wm, ws, lerr);
}
// the same with 8-bit H and E (batches whose scores cannot exceed 255: half the shared memory per thread, more warps per SM)
DEV void body_extend_q8(const BatchParams& P, i64 n, u32* counter, i64 tid)
{
	u32* const lerr = thread_err_word(P, tid); *lerr = 0;
	ThreadExtMem tm; WarpMem wm; WarpScratch ws; make_thread_ext_ctx<u16>(P, tid, tm, wm, ws, lerr, true);
	thread_extend_queue<false, u16>(P, n, counter, tm, wm, ws, lerr);
}
// heavy reads, step 1 (warp per read): the reference window of every chain and one extension task per kept seed
template <int W>
DEV void wbody_heavy_prep(const BatchParams& P, i64 item, WarpMem& wm, WarpScratch& ws)
{
	LANE_DECL;
	(void)wm; (void)ws;
	const i64 r = (i64)P.heavy_list[item];
	const u64 t0 = P.seed_off[r];
	const int nck = (int)P.n_ck[r];
	const int l_query = (int)(P.seq_off[r + 1] - P.seq_off[r] - 1);
	const i64 l_pac = P.x.l_pac;
	FOR_LANES {
		for (int c = lane; c < nck; c += W) {
			const ChainK ck = P.ck[t0 + c];
			const int cn = (int)ck.n_seeds;
			if (cn == 0) continue;
			const Seed* seeds = P.ks + t0 + ck.seed_start;
			i64 rmax0 = l_pac << 1, rmax1 = 0;
			for (int t = 0; t < cn; ++t) {
				const Seed& sd = seeds[t];
				i64 b = sd.rbeg - (sd.qbeg + cal_max_gap(P.opt, sd.qbeg));
				i64 e = sd.rbeg + sd.len + ((l_query - sd.qbeg - sd.len) + cal_max_gap(P.opt, l_query - sd.qbeg - sd.len));
				rmax0 = rmax0 < b ? rmax0 : b;
				rmax1 = rmax1 > e ? rmax1 : e;
			}
			rmax0 = rmax0 > 0 ? rmax0 : 0;
			rmax1 = rmax1 < l_pac << 1 ? rmax1 : l_pac << 1;
			if (rmax0 < l_pac && l_pac < rmax1) { if (seeds[0].rbeg < l_pac) rmax1 = l_pac; else rmax0 = l_pac; }
			ref_clip_window(P.x, rmax0, seeds[0].rbeg, rmax1);
			P.ck_rmax[2 * (t0 + c)] = rmax0; P.ck_rmax[2 * (t0 + c) + 1] = rmax1;
			const u32 base = atomic_add_u32(P.n_ext_tasks, (u32)cn);
			for (int t = 0; t < cn; ++t) {
				ExtTask tk; tk.r = (u32)r; tk.c = (u32)c; tk.t = (u32)t; tk.pad_ = 0; P.ext_tasks[base + t] = tk;
				// sort key of a side: query length, then the kind of read (tasks of repeat-derived reads end at scattered rows; those
				// of ordinary reads mostly run to the end of the read: kept apart, a warp's 32 tasks stay in step longer)
				if (P.ext_sides) { P.ext_key_l[base + t] = (u32)seeds[t].qbeg << 1 | 1u; P.ext_key_r[base + t] = (u32)(l_query - seeds[t].qbeg - seeds[t].len) << 1 | 1u; }
			}
		}
	}
}
// light reads (thread per read): the reference window of every chain and ONE task per chain, its best-scoring seed --
// the first one mem_chain2aln visits.  Further seeds of a chain are almost always contained in that seed's region; the
// replay sends the read to the warp kernel when one is not.  Reads with many seeds or chains go on the heavy list.
DEV void body_ext_prep(const BatchParams& P, i64 r)
{
	const int nck = (int)P.n_ck[r], nks = (int)P.n_ks[r];
	if (nck == 0) { P.n_reg[r] = 0; return; }
	if (nks > P.light_max_seeds || nck > P.light_max_chains) { const u32 slot = atomic_add_u32(P.n_heavy, 1u); P.heavy_list[slot] = (u32)r; return; }
	const u64 t0 = P.seed_off[r];
	const int l_query = (int)(P.seq_off[r + 1] - P.seq_off[r] - 1);
	const i64 l_pac = P.x.l_pac;
	int n_t = 0;
	for (int c = 0; c < nck; ++c) if (P.ck[t0 + c].n_seeds) ++n_t;
	u32 base = n_t ? atomic_add_u32(P.n_ext_tasks, (u32)n_t) : 0u;
	for (int c = 0; c < nck; ++c) {
		const ChainK ck = P.ck[t0 + c];
		const int cn = (int)ck.n_seeds;
		if (cn == 0) continue;
		const Seed* seeds = P.ks + t0 + ck.seed_start;
		i64 rmax0 = l_pac << 1, rmax1 = 0;
		int best = 0; u64 best_key = 0;
		for (int t = 0; t < cn; ++t) {
			const Seed& sd = seeds[t];
			i64 b = sd.rbeg - (sd.qbeg + cal_max_gap(P.opt, sd.qbeg));
			i64 e = sd.rbeg + sd.len + ((l_query - sd.qbeg - sd.len) + cal_max_gap(P.opt, l_query - sd.qbeg - sd.len));
			rmax0 = rmax0 < b ? rmax0 : b;
			rmax1 = rmax1 > e ? rmax1 : e;
			const u64 key = (u64)(u32)sd.score << 32 | (u32)t; // mem_chain2aln visits the seeds by descending (score, index)
			if (key >= best_key) { best_key = key; best = t; }
		}
		rmax0 = rmax0 > 0 ? rmax0 : 0;
		rmax1 = rmax1 < l_pac << 1 ? rmax1 : l_pac << 1;
		if (rmax0 < l_pac && l_pac < rmax1) { if (seeds[0].rbeg < l_pac) rmax1 = l_pac; else rmax0 = l_pac; }
		ref_clip_window(P.x, rmax0, seeds[0].rbeg, rmax1);
		P.ck_rmax[2 * (t0 + c)] = rmax0; P.ck_rmax[2 * (t0 + c) + 1] = rmax1;
		ExtTask tk; tk.r = (u32)r; tk.c = (u32)c; tk.t = (u32)best; tk.pad_ = 0;
		P.ext_tasks[base] = tk;
		P.ext_key_l[base] = (u32)seeds[best].qbeg << 1; P.ext_key_r[base] = (u32)(l_query - seeds[best].qbeg - seeds[best].len) << 1;
		++base;
	}
}
// Shared memory of a side launch: per thread a DP row of rw cells and the rw query codes of the side, both interleaved
// over the threads of the block (element k at (k * threads + tid) * element size: bank tid % 32 for 4-byte elements).
// The launches are split by query-length class, so short sides -- the bulk -- run with small rows and many resident warps.
HOSTDEV size_t side_smem_bytes(int block_threads, int rw, int cell) { return (size_t)block_threads * ((((size_t)rw * cell + 3) & ~(size_t)3) + (size_t)((rw + 3) & ~3)); }
#ifdef B200_SIM
static u32 g_sim_side_row[65536]; static u8 g_sim_side_q[65536];
#endif
template <class CELL>
DEV void make_side_mem(const BatchParams& P, ThreadExtMem& tm)
{
	const int rw = P.ext_rw;
#ifdef B200_SIM
	tm.row = (void*)g_sim_side_row; tm.row_stride = 1; tm.row_cap = rw < 65536 ? rw : 65536;
	tm.rd = g_sim_side_q; tm.rd_stride = 1; tm.rd_cap = tm.row_cap;
#else
	extern __shared__ __align__(16) u8 tsm_[];
	const int st = (int)blockDim.x;
	const size_t row_bytes = (((size_t)rw * sizeof(CELL)) + 3) & ~(size_t)3;
	tm.row = (void*)(tsm_ + (size_t)threadIdx.x * sizeof(CELL)); tm.row_stride = st; tm.row_cap = rw;
	tm.rd = tsm_ + (size_t)st * row_bytes + threadIdx.x; tm.rd_stride = st; tm.rd_cap = rw;
#endif
	tm.rowpack = P.rowpack;
}
// the side launches (extend_thread.cuh): left, then right
DEV void body_ext_left_q(const BatchParams& P, i64 n, u32* counter, i64 tid) { (void)n; (void)tid; ThreadExtMem tm; make_side_mem<u32>(P, tm); thread_extend_sides<0, u32>(P, counter, tm); }
DEV void body_ext_right_q(const BatchParams& P, i64 n, u32* counter, i64 tid) { (void)n; (void)tid; ThreadExtMem tm; make_side_mem<u32>(P, tm); thread_extend_sides<1, u32>(P, counter, tm); }
DEV void body_ext_left_q8(const BatchParams& P, i64 n, u32* counter, i64 tid) { (void)n; (void)tid; ThreadExtMem tm; make_side_mem<u16>(P, tm); thread_extend_sides<0, u16>(P, counter, tm); }
DEV void body_ext_right_q8(const BatchParams& P, i64 n, u32* counter, i64 tid) { (void)n; (void)tid; ThreadExtMem tm; make_side_mem<u16>(P, tm); thread_extend_sides<1, u16>(P, counter, tm); }
// replay of the light reads (thread per read): mem_chain2aln's decisions over the finished extensions, then region
// de-duplication.  A missing extension, or a patching alignment beyond the thread-level rows, hands the read to the
// warp kernel (heavy list), which redoes it from the chains.
DEV void body_ext_replay(const BatchParams& P, i64 r, i64 tid)
{
	const int nck = (int)P.n_ck[r], nks = (int)P.n_ks[r];
	if (nck == 0 || nks > P.light_max_seeds || nck > P.light_max_chains) return; // nothing to do / a heavy read
	u32* const lerr = thread_err_word(P, tid); *lerr = 0;
	WarpMem wm; WarpScratch ws; make_thread_ctx(P, tid, wm, ws, lerr, false);
	const u64 t0 = P.seed_off[r];
	const int len = (int)(P.seq_off[r + 1] - P.seq_off[r] - 1);
	const u8* seq = P.seq + P.seq_off[r];
	int n_av = 0; const int av_cap = nks;
	Reg* av = P.regs_tmp + t0;
	for (int c = 0; c < nck && !err_peek(lerr); ++c) {
		const ChainK ck = P.ck[t0 + c];
		warp_chain2aln<1>(P.x, P.opt, wm, ws, P.cnt, len, seq, ck, P.ks + t0 + ck.seed_start, P.srt + t0, av, &n_av, av_cap, P.ext_res + t0 + ck.seed_start, true);
	}
	int n2 = 0;
	if (!err_peek(lerr)) n2 = warp_sort_dedup_patch<1>(P.x, P.opt, wm, ws, P.cnt, seq, n_av, av);
	if (err_peek(lerr) & ERR_NEED_WARP) { const u32 slot = atomic_add_u32(P.n_heavy, 1u); P.heavy_list[slot] = (u32)r; return; }
	if (err_peek(lerr)) atomic_or_u32(P.err, err_peek(lerr));
	P.n_reg[r] = (u32)n2;
}
// heavy reads, step 2 (queue of single-seed tasks, thread-level DP)
DEV void body_extend_task_q(const BatchParams& P, i64 n, u32* counter, i64 tid)
{
	(void)n;
	u32* const lerr = thread_err_word(P, tid); *lerr = 0;
	ThreadExtMem tm; WarpMem wm; WarpScratch ws; make_thread_ext_ctx<u32>(P, tid, tm, wm, ws, lerr, false);
	thread_extend_queue<true, u32>(P, (i64)*P.n_ext_tasks, counter, tm, wm, ws, lerr);
}
DEV void body_extend_task_q8(const BatchParams& P, i64 n, u32* counter, i64 tid)
{
	(void)n;
	u32* const lerr = thread_err_word(P, tid); *lerr = 0;
	ThreadExtMem tm; WarpMem wm; WarpScratch ws; make_thread_ext_ctx<u16>(P, tid, tm, wm, ws, lerr, false);
	thread_extend_queue<true, u16>(P, (i64)*P.n_ext_tasks, counter, tm, wm, ws, lerr);
}
template <int W>
DEV void wbody_extend_heavy(const BatchParams& P, i64 item, WarpMem& wm, WarpScratch& ws) { wbody_extend<W>(P, (i64)P.heavy_list[item], wm, ws); }

DEV void body_save_regs(const BatchParams& P, i64 r)
{
	const Reg* s = P.regs + P.reg_off[r]; Reg* d = P.snap + P.reg_off[r];
	int n = (int)P.n_reg[r];
	for (int i = 0; i < n; ++i) d[i] = s[i];
}
DEV void restore_regs(const BatchParams& P, i64 r, int lane, int W)
{
	const Reg* s = P.snap + P.reg_off[r]; Reg* d = P.regs + P.reg_off[r];
	int n = (int)P.n_reg[r];
	for (int i = lane; i < n; i += W) d[i] = s[i];
}
DEV void body_final_pe_t(const BatchParams& P, i64 pr, i64 tid)
{
	u32* const lerr = thread_err_word(P, tid); *lerr = 0;
	WarpMem wm; WarpScratch ws; make_thread_ctx(P, tid, wm, ws, lerr);
	wbody_final_pe<1>(P, pr, wm, ws);
	if (err_peek(lerr)) { u32 slot = atomic_add_u32(P.n_redo, 1u); P.redo_list[slot] = (u32)pr; }
}
DEV void body_final_se_t(const BatchParams& P, i64 r, i64 tid)
{
	u32* const lerr = thread_err_word(P, tid); *lerr = 0;
	WarpMem wm; WarpScratch ws; make_thread_ctx(P, tid, wm, ws, lerr);
	wbody_final_se<1>(P, r, wm, ws);
	if (err_peek(lerr)) { u32 slot = atomic_add_u32(P.n_redo, 1u); P.redo_list[slot] = (u32)r; }
}
// cost class of an alignment task (sort key): 0 blank, 1 ungapped (a direct sum), else ~ reference rows x band columns
DEV u32 ga_cost_key(const MemOpt& opt, const GaTask& t)
{
	const int l_query = t.qe - t.qb; const i64 rlen = t.re - t.rb;
	if (l_query <= 0 || rlen <= 0) return 0u;
	if ((i64)l_query == rlen && t.w2 == 0) return 1u;
	int w = t.w2 < opt.w << 2 ? t.w2 : opt.w << 2;
	const i64 dl = rlen > l_query ? rlen - l_query : l_query - rlen;
	if (w < dl + 3) w = (int)(dl + 3);
	const i64 band = 2 * (i64)w + 1 < l_query ? 2 * (i64)w + 1 : l_query;
	const i64 cost = (rlen * band) >> 6;
	return 2u + (u32)(cost < 65000 ? cost : 65000);
}
// the tasks the thread-level kernel could not take (band beyond its shared-memory ring, traceback beyond its scratch),
// one warp each: mem_reg2aln's band loop around bwa_gen_cigar2, result into the task's GaRes
template <int W>
DEV void wbody_galign_fb(const BatchParams& P, i64 item, WarpMem& wm, WarpScratch& ws)
{
	LANE_DECL;
	const u32 tid = P.ga_fb[item];
	const GaTask tk = P.ga_tasks[tid];
	const MemOpt& opt = P.opt;
	int w2 = tk.w2, i = 0, score = 0, last_sc = -(1 << 30), NM = -1, nc = 0, ml = 0;
	bool ok = true;
	do {
		w2 = w2 < opt.w << 2 ? w2 : opt.w << 2;
		ok = warp_gen_cigar2<W>(P.x, opt, wm, ws, P.cnt, w2, tk.qe - tk.qb, P.seq + tk.qoff + tk.qb, tk.rb, tk.re, &score, true, &nc, &NM, &ml);
		if (!ok || score == last_sc || w2 == opt.w << 2) break;
		last_sc = score;
		w2 <<= 1;
	} while (++i < 3 && score < tk.truesc - opt.a);
	GaRes* o = &P.ga_res[tid];
	if (!ok || nc > GA_CIG_CAP || ml > GA_MD_CAP) return; // (valid stays 0: the item is redone as a whole)
	FOR_LANES { for (int k = lane; k < nc; k += W) o->cigar[k] = ws.cigar[k]; for (int k = lane; k < ml; k += W) o->md[k] = ws.md[k]; }
	ONE_LANE { o->score = score; o->NM = NM; o->nc = nc; o->ml = ml; o->valid = 1; }
	WARP_SYNC();
}

// ---- two-pass finalize (galign_thread.cuh): collect the alignment requests of an item, later consume their results
DEV void final_collect_item(const BatchParams& P, i64 item, i64 tid, bool pe)
{
	// An item with many hits (reads from repeats) costs a thread quadratic work in both passes and holds its warp's other 31
	// lanes back: it goes straight to the warp-level kernel, which runs the decision logic once with its scans spread over the lanes.
	{
		const u32 nh = pe ? P.n_reg[item << 1] + P.n_reg[(item << 1) + 1] : P.n_reg[item];
		if (nh > (u32)P.final_heavy_hits) { P.ga_item0[item] = 0; P.ga_itemn[item] = 0xffffffffu; return; }
	}
	u32* const sink = thread_err_word(P, tid); *sink = 0; // limits hit while collecting are not real: no alignment has been computed yet
	WarpMem wm; WarpScratch ws; make_thread_ctx(P, tid, wm, ws, sink);
	GaTask lt[GA_ITEM_MAX];
	FinalMode fm; fm.mode = FIN_COLLECT; fm.lt = lt; fm.lt_n = 0; fm.res = nullptr; fm.res_n = 0; fm.res_k = 0;
	if (pe) wbody_final_pe<1>(P, item, wm, ws, &fm); else wbody_final_se<1>(P, item, wm, ws, &fm);
	u32 n = (u32)fm.lt_n, base = 0;
	bool direct = n > (u32)GA_ITEM_MAX;
	if (!direct && n) {
		base = atomic_add_u32(P.n_ga_tasks, n);
		if ((u64)base + n > (u64)P.ga_cap) { // the pool is full: a warp redoes the item.  The counter stays above ga_cap; the task kernel runs
			direct = true;                    // min(counter, ga_cap) slots, so the slots this item reserved below the cap are blanked (l_query = 0: skipped)
			GaTask z; memset(&z, 0, sizeof(z));
			for (u64 k = base; k < (u64)P.ga_cap; ++k) { P.ga_tasks[k] = z; if (P.ga_key) P.ga_key[k] = 0u; }
		}
	}
	if (!direct) for (u32 k = 0; k < n; ++k) { P.ga_tasks[base + k] = lt[k]; if (P.ga_key) P.ga_key[base + k] = ga_cost_key(P.opt, lt[k]); }
	P.ga_item0[item] = base; P.ga_itemn[item] = direct ? 0xffffffffu : n;
}
DEV void final_consume_item(const BatchParams& P, i64 item, i64 tid, bool pe)
{
	const u32 n = P.ga_itemn[item];
	bool redo = n == 0xffffffffu;
	if (!redo) { if (pe) { restore_regs(P, item << 1, 0, 1); restore_regs(P, (item << 1) + 1, 0, 1); } else restore_regs(P, item, 0, 1); } // the collect pass edited them (the redo kernel restores them itself)
	if (!redo) {
		u32* const lerr = thread_err_word(P, tid); *lerr = 0;
		WarpMem wm; WarpScratch ws; make_thread_ctx(P, tid, wm, ws, lerr);
		FinalMode fm; fm.mode = FIN_CONSUME; fm.lt = nullptr; fm.lt_n = 0; fm.res = P.ga_res + P.ga_item0[item]; fm.res_n = (int)n; fm.res_k = 0;
		if (pe) wbody_final_pe<1>(P, item, wm, ws, &fm); else wbody_final_se<1>(P, item, wm, ws, &fm);
		redo = err_peek(lerr) != 0 || fm.res_k != (int)n;
	}
	if (redo) { u32 slot = atomic_add_u32(P.n_redo, 1u); P.redo_list[slot] = (u32)item; }
}
DEV void body_final_pe_collect(const BatchParams& P, i64 pr, i64 tid) { final_collect_item(P, pr, tid, true); }
DEV void body_final_se_collect(const BatchParams& P, i64 r, i64 tid) { final_collect_item(P, r, tid, false); }
DEV void body_final_pe_consume(const BatchParams& P, i64 pr, i64 tid) { final_consume_item(P, pr, tid, true); }
DEV void body_final_se_consume(const BatchParams& P, i64 r, i64 tid) { final_consume_item(P, r, tid, false); }
// shared memory of the alignment-task kernel: one ring of THREAD_RING_N packed H/E words per thread + the query row
HOSTDEV size_t galign_block_smem_bytes(int block_threads, int wm_cap) { return (size_t)block_threads * (THREAD_RING_N * 4 + (size_t)((wm_cap + 2 + 3) & ~3)); }
DEV void body_galign_task_q(const BatchParams& P, i64 n, u32* counter, i64 tid)
{
	(void)n;
	u32* const lerr = thread_err_word(P, tid); *lerr = 0;
	WarpMem wm; WarpScratch ws; make_thread_ctx(P, tid, wm, ws, lerr);
#ifndef B200_SIM
	{
		extern __shared__ __align__(16) u8 tsm_[];
		wm.ringH = (i32*)tsm_ + threadIdx.x; wm.ringE = nullptr; wm.ring_n = THREAD_RING_N; wm.ring_stride = (int)blockDim.x;
		wm.q = tsm_ + (size_t)blockDim.x * THREAD_RING_N * 4 + threadIdx.x; wm.q_stride = (int)blockDim.x;
	}
#endif
	const u32 n_res = *P.n_ga_tasks; // (reservations may have run past the pool: final_collect_item)
	thread_galign_queue(P, (i64)(n_res < P.ga_cap ? n_res : P.ga_cap), counter, wm, ws, lerr);
}
template <int W>
DEV void wbody_final_pe_redo(const BatchParams& P, i64 item, WarpMem& wm, WarpScratch& ws)
{
	LANE_DECL;
	const i64 pr = (i64)P.redo_list[item];
	FOR_LANES { restore_regs(P, pr << 1, lane, W); restore_regs(P, (pr << 1) + 1, lane, W); }
	WARP_SYNC();
	wbody_final_pe<W>(P, pr, wm, ws);
}
template <int W>
DEV void wbody_final_se_redo(const BatchParams& P, i64 item, WarpMem& wm, WarpScratch& ws)
{
	LANE_DECL;
	const i64 r = (i64)P.redo_list[item];
	FOR_LANES { restore_regs(P, r, lane, W); }
	WARP_SYNC();
	wbody_final_se<W>(P, r, wm, ws);
}

// ---- output stream (fmt_BAMish, jnibwa.c:43-97): size per read, then write --------------------------
DEV u32 rec_words(const AlnRec& a)
{
	u32 n = 1;
	if (!(a.flag & 0x4)) n += 8 + (u32)a.n_cigar + (u32)((a.md_len + 3) >> 2) + (u32)((a.xa_len + 3) >> 2);
	if ((a.flag & 0x9) == 1) n += 3;
	return n;
}

DEV void body_out_size(const BatchParams& P, i64 r)
{
	if ((P.opt.flag & MEM_F_PE) && r >= ((i64)P.n_units << 1)) { P.out_words[r] = 0; return; } // odd trailing read (hazard H11)
	u32 n = 1;
	const AlnRec* a = P.recs + P.reg_off[r];
	for (u32 k = 0; k < P.n_rec[r]; ++k) n += rec_words(a[k]);
	P.out_words[r] = n;
}

DEV void body_out_write(const BatchParams& P, i64 r)
{
	if (P.out_words[r] == 0) return;
	u32* o = P.out + P.out_off[r];
	const AlnRec* a = P.recs + P.reg_off[r];
	u32 nr = P.n_rec[r];
	*o++ = nr;
	for (u32 k = 0; k < nr; ++k) {
		const AlnRec& p = a[k];
		u32 fm = (u32)p.flag;
		if (p.flag & 0x10000) fm |= 0x100;
		*o++ = (fm << 16) | (p.mapq & 0xff);
		if (!(p.flag & 0x4)) {
			*o++ = (u32)p.rid; *o++ = (u32)p.pos; *o++ = (u32)p.NM; *o++ = (u32)p.score; *o++ = (u32)p.sub; *o++ = (u32)p.n_cigar;
			const u32* v = P.var + p.var_off;
			for (int c = 0; c < p.n_cigar; ++c) { u32 lo = v[c]; if ((lo & 0xf) > 2) ++lo; *o++ = lo; } // MIDSH -> BAM numbering
			*o++ = (u32)p.md_len;
			for (int c = 0; c < (p.md_len + 3) >> 2; ++c) *o++ = v[p.n_cigar + c];
			*o++ = (u32)p.xa_len;
			const u32* xv = P.var + p.xa_off;
			for (int c = 0; c < (p.xa_len + 3) >> 2; ++c) *o++ = xv[c];
		}
		if ((p.flag & 0x9) == 1) { *o++ = (u32)p.m_rid; *o++ = (u32)p.m_pos; *o++ = (u32)p.tlen; }
	}
}
