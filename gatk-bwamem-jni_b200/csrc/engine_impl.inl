// engine_impl.inl -- host-side orchestration of one index handle and of createAlignments batches.
// Included by engine.cu (the product: CUDA runtime, sm_100a kernels) and by tests/cudasim (kernel-debug
// harness on the host).  The including file provides the small runtime vocabulary used below:
//   rt_malloc / rt_free / rt_h2d / rt_d2h / rt_memset / rt_sync / rt_scan_u32 and the three launchers
//   RT_THREADS(body, P, n)  RT_THREADS_TID(body, P, n)  RT_WARPS(wbody, P, n)
// plus rt_thread_slots() / rt_warp_slots() (how many scratch slots the launchers may touch).
//
// Drop-in boundary: jnibwa_openIndex / destroyIndex / getRefContigNames / createAlignments of
// /root/reference/src/main/c/jnibwa.h:11-16 (definitions at jnibwa.c:154-235).

#include <cmath>
#include <ctime>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <mutex>
#include <condition_variable>
#include <chrono>
#include <thread>
#include <string>
#include <vector>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include "bwa_image.h"
#include "pestat_host.h"

namespace b200 {

#ifndef B200_SIM
void scan_reads(const char* base, uint32_t n, uint64_t* off, int* max_len, std::atomic<uint64_t>* progress); // host_scan.cpp
#endif

bool ImageView::parse(const uint8_t* m, size_t l)
{
	size_t k = 0;
	if (l < 1120) return false;
	memcpy(&primary, m, 8); memcpy(L2, m + 8, 40); memcpy(&seq_len, m + 48, 8); memcpy(&bwt_size, m + 56, 8);
	memcpy(&sa_intv, m + 72 + 1024, 4); memcpy(&n_sa, m + 72 + 1024 + 8, 8);
	k = 1120;
	if (k + bwt_size * 4 > l) return false;
	bwt = (const uint32_t*)(m + k); k += bwt_size * 4;
	if (k + n_sa * 8 > l) return false;
	sa = (const uint64_t*)(m + k); k += n_sa * 8;
	if (k + 48 > l) return false;
	memcpy(&l_pac, m + k, 8); memcpy(&n_seqs, m + k + 8, 4); memcpy(&n_holes, m + k + 24, 4);
	k += 48;
	k += (size_t)n_holes * 16;
	if (n_seqs < 0 || k + (size_t)n_seqs * 40 > l) return false;
	contigs.resize((size_t)n_seqs);
	for (int i = 0; i < n_seqs; ++i) {
		const uint8_t* p = m + k + (size_t)i * 40;
		memcpy(&contigs[i].offset, p, 8); memcpy(&contigs[i].len, p + 8, 4); memcpy(&contigs[i].n_ambs, p + 12, 4);
		memcpy(&contigs[i].gi, p + 16, 4); memcpy(&contigs[i].is_alt, p + 20, 4);
	}
	k += (size_t)n_seqs * 40;
	for (int i = 0; i < n_seqs; ++i) { // strings are delimited by NUL, never by fixed offsets
		size_t e = k; while (e < l && m[e]) ++e; if (e >= l) return false;
		contigs[i].name.assign((const char*)m + k, e - k); k = e + 1;
		e = k; while (e < l && m[e]) ++e; if (e >= l) return false;
		contigs[i].anno.assign((const char*)m + k, e - k); k = e + 1;
	}
	pac = m + k; k += (size_t)(l_pac / 4 + 1);
	return k == l; // the reference asserts that the walk ends exactly at the file size
}

// development aid (B200_HOSTPROF=1): wall-clock marks of the host side of one call, printed to stderr
struct HostProf {
	bool on = false;
	std::vector<std::pair<const char*, double>> marks; double t0 = 0;
	static double now() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
	void begin() { if (on) { marks.clear(); t0 = now(); } }
	void mark(const char* nm) { if (on) marks.push_back({ nm, now() }); }
	void report() { if (!on) return; double p = t0; fprintf(stderr, "[hostprof]"); for (auto& m : marks) { fprintf(stderr, " %s=%.2f", m.first, m.second - p); p = m.second; } fprintf(stderr, " total=%.2f ms\n", p - t0); }
};

// Everything the environment can say about a handle, read ONCE by jnibwa_openIndex (never per batch).  The BWA_B200_*
// variables are the deployment knobs of INTEGRATION.md; B200_HOSTPROF / B200_DEBUG_STATS are reporting aids.  The
// remaining B200_* variables force kernel variants and capacity limits for the parity tests and for tuning runs: they
// change which kernels a call launches, so they are honoured only when B200_TEST_HOOKS=1 is set as well.
struct EngineConfig {
	int sa_intv = 4;             // BWA_B200_SA_INTV: suffix-array samples kept on the device (8 B per 4 rows = 12.4 GB for a human genome)
	int contexts = 3;            // BWA_B200_CONTEXTS: batches in flight per device
	int seed_pool_gb = 20;       // BWA_B200_SEED_POOL_GB: largest candidate pool of the seeding stage; larger batches are seeded in slices of reads (each slice costs ~1 % of the step)
	int kmer_k = 12;             // BWA_B200_KMER: longest string of the prefix tables (0: none; 12 = 358 MB)
	bool hostprof = false, debug_stats = false;
	// ---- test hooks / tuning aids (B200_TEST_HOOKS=1)
	bool no_stage_overlap = false, seed_one_launch = false, seed_fused_rounds = false, tiny_seed_pool = false, tiny_seed_tasks = false, seed_no_overlap = false;
	bool chain_all_warp = false, chain_all_thread = false, chain_no_overlap = false, tiny_scratch = false, all_heavy = false, all_light = false;
	bool warp_extend = false, ext_wide_cells = false, no_ext_tasks = false, rescue_fallback = false, final_one_pass = false, final_ring = false, tiny_ga_cap = false, tiny_intv_cap = false, tiny_warp_scratch = false;
	bool ext_legacy = false, ext_no_sort = false, force_long_mode = false, rescue_warp_sw = false, ga_no_sort = false, ext_task_machine = false;
	int seed_lcap = 0, seed_shared_read = -1, tx_chunk = -1, light_seeds = -1, light_chains = -1, final_heavy_hits = -1, occ_sb_shift = 0, seed_slices = 0;
	static bool flag(const char* nm) { return getenv(nm) != nullptr; }
	static int num(const char* nm, int dflt) { const char* e = getenv(nm); return e ? atoi(e) : dflt; }
	void read_env()
	{
		sa_intv = num("BWA_B200_SA_INTV", sa_intv);
		contexts = num("BWA_B200_CONTEXTS", contexts); if (contexts < 1) contexts = 1; if (contexts > 8) contexts = 8;
		seed_pool_gb = num("BWA_B200_SEED_POOL_GB", seed_pool_gb); if (seed_pool_gb < 1) seed_pool_gb = 1;
		kmer_k = num("BWA_B200_KMER", kmer_k); if (kmer_k < 2) kmer_k = 0; if (kmer_k > 13) kmer_k = 13;
		hostprof = flag("B200_HOSTPROF"); debug_stats = flag("B200_DEBUG_STATS");
		const char* th = getenv("B200_TEST_HOOKS");
		if (!th || atoi(th) == 0) return;
		no_stage_overlap = flag("B200_NO_STAGE_OVERLAP"); seed_one_launch = flag("B200_SEED_ONE_LAUNCH"); seed_fused_rounds = flag("B200_SEED_FUSED_ROUNDS");
		tiny_seed_pool = flag("B200_TEST_TINY_SEED_POOL"); tiny_seed_tasks = flag("B200_TEST_TINY_SEED_TASKS"); seed_no_overlap = flag("B200_SEED_NO_OVERLAP");
		chain_all_warp = flag("B200_TEST_CHAIN_ALL_WARP"); chain_all_thread = flag("B200_TEST_CHAIN_ALL_THREAD"); chain_no_overlap = flag("B200_CHAIN_NO_OVERLAP");
		tiny_scratch = flag("B200_TEST_TINY_SCRATCH"); all_heavy = flag("B200_TEST_ALL_HEAVY"); all_light = flag("B200_TEST_ALL_LIGHT");
		warp_extend = flag("B200_WARP_EXTEND"); ext_wide_cells = flag("B200_EXT_WIDE_CELLS"); no_ext_tasks = flag("B200_NO_EXT_TASKS");
		rescue_fallback = flag("B200_TEST_RESCUE_FALLBACK"); final_one_pass = flag("B200_FINAL_ONE_PASS"); final_ring = flag("B200_TEST_FINAL_RING");
		tiny_ga_cap = flag("B200_TEST_TINY_GA_CAP"); tiny_intv_cap = flag("B200_TEST_TINY_INTV_CAP"); tiny_warp_scratch = flag("B200_TEST_TINY_WARP_SCRATCH"); ext_legacy = flag("B200_EXT_LEGACY"); ext_no_sort = flag("B200_EXT_NO_SORT"); force_long_mode = flag("B200_TEST_LONG_MODE"); rescue_warp_sw = flag("B200_RESCUE_WARP_SW"); ga_no_sort = flag("B200_GA_NO_SORT"); ext_task_machine = flag("B200_EXT_TASK_MACHINE");
		seed_lcap = num("B200_SEED_LCAP", 0); seed_shared_read = num("B200_SEED_SHARED_READ", -1); tx_chunk = num("B200_TX_CHUNK", -1);
		light_seeds = num("B200_LIGHT_SEEDS", -1); light_chains = num("B200_LIGHT_CHAINS", -1); final_heavy_hits = num("B200_FINAL_HEAVY_HITS", -1);
		seed_slices = num("B200_TEST_SEED_SLICES", 0);
		occ_sb_shift = num("B200_TEST_OCC_SB", 0); if (occ_sb_shift < 0 || occ_sb_shift > 40) occ_sb_shift = 0; // force super-block bases every 2^n sectors
	}
};

struct DBuf { // growable device buffer
	void* p = nullptr; size_t cap = 0;
	bool ensure(size_t bytes) {
		if (bytes <= cap) return true;
		if (p) rt_free(p);
		const size_t slack = bytes / 4 < ((size_t)512 << 20) ? bytes / 4 : ((size_t)512 << 20); // (grow-only: room for the next, slightly larger batch)
		size_t want = bytes + slack + 256;
		p = rt_malloc(want); cap = p ? want : 0;
		return p != nullptr;
	}
	void release() { if (p) rt_free(p); p = nullptr; cap = 0; }
	template <class T> T* as() { return (T*)p; }
};

#define LOG_TAB_N 65536

// mem_opt_init of un-vendored lh3/bwa (called at …BwaMemIndex.c:91); defaults per SURVEY.md App. B.1.
// The block is malloc'd: Java releases it through destroyByteBuffer == free (…BwaMemIndex.c:157-160).
inline MemOpt* default_options()
{
	MemOpt* o = (MemOpt*)calloc(1, sizeof(MemOpt));
	if (!o) return nullptr;
	o->a = 1; o->b = 4; o->o_del = o->o_ins = 6; o->e_del = o->e_ins = 1; o->w = 100; o->T = 30; o->zdrop = 100;
	o->pen_unpaired = 17; o->pen_clip5 = o->pen_clip3 = 5; o->max_mem_intv = 20; o->min_seed_len = 19; o->split_width = 10;
	o->max_occ = 500; o->max_chain_gap = 10000; o->max_ins = 10000; o->mask_level = 0.50f; o->drop_ratio = 0.50f;
	o->XA_drop_ratio = 0.80f; o->split_factor = 1.5f; o->chunk_size = 10000000; o->n_threads = 1; o->max_XA_hits = 5;
	o->max_XA_hits_alt = 200; o->max_matesw = 50; o->mask_level_redun = 0.95f; o->min_chain_weight = 0;
	o->max_chain_extend = 1 << 30; o->mapQ_coef_len = 50; o->mapQ_coef_fac = (int)log((double)o->mapQ_coef_len);
	int k = 0;
	for (int i = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) o->mat[k++] = i == j ? (i8)o->a : (i8)-o->b; o->mat[k++] = -1; }
	for (int j = 0; j < 5; ++j) o->mat[k++] = -1;
	return o;
}

// jnibwa_getRefContigNames (jnibwa.c:174-195): i32 n, then { i32 len, bytes } per contig, no NULs
inline void* contig_names_blob(const ImageView& img, size_t* size)
{
	size_t sz = 4 + 4 * (size_t)img.n_seqs;
	for (auto& c : img.contigs) sz += c.name.size();
	char* buf = (char*)malloc(sz);
	if (!buf) return nullptr;
	int32_t n = img.n_seqs;
	memcpy(buf, &n, 4);
	char* p = buf + 4;
	for (auto& c : img.contigs) { int32_t l = (int32_t)c.name.size(); memcpy(p, &l, 4); p += 4; memcpy(p, c.name.data(), (size_t)l); p += l; }
	*size = sz;
	return buf;
}

struct Engine;
// The index resident in the HBM of ONE device (uploaded once by jnibwa_openIndex; read-only afterwards, shared by every
// context of that device).
struct Replica {
	int dev = 0;
	DevIndex dx;
	DBuf d_bwt, d_occ_sb, d_sa, d_sa_dense, d_pac, d_ctg, d_log, d_names, d_name_off, d_kmer;
	int busy = 0; // contexts of this device inside a call (guarded by Engine::mu)
	void release_all() { DBuf* all[] = { &d_bwt, &d_occ_sb, &d_sa, &d_sa_dense, &d_pac, &d_ctg, &d_log, &d_names, &d_name_off, &d_kmer }; for (DBuf* b : all) b->release(); }
};

// One batch in flight: a runtime context (streams, events, timer: engine.cu RtCtx) on one device, the grow-only batch
// workspace in that device's HBM, the pinned staging of the request and the staged request itself.  A createAlignments
// call borrows a Batch for its whole duration; callers on different Batches overlap freely (stage of call N+1 beside
// the kernels of call N beside the result copy of call N-1), which is the reference's threading contract: one
// BwaMemIndex shared by any number of threads, each with its own BwaMemAligner (BwaMemAligner.java:18, BwaMemIndex.java:18,26-27).
struct Batch {
	Engine* eng = nullptr; Replica* rep = nullptr; RtCtx* rt = nullptr;
	bool in_use = false; // guarded by Engine::mu
	int scratch_scale = 1; // multiplier of the warp-level scratch capacities: doubled when an item of a batch outgrew one (Batch::run)
	int intv_scale = 1;  // interval-list capacity multiplier: doubled when a batch overflowed its lists (Batch::run)
	HostProf hp;
	// batch workspace
	DBuf b_blob, b_seq, b_seqoff, b_intv, b_nintv, b_nseed, b_smem, b_seedoff, b_rbeg, b_S, b_C, b_ord, b_tmp32, b_ck, b_ks, b_srt, b_regtmp,
	     b_nck, b_nks, b_nreg, b_regcap, b_regoff, b_regs, b_snap, b_ztmp, b_pairv, b_recs, b_nrec, b_cand, b_pairtab, b_var, b_misc,
	     b_outwords, b_outoff, b_out, b_wsz, b_wscig, b_wsmd, b_wsxa, b_wsbrow, b_nb, b_swtask, b_swstate, b_swlist, b_tbuf, b_redo, b_heavy, b_multi, b_cand2, b_specstate, b_spectask, b_nreg0, b_rsheavy, b_wsrt, b_wsrkey, b_wsrrank, b_exttask, b_extres, b_ckrmax, b_chainheavy, b_gatask, b_gares, b_gaitem0, b_gaitemn, b_seedpool, b_seedused, b_seedtask, b_seedfb,
	     b_xtkey, b_xtkey2, b_xtidx, b_xtidx2, b_wsrows, b_flttask, b_tswhmax, b_tswbrow, b_swrev, b_swfb, b_gakey, b_gakey2, b_gaiota, b_gaorder, b_gafb;
	Counters last_counters; uint64_t last_n_seed = 0; int n_launches = 0;
	std::string last_error;
	// staged request
	bool st_valid = false; uint32_t st_n_reads = 0; int st_max_len = 0; uint64_t st_blob_len = 0; std::vector<uint64_t> st_seq_off;
#ifndef B200_SIM
	enum { NUP = 4 };
	cudaStream_t up_stream[NUP] = { nullptr, nullptr, nullptr, nullptr };
	cudaEvent_t up_done[NUP] = { nullptr, nullptr, nullptr, nullptr };
	char* pin = nullptr; uint64_t pin_cap = 0; // pinned staging of the request bytes, grow-only like the device buffer
	char* pin_out = nullptr; uint64_t pin_out_cap = 0; std::vector<cudaEvent_t> out_done; // pinned staging of the record stream, one event per chunk
#endif
	bool fetch_result(void* dst, const void* dev_src, size_t bytes);
	bool stage(const char* blob, bool wait = true);
	void* run(const MemOpt* opt, const PeStat* pes0, size_t* out_size, bool host_output);
	void* align(const MemOpt* opt, const PeStat* pes0, char* blob, size_t* out_size);
	void release_all();
	size_t workspace_bytes(); // device bytes this context holds
};

struct Engine {
	// host view of the image
	void* map = nullptr; size_t map_len = 0;
	ImageView img;
	EngineConfig cfg;
	std::vector<Replica*> reps;   // one per device of the pool (BWA_B200_DEVICES; default: the opening thread's device)
	std::vector<Batch*> ctxs;     // cfg.contexts per replica
	std::mutex mu; std::condition_variable cv;
	std::string last_error;       // of the most recent failed call on this handle
	Counters last_counters; uint64_t last_n_seed = 0; int n_launches = 0; // of the most recent call (instrumentation)
	Batch* last_batch = nullptr;

	bool open_from_fd(int fd, const std::vector<int>& devices);
	bool open_from_fd(int fd) { return open_from_fd(fd, std::vector<int>(1, rt_current_device())); }
	bool upload(Replica& r, const Replica* from);
	void close_all();
	// borrow the idle context of the least busy device (blocks while all are busy); `want` pins a particular one
	Batch* acquire(Batch* want = nullptr);
	void release(Batch* b);
	void* align(const MemOpt* opt, const PeStat* pes0, char* blob, size_t* out_size);
};

inline Batch* Engine::acquire(Batch* want)
{
	std::unique_lock<std::mutex> lock(mu);
	for (;;) {
		Batch* best = nullptr;
		if (want) { if (!want->in_use) best = want; }
		else for (Batch* b : ctxs) if (!b->in_use && (!best || b->rep->busy < best->rep->busy)) best = b;
		if (best) { best->in_use = true; ++best->rep->busy; rt_bind(best->rt); return best; }
		if (ctxs.empty()) return nullptr;
		cv.wait(lock);
	}
}
inline void Engine::release(Batch* b)
{
	{
		std::lock_guard<std::mutex> lock(mu);
		if (!b->last_error.empty()) last_error = b->last_error;
		last_counters = b->last_counters; last_n_seed = b->last_n_seed; n_launches = b->n_launches; last_batch = b;
		b->in_use = false; --b->rep->busy;
	}
	cv.notify_one();
}
inline void* Engine::align(const MemOpt* opt, const PeStat* pes0, char* blob, size_t* out_size)
{
	Batch* b = acquire();
	if (!b) { last_error = "index handle has no device context"; return nullptr; }
	void* r = b->align(opt, pes0, blob, out_size);
	release(b);
	return r;
}

inline bool Engine::open_from_fd(int fd, const std::vector<int>& devices)
{
	cfg.read_env();
	struct stat st;
	if (fstat(fd, &st) == -1) { close(fd); return false; }
	void* m = mmap(0, st.st_size, PROT_READ, MAP_SHARED, fd, 0);
	close(fd); // the reference takes ownership of fd and closes it (jnibwa.c:158)
	if (m == MAP_FAILED) return false;
	map = m; map_len = (size_t)st.st_size;
	if (!img.parse((const uint8_t*)map, map_len)) { last_error = "malformed index image"; return false; }
	if (img.sa_intv <= 0 || (img.sa_intv & (img.sa_intv - 1))) { last_error = "SA interval must be a power of two"; return false; }
	if (devices.empty()) { last_error = "no device"; return false; }
	for (size_t k = 0; k < devices.size(); ++k) {
		Replica* r = new Replica(); r->dev = devices[k];
		reps.push_back(r);
		for (int c = 0; c < cfg.contexts; ++c) {
			RtCtx* rt = rt_ctx_create(r->dev);
			if (!rt) { last_error = "cannot create a device context"; return false; }
			Batch* b = new Batch(); b->eng = this; b->rep = r; b->rt = rt; b->hp.on = cfg.hostprof;
			ctxs.push_back(b);
		}
		rt_bind(ctxs[k * (size_t)cfg.contexts]->rt);
		if (!upload(*r, k ? reps[0] : nullptr)) return false;
	}
	return true;
}

// Replica 0 takes the image from the host (and re-lays it out on the device); the others copy the finished arrays
// device to device (NVLink), so a pool of N GPUs costs one host upload.
inline bool Engine::upload(Replica& R, const Replica* from)
{
	DBuf& d_bwt = R.d_bwt; DBuf& d_sa = R.d_sa; DBuf& d_sa_dense = R.d_sa_dense; DBuf& d_pac = R.d_pac; DBuf& d_ctg = R.d_ctg; DBuf& d_log = R.d_log;
	DBuf& d_names = R.d_names; DBuf& d_name_off = R.d_name_off; DBuf& d_kmer = R.d_kmer; DevIndex& dx = R.dx;
	size_t bwt_bytes = (size_t)img.bwt_size * 4;
	const uint64_t n_blocks = (img.seq_len + 127) / 128;
	const size_t pac_bytes = (size_t)(img.l_pac / 4 + 1);
	// The device keeps u32 counts per sector (fm.cuh).  A text in which one base occurs 2^32 times or more gets super-block
	// bases (u64, every 2^31 rows) that the counts are relative to; they are read off the image's own blocks here.
	std::vector<uint64_t> occ_sb; int occ_sb_shift = 0;
	{
		bool wide = false;
		for (int c = 0; c < 4; ++c) wide = wide || img.L2[c + 1] - img.L2[c] >= ((uint64_t)1 << 32);
		if (wide) occ_sb_shift = 25;
		if (cfg.occ_sb_shift >= 1) occ_sb_shift = cfg.occ_sb_shift;
		if (occ_sb_shift) {
			const uint64_t n_sb = ((2 * n_blocks) >> occ_sb_shift) + 1;
			occ_sb.resize((size_t)n_sb * 4);
			for (uint64_t t = 0; t < n_sb; ++t) fm_image_counts((const u32*)img.bwt, (u64)img.bwt_size, (t << occ_sb_shift) >> 1, &occ_sb[(size_t)t * 4]);
			if (!R.d_occ_sb.ensure(occ_sb.size() * 8)) { last_error = "device allocation failed"; return false; }
			rt_h2d(R.d_occ_sb.p, occ_sb.data(), occ_sb.size() * 8);
		}
	}
	if (from) {
		const bool dense = from->d_sa_dense.p != nullptr;
		const size_t sa_bytes = dense ? from->d_sa_dense.cap : (size_t)img.n_sa * 8;
		bool ok = d_bwt.ensure((size_t)n_blocks * 64 + 64) && (dense ? d_sa_dense : d_sa).ensure(sa_bytes) && d_pac.ensure(pac_bytes + 16);
		if (!ok) { last_error = "device allocation failed"; return false; }
		ok = rt_peer_copy(d_bwt.p, R.dev, from->d_bwt.p, from->dev, (size_t)n_blocks * 64)
		     && rt_peer_copy((dense ? d_sa_dense : d_sa).p, R.dev, (dense ? from->d_sa_dense : from->d_sa).p, from->dev, dense ? (size_t)((img.seq_len + 1 + from->dx.sa_intv - 1) / from->dx.sa_intv) * 8 : sa_bytes)
		     && rt_peer_copy(d_pac.p, R.dev, from->d_pac.p, from->dev, pac_bytes);
		if (!ok) { last_error = "peer copy of the index failed"; return false; }
	} else {
		{ // the occ-interleaved BWT goes up as it is in the image, then every 64-byte block is re-laid out for the device
		  // (two 32-byte sectors: u32 counts, symbols as two bit-planes; fm.cuh); the staging copy is released
			DBuf raw;
			if (!raw.ensure(bwt_bytes + 64) || !d_bwt.ensure((size_t)n_blocks * 64 + 64)) { raw.release(); last_error = "device allocation failed"; return false; }
			rt_h2d(raw.p, img.bwt, bwt_bytes);
			RT_RELAYOUT((const u32*)raw.p, (u64)img.bwt_size, (u32*)d_bwt.p, n_blocks, occ_sb_shift ? (const u64*)R.d_occ_sb.p : (const u64*)nullptr, occ_sb_shift);
			rt_sync();
			raw.release();
		}
		if (!d_sa.ensure((size_t)img.n_sa * 8) || !d_pac.ensure(pac_bytes + 16)) { last_error = "device allocation failed"; return false; }
		rt_h2d(d_sa.p, img.sa, (size_t)img.n_sa * 8);
		rt_h2d(d_pac.p, img.pac, pac_bytes);
	}
	std::vector<int64_t> off((size_t)img.n_seqs + 1);
	for (int i = 0; i < img.n_seqs; ++i) off[i] = img.contigs[i].offset;
	off[img.n_seqs] = img.l_pac;
	std::vector<double> lt(LOG_TAB_N);
	for (int i = 0; i < LOG_TAB_N; ++i) lt[i] = log((double)i); // glibc values (hazard H5)
	std::string names; std::vector<uint32_t> noff;
	for (auto& c : img.contigs) { noff.push_back((uint32_t)names.size()); names += c.name; names.push_back('\0'); }
	noff.push_back((uint32_t)names.size());
	std::vector<u8> alt((size_t)img.n_seqs + 1, 0);
	for (int i = 0; i < img.n_seqs; ++i) alt[i] = img.contigs[i].is_alt ? 1 : 0;
	if (!(d_ctg.ensure(off.size() * 8 + alt.size()) && d_log.ensure(lt.size() * 8) && d_names.ensure(names.size() + 1) && d_name_off.ensure(noff.size() * 4))) { last_error = "device allocation failed"; return false; }
	rt_h2d(d_ctg.p, off.data(), off.size() * 8);
	rt_h2d((u8*)d_ctg.p + off.size() * 8, alt.data(), alt.size());
	rt_h2d(d_log.p, lt.data(), lt.size() * 8);
	rt_h2d(d_names.p, names.data(), names.size());
	rt_h2d(d_name_off.p, noff.data(), noff.size() * 4);
	memset(&dx, 0, sizeof(dx));
	dx.primary = img.primary; memcpy(dx.L2, img.L2, 40); dx.seq_len = img.seq_len;
	dx.bwt = (const uint4*)d_bwt.p; dx.occ_sb = occ_sb_shift ? (const u64*)R.d_occ_sb.p : nullptr; dx.occ_sb_shift = occ_sb_shift; dx.sa = (const u64*)d_sa.p; dx.sa_intv = (u32)img.sa_intv;
	dx.l_pac = img.l_pac; dx.n_seqs = img.n_seqs; dx.ctg_off = (const i64*)d_ctg.p; dx.pac = (const u8*)d_pac.p;
	dx.ctg_alt = (const u8*)d_ctg.p + off.size() * 8;
	dx.log_tab = (const double*)d_log.p; dx.log_tab_n = LOG_TAB_N;
	rt_sync();
	if (from) {
		if (from->d_sa_dense.p) { dx.sa = d_sa_dense.as<u64>(); dx.sa_intv = from->dx.sa_intv; }
		if (from->dx.kmer_tab) {
			const size_t kb = (size_t)kmer_level_off(from->dx.kmer_k + 1) * 16;
			if (d_kmer.ensure(kb) && rt_peer_copy(d_kmer.p, R.dev, from->d_kmer.p, from->dev, kb)) { rt_sync(); dx.kmer_tab = d_kmer.as<uint4>(); dx.kmer_k = from->dx.kmer_k; }
		}
		return g_last_cuda_ok();
	}
	// Spend HBM on the suffix array: the image samples every 32nd row (bwt_sa then walks ~15.5 LF steps per lookup, each
	// a random 64-byte read); one LF-walk per 8th row at load time gives an 8-row sampling (6 GB at 3 Gbp) and ~3.5 steps.
	// SA values are layout-independent, so results are unchanged.  BWA_B200_SA_INTV=32 keeps the image's sampling.
	{
		const int want = cfg.sa_intv;
		if (want >= 1 && want < (int)dx.sa_intv && (want & (want - 1)) == 0) {
			const uint64_t n_rows = img.seq_len + 1, n_out = (n_rows + want - 1) / want;
			if (d_sa_dense.ensure((size_t)n_out * 8)) {
				RT_DENSIFY(dx, (u64)n_rows, (u32)want, d_sa_dense.as<u64>());
				rt_sync();
				dx.sa = d_sa_dense.as<u64>(); dx.sa_intv = (u32)want;
				d_sa.release();
			}
		}
	}
	// Spend HBM on the first steps of every forward scan too: the bidirectional interval of every string of 2..K bases
	// (4^K * 16 * 4/3 bytes: 358 MB at K = 12), built level by level from the index itself, so a scan's first K - 1
	// extensions are one 16-byte read each (mostly L2 hits) instead of two 64-byte block walks.  Same intervals, so
	// results are unchanged.  Only worth it when the index is much larger than the tables.
	if (cfg.kmer_k >= 2 && img.seq_len >= ((uint64_t)1 << (2 * cfg.kmer_k + 2))) {
		const int K = cfg.kmer_k;
		if (d_kmer.ensure((size_t)kmer_level_off(K + 1) * 16)) {
			for (int L1 = 2; L1 <= K; ++L1) RT_KMER_LEVEL(dx, d_kmer.as<uint4>(), L1, (u64)1 << (2 * L1));
			rt_sync();
			if (g_last_cuda_ok()) { dx.kmer_tab = d_kmer.as<uint4>(); dx.kmer_k = K; }
		}
	}
	return g_last_cuda_ok();
}

inline size_t Batch::workspace_bytes()
{
	DBuf* all[] = { &b_blob, &b_seq, &b_seqoff, &b_intv, &b_nintv, &b_nseed, &b_smem,
	                &b_seedoff, &b_rbeg, &b_S, &b_C, &b_ord, &b_tmp32, &b_ck, &b_ks, &b_srt, &b_regtmp, &b_nck, &b_nks, &b_nreg, &b_regcap, &b_regoff,
	                &b_regs, &b_snap, &b_ztmp, &b_pairv, &b_recs, &b_nrec, &b_cand, &b_pairtab, &b_var, &b_misc, &b_outwords, &b_outoff, &b_out,
	                &b_wsz, &b_wscig, &b_wsmd, &b_wsxa, &b_wsbrow, &b_nb, &b_swtask, &b_swstate, &b_swlist, &b_tbuf, &b_redo, &b_heavy, &b_multi, &b_cand2, &b_specstate, &b_spectask, &b_nreg0, &b_rsheavy, &b_wsrt, &b_wsrkey, &b_wsrrank, &b_exttask, &b_extres, &b_ckrmax, &b_chainheavy, &b_gatask, &b_gares, &b_gaitem0, &b_gaitemn, &b_seedpool, &b_seedused, &b_seedtask, &b_seedfb,
	                &b_xtkey, &b_xtkey2, &b_xtidx, &b_xtidx2, &b_wsrows, &b_flttask, &b_tswhmax, &b_tswbrow, &b_swrev, &b_swfb, &b_gakey, &b_gakey2, &b_gaiota, &b_gaorder, &b_gafb };
	size_t n = 0;
	for (DBuf* b : all) n += b->cap;
	if (eng->cfg.debug_stats) { // the largest buffers by name
		static const char* const names[] = { "b_blob", "b_seq", "b_seqoff", "b_intv", "b_nintv", "b_nseed", "b_smem", "b_seedoff", "b_rbeg", "b_S", "b_C", "b_ord", "b_tmp32", "b_ck", "b_ks", "b_srt", "b_regtmp", "b_nck", "b_nks", "b_nreg", "b_regcap", "b_regoff", "b_regs", "b_snap", "b_ztmp", "b_pairv", "b_recs", "b_nrec", "b_cand", "b_pairtab", "b_var", "b_misc", "b_outwords", "b_outoff", "b_out", "b_wsz", "b_wscig", "b_wsmd", "b_wsxa", "b_wsbrow", "b_nb", "b_swtask", "b_swstate", "b_swlist", "b_tbuf", "b_redo", "b_heavy", "b_multi", "b_cand2", "b_specstate", "b_spectask", "b_nreg0", "b_rsheavy", "b_wsrt", "b_wsrkey", "b_wsrrank", "b_exttask", "b_extres", "b_ckrmax", "b_chainheavy", "b_gatask", "b_gares", "b_gaitem0", "b_gaitemn", "b_seedpool", "b_seedused", "b_seedtask", "b_seedfb", "b_xtkey", "b_xtkey2", "b_xtidx", "b_xtidx2", "b_wsrows", "b_flttask", "b_tswhmax", "b_tswbrow", "b_swrev", "b_swfb", "b_gakey", "b_gakey2", "b_gaiota", "b_gaorder", "b_gafb" };
		for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) if (all[k]->cap >= ((size_t)1 << 29)) fprintf(stderr, "[bwamem-b200]   %-12s %.2f GB\n", names[k], (double)all[k]->cap / 1e9);
	}
	return n;
}
inline void Batch::release_all()
{
	DBuf* all[] = { &b_blob, &b_seq, &b_seqoff, &b_intv, &b_nintv, &b_nseed, &b_smem,
	                &b_seedoff, &b_rbeg, &b_S, &b_C, &b_ord, &b_tmp32, &b_ck, &b_ks, &b_srt, &b_regtmp, &b_nck, &b_nks, &b_nreg, &b_regcap, &b_regoff,
	                &b_regs, &b_snap, &b_ztmp, &b_pairv, &b_recs, &b_nrec, &b_cand, &b_pairtab, &b_var, &b_misc, &b_outwords, &b_outoff, &b_out,
	                &b_wsz, &b_wscig, &b_wsmd, &b_wsxa, &b_wsbrow, &b_nb, &b_swtask, &b_swstate, &b_swlist, &b_tbuf, &b_redo, &b_heavy, &b_multi, &b_cand2, &b_specstate, &b_spectask, &b_nreg0, &b_rsheavy, &b_wsrt, &b_wsrkey, &b_wsrrank, &b_exttask, &b_extres, &b_ckrmax, &b_chainheavy, &b_gatask, &b_gares, &b_gaitem0, &b_gaitemn, &b_seedpool, &b_seedused, &b_seedtask, &b_seedfb,
	                &b_xtkey, &b_xtkey2, &b_xtidx, &b_xtidx2, &b_wsrows, &b_flttask, &b_tswhmax, &b_tswbrow, &b_swrev, &b_swfb, &b_gakey, &b_gakey2, &b_gaiota, &b_gaorder, &b_gafb };
	for (DBuf* b : all) b->release();
#ifndef B200_SIM
	for (int t = 0; t < NUP; ++t) { if (up_stream[t]) cudaStreamDestroy(up_stream[t]); if (up_done[t]) cudaEventDestroy(up_done[t]); up_stream[t] = nullptr; up_done[t] = nullptr; }
	rt_host_free(pin); pin = nullptr; pin_cap = 0;
	rt_host_free(pin_out); pin_out = nullptr; pin_out_cap = 0;
	for (cudaEvent_t e : out_done) cudaEventDestroy(e);
	out_done.clear();
#endif
}

inline void Engine::close_all()
{
	for (Batch* b : ctxs) { rt_bind(b->rt); b->release_all(); }
	for (size_t k = 0; k < reps.size(); ++k) { if (k * (size_t)cfg.contexts < ctxs.size()) rt_bind(ctxs[k * (size_t)cfg.contexts]->rt); reps[k]->release_all(); }
	for (Batch* b : ctxs) { rt_ctx_destroy(b->rt); delete b; }
	for (Replica* r : reps) delete r;
	ctxs.clear(); reps.clear();
	rt_bind(nullptr);
	if (map) munmap(map, map_len);
	map = nullptr;
}

// misc block layout (u32 words): [20] warp-replayed rescue reads  [0] err  [1] work counter  [2] var cursor  [3] rescue task count  [4] max nb  [5] redo count  [6] heavy-read count  [7] multi-hit read count  [8..] Counters
// Stage a request: one host pass over the blob for the read boundaries, then upload of bases and offsets.
// The wire format carries no byte length (u32 count, then NUL-terminated reads; BwaMemAligner.java:198-209), so the read
// boundaries have to be found by a sequential scan that must not look past the last NUL.  While the host scans, a
// helper thread already uploads the scanned prefix into the (grow-only) device buffer of the previous calls; a
// request larger than that buffer falls back to scan-then-upload.
inline bool Batch::stage(const char* blob_in, bool wait)
{
	st_valid = false;
	uint32_t n_reads; memcpy(&n_reads, blob_in, 4);
	const char* base = blob_in + 4;
	st_seq_off.resize((size_t)n_reads + 1);
	int max_len = 0;
	uint64_t uploaded = 0;
#ifndef B200_SIM
	// While the host scans, NUP helper threads move the scanned prefix in alternating 4 MiB chunks: each copies its chunk
	// into a pinned staging buffer (a pageable cudaMemcpy would be staged by the driver on ONE thread at ~9 GB/s) and
	// queues a true DMA on its own stream.
	std::atomic<uint64_t> scanned(0);
	std::atomic<int> state(0); // 0 scanning, 1 scan finished, 2 abandon
	const int dev = rep->dev;
	const uint64_t dcap = b_blob.cap;
	char* dptr = (char*)b_blob.p;
	const uint64_t CH = 4u << 20;
	std::thread copier[NUP]; uint64_t reached[NUP] = { 0, 0, 0, 0 };
	bool overlap = n_reads >= 4096 && dcap > 0 && !eng->cfg.no_stage_overlap;
	if (overlap && pin_cap < dcap) {
		rt_host_free(pin);
		pin = (char*)rt_host_alloc(dcap); pin_cap = pin ? dcap : 0;
		if (!pin) overlap = false;
	}
	if (overlap) {
		for (int t = 0; t < NUP && overlap; ++t) if (!up_stream[t]) {
			if (!cu_ok(cudaStreamCreateWithFlags(&up_stream[t], cudaStreamNonBlocking), "cudaStreamCreate") || !cu_ok(cudaEventCreateWithFlags(&up_done[t], cudaEventDisableTiming), "cudaEventCreate")) overlap = false;
		}
	}
	std::atomic<int> copy_failed(0);
	if (overlap) {
		for (int t = 0; t < NUP; ++t) copier[t] = std::thread([&, t]() {
			cudaSetDevice(dev);
			for (uint64_t c = (uint64_t)t; ; c += NUP) { // chunk c = bytes [c*CH, (c+1)*CH); reached[t] = first chunk start this thread has NOT sent
				const uint64_t lo = c * CH;
				uint64_t hi = lo + CH;
				reached[t] = lo;
				for (;;) {
					const int stt = state.load(std::memory_order_acquire);
					const uint64_t avail = scanned.load(std::memory_order_acquire);
					if (stt == 2 || lo >= dcap) goto out;
					if (avail >= hi && hi <= dcap) break;
					if (stt == 1) { hi = avail < dcap ? avail : dcap; break; } // the last, partial chunk
					std::this_thread::sleep_for(std::chrono::microseconds(20)); // (a spinning helper would take a core from another caller's scan)
				}
				if (hi <= lo) break;
				memcpy(pin + lo, base + lo, hi - lo);
				if (cudaMemcpyAsync(dptr + lo, pin + lo, hi - lo, cudaMemcpyHostToDevice, up_stream[t]) != cudaSuccess) { copy_failed.store(1); break; }
				reached[t] = lo + NUP * CH;
				if (hi - lo < CH) break;
			}
		out:
			if (cudaEventRecord(up_done[t], up_stream[t]) != cudaSuccess) copy_failed.store(1);
		});
	}
	scan_reads(base, n_reads, st_seq_off.data(), &max_len, overlap ? &scanned : nullptr);
#else
	{
		const char* p = base;
		for (uint32_t i = 0; i < n_reads; ++i) {
			size_t l = strlen(p);
			st_seq_off[i] = (uint64_t)(p - base);
			if ((int)l > max_len) max_len = (int)l;
			p += l + 1;
		}
		st_seq_off[n_reads] = (uint64_t)(p - base);
	}
#endif
	st_n_reads = n_reads; st_max_len = max_len; st_blob_len = st_seq_off[n_reads];
#ifndef B200_SIM
	if (overlap) {
		scanned.store(st_blob_len, std::memory_order_release);
		state.store(st_blob_len + 16 <= dcap ? 1 : 2, std::memory_order_release);
		for (int t = 0; t < NUP; ++t) copier[t].join();
		if (copy_failed.load()) { cudaGetLastError(); last_error = "staging copy failed"; return false; }
		if (st_blob_len + 16 <= dcap) {
			// the chunks alternate between the threads: everything below the lowest of the frontiers is on its way
			uploaded = reached[0];
			for (int t = 1; t < NUP; ++t) uploaded = reached[t] < uploaded ? reached[t] : uploaded;
			if (uploaded > st_blob_len) uploaded = st_blob_len;
			for (int t = 0; t < NUP; ++t) cudaStreamWaitEvent(rt_stream(), up_done[t], 0);
		}
	}
#endif
	if (n_reads) {
		if (!(b_blob.ensure(st_blob_len + 16) && b_seqoff.ensure(((size_t)n_reads + 1) * 8))) { last_error = "device allocation failed"; return false; }
		if (uploaded < st_blob_len) rt_h2d((char*)b_blob.p + uploaded, base + uploaded, st_blob_len - uploaded);
		rt_h2d(b_seqoff.p, st_seq_off.data(), st_seq_off.size() * 8);
		if (wait) rt_sync();
	}
	st_valid = true;
	return true;
}

// The record stream back to the caller's malloc'd buffer (Java frees it with free(): ...BwaMemIndex.c:157-160).  A copy from
// the device straight into fresh pageable memory runs at ~2 GB/s (the driver stages it and every page faults on first
// touch); instead the stream comes down in chunks into the context's pinned buffer at DMA speed while helper threads copy
// the chunks that have arrived into the destination, faulting its pages in parallel.
inline bool Batch::fetch_result(void* dst, const void* dev_src, size_t bytes)
{
#ifndef B200_SIM
	const size_t CH = (size_t)4 << 20;
	enum { NDOWN = 4 };
	if (bytes >= 2 * CH) {
		if (pin_out_cap < bytes) {
			rt_host_free(pin_out);
			const size_t want = bytes + bytes / 4;
			pin_out = (char*)rt_host_alloc(want); pin_out_cap = pin_out ? want : 0;
		}
		const size_t n_ch = (bytes + CH - 1) / CH;
		bool ok = pin_out != nullptr;
		while (ok && out_done.size() < n_ch) { cudaEvent_t e; ok = cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess; if (ok) out_done.push_back(e); }
		if (ok) {
			for (size_t c = 0; c < n_ch && ok; ++c) {
				const size_t lo = c * CH, len = bytes - lo < CH ? bytes - lo : CH;
				ok = cu_ok(cudaMemcpyAsync(pin_out + lo, (const char*)dev_src + lo, len, cudaMemcpyDeviceToHost, rt_stream()), "d2h") && cu_ok(cudaEventRecord(out_done[c], rt_stream()), "d2h event");
			}
			if (ok) {
				std::atomic<int> failed(0);
				const int dev = rep->dev;
				std::thread th[NDOWN];
				for (int t = 0; t < NDOWN; ++t) th[t] = std::thread([&, t]() {
					cudaSetDevice(dev);
					for (size_t c = (size_t)t; c < n_ch; c += NDOWN) {
						if (cudaEventSynchronize(out_done[c]) != cudaSuccess) { failed.store(1); return; }
						const size_t lo = c * CH, len = bytes - lo < CH ? bytes - lo : CH;
						memcpy((char*)dst + lo, pin_out + lo, len);
					}
				});
				for (int t = 0; t < NDOWN; ++t) th[t].join();
				if (!failed.load()) return true;
			}
			cudaGetLastError();
			rt_sync();
			return false;
		}
		cudaGetLastError(); // (no pinned memory to be had: the plain copy below)
	}
#endif
	rt_d2h(dst, dev_src, bytes);
	return true;
}

inline void* Batch::align(const MemOpt* opt_in, const PeStat* pes0, char* blob_in, size_t* out_size)
{
	last_error.clear();
	*out_size = 0;
	hp.begin();
	if (!stage(blob_in, false)) return nullptr; // same stream: the kernels queue up behind the upload
	hp.mark("stage");
	void* r = run(opt_in, pes0, out_size, true);
	hp.mark("d2h_out");
	hp.report();
	return r;
}

// Run the device pipeline on the staged request.  host_output: copy the record stream into a malloc'd buffer (the
// reference contract); otherwise it stays in device memory (bench.py's HBM-resident measurement) and a non-null
// sentinel is returned.
inline void* Batch::run(const MemOpt* opt_in, const PeStat* pes0, size_t* out_size, bool host_output)
{
	*out_size = 0;
	if (!st_valid) { last_error = "no staged request"; return nullptr; }
	rt_timer_begin();
	MemOpt opt = *opt_in; // the options live in a caller-owned buffer (…BwaMemIndex.c:146): snapshot them
	const EngineConfig& cfg = eng->cfg; const ImageView& img = eng->img; const DevIndex& dx = rep->dx;
	DBuf& d_names = rep->d_names; DBuf& d_name_off = rep->d_name_off;
	const uint32_t n_reads = st_n_reads;
	const int max_len = st_max_len;
	const uint64_t blob_len = st_blob_len;
	const bool pe = (opt.flag & MEM_F_PE) != 0;
	const int n_units = pe ? (int)(n_reads >> 1) : (int)n_reads;
	if (n_reads == 0) { void* r = malloc(1); return r; }
	// ---- option / shape validation: fail loudly instead of approximating
	if (opt.e_del <= 0 || opt.e_ins <= 0 || opt.a <= 0 || opt.max_occ <= 0 || opt.min_seed_len <= 0) { last_error = "unsupported option values"; return nullptr; }
	if (max_len > 65535) { last_error = "reads longer than 65535 bases are not supported"; return nullptr; } // (query coordinates are packed in 16 bits in the chain filter)
	// mem_flt_chained_seeds (B10) acts on a read of l bases unless min_l > 0.05 l, min_l = 1.1 min_chain_weight or 5.5 ln l.
	// 5.5 ln l - 0.05 l is concave, so with default options the test fails first for the LONGEST read: when it holds
	// for max_len (and for min_seed_len) no read of the batch needs the filter and its launches are skipped.
	bool flt_possible = false;
	if (max_len >= opt.min_seed_len) {
		const double min_l = opt.min_chain_weight ? 1.1f * opt.min_chain_weight : 5.5f * log((double)max_len);
		const double at_min = opt.min_chain_weight ? min_l : 5.5f * log((double)opt.min_seed_len);
		flt_possible = !(min_l > 0.05f * max_len) || !(at_min > 0.05f * opt.min_seed_len);
	}
	// Batches with long reads (assembled contigs of 0.5-50 kbp, BwaMemAligner.java:140-146): the per-thread / per-warp
	// scratch grows with the read length (candidate lists, DP rows, traceback), so such batches run every read on the
	// warp-level kernels with rows in global memory, on as many slots as a fixed scratch budget allows.
	const bool long_mode = max_len > 1000 || cfg.force_long_mode;
	const int wm_cap = ((max_len + 15) / 16) * 16 + 32;
	// SMEMs of a read.  A read from a family of many diverged repeat copies has a new SMEM every base or two: up to
	// len - min_seed_len from round 1 and as many again from a re-seeding scan (185 seen for 150 bases on the 3 Gbp bench
	// genome), so 2 len + 64; long reads: one per ~6 bases.  A read that still overflows has its list clamped
	// (kernels.cuh SeedSrc::next), the flag is read right after the seeding stage, and the whole batch is run again with
	// twice the capacity, which this context then keeps (intv_scale).
	const int intv_cap0 = cfg.tiny_intv_cap ? 4 : (max_len > 512 ? 512 + (max_len - 512) / 6 : 2 * (max_len < 64 ? 64 : max_len) + 64);
	const int intv_cap = intv_cap0 * intv_scale;
	const int L_ = max_len > 0 ? max_len : 1;
	// traceback of the banded global alignment: n_col <= min(l_query, 2 w + 1) bytes per reference row, w <= max(4 opt.w, dl + 3)
	// These capacities of the warp-level scratch are bounds for ordinary data, not proofs (an XA string of 200 ALT hits on
	// contigs with long names, a CIGAR of a pathological alignment).  An item that outgrows one raises a flag and is cut
	// short; the flags are read at the end of the batch and the WHOLE batch is run again with every one of them doubled
	// (scratch_scale, kept by the context) instead of failing the call.  (B200_TEST_TINY_WARP_SCRATCH starts at 1/64.)
	const int sdiv = cfg.tiny_warp_scratch ? 64 : 1;
	const size_t ws_z_cap = std::max<size_t>(64, (size_t)((L_ < 8 * opt.w + 17 ? L_ : 8 * opt.w + 17) + 4) * (size_t)(L_ + 8 * opt.w + 64) * (size_t)scratch_scale / sdiv);
	const int ws_cigar_cap = std::max(4, (3 * L_ + 4 * opt.w + 64) * scratch_scale / sdiv), ws_md_cap = std::max(8, (4 * L_ + 8 * opt.w + 128) * scratch_scale / sdiv),
	          ws_xa_cap = std::max(64, (8192 + (long_mode ? 8 * L_ : 0)) * scratch_scale / sdiv), ws_rt_cap = 1024;
	rt_set_slot_caps(0, 0);
	if (long_mode) {
		const size_t per_thread = (size_t)(max_len + 1) * sizeof(Intv) + thread_scratch_bytes(wm_cap, 64, 48, 192, 1024);
		const size_t per_warp = ws_z_cap + (size_t)ws_cigar_cap * 4 + ws_md_cap + ws_xa_cap + (size_t)ws_rt_cap * (sizeof(Reg) + 12) + warp_smem_bytes(wm_cap) + (size_t)(max_len + 10064) * 8;
		const size_t budget_t = (size_t)6 << 30, budget_w = (size_t)12 << 30;
		int64_t tb = (int64_t)(budget_t / per_thread / THREADS_PER_BLOCK), wb = (int64_t)(budget_w / per_warp / WARPS_PER_BLOCK);
		rt_set_slot_caps(tb < 1 ? 1 : tb, wb < 1 ? 1 : wb);
	}

	BatchParams P;
	memset(&P, 0, sizeof(P));
	P.x = dx; P.opt = opt; P.n_reads = (int)n_reads; P.n_units = n_units; P.blob_len = blob_len;
	P.ctg_names = (const char*)d_names.p; P.ctg_name_off = (const u32*)d_name_off.p;
	bool ok = true;
	ok &= b_seq.ensure(blob_len + 16);
	ok &= b_intv.ensure((size_t)n_reads * intv_cap * sizeof(Intv)) && b_nintv.ensure((size_t)n_reads * 4) && b_nseed.ensure((size_t)n_reads * 4);
	const int64_t tslots = rt_thread_slots(), wslots = rt_warp_slots();
	ok &= b_smem.ensure((size_t)tslots * (size_t)(max_len + 1) * sizeof(Intv));
	ok &= b_seedoff.ensure(((size_t)n_reads + 1) * 8) && b_nck.ensure((size_t)n_reads * 4) && b_nks.ensure((size_t)n_reads * 4) && b_nreg.ensure((size_t)n_reads * 4);
	ok &= b_regcap.ensure((size_t)n_reads * 4) && b_regoff.ensure(((size_t)n_reads + 1) * 8) && b_nrec.ensure((size_t)n_reads * 4);
	ok &= b_outwords.ensure((size_t)n_reads * 4) && b_outoff.ensure(((size_t)n_reads + 1) * 8) && b_misc.ensure(1024);
	if (!ok) { last_error = "device allocation failed"; return nullptr; }
	rt_memset(b_misc.p, 0, 1024);
	P.err = b_misc.as<u32>(); P.work_counter = b_misc.as<u32>() + 1; P.var_cursor = b_misc.as<u32>() + 2; P.cnt = (Counters*)(b_misc.as<u32>() + 8);
	P.blob = b_blob.as<u8>(); P.seq = b_seq.as<u8>(); P.seq_off = b_seqoff.as<u64>();
	P.intv = b_intv.as<Intv>(); P.intv_cap = intv_cap; P.n_intv = b_nintv.as<u32>(); P.n_seed = b_nseed.as<u32>();
	P.smem_scratch = b_smem.as<Intv>(); P.smem_cap = max_len + 1;
	P.seed_shared = (max_len <= 255 && img.seq_len < (1ULL << 40)) ? 1 : 0;
	P.seed_off = b_seedoff.as<u64>(); P.n_ck = b_nck.as<u32>(); P.n_ks = b_nks.as<u32>(); P.n_reg = b_nreg.as<u32>();
	P.reg_cap = b_regcap.as<u32>(); P.reg_off = b_regoff.as<u64>(); P.n_rec = b_nrec.as<u32>();
	P.out_words = b_outwords.as<u32>(); P.out_off = b_outoff.as<u64>();
	P.wm_cap = wm_cap; P.wm_smem_cap = wm_cap;
	if (long_mode) { // the DP rows of a warp slot in global memory (kernels.cuh make_warp_ctx), for every warp kernel of the batch
		P.ws_rows_stride = (warp_smem_bytes(wm_cap) + 255) & ~(size_t)255;
		if (!b_wsrows.ensure((size_t)wslots * P.ws_rows_stride)) { last_error = "device allocation failed (warp rows)"; return nullptr; }
		P.ws_rows = b_wsrows.as<u8>(); P.wm_smem_cap = 0;
	}

	// ---- stage 1: encode + seed
	RT_THREADS(body_encode, P, (int64_t)blob_len);
	P.seed_lcap = 8; P.seed_shared_read = 0; // few candidates survive per position; the read is walked sequentially (L1)
	if (cfg.seed_lcap >= 1 && cfg.seed_lcap <= SEED_LCAP) P.seed_lcap = cfg.seed_lcap; // tuning aids
	if (cfg.seed_shared_read >= 0) P.seed_shared_read = cfg.seed_shared_read;
	{
		const size_t ssm = P.seed_shared ? seed_block_smem_bytes(THREADS_PER_BLOCK, max_len, P.seed_lcap, P.seed_shared_read) : 0;
		if (cfg.seed_one_launch) RT_THREADS_QUEUE_SMEM(body_seed_q, P, (int64_t)n_reads, ssm);
		else if (!P.seed_shared || cfg.seed_fused_rounds) { // rounds 1/2 in one machine (two kinds of steps), round 3 alone
			RT_THREADS_QUEUE_SMEM(body_seed12_q, P, (int64_t)n_reads, ssm);
			RT_THREADS_QUEUE_SMEM(body_seed3_q, P, (int64_t)n_reads, ssm);
		} else {
			// every launch has ONE kind of extension step: forward scans (lane per read), backward sweeps (lane per scan),
			// for round 1 and again for round 2, then round 3 (forward-only seeds; it also orders and counts the list)
			// Candidates of all scans of a round: 6 x length + 16 per read (14 KB at 150 bases).  Ordinary reads use a tenth of
			// it, but the few made of low-complexity sequence need it all -- with a third of it (measured) 20-30 reads per
			// million fall to the one-machine kernel below, where a single thread takes 30 ms over one of them.  So the cap
			// stays and the BATCH is seeded in slices of reads instead, as many as keep the pool under cfg.seed_pool_gb
			// (default 20 GB: 500 000 pairs of 150 bases need 17 GB and run as one slice, of 250 bases 23 GB and run as two; rounds 1/2
			// of a slice are complete before the next begins; measured cost: 2.4 ms of a 105 ms step for two slices).
			P.seed_pool_cap = (u32)(6 * max_len + 16);
			if (cfg.tiny_seed_pool) P.seed_pool_cap = 24;
			int n_slices = (int)(((size_t)n_reads * P.seed_pool_cap * 16 + ((size_t)cfg.seed_pool_gb << 30) - 1) / ((size_t)cfg.seed_pool_gb << 30));
			if (cfg.seed_slices > 0) n_slices = cfg.seed_slices;
			if (n_slices < 1) n_slices = 1;
			if ((int64_t)n_slices > (int64_t)n_reads) n_slices = (int)n_reads;
			const int64_t slice_reads = ((int64_t)n_reads + n_slices - 1) / n_slices;
			P.seed_task_cap = (u32)std::min<uint64_t>((uint64_t)0x7fffffff, (uint64_t)slice_reads * 24 + 1024); // one task per forward scan (a scan per sequencing error, roughly); reads whose scans do not fit are redone below
			if (cfg.tiny_seed_tasks) P.seed_task_cap = 48;
			ok = b_seedpool.ensure((size_t)slice_reads * P.seed_pool_cap * 16) && b_seedused.ensure((size_t)n_reads * 4) && b_seedtask.ensure((size_t)P.seed_task_cap * sizeof(SeedTask)) && b_seedfb.ensure(((size_t)n_reads + 1) * 4);
			if (!ok) { last_error = "device allocation failed (seeding pools)"; return nullptr; }
			P.seed_pool = b_seedpool.as<uint4>(); P.seed_pool_used = b_seedused.as<u32>(); P.seed_tasks = b_seedtask.as<SeedTask>(); P.n_seed_tasks = b_misc.as<u32>() + 24; // words 24..27: the 64-bit reservation counter and the counts of a full list (fm.cuh seed_task_reserve)
			P.seed_fb = b_seedfb.as<u32>(); P.n_seed_fb = b_misc.as<u32>() + 28;
			rt_memset(P.n_intv, 0, (size_t)n_reads * 4);
			// The scans with many candidates (a warp each: few, long-running warps) run on the auxiliary stream beside the
			// lane-per-scan sweep of the others; both only append to per-read lists that round 3 orders afterwards.
			const bool overlap = !cfg.seed_no_overlap;
			for (int64_t r0 = 0; r0 < (int64_t)n_reads; r0 += slice_reads) {
				const int64_t nr = std::min<int64_t>(slice_reads, (int64_t)n_reads - r0);
				P.seed_read_base = r0;
				if (r0) rt_memset(P.n_seed_tasks, 0, 16);
				RT_THREADS_QUEUE_SMEM(body_seedF1_q, P, nr, 0);
				if (overlap) rt_fork_aux();
				RT_WARPS_DEVN(wbody_seedB_heavy, P, P.n_seed_tasks + 1, nr);
				if (overlap) rt_back_to_main();
				RT_THREADS_QUEUE_SMEM(body_seedB_q, P, (int64_t)P.seed_task_cap, seed_bwd_smem_bytes(THREADS_PER_BLOCK));
				if (overlap) rt_join_aux("wbody_seedB_heavy");
				rt_memset(P.n_seed_tasks, 0, 16);
				RT_THREADS_QUEUE_SMEM(body_seedF2_q, P, nr, 0);
				if (overlap) rt_fork_aux();
				RT_WARPS_DEVN(wbody_seedB_heavy, P, P.n_seed_tasks + 1, nr);
				if (overlap) rt_back_to_main();
				RT_THREADS_QUEUE_SMEM(body_seedB2_q, P, (int64_t)P.seed_task_cap, seed_bwd_smem_bytes(THREADS_PER_BLOCK));
				if (overlap) rt_join_aux("wbody_seedB_heavy");
			}
			P.seed_read_base = 0;
			RT_THREADS_QUEUE_SMEM(body_seed12_fb_q, P, (int64_t)n_reads, ssm); // (usually an empty list)
			RT_THREADS_QUEUE_SMEM(body_seed3_q, P, (int64_t)n_reads, ssm);
		}
	}
	hp.mark("launch1");
	if (cfg.debug_stats) { // the longest interval list of the batch (capacity: intv_cap per read)
		std::vector<uint32_t> ni((size_t)n_reads);
		rt_d2h(ni.data(), P.n_intv, (size_t)n_reads * 4);
		size_t am = 0; for (size_t r = 0; r < (size_t)n_reads; ++r) if (ni[r] > ni[am]) am = r;
		std::vector<uint64_t> so(2); rt_d2h(so.data(), P.seq_off + am, 16);
		std::string rd((size_t)(so[1] - so[0]), '\0'); rt_d2h(&rd[0], P.seq + so[0], rd.size());
		for (char& c : rd) c = (unsigned char)c < 4 ? "ACGT"[(int)c] : 'N';
		fprintf(stderr, "[bwamem-b200] longest interval list: %u of %d (read %zu: %s)\n", ni[am], intv_cap, am, rd.c_str());
	}
	uint64_t T = rt_scan_u32(P.n_seed, P.seed_off, n_reads);
	if (!g_last_cuda_ok()) { last_error = "device error or allocation failure (scan / sort scratch)"; return nullptr; }
	hp.mark("seed_sync");
	{ // interval lists that did not fit: again, with twice the room
		uint32_t e0 = 0;
		rt_d2h(&e0, P.err, 4);
		if ((e0 & ERR_INTV_OVERFLOW) && intv_scale < 65536) {
			intv_scale *= 2;
			if (cfg.debug_stats) fprintf(stderr, "[bwamem-b200] interval lists overflowed %d entries per read: re-running the batch with %d\n", intv_cap, intv_cap0 * intv_scale);
			return run(opt_in, pes0, out_size, host_output);
		}
	}
	P.n_seed_total = T; last_n_seed = T;
	{
		size_t Tn = (size_t)T + 1;
		ok = b_rbeg.ensure(Tn * 8) && b_S.ensure(Tn * sizeof(Seed)) && b_C.ensure(Tn * sizeof(ChainW)) && b_ord.ensure(Tn * 4) && b_tmp32.ensure(Tn * 4)
		     && b_ck.ensure(Tn * sizeof(ChainK)) && b_ks.ensure(Tn * sizeof(Seed)) && b_srt.ensure(Tn * 8) && b_regtmp.ensure(Tn * sizeof(Reg));
		if (!ok) { last_error = "device allocation failed (seed pools)"; return nullptr; }
	}
	P.sd_rbeg = b_rbeg.as<i64>(); P.S = b_S.as<Seed>(); P.C = b_C.as<ChainW>(); P.ord = b_ord.as<u32>(); P.tmp32 = b_tmp32.as<u32>();
	P.ck = b_ck.as<ChainK>(); P.ks = b_ks.as<Seed>(); P.srt = b_srt.as<u64>(); P.regs_tmp = b_regtmp.as<Reg>();
	// ---- stage 2: SA lookups, chaining
	RT_THREADS(body_sal, P, (int64_t)T);
	if (!b_chainheavy.ensure(((size_t)n_reads + 1) * 4)) { last_error = "device allocation failed"; return nullptr; }
	P.chain_heavy = b_chainheavy.as<u32>(); P.n_chain_heavy = b_misc.as<u32>() + 22;
	P.chain_heavy_seeds = cfg.chain_all_warp ? -1 : (cfg.chain_all_thread ? (1 << 30) : CHAIN_HEAVY_SEEDS);
	if (cfg.chain_no_overlap) { P.chain_listed = 0; RT_THREADS(body_chain, P, (int64_t)n_reads); RT_WARPS_DEVN(wbody_chain_heavy, P, P.n_chain_heavy, (int64_t)n_reads); }
	else { // the warp-per-read kernel (few long-running warps) on the auxiliary stream beside the thread-per-read one
		P.chain_listed = 1;
		RT_THREADS(body_chain_classify, P, (int64_t)n_reads);
		rt_fork_aux();
		RT_WARPS_DEVN(wbody_chain_heavy, P, P.n_chain_heavy, (int64_t)n_reads);
		rt_back_to_main();
		RT_THREADS(body_chain, P, (int64_t)n_reads);
		rt_join_aux("wbody_chain_heavy");
	}
	// ---- warp scratch
	{
		P.ws_z_cap = ws_z_cap; P.ws_cigar_cap = ws_cigar_cap; P.ws_md_cap = ws_md_cap; P.ws_xa_cap = ws_xa_cap;
		ok = b_wsz.ensure((size_t)wslots * P.ws_z_cap) && b_wscig.ensure((size_t)wslots * P.ws_cigar_cap * 4) && b_wsmd.ensure((size_t)wslots * P.ws_md_cap) && b_wsxa.ensure((size_t)wslots * P.ws_xa_cap);
		if (!ok) { last_error = "device allocation failed (warp scratch)"; return nullptr; }
		P.ws_z = b_wsz.as<u8>(); P.ws_cigar = b_wscig.as<u32>(); P.ws_md = b_wsmd.as<char>(); P.ws_xa = b_wsxa.as<char>();
		P.ws_brow = nullptr; P.ws_brow_cap = 0;
		P.ws_rt_cap = ws_rt_cap; // longer region lists take the scalar de-duplication
		ok = b_wsrt.ensure((size_t)wslots * P.ws_rt_cap * sizeof(Reg)) && b_wsrkey.ensure((size_t)wslots * (P.ws_rt_cap + 2) * 8) && b_wsrrank.ensure((size_t)wslots * (P.ws_rt_cap + 2) * 4);
		if (!ok) { last_error = "device allocation failed (warp scratch)"; return nullptr; }
		P.ws_rt = b_wsrt.as<Reg>(); P.ws_rkey = b_wsrkey.as<i64>(); P.ws_rrank = b_wsrrank.as<u32>();
	}
	// ---- thread scratch for the thread-per-item kernels (extension, finalize) and the hand-over lists
	{
		P.t_z_cap = 8192; P.t_cigar_cap = 48; P.t_md_cap = 192; P.t_xa_cap = 1024;
		if (cfg.tiny_scratch) { P.t_z_cap = 1024; P.t_cigar_cap = 6; P.t_md_cap = 24; P.t_xa_cap = 80; } // tests: force hand-overs
		if (long_mode) P.t_z_cap = 64; // (no thread-level alignment runs in this mode)
		P.t_stride = thread_scratch_bytes(P.wm_cap, P.t_z_cap, P.t_cigar_cap, P.t_md_cap, P.t_xa_cap);
		ok = b_tbuf.ensure((size_t)tslots * P.t_stride) && b_redo.ensure(((size_t)n_reads + 1) * 4) && b_heavy.ensure(((size_t)n_reads + 1) * 4);
		if (!ok) { last_error = "device allocation failed (thread scratch)"; return nullptr; }
		P.t_buf = b_tbuf.as<u8>(); P.redo_list = b_redo.as<u32>(); P.n_redo = b_misc.as<u32>() + 5;
		P.heavy_list = b_heavy.as<u32>(); P.n_heavy = b_misc.as<u32>() + 6;
		P.light_max_seeds = 12; P.light_max_chains = 4; P.tx_chunk = 32;
		if (cfg.tx_chunk >= 0) P.tx_chunk = cfg.tx_chunk; // tuning aid (0: whole rows)
		if (cfg.light_seeds >= 0) P.light_max_seeds = cfg.light_seeds; // tuning aids
		if (cfg.light_chains >= 0) P.light_max_chains = cfg.light_chains;
		if (cfg.all_heavy) P.light_max_seeds = -1;
		if (cfg.all_light) { P.light_max_seeds = 1 << 30; P.light_max_chains = 1 << 30; }
	}
	for (int t = 0; t < 5; ++t) { P.rowpack[t] = 0; for (int c2 = 0; c2 < 5; ++c2) P.rowpack[t] |= (u64)(u8)opt.mat[t * 5 + c2] << (8 * c2); }
	// ---- mem_flt_chained_seeds (long reads): re-score the seeds of the kept chains, drop the weak ones (kernels.cuh)
	if (flt_possible) {
		if (!b_flttask.ensure(((size_t)T + 1) * 8)) { last_error = "device allocation failed (seed filter)"; return nullptr; }
		P.flt_tasks = b_flttask.as<u32>(); P.n_flt_tasks = b_misc.as<u32>() + 29;
		RT_THREADS(body_flt_gen, P, (int64_t)n_reads);
		RT_WARPS_DEVN(wbody_flt_sw, P, P.n_flt_tasks, (int64_t)T);
		RT_THREADS(body_flt_apply, P, (int64_t)n_reads);
	}
	// ---- stage 3: extension + de-duplication (thread per light read with shared-memory DP rows, warp per heavy read)
	if (cfg.warp_extend || long_mode) RT_WARPS(wbody_extend, P, (int64_t)n_reads); // the all-warp formulation (long reads; kept for comparison otherwise)
	else {
		int rw = max_len - opt.min_seed_len + 2;
		// cells: H and E as 8-bit halves when no score of the batch can exceed 255 (the per-read test in the kernel is the
		// same bound), else as 16-bit halves.  Narrow cells halve the shared memory per thread: more resident warps.
		int mxs = 0; for (int k = 0; k < 25; ++k) mxs = mxs > opt.mat[k] ? mxs : (int)opt.mat[k];
		const bool narrow = !cfg.ext_wide_cells && (int64_t)max_len * (mxs > opt.a ? mxs : opt.a) <= 255;
		const int cell = narrow ? 2 : 4;
		if (!narrow && rw * cell < 4 * THREAD_RING_N) rw = 4 * THREAD_RING_N / cell; // wide cells: the patching rings reuse the row (at least THREAD_RING_N / 2 slots)
		const size_t budget = 200 * 1024; // per block of THREADS_PER_BLOCK threads
		const size_t qb = (size_t)((wm_cap + 2 + 3) & ~3);
		if (thread_ext_smem_bytes(THREADS_PER_BLOCK, rw, wm_cap, cell) > budget) { // long reads: shorter rows, long extensions go to the warp kernel
			size_t fit = budget / THREADS_PER_BLOCK > qb ? (budget / THREADS_PER_BLOCK - qb) / cell : 0;
			rw = fit * cell >= (size_t)4 * THREAD_RING_N ? (int)fit : 0; // (narrow cells never get here: their scores bound max_len by 255)
		}
		if (rw == 0) RT_WARPS(wbody_extend, P, (int64_t)n_reads);
		else if (!cfg.ext_legacy) {
			// uniform side tasks (extend_thread.cuh): prep -> left sides -> right sides -> replay; see the comment there
			P.t_ext_rw = rw;
			const size_t Tn = (size_t)T + 1;
			ok = b_exttask.ensure(Tn * sizeof(ExtTask)) && b_extres.ensure(Tn * sizeof(ExtRes)) && b_ckrmax.ensure(Tn * 16) && b_xtkey.ensure(Tn * 8) && b_xtkey2.ensure(Tn * 8) && b_xtidx.ensure(Tn * 4) && b_xtidx2.ensure(Tn * 8);
			if (!ok) { last_error = "device allocation failed (extension tasks)"; return nullptr; }
			P.ext_tasks = b_exttask.as<ExtTask>(); P.ext_res = b_extres.as<ExtRes>(); P.ck_rmax = b_ckrmax.as<i64>(); P.n_ext_tasks = b_misc.as<u32>() + 21;
			P.ext_key_l = b_xtkey.as<u32>(); P.ext_key_r = b_xtkey.as<u32>() + Tn; P.ext_order_l = nullptr; P.ext_order_r = nullptr;
			P.ext_sides = 1;
			rt_memset(P.ext_res, 0, Tn * sizeof(ExtRes)); // valid = 0: "not extended ahead of time"
			RT_THREADS(body_ext_prep, P, (int64_t)n_reads);
			RT_WARPS_DEVN(wbody_heavy_prep, P, P.n_heavy, (int64_t)n_reads);
			uint32_t n_t = 0;
			rt_d2h(&n_t, P.n_ext_tasks, 4);
			hp.mark("ext_prep_sync");
			// classes of query length: a launch per class and side, rows sized for the class (short sides are the bulk: small
			// rows let many warps be resident); without the ordering there is one class with the longest rows
			enum { NCLS = 4 };
			const uint32_t thr[NCLS] = { 24u, 48u, 96u, 0x7fffffffu }; // a class holds the sides with fewer than thr bases (and at least the class below's)
			uint32_t thr_key[NCLS]; for (int k = 0; k < NCLS; ++k) thr_key[k] = thr[k] << 1; // (keys are query length * 2 + kind of read)
			uint32_t bnd[2 * NCLS];
			for (int k = 0; k < 2 * NCLS; ++k) bnd[k] = (k % NCLS == NCLS - 1) ? n_t : 0u; // (unordered: everything in the last class)
			if (n_t && !cfg.ext_no_sort) { // order the tasks of each side by query length: the lanes of a warp then run rows of similar width
				int bits = 2; while ((1 << (bits - 1)) <= max_len && bits < 31) ++bits;
				rt_iota_u32(b_xtidx.as<u32>(), n_t);
				rt_sort_pairs_u32(P.ext_key_l, b_xtkey2.as<u32>(), b_xtidx.as<u32>(), b_xtidx2.as<u32>(), n_t, bits);
				rt_sort_pairs_u32(P.ext_key_r, b_xtkey2.as<u32>() + Tn, b_xtidx.as<u32>(), b_xtidx2.as<u32>() + Tn, n_t, bits);
				if (!g_last_cuda_ok()) { last_error = "device error or allocation failure (scan / sort scratch)"; return nullptr; } // (an unsorted order array would send the side kernels to arbitrary tasks)
				P.ext_order_l = b_xtidx2.as<u32>(); P.ext_order_r = b_xtidx2.as<u32>() + Tn;
				// class boundaries in the two sorted key arrays (first position whose key is >= thr)
				rt_lower_bounds_u32(b_xtkey2.as<u32>(), n_t, thr_key, NCLS - 1, b_xtidx.as<u32>());
				rt_lower_bounds_u32(b_xtkey2.as<u32>() + Tn, n_t, thr_key, NCLS - 1, b_xtidx.as<u32>() + NCLS);
				rt_d2h(bnd, b_xtidx.p, sizeof(bnd));
				bnd[NCLS - 1] = n_t; bnd[2 * NCLS - 1] = n_t;
				hp.mark("ext_class_sync");
			}
			if (cfg.ext_task_machine) { // comparison route: both sides of a task by the per-lane machine of the read-level kernel (extend_thread.cuh, TASKS = true)
				P.ext_sides = 0;
				const size_t task_smem = thread_ext_smem_bytes(THREADS_PER_BLOCK, rw, wm_cap, cell);
				if (narrow) RT_THREADS_QUEUE_SMEM(body_extend_task_q8, P, (int64_t)n_t, task_smem);
				else RT_THREADS_QUEUE_SMEM(body_extend_task_q, P, (int64_t)n_t, task_smem);
			} else
			for (int side = 0; side < 2; ++side) {
				for (int k = NCLS - 1; k >= 0; --k) { // longest class first
					const uint32_t lo = k ? bnd[side * NCLS + k - 1] : 0u, hi = bnd[side * NCLS + k];
					if (hi <= lo) continue;
					int crw = thr[k] < (uint32_t)rw ? (int)thr[k] : rw; // sides of < thr bases need thr cells (H[0..qlen])
					P.ext_lo = lo; P.ext_hi = hi; P.ext_rw = crw;
					const size_t sm = side_smem_bytes(THREADS_PER_BLOCK, crw, cell);
					if (narrow) { if (side == 0) RT_THREADS_QUEUE_SMEM(body_ext_left_q8, P, (int64_t)(hi - lo), sm); else RT_THREADS_QUEUE_SMEM(body_ext_right_q8, P, (int64_t)(hi - lo), sm); }
					else { if (side == 0) RT_THREADS_QUEUE_SMEM(body_ext_left_q, P, (int64_t)(hi - lo), sm); else RT_THREADS_QUEUE_SMEM(body_ext_right_q, P, (int64_t)(hi - lo), sm); }
				}
			}
			RT_THREADS_TID(body_ext_replay, P, (int64_t)n_reads);
			P.use_ext_res = 1;
			RT_WARPS_DEVN(wbody_extend_heavy, P, P.n_heavy, (int64_t)n_reads); // heavy reads, and light ones the replay could not finish
			P.use_ext_res = 0;
		} else {
			P.t_ext_rw = rw;
			const size_t ext_smem = thread_ext_smem_bytes(THREADS_PER_BLOCK, rw, wm_cap, cell), task_smem = ext_smem;
			if (narrow) RT_THREADS_QUEUE_SMEM(body_extend_q8, P, (int64_t)n_reads, ext_smem);
			else RT_THREADS_QUEUE_SMEM(body_extend_q, P, (int64_t)n_reads, ext_smem);
			if (cfg.no_ext_tasks) RT_WARPS_DEVN(wbody_extend_heavy, P, P.n_heavy, (int64_t)n_reads); // each heavy read extended by its own warp
			else {
				// heavy reads: every kept seed becomes an extension task for the thread-level queue, then a warp per read
				// replays the containment decisions over the finished extensions and de-duplicates the regions
				size_t Tn = (size_t)T + 1;
				if (!(b_exttask.ensure(Tn * sizeof(ExtTask)) && b_extres.ensure(Tn * sizeof(ExtRes)) && b_ckrmax.ensure(Tn * 16))) { last_error = "device allocation failed (extension tasks)"; return nullptr; }
				P.ext_tasks = b_exttask.as<ExtTask>(); P.ext_res = b_extres.as<ExtRes>(); P.ck_rmax = b_ckrmax.as<i64>(); P.n_ext_tasks = b_misc.as<u32>() + 21;
				RT_WARPS_DEVN(wbody_heavy_prep, P, P.n_heavy, (int64_t)n_reads);
				if (narrow) RT_THREADS_QUEUE_SMEM(body_extend_task_q8, P, (int64_t)T, task_smem);
				else RT_THREADS_QUEUE_SMEM(body_extend_task_q, P, (int64_t)T, task_smem);
				P.use_ext_res = 1;
				RT_WARPS_DEVN(wbody_extend_heavy, P, P.n_heavy, (int64_t)n_reads);
				P.use_ext_res = 0;
			}
		}
	}
	// ---- paired-end: insert-size statistics (the one batch-global step; host)
	for (int d = 0; d < 4; ++d) { memset(&P.pes[d], 0, sizeof(PeStat)); P.pes[d].failed = 1; }
	if (pe) {
		if (pes0) memcpy(P.pes, pes0, 4 * sizeof(PeStat));
		else {
			if (!b_cand.ensure((size_t)(n_units + 1) * 8)) { last_error = "device allocation failed"; return nullptr; }
			P.pe_cand = b_cand.as<i64>();
			RT_THREADS(body_pe_cand, P, (int64_t)n_units);
			// sorted on the device (keys dir<<48 | isize; -1 = no candidate sorts last as unsigned), summed on the host
			if (!b_cand2.ensure((size_t)(n_units + 1) * 8)) { last_error = "device allocation failed"; return nullptr; }
			rt_sort_u64((const uint64_t*)P.pe_cand, b_cand2.as<uint64_t>(), (uint64_t)n_units);
			if (!g_last_cuda_ok()) { last_error = "device error or allocation failure (scan / sort scratch)"; return nullptr; }
			std::vector<int64_t> cand((size_t)n_units);
			rt_d2h(cand.data(), b_cand2.p, (size_t)n_units * 8);
			hp.mark("extend_sync");
			host_mem_pestat(opt, cand.data(), n_units, P.pes, true);
			hp.mark("pestat");
		}
		std::vector<double> tab; int64_t off[4];
		if (!host_pair_table(opt, P.pes, tab, off)) { last_error = "insert-size range too wide to tabulate"; return nullptr; }
		if (!b_pairtab.ensure(tab.size() * 8 + 8)) { last_error = "device allocation failed"; return nullptr; }
		if (!tab.empty()) rt_h2d(b_pairtab.p, tab.data(), tab.size() * 8);
		P.pair_tab = b_pairtab.as<double>();
		int max_win = 0;
		for (int d = 0; d < 4; ++d) { P.pair_tab_off[d] = off[d]; if (!P.pes[d].failed) { int wd = P.pes[d].high - P.pes[d].low + 1; if (wd > max_win) max_win = wd; } }
		P.ws_brow_cap = std::max(16, (max_win + max_len + 64) * scratch_scale / sdiv);
		if (!b_wsbrow.ensure((size_t)wslots * P.ws_brow_cap * 8)) { last_error = "device allocation failed"; return nullptr; }
		P.ws_brow = b_wsbrow.as<u64>();
	}
	// ---- region capacity, offsets, compaction
	RT_THREADS(body_reg_cap, P, (int64_t)n_reads);
	uint64_t R = rt_scan_u32(P.reg_cap, P.reg_off, n_reads);
	if (!g_last_cuda_ok()) { last_error = "device error or allocation failure (scan / sort scratch)"; return nullptr; }
	hp.mark("regcap_sync");
	{
		size_t Rn = (size_t)R + 2;
		ok = b_regs.ensure(Rn * sizeof(Reg)) && b_snap.ensure(Rn * sizeof(Reg)) && b_ztmp.ensure(Rn * 4) && b_pairv.ensure(Rn * sizeof(Pair64)) && b_recs.ensure(Rn * sizeof(AlnRec));
		P.var_cap = (u32)std::min<uint64_t>((uint64_t)0xfffffff0u, ((uint64_t)R * 24 + (uint64_t)n_reads * 8 + 4096 + (long_mode ? 2 * blob_len : 0)) * (uint64_t)scratch_scale / (uint64_t)sdiv); // CIGAR + MD + XA words of the records
		ok = ok && b_var.ensure((size_t)P.var_cap * 4);
		if (!ok) { last_error = "device allocation failed (region pools)"; return nullptr; }
	}
	P.regs = b_regs.as<Reg>(); P.snap = b_snap.as<Reg>(); P.ztmp = b_ztmp.as<u32>(); P.pairv = b_pairv.as<Pair64>(); P.recs = b_recs.as<AlnRec>(); P.var = b_var.as<u32>();
	RT_THREADS(body_compact, P, (int64_t)n_reads);
	// ---- mate rescue as a task pipeline (rescue.cuh): snapshot, ONE speculative gen -> sw -> replay over all rescuing
	// hits; reads whose replay needs an alignment that was not speculated take the round-by-round kernels (rare)
	if (pe && !(opt.flag & MEM_F_NO_RESCUE)) {
		size_t nr = (size_t)n_reads + 2, Rn = (size_t)R + 2;
		ok = b_nb.ensure(nr * 4) && b_swtask.ensure((Rn > nr * 4 ? Rn : nr * 4) * sizeof(SwTask)) && b_swstate.ensure(nr * 4) && b_swlist.ensure(nr * 4 * 4)
		     && b_multi.ensure(nr * 4) && b_specstate.ensure(Rn * 4) && b_spectask.ensure(Rn * 4 * 4) && b_nreg0.ensure(nr * 4) && b_rsheavy.ensure(nr * 4);
		if (!ok) { last_error = "device allocation failed (rescue pools)"; return nullptr; }
		P.nb = b_nb.as<u32>(); P.sw_tasks = b_swtask.as<SwTask>(); P.sw_state = b_swstate.as<u8>(); P.sw_list = b_swlist.as<u32>();
		P.sw_n_list = b_misc.as<u32>() + 3; P.max_nb = b_misc.as<u32>() + 4; P.n_multi = b_misc.as<u32>() + 7;
		P.multi = b_multi.as<u32>(); P.spec_state = b_specstate.as<u8>(); P.spec_task = b_spectask.as<u32>(); P.n_reg0 = b_nreg0.as<u32>();
		P.force_fb = cfg.rescue_fallback ? 1 : 0; // tests: every rescuing read takes the round-by-round path
		P.rs_heavy = b_rsheavy.as<u32>(); P.n_rs_heavy = b_misc.as<u32>() + 20;
		RT_THREADS(body_rescue_snap, P, (int64_t)n_reads);
		RT_THREADS(body_rescue_classify, P, (int64_t)n_reads);
		// the task pool holds at most sum(nb * live orientations) <= R entries (that head-room is part of reg_cap)
		RT_THREADS(body_rescue_gen_all, P, (int64_t)n_reads);
		if (cfg.rescue_warp_sw || long_mode) RT_WARPS_DEVN(wbody_rescue_sw, P, P.sw_n_list, (int64_t)R); // the warp-per-task formulation (comparison, long reads)
		else {
			// thread-per-task alignments with two cells per register (sw_thread.cuh): forward pass of every task, reverse pass of
			// those that scored, the warp-per-task kernel for the tasks the byte formulation does not take
			int shift = 127, mmax = 0;
			for (int k = 0; k < 25; ++k) { if (opt.mat[k] < shift) shift = opt.mat[k]; if (opt.mat[k] > mmax) mmax = opt.mat[k]; }
			shift = (256 - (shift & 0xff)) & 0xff; // uint8 arithmetic of the profile builder (ksw_qinit)
			P.sw_shift = shift; P.sw_mmax = mmax;
			for (int t = 0; t < 4; ++t) { P.swtab[t] = 0; for (int c = 0; c < 6; ++c) P.swtab[t] |= (u64)(u8)((c < 5 ? (int)opt.mat[t * 5 + c] : 0) + shift) << (8 * c); }
			P.tsw_capw = ((max_len + 15) / 16) * 8;
			if (P.tsw_capw > 160) P.tsw_capw = 160; // (queries of up to 320 bases; longer ones take the warp kernel)
			P.tsw_brow_cap = P.ws_brow_cap;
			ok = b_tswhmax.ensure((size_t)((tslots + 31) / 32 * 32) * P.tsw_capw * 4 + 4096) && b_tswbrow.ensure((size_t)tslots * P.tsw_brow_cap * 8 + 64) && b_swrev.ensure(Rn * 4) && b_swfb.ensure(Rn * 4);
			if (!ok) { last_error = "device allocation failed (rescue alignment scratch)"; return nullptr; }
			P.tsw_hmax = b_tswhmax.as<u32>(); P.tsw_brow = b_tswbrow.as<u64>(); P.sw_rev = b_swrev.as<u32>(); P.sw_fb = b_swfb.as<u32>();
			P.n_sw_rev = b_misc.as<u32>() + 30; P.n_sw_fb = b_misc.as<u32>() + 31;
			const size_t sm = tsw_block_smem_bytes(THREADS_PER_BLOCK, P.tsw_capw);
			RT_THREADS_QUEUE_SMEM(body_tsw_fwd_q, P, (int64_t)R, sm);
			RT_THREADS_QUEUE_SMEM(body_tsw_rev_q, P, (int64_t)R, sm);
			RT_WARPS_DEVN(wbody_rescue_sw_fb, P, P.n_sw_fb, (int64_t)R);
		}
		RT_WARPS_DEVN(wbody_rescue_apply_heavy, P, P.n_rs_heavy, (int64_t)n_reads); // long chains first: they are the tail
		RT_THREADS(body_rescue_apply_all, P, (int64_t)n_reads);
		uint32_t hdr[8];
		rt_d2h(hdr, b_misc.as<u32>(), sizeof(hdr));
		const uint32_t max_nb = hdr[4], n_fb = hdr[7];
		hp.mark("rescue_sync");
		if (n_fb) {
			RT_THREADS(body_rescue_fb_restore, P, (int64_t)n_fb);
			int rounds = (int)max_nb < opt.max_matesw ? (int)max_nb : opt.max_matesw;
			for (int k = 0; k < rounds; ++k) {
				P.round = k;
				rt_memset(P.sw_n_list, 0, 4);
				RT_THREADS(body_rescue_gen_multi, P, (int64_t)n_fb);
				RT_WARPS_DEVN(wbody_rescue_sw, P, P.sw_n_list, (int64_t)n_reads * 4);
				RT_THREADS(body_rescue_apply_multi, P, (int64_t)n_fb);
			}
		}
		P.rescued = 1;
	}
	// ---- finalize: a thread per read / pair; items that exceed the thread-level scratch are redone by a warp
	{
		RT_THREADS(body_save_regs, P, (int64_t)n_reads);
		const int64_t n_items = pe ? (int64_t)n_units : (int64_t)n_reads;
		if (long_mode) { // every item straight to the warp-level kernel
			rt_iota_u32(P.redo_list, (uint64_t)n_items);
			const uint32_t ni = (uint32_t)n_items; rt_h2d(P.n_redo, &ni, 4); rt_sync();
		}
		if (long_mode) {}
		else if (cfg.final_one_pass) { // every item computes its own alignments (kept for comparison and tests)
			if (pe) RT_THREADS_TID_RING(body_final_pe_t, P, n_items); else RT_THREADS_TID_RING(body_final_se_t, P, n_items);
		} else {
			// two passes around a queue of alignment tasks (galign_thread.cuh): collect -> tasks -> consume
			P.ga_cap = (u32)std::min<uint64_t>((uint64_t)0x7fffffff, (uint64_t)n_reads * 4 + 1024); // items that find the pool full are redone by a warp
			if (cfg.tiny_ga_cap) P.ga_cap = 40;
			ok = b_gatask.ensure((size_t)P.ga_cap * sizeof(GaTask)) && b_gares.ensure((size_t)P.ga_cap * sizeof(GaRes)) && b_gaitem0.ensure(((size_t)n_items + 1) * 4) && b_gaitemn.ensure(((size_t)n_items + 1) * 4);
			if (!ok) { last_error = "device allocation failed (alignment tasks)"; return nullptr; }
			P.ga_tasks = b_gatask.as<GaTask>(); P.ga_res = b_gares.as<GaRes>(); P.ga_item0 = b_gaitem0.as<u32>(); P.ga_itemn = b_gaitemn.as<u32>();
			P.n_ga_tasks = b_misc.as<u32>() + 23;
			P.final_heavy_hits = cfg.final_heavy_hits >= 0 ? cfg.final_heavy_hits : 24;
			ok = b_gakey.ensure((size_t)P.ga_cap * 4) && b_gakey2.ensure((size_t)P.ga_cap * 4) && b_gaiota.ensure((size_t)P.ga_cap * 4) && b_gaorder.ensure((size_t)P.ga_cap * 4) && b_gafb.ensure((size_t)P.ga_cap * 4);
			if (!ok) { last_error = "device allocation failed (alignment tasks)"; return nullptr; }
			P.ga_key = cfg.ga_no_sort ? nullptr : b_gakey.as<u32>(); P.ga_order = nullptr; P.ga_fb = b_gafb.as<u32>(); P.n_ga_fb = b_misc.as<u32>() + 32;
			// (neither pass computes an alignment itself: no shared-memory rings, which lets a third block onto each SM)
			const bool fin_ring = cfg.final_ring;
			if (fin_ring) { if (pe) RT_THREADS_TID_RING(body_final_pe_collect, P, n_items); else RT_THREADS_TID_RING(body_final_se_collect, P, n_items); }
			else if (pe) RT_THREADS_TID(body_final_pe_collect, P, n_items); else RT_THREADS_TID(body_final_se_collect, P, n_items);
			rt_memset(P.var_cursor, 0, 4); // payloads written while collecting are discarded
			if (P.ga_key) { // order the tasks by cost (ungapped sums first, then by band area): the lanes of a warp then work on similar tasks
				uint32_t n_ga = 0;
				rt_d2h(&n_ga, P.n_ga_tasks, 4);
				hp.mark("ga_sync");
				if (n_ga > P.ga_cap) n_ga = P.ga_cap;
				if (n_ga) {
					rt_iota_u32(b_gaiota.as<u32>(), n_ga);
					rt_sort_pairs_u32(P.ga_key, b_gakey2.as<u32>(), b_gaiota.as<u32>(), b_gaorder.as<u32>(), n_ga, 17);
					if (!g_last_cuda_ok()) { last_error = "device error or allocation failure (scan / sort scratch)"; return nullptr; }
					P.ga_order = b_gaorder.as<u32>();
				}
			}
			RT_THREADS_QUEUE_SMEM(body_galign_task_q, P, (int64_t)P.ga_cap, galign_block_smem_bytes(THREADS_PER_BLOCK, P.wm_cap));
			RT_WARPS_DEVN(wbody_galign_fb, P, P.n_ga_fb, (int64_t)P.ga_cap); // bands / tracebacks beyond the thread-level kernel
			if (fin_ring) { if (pe) RT_THREADS_TID_RING(body_final_pe_consume, P, n_items); else RT_THREADS_TID_RING(body_final_se_consume, P, n_items); }
			else if (pe) RT_THREADS_TID(body_final_pe_consume, P, n_items); else RT_THREADS_TID(body_final_se_consume, P, n_items);
		}
		if (pe) RT_WARPS_DEVN(wbody_final_pe_redo, P, P.n_redo, n_items); else RT_WARPS_DEVN(wbody_final_se_redo, P, P.n_redo, n_items);
	}
	// ---- output stream
	RT_THREADS(body_out_size, P, (int64_t)n_reads);
	uint64_t W = rt_scan_u32(P.out_words, P.out_off, n_reads);
	if (!g_last_cuda_ok()) { last_error = "device error or allocation failure (scan / sort scratch)"; return nullptr; }
	hp.mark("final_sync");
	if (!b_out.ensure((size_t)W * 4 + 16)) { last_error = "device allocation failed (output)"; return nullptr; }
	P.out = b_out.as<u32>();
	RT_THREADS(body_out_write, P, (int64_t)n_reads);
	uint32_t misc[40];
	rt_d2h(misc, b_misc.p, sizeof(misc));
	memcpy(&last_counters, misc + 8, sizeof(Counters));
	n_launches = rt_launches();
	if (!g_last_cuda_ok()) { last_error = "a CUDA call of this batch failed (see stderr for the stage)"; return nullptr; } // (refused launches included: rt_mark)
	if (cfg.debug_stats) fprintf(stderr, "[bwamem-b200] workspace of this context: %.2f GB\n", (double)workspace_bytes() / 1e9);
	if (cfg.debug_stats) fprintf(stderr, "[bwamem-b200] reads %u: finalize redo %u, alignment tasks %u, extension-heavy reads %u / tasks %u, chain-heavy %u, rescue warp-replayed %u / fallback %u, seed one-machine reads %u\n",
	                                        n_reads, misc[5], misc[23], misc[6], misc[21], misc[22], misc[20], misc[7], misc[28]);
	{
		const uint32_t scratch_flags = ERR_Z_OVERFLOW | ERR_CIGAR_OVERFLOW | ERR_MD_OVERFLOW | ERR_XA_OVERFLOW | ERR_BROW_OVERFLOW | ERR_VAR_OVERFLOW;
		if (misc[0] && !(misc[0] & ~scratch_flags) && scratch_scale < 64) { // an item outgrew the warp-level scratch: once more, with twice the room
			scratch_scale *= 2;
			if (cfg.debug_stats) fprintf(stderr, "[bwamem-b200] warp-level scratch overflowed (flags 0x%x): re-running the batch with %dx the capacities\n", misc[0], scratch_scale);
			return run(opt_in, pes0, out_size, host_output);
		}
	}
	if (misc[0]) {
		char buf[128]; snprintf(buf, sizeof buf, "device-side capacity/limit error flags 0x%x", misc[0]);
		last_error = buf;
		return nullptr;
	}
	if (!host_output) { *out_size = (size_t)W * 4; rt_sync(); return (void*)b_out.p; }
	hp.mark("write_sync");
	void* res = malloc((size_t)W * 4 + 4);
	if (!res) { last_error = "host allocation failed"; return nullptr; }
	if (!fetch_result(res, P.out, (size_t)W * 4)) { free(res); last_error = "result copy failed"; return nullptr; }
	*out_size = (size_t)W * 4;
	return res;
}

} // namespace b200
