// sim_engine.cpp -- KERNEL-DEBUG HARNESS (test tooling, never shipped, never loaded by the package).
//
// Compiles the engine's kernel source (gatk-bwamem-jni_b200/csrc/*.cuh) and its host orchestration
// (engine_impl.inl) with -DB200_SIM for the host, so that the exact kernel logic can be executed under
// gdb / ASan / UBSan on the CPU-only development box and compared with the oracle before GPU time is
// spent.  A warp variable becomes an array of 32 lanes and collectives become loops (see simt.cuh).
// The product library libbwa_b200.so is built by nvcc only and contains none of this.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#define B200_SIM 1
#include "simt.cuh"

static void* rt_malloc(size_t n) { void* p = malloc(n ? n : 1); if (p) memset(p, 0xA5, n ? n : 1); return p; } // poisoned like cudaMalloc (not zeroed)
static void rt_free(void* p) { free(p); }
static void rt_h2d(void* d, const void* s, size_t n) { memcpy(d, s, n); }
static void rt_d2h(void* d, const void* s, size_t n) { memcpy(d, s, n); }
static void rt_memset(void* d, int v, size_t n) { memset(d, v, n); }
static void rt_sync() {}
struct RtCtx { int dev = 0; };
static RtCtx* rt_ctx_create(int dev) { RtCtx* c = new RtCtx(); c->dev = dev; return c; }
static void rt_ctx_destroy(RtCtx* c) { delete c; }
static bool rt_bind(RtCtx*) { return true; }
static int rt_current_device() { return 0; }
static int rt_launches() { return 0; }
static bool g_last_cuda_ok() { return true; }
static bool rt_peer_copy(void* dst, int, const void* src, int, size_t n) { memcpy(dst, src, n); return true; }
static void* rt_host_alloc(size_t n) { return malloc(n ? n : 1); }
static void rt_host_free(void* p) { free(p); }
static bool cu_ok(int, const char*) { return true; }
static void rt_sort_pairs_u32(const uint32_t* kin, uint32_t* kout, const uint32_t* vin, uint32_t* vout, uint64_t n, int)
{
	std::vector<uint32_t> ord(n); for (uint64_t i = 0; i < n; ++i) ord[i] = (uint32_t)i;
	std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return kin[a] < kin[b]; });
	for (uint64_t i = 0; i < n; ++i) { kout[i] = kin[ord[i]]; vout[i] = vin[ord[i]]; }
}
static void rt_iota_u32(uint32_t* p, uint64_t n) { for (uint64_t i = 0; i < n; ++i) p[i] = (uint32_t)i; }
static void rt_lower_bounds_u32(const uint32_t* keys, uint64_t n, const uint32_t* thr, int n_thr, uint32_t* out)
{
	for (int k = 0; k < n_thr; ++k) out[k] = (uint32_t)(std::lower_bound(keys, keys + n, thr[k]) - keys);
}
static void rt_set_slot_caps(int64_t, int64_t) {}
static void rt_fork_aux() {}
static void rt_back_to_main() {}
static void rt_join_aux(const char*) {}
static int64_t rt_thread_slots() { return 1; }
static int64_t rt_warp_slots() { return 1; }
static void rt_sort_u64(const uint64_t* in, uint64_t* out, uint64_t n) { memcpy(out, in, n * 8); std::sort(out, out + n); }
static uint64_t rt_scan_u32(const uint32_t* in, uint64_t* out, uint64_t n) { uint64_t s = 0; for (uint64_t i = 0; i < n; ++i) { out[i] = s; s += in[i]; } out[n] = s; return s; }

#include "kernels.cuh"

#define RT_THREADS(body, P, n) do { for (int64_t i_ = 0; i_ < (int64_t)(n); ++i_) body(P, i_); } while (0)
#define RT_THREADS_TID(body, P, n) do { for (int64_t i_ = 0; i_ < (int64_t)(n); ++i_) body(P, i_, 0); } while (0)
#define RT_THREADS_TID_RING(body, P, n) RT_THREADS_TID(body, P, n)
#define RT_THREADS_QUEUE_SMEM(body, P, n, bytes) do { u32 c_ = 0; body(P, (int64_t)(n), &c_, 0); } while (0)
#define THREADS_PER_BLOCK 128
#define WARPS_PER_BLOCK 4
#define RT_WARPS(wbody, P, n) do { \
	std::vector<uint8_t> sm_(warp_smem_bytes((P).wm_cap) + 64); \
	WarpMem wm_; WarpScratch ws_; make_warp_ctx(P, 0, sm_.data(), wm_, ws_); \
	for (int64_t i_ = 0; i_ < (int64_t)(n); ++i_) wbody<32>(P, i_, wm_, ws_); } while (0)

static void rt_timer_begin() {}
#define RT_KMER_LEVEL(x, tab, L1, n) do { for (u64 e_ = 0; e_ < (u64)(n); ++e_) kmer_build_entry((x), (tab), (L1), e_); } while (0)
#define RT_DENSIFY(x, n_rows, dense, out) do { for (u64 t_ = 0; t_ < ((u64)(n_rows) + (dense) - 1) / (dense); ++t_) (out)[t_] = fm_sa((x), t_ * (dense)); } while (0)
#define RT_RELAYOUT(src, size, dst, nb, sb, sbs) do { for (u64 b_ = 0; b_ < (u64)(nb); ++b_) fm_relayout_block((src), (size), (dst), b_, (sb), (sbs)); } while (0)
#define RT_WARPS_DEVN(wbody, P, n_ptr, n_max) do { \
	std::vector<uint8_t> sm_(warp_smem_bytes((P).wm_cap) + 64); \
	WarpMem wm_; WarpScratch ws_; make_warp_ctx(P, 0, sm_.data(), wm_, ws_); \
	for (int64_t i_ = 0; i_ < (int64_t)*(n_ptr); ++i_) wbody<32>(P, i_, wm_, ws_); } while (0)
#include "engine_impl.inl"

using namespace b200;

extern "C" {
void* sim_openIndex(int fd) { Engine* e = new Engine(); if (!e->open_from_fd(fd)) { fprintf(stderr, "[sim] openIndex: %s\n", e->last_error.c_str()); e->close_all(); delete e; return nullptr; } return e; }
int sim_destroyIndex(void* h) { Engine* e = (Engine*)h; e->close_all(); delete e; return 0; }
void* sim_createAlignments(void* h, void* opt, void* pes, char* blob, size_t* size)
{
	Engine* e = (Engine*)h;
	void* r = e->align((const MemOpt*)opt, (const PeStat*)pes, blob, size);
	if (!r) fprintf(stderr, "[sim] createAlignments: %s\n", e->last_error.c_str());
	return r;
}
void sim_free(void* p) { free(p); }
void* sim_opt_init() { return default_options(); }
void* sim_getRefContigNames(void* h, size_t* size) { return contig_names_blob(((Engine*)h)->img, size); }
void sim_sdp_stats(long* out) { memcpy(out, g_sim_sdp, sizeof g_sim_sdp); }
void sim_last_counters(void* h, uint64_t* out) { memcpy(out, &((Engine*)h)->last_counters, sizeof(Counters)); }
}
