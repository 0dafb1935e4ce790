// engine.cu -- the product: CUDA runtime binding of the engine (sm_100a) and its C-ABI.
//
// Exports the C entry points of /root/reference/src/main/c/jnibwa.h:11-16 with identical signatures
// (declared, with the reference line each replaces, in include/bwamem_b200.h), mem_opt_init (called by
// …BwaMemIndex.c:91), the allocator counterpart of destroyByteBuffer (…BwaMemIndex.c:157-160) and a few
// b200_* instrumentation calls used by bench.py.  No CPU fallback: every call that needs the device fails
// (NULL / 0) when no CUDA device is usable.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "simt.cuh"

// ---- runtime vocabulary expected by engine_impl.inl ------------------------------------------------
static thread_local cudaError_t g_last_cuda = cudaSuccess;
static inline bool cu_ok(cudaError_t e, const char* what)
{
	if (e != cudaSuccess) { g_last_cuda = e; fprintf(stderr, "[bwamem-b200] CUDA error in %s: %s\n", what, cudaGetErrorString(e)); return false; }
	return true;
}
static cudaStream_t g_stream = nullptr; // one stream per process (one process per GPU)
static int g_sm_count = 0;
// A second stream for a kernel that is independent of the next one on the main stream (a few long-running warps
// beside a throughput kernel): rt_fork_aux() .. launches .. rt_back_to_main() .. launches .. rt_join_aux().
static cudaStream_t g_aux_stream = nullptr, g_stream_override = nullptr;
static cudaEvent_t g_ev_fork = nullptr, g_ev_join = nullptr;
static cudaStream_t rt_stream()
{
	if (g_stream_override) return g_stream_override;
	if (!g_stream) { cu_ok(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking), "cudaStreamCreate"); }
	return g_stream;
}
static void rt_fork_aux()
{
	cudaStream_t main_s = rt_stream();
	if (!g_aux_stream) {
		int prio_low = 0, prio_high = 0; // highest priority: the few blocks of the auxiliary kernel are placed before the grid that fills the rest of the GPU
		cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high);
		cu_ok(cudaStreamCreateWithPriority(&g_aux_stream, cudaStreamNonBlocking, prio_high), "cudaStreamCreate");
		cudaEventCreateWithFlags(&g_ev_fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&g_ev_join, cudaEventDisableTiming);
	}
	cudaEventRecord(g_ev_fork, main_s); cudaStreamWaitEvent(g_aux_stream, g_ev_fork, 0);
	g_stream_override = g_aux_stream;
}
static void rt_back_to_main() { cudaEventRecord(g_ev_join, g_aux_stream); g_stream_override = nullptr; }
static int rt_sm_count()
{
	if (!g_sm_count) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev); if (g_sm_count <= 0) g_sm_count = 148; }
	return g_sm_count;
}
static void* rt_malloc(size_t n) { void* p = nullptr; if (!cu_ok(cudaMalloc(&p, n ? n : 1), "cudaMalloc")) return nullptr; return p; }
static void rt_free(void* p) { cudaFree(p); }
static void rt_h2d(void* d, const void* s, size_t n) { if (n) cu_ok(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, rt_stream()), "h2d"); }
static void rt_d2h(void* d, const void* s, size_t n) { if (n) { cu_ok(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, rt_stream()), "d2h"); cu_ok(cudaStreamSynchronize(rt_stream()), "d2h sync"); } }
static void rt_memset(void* d, int v, size_t n) { cu_ok(cudaMemsetAsync(d, v, n, rt_stream()), "memset"); }
static void rt_sync() { cu_ok(cudaStreamSynchronize(rt_stream()), "sync"); }

#define THREADS_PER_BLOCK 128
#define WARPS_PER_BLOCK 4
#define SEED_BLOCKS_PER_SM 8
#define WARP_BLOCKS_PER_SM 8
static int64_t rt_thread_slots() { return (int64_t)rt_sm_count() * SEED_BLOCKS_PER_SM * THREADS_PER_BLOCK; }
static int64_t rt_warp_slots() { return (int64_t)rt_sm_count() * WARP_BLOCKS_PER_SM * WARPS_PER_BLOCK; }

#ifdef B200_DEBUG_CKPT
__device__ volatile unsigned* g_ckpt = nullptr;
static unsigned* g_ckpt_host = nullptr;
#endif
#include "kernels.cuh"

// exclusive scan of u32 counts into u64 offsets (out[n] = total); returns the total
static void* g_scan_tmp = nullptr; static size_t g_scan_tmp_bytes = 0;
struct U32ToU64 { __host__ __device__ u64 operator()(const u32& v) const { return (u64)v; } };
__global__ void k_scan_close(const u32* in, u64* out, u64 n) { out[n] = n ? out[n - 1] + in[n - 1] : 0; }
static uint64_t rt_scan_u32(const uint32_t* in, uint64_t* out, uint64_t n)
{
	if (n) {
		cub::TransformInputIterator<u64, U32ToU64, const u32*> it(in, U32ToU64());
		size_t need = 0;
		cub::DeviceScan::ExclusiveSum(nullptr, need, it, out, (int64_t)n, rt_stream());
		if (need > g_scan_tmp_bytes) { if (g_scan_tmp) cudaFree(g_scan_tmp); cudaMalloc(&g_scan_tmp, need + 1024); g_scan_tmp_bytes = need + 1024; }
		cub::DeviceScan::ExclusiveSum(g_scan_tmp, need, it, out, (int64_t)n, rt_stream());
	}
	k_scan_close<<<1, 1, 0, rt_stream()>>>(in, out, n);
	uint64_t total = 0;
	rt_d2h(&total, out + n, 8);
	return total;
}

static void rt_sort_u64(const uint64_t* in, uint64_t* out, uint64_t n)
{
	if (!n) return;
	size_t need = 0;
	cub::DeviceRadixSort::SortKeys(nullptr, need, in, out, (int64_t)n, 0, 64, rt_stream());
	if (need > g_scan_tmp_bytes) { if (g_scan_tmp) cudaFree(g_scan_tmp); cudaMalloc(&g_scan_tmp, need + 1024); g_scan_tmp_bytes = need + 1024; }
	cub::DeviceRadixSort::SortKeys(g_scan_tmp, need, in, out, (int64_t)n, 0, 64, rt_stream());
}

// ---- launchers ---------------------------------------------------------------------------------------
// Thread kernels: grid-stride over work items.  Warp kernels: one warp per item, items handed out by an
// atomic counter so that long items (repetitive reads) do not stall a statically assigned neighbour.
#define DEF_THREAD_KERNEL(body) \
	__global__ void __launch_bounds__(THREADS_PER_BLOCK) k_##body(const __grid_constant__ BatchParams P, i64 n) { \
		for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) body(P, i); }
#define DEF_THREAD_TID_KERNEL(body) \
	__global__ void __launch_bounds__(THREADS_PER_BLOCK) k_##body(const __grid_constant__ BatchParams P, i64 n) { \
		const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x; \
		for (i64 i = tid; i < n; i += (i64)gridDim.x * blockDim.x) body(P, i, tid); }
// persistent thread kernel: every thread takes work items from a batch-wide counter inside the body
#define DEF_THREAD_QUEUE_KERNEL(body) \
	__global__ void __launch_bounds__(THREADS_PER_BLOCK) k_##body(const __grid_constant__ BatchParams P, i64 n, u32* counter) { \
		body(P, n, counter, (i64)blockIdx.x * blockDim.x + threadIdx.x); }
#define DEF_THREAD_QUEUE_KERNEL_MB(body, minb) \
	__global__ void __launch_bounds__(THREADS_PER_BLOCK, minb) k_##body(const __grid_constant__ BatchParams P, i64 n, u32* counter) { \
		body(P, n, counter, (i64)blockIdx.x * blockDim.x + threadIdx.x); }
#define DEF_WARP_KERNEL(wbody) \
	__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) k_##wbody(const __grid_constant__ BatchParams P, i64 n, u32* counter) { \
		extern __shared__ __align__(16) u8 smem_[]; \
		const int wib = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31); \
		const i64 slot = (i64)blockIdx.x * WARPS_PER_BLOCK + wib; \
		WarpMem wm; WarpScratch ws; \
		make_warp_ctx(P, slot, smem_ + (size_t)wib * ((warp_smem_bytes(P.wm_cap) + 15) & ~(size_t)15), wm, ws); \
		for (;;) { \
			u32 item = 0; \
			if (lane == 0) item = atomicAdd(counter, 1u); \
			item = __shfl_sync(FULLMASK, item, 0); \
			if ((i64)item >= n) break; \
			wbody<32>(P, (i64)item, wm, ws); \
			__syncwarp(); \
		} }

// warp kernel whose item count lives in device memory (written by the preceding kernel of the same stream)
#define DEF_WARP_KERNEL_DEVN(wbody) DEF_WARP_KERNEL_DEVN_LB(wbody, __launch_bounds__(WARPS_PER_BLOCK * 32))
#define DEF_WARP_KERNEL_DEVN_MB(wbody, minb) DEF_WARP_KERNEL_DEVN_LB(wbody, __launch_bounds__(WARPS_PER_BLOCK * 32, minb))
#define DEF_WARP_KERNEL_DEVN_LB(wbody, bounds) \
	__global__ void bounds k_##wbody(const __grid_constant__ BatchParams P, const u32* n_ptr, u32* counter) { \
		extern __shared__ __align__(16) u8 smem_[]; \
		const int wib = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31); \
		const i64 slot = (i64)blockIdx.x * WARPS_PER_BLOCK + wib; \
		const i64 n = (i64)*n_ptr; \
		WarpMem wm; WarpScratch ws; \
		make_warp_ctx(P, slot, smem_ + (size_t)wib * ((warp_smem_bytes(P.wm_cap) + 15) & ~(size_t)15), wm, ws); \
		for (;;) { \
			u32 item = 0; \
			if (lane == 0) item = atomicAdd(counter, 1u); \
			item = __shfl_sync(FULLMASK, item, 0); \
			if ((i64)item >= n) break; \
			wbody<32>(P, (i64)item, wm, ws); \
			__syncwarp(); \
		} }

DEF_THREAD_KERNEL(body_encode)
#ifndef SEED_MIN_BLOCKS
#define SEED_MIN_BLOCKS 5
#endif
__global__ void __launch_bounds__(THREADS_PER_BLOCK, SEED_MIN_BLOCKS) k_body_seed_q(const __grid_constant__ BatchParams P, i64 n, u32* counter) { body_seed_q(P, n, counter, (i64)blockIdx.x * blockDim.x + threadIdx.x); }
DEF_THREAD_QUEUE_KERNEL(body_seedF1_q)
DEF_THREAD_QUEUE_KERNEL(body_seedF2_q)
DEF_THREAD_QUEUE_KERNEL(body_seedB_q)
DEF_THREAD_QUEUE_KERNEL(body_seedB2_q)
DEF_WARP_KERNEL_DEVN(wbody_seedB_heavy)
__global__ void __launch_bounds__(THREADS_PER_BLOCK, SEED_MIN_BLOCKS) k_body_seed12_q(const __grid_constant__ BatchParams P, i64 n, u32* counter) { body_seed12_q(P, n, counter, (i64)blockIdx.x * blockDim.x + threadIdx.x); }
__global__ void __launch_bounds__(THREADS_PER_BLOCK, SEED_MIN_BLOCKS) k_body_seed12_fb_q(const __grid_constant__ BatchParams P, i64 n, u32* counter) { body_seed12_fb_q(P, n, counter, (i64)blockIdx.x * blockDim.x + threadIdx.x); }
__global__ void __launch_bounds__(THREADS_PER_BLOCK, SEED_MIN_BLOCKS) k_body_seed3_q(const __grid_constant__ BatchParams P, i64 n, u32* counter) { body_seed3_q(P, n, counter, (i64)blockIdx.x * blockDim.x + threadIdx.x); }
DEF_THREAD_KERNEL(body_sal)
DEF_THREAD_KERNEL(body_chain)
DEF_THREAD_KERNEL(body_chain_classify)
DEF_WARP_KERNEL_DEVN(wbody_chain_heavy)
DEF_THREAD_KERNEL(body_pe_cand)
DEF_THREAD_KERNEL(body_reg_cap)
DEF_THREAD_KERNEL(body_compact)
DEF_THREAD_KERNEL(body_out_size)
DEF_THREAD_KERNEL(body_out_write)
DEF_THREAD_KERNEL(body_rescue_snap)
DEF_THREAD_KERNEL(body_rescue_gen_all)
DEF_THREAD_KERNEL(body_rescue_apply_all)
DEF_THREAD_KERNEL(body_rescue_fb_restore)
DEF_THREAD_KERNEL(body_rescue_classify)
DEF_WARP_KERNEL_DEVN(wbody_rescue_apply_heavy)
DEF_THREAD_KERNEL(body_rescue_gen_multi)
DEF_THREAD_KERNEL(body_rescue_apply_multi)
#ifndef RESCUE_SW_MIN_BLOCKS
#define RESCUE_SW_MIN_BLOCKS 4 // 128 registers: four blocks (16 warps) per SM; the kernel is bound by the latency of its row recurrence
#endif
DEF_WARP_KERNEL_DEVN_MB(wbody_rescue_sw, RESCUE_SW_MIN_BLOCKS)
DEF_WARP_KERNEL(wbody_extend)
DEF_THREAD_QUEUE_KERNEL_MB(body_extend_q, 2) // two blocks per SM fit the shared DP rows: let the state machine keep its state in registers
DEF_THREAD_QUEUE_KERNEL(body_extend_task_q)
// 8-bit cells: the shared rows of four blocks fit an SM; the read-level machine needs ~170 registers, so three of its blocks do
DEF_THREAD_QUEUE_KERNEL_MB(body_extend_q8, 3)
DEF_THREAD_QUEUE_KERNEL_MB(body_extend_task_q8, 4)
DEF_WARP_KERNEL_DEVN(wbody_heavy_prep)
DEF_WARP_KERNEL_DEVN(wbody_extend_heavy)
DEF_THREAD_KERNEL(body_save_regs)
DEF_THREAD_TID_KERNEL(body_final_pe_t)
DEF_THREAD_TID_KERNEL(body_final_se_t)
DEF_THREAD_TID_KERNEL(body_final_pe_collect)
DEF_THREAD_TID_KERNEL(body_final_se_collect)
DEF_THREAD_TID_KERNEL(body_final_pe_consume)
DEF_THREAD_TID_KERNEL(body_final_se_consume)
DEF_THREAD_QUEUE_KERNEL(body_galign_task_q)
DEF_WARP_KERNEL_DEVN(wbody_final_pe_redo)
DEF_WARP_KERNEL_DEVN(wbody_final_se_redo)

static u32* g_counters = nullptr; static int g_counter_next = 0;
static u32* next_counter()
{
	if (!g_counters) { cudaMalloc(&g_counters, 64 * sizeof(u32)); }
	u32* c = g_counters + (g_counter_next++ & 63);
	cudaMemsetAsync(c, 0, sizeof(u32), rt_stream());
	return c;
}

struct StageTimer { // CUDA events on the launching stream around every stage of the last batch
	static const int MAXS = 512;
	cudaEvent_t ev[MAXS + 1]; const char* name[MAXS]; int n = 0; bool init = false;
	void begin() {
#ifdef B200_DEBUG_CKPT
		if (!g_ckpt_host) { cudaHostAlloc(&g_ckpt_host, 1024, cudaHostAllocMapped); memset(g_ckpt_host, 0, 1024); unsigned* d = nullptr; cudaHostGetDevicePointer(&d, g_ckpt_host, 0); cudaMemcpyToSymbol(g_ckpt, &d, sizeof(d)); }
#endif
		if (!init) { for (auto& e : ev) cudaEventCreate(&e); init = true; } n = 0; cudaEventRecord(ev[0], rt_stream()); }
	void mark(const char* nm)
	{
		if (g_stream_override) return; // a kernel on the auxiliary stream: its time is inside the stages of the main stream it overlaps
		if (n < MAXS) { name[n] = nm; ++n; cudaEventRecord(ev[n], rt_stream()); }
		static const bool trace = getenv("B200_TRACE") != nullptr; // development aid: synchronise and report after every stage
		if (trace) {
#ifdef B200_DEBUG_CKPT
			for (int t = 0; t < 100 && cudaStreamQuery(rt_stream()) == cudaErrorNotReady; ++t) usleep(100000);
			if (cudaStreamQuery(rt_stream()) == cudaErrorNotReady) {
				fprintf(stderr, "[bwamem-b200] stage %s HUNG; checkpoints (warp: code val):", nm);
				for (int w = 0; w < 16; ++w) fprintf(stderr, " %d:%u/%u", w, g_ckpt_host[2 * w], g_ckpt_host[2 * w + 1]);
				fprintf(stderr, "\n"); fflush(stderr); _exit(3);
			}
#endif
			cudaError_t e = cudaStreamSynchronize(rt_stream()); fprintf(stderr, "[bwamem-b200] stage %s: %s\n", nm, cudaGetErrorString(e)); fflush(stderr);
		}
	}
};
static StageTimer g_timer;
// the main stream waits for the auxiliary one; the stage table shows the wait (the part of the auxiliary kernel that its
// neighbour did not cover) under the auxiliary kernel's name
static void rt_join_aux(const char* nm) { cudaStreamWaitEvent(rt_stream(), g_ev_join, 0); g_timer.mark(nm); }
static int g_launches = 0;

#define RT_THREADS(body, P, n) do { i64 n_ = (i64)(n); if (n_ > 0) { \
	i64 nb_ = (n_ + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK; i64 cap_ = (i64)rt_sm_count() * 16; if (nb_ > cap_) nb_ = cap_; \
	k_##body<<<(unsigned)nb_, THREADS_PER_BLOCK, 0, rt_stream()>>>(P, n_); ++g_launches; } g_timer.mark(#body); } while (0)
#define RT_THREADS_TID(body, P, n) do { i64 n_ = (i64)(n); if (n_ > 0) { \
	i64 nb_ = (n_ + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK; i64 cap_ = (i64)rt_sm_count() * SEED_BLOCKS_PER_SM; if (nb_ > cap_) nb_ = cap_; \
	k_##body<<<(unsigned)nb_, THREADS_PER_BLOCK, 0, rt_stream()>>>(P, n_); ++g_launches; } g_timer.mark(#body); } while (0)
// thread kernels whose threads use the shared-memory rings (make_thread_ctx)
#define RT_THREADS_TID_RING(body, P, n) do { i64 n_ = (i64)(n); if (n_ > 0) { \
	i64 nb_ = (n_ + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK; i64 cap_ = (i64)rt_sm_count() * SEED_BLOCKS_PER_SM; if (nb_ > cap_) nb_ = cap_; \
	size_t tsm_ = thread_block_smem_bytes(THREADS_PER_BLOCK, (P).wm_cap); \
	{ static size_t set_ = 0; if (tsm_ > set_) { cudaFuncSetAttribute(k_##body, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm_); set_ = tsm_; } } \
	k_##body<<<(unsigned)nb_, THREADS_PER_BLOCK, tsm_, rt_stream()>>>(P, n_); ++g_launches; } g_timer.mark(#body); } while (0)
// persistent queue kernels with an explicit dynamic shared-memory size (as many blocks per SM as that size allows)
#define RT_THREADS_QUEUE_SMEM(body, P, n, bytes) do { i64 n_ = (i64)(n); if (n_ > 0) { \
	size_t tsm_ = (bytes); \
	i64 nb_ = (n_ + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK; i64 per_sm_ = (i64)((227 * 1024) / (tsm_ + 1024)); if (per_sm_ < 1) per_sm_ = 1; if (per_sm_ > SEED_BLOCKS_PER_SM) per_sm_ = SEED_BLOCKS_PER_SM; \
	i64 cap_ = (i64)rt_sm_count() * per_sm_; if (nb_ > cap_) nb_ = cap_; \
	{ static size_t set_ = 0; if (tsm_ > set_) { cudaFuncSetAttribute(k_##body, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm_); set_ = tsm_; } } \
	k_##body<<<(unsigned)nb_, THREADS_PER_BLOCK, tsm_, rt_stream()>>>(P, n_, next_counter()); ++g_launches; } g_timer.mark(#body); } while (0)
#define RT_WARPS(wbody, P, n) do { i64 n_ = (i64)(n); if (n_ > 0) { \
	size_t sm_ = ((warp_smem_bytes((P).wm_cap) + 15) & ~(size_t)15) * WARPS_PER_BLOCK; \
	i64 nb_ = (n_ + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK; i64 cap_ = (i64)rt_sm_count() * WARP_BLOCKS_PER_SM; if (nb_ > cap_) nb_ = cap_; \
	{ static size_t set_ = 0; if (sm_ > set_) { cudaFuncSetAttribute(k_##wbody, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_); set_ = sm_; } } \
	k_##wbody<<<(unsigned)nb_, WARPS_PER_BLOCK * 32, sm_, rt_stream()>>>(P, n_, next_counter()); ++g_launches; } g_timer.mark(#wbody); } while (0)

#define RT_WARPS_DEVN(wbody, P, n_ptr, n_max) do { \
	size_t sm_ = ((warp_smem_bytes((P).wm_cap) + 15) & ~(size_t)15) * WARPS_PER_BLOCK; \
	i64 nb_ = (i64)rt_sm_count() * WARP_BLOCKS_PER_SM; \
	{ static size_t set_ = 0; if (sm_ > set_) { cudaFuncSetAttribute(k_##wbody, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_); set_ = sm_; } } \
	k_##wbody<<<(unsigned)nb_, WARPS_PER_BLOCK * 32, sm_, rt_stream()>>>(P, n_ptr, next_counter()); ++g_launches; g_timer.mark(#wbody); } while (0)

__global__ void k_relayout(const u32* src, u64 bwt_size, u32* dst, u64 n_blocks)
{
	for (u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (u64)gridDim.x * blockDim.x) fm_relayout_block(src, bwt_size, dst, b);
}
// denser suffix-array samples at load time: row k (a multiple of `dense`) walks to the next row the image samples
__global__ void k_densify_sa(DevIndex x, u64 n_rows, u32 dense, u64* out)
{
	const u64 n_out = (n_rows + dense - 1) / dense;
	for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < n_out; t += (u64)gridDim.x * blockDim.x) out[t] = fm_sa(x, t * dense);
}
#define RT_DENSIFY(x, n_rows, dense, out) do { k_densify_sa<<<(unsigned)(rt_sm_count() * 16), 256, 0, rt_stream()>>>((x), (n_rows), (dense), (out)); cu_ok(cudaGetLastError(), "densify"); } while (0)
#define RT_RELAYOUT(src, size, dst, nb) do { k_relayout<<<(unsigned)(rt_sm_count() * 16), 256, 0, rt_stream()>>>((src), (size), (dst), (nb)); cu_ok(cudaGetLastError(), "relayout"); } while (0)

static void rt_timer_begin() { g_timer.begin(); g_launches = 0; }

#include "engine_impl.inl"

using namespace b200;

// ---- C-ABI (include/bwamem_b200.h) ---------------------------------------------------------------------
extern "C" {

void* jnibwa_openIndex(int fd)
{
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
		fprintf(stderr, "[bwamem-b200] no CUDA device: this engine has no CPU fallback\n");
		if (fd >= 0) close(fd);
		return nullptr;
	}
	if (const char* dv = getenv("BWA_B200_DEVICE")) { // the Java API has no notion of a device
		if (cudaSetDevice(atoi(dv)) != cudaSuccess) { fprintf(stderr, "[bwamem-b200] BWA_B200_DEVICE=%s is not a usable device\n", dv); if (fd >= 0) close(fd); return nullptr; }
	}
	std::lock_guard<std::mutex> lock(Engine::device_mutex());
	Engine* e = new Engine();
	if (!e->open_from_fd(fd)) { fprintf(stderr, "[bwamem-b200] openIndex: %s\n", e->last_error.c_str()); e->close_all(); delete e; return nullptr; }
	return e;
}

int jnibwa_destroyIndex(void* h)
{
	if (!h) return 0;
	std::lock_guard<std::mutex> lock(Engine::device_mutex());
	Engine* e = (Engine*)h;
	e->close_all();
	delete e;
	return 0;
}

void* jnibwa_getRefContigNames(void* h, size_t* size) { return h ? contig_names_blob(((Engine*)h)->img, size) : nullptr; }

void* jnibwa_createAlignments(void* h, void* opt, void* pes, char* seqs, size_t* size)
{
	if (!h || !opt || !seqs || !size) return nullptr;
	Engine* e = (Engine*)h;
	g_last_cuda = cudaSuccess;
	void* r = e->align((const MemOpt*)opt, (const PeStat*)pes, seqs, size);
	e->n_launches = g_launches;
	if (r && (g_last_cuda != cudaSuccess || cudaGetLastError() != cudaSuccess)) { free(r); r = nullptr; e->last_error = "CUDA error during the batch"; }
	if (!r) fprintf(stderr, "[bwamem-b200] createAlignments: %s\n", e->last_error.c_str());
	return r;
}

void* mem_opt_init(void) { return default_options(); }
void jnibwa_free(void* p) { free(p); }
const char* jnibwa_getVersion(void) { return "bwamem-b200 0.1 (sm_100a; bwa-mem semantics of lh3/bwa cb95061)"; }

// ---- instrumentation for bench.py ------------------------------------------------------------------
// HBM-resident measurement: stage a request (host parse + upload, untimed by the caller), then run the device
// pipeline on it leaving the record stream in device memory.  Returns the stream's byte size, or -1.
int b200_stageBatch(void* h, const char* seqs)
{
	Engine* e = (Engine*)h;
	std::lock_guard<std::mutex> lock(Engine::device_mutex());
	return e->stage(seqs) ? 0 : 1;
}
long long b200_alignStaged(void* h, void* opt, void* pes)
{
	Engine* e = (Engine*)h;
	std::lock_guard<std::mutex> lock(Engine::device_mutex());
	size_t size = 0;
	g_last_cuda = cudaSuccess;
	void* r = e->run((const MemOpt*)opt, (const PeStat*)pes, &size, false);
	e->n_launches = g_launches;
	if (!r || g_last_cuda != cudaSuccess) { fprintf(stderr, "[bwamem-b200] alignStaged: %s\n", e->last_error.c_str()); return -1; }
	return (long long)size;
}
// device time of the last batch: first to last stage event on the engine's stream (ms)
float b200_lastBatchMs(void* h)
{
	(void)h; float t = 0;
	if (g_timer.n > 0) cudaEventElapsedTime(&t, g_timer.ev[0], g_timer.ev[g_timer.n]);
	return t;
}
int b200_setDevice(int dev) { return cudaSetDevice(dev) == cudaSuccess ? 0 : 1; }
// test hook (tests/test_smem_layout.py): where thread `tid` of a block of `threads` keeps its DP row, read codes and
// patching rings in the dynamic shared memory of the thread-level extension kernels.
// out = { row offset, row element stride (bytes), query offset, query stride, ringH offset, ringE offset, ring stride, ring slots, block bytes }
void b200_extLayout(int threads, int tid, int rw, int wm_cap, int cell, int patches, uint64_t* out)
{
	const ExtLayout L = thread_ext_layout(threads, tid, rw, wm_cap, cell, patches != 0);
	out[0] = L.row_off; out[1] = (uint64_t)threads * cell; out[2] = L.q_off; out[3] = (uint64_t)threads;
	out[4] = L.ringH_off; out[5] = L.ringE_off; out[6] = (uint64_t)threads * 4; out[7] = (uint64_t)L.ring_n;
	out[8] = thread_ext_smem_bytes(threads, rw, wm_cap, cell);
}
const char* b200_lastError(void* h) { return h ? ((Engine*)h)->last_error.c_str() : ""; }
// copies up to cap stage times (ms) of the last batch; names[i] points at static strings; returns the stage count
int b200_lastStageTimes(void* h, float* ms, const char** names, int cap)
{
	(void)h;
	int n = 0; // aggregated by stage name, in order of first appearance
	for (int i = 0; i < g_timer.n; ++i) {
		float t = 0; cudaEventElapsedTime(&t, g_timer.ev[i], g_timer.ev[i + 1]);
		int j = 0;
		for (; j < n; ++j) if (strcmp(names[j], g_timer.name[i]) == 0) break;
		if (j == n) { if (n >= cap) continue; names[n] = g_timer.name[i]; ms[n] = 0; ++n; }
		ms[j] += t;
	}
	return n;
}
// counters: cells_ext, cells_glb, cells_msw, n_ext, n_glb, n_msw, n_seed_lookups, n_launches
void b200_lastCounters(void* h, uint64_t* out)
{
	Engine* e = (Engine*)h;
	memcpy(out, &e->last_counters, sizeof(Counters));
	out[6] = e->last_n_seed;
	out[7] = (uint64_t)e->n_launches;
}

} // extern "C"
