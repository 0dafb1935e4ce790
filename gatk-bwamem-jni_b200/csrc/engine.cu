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
#include <atomic>
#include <mutex>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "simt.cuh"

// ---- runtime vocabulary expected by engine_impl.inl ------------------------------------------------
// Every createAlignments call runs inside a CONTEXT (RtCtx): its own streams, events, stage timer, launch counter and
// scan scratch on one device.  A host thread binds a context for the duration of a call (rt_bind sets the CUDA device
// of the thread -- the current device is per host thread -- and the thread-local pointer the rt_* helpers read), so
// any number of Java threads may be inside the library at once, each on its own context, on any GPU of the pool.
#define RT_MAX_DEV 16
static thread_local cudaError_t g_last_cuda = cudaSuccess;
static inline bool cu_ok(cudaError_t e, const char* what)
{
	if (e != cudaSuccess) { g_last_cuda = e; fprintf(stderr, "[bwamem-b200] CUDA error in %s: %s\n", what, cudaGetErrorString(e)); return false; }
	return true;
}
struct RtCtx {
	int dev = 0, sm_count = 148;
	cudaStream_t stream = nullptr;   // the context's main stream
	// A second stream for a kernel that is independent of the next one on the main stream (a few long-running warps
	// beside a throughput kernel): rt_fork_aux() .. launches .. rt_back_to_main() .. launches .. rt_join_aux().
	cudaStream_t aux = nullptr, override_ = nullptr;
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
	void* scan_tmp = nullptr; size_t scan_tmp_bytes = 0;
	unsigned* counters = nullptr; int counter_next = 0;
	int launches = 0;
	// per-batch limit on resident blocks of the kernels that own per-thread / per-warp scratch (0: SEED_BLOCKS_PER_SM /
	// WARP_BLOCKS_PER_SM blocks per SM).  Batches of long reads need megabytes of scratch per slot: they run with fewer slots.
	int thread_block_cap = 0, warp_block_cap = 0;
	// CUDA events on the launching stream around every stage of the last batch
	static const int MAXS = 512;
	cudaEvent_t ev[MAXS + 1]; const char* name[MAXS]; int n_marks = 0; bool timer_init = false;
	bool trace = false;
};
static thread_local RtCtx* t_rt = nullptr;
static RtCtx* rt_ctx_create(int dev)
{
	if (!cu_ok(cudaSetDevice(dev), "cudaSetDevice")) return nullptr;
	RtCtx* c = new RtCtx();
	c->dev = dev;
	int prio_low = 0, prio_high = 0; // highest priority: the few blocks of the auxiliary kernel are placed before the grid that fills the rest of the GPU
	bool ok = cu_ok(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute");
	if (c->sm_count <= 0) c->sm_count = 148;
	ok = ok && cu_ok(cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high), "cudaDeviceGetStreamPriorityRange");
	ok = ok && cu_ok(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate");
	ok = ok && cu_ok(cudaStreamCreateWithPriority(&c->aux, cudaStreamNonBlocking, prio_high), "cudaStreamCreate");
	ok = ok && cu_ok(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming), "cudaEventCreate");
	ok = ok && cu_ok(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming), "cudaEventCreate");
	ok = ok && cu_ok(cudaMalloc(&c->counters, 64 * sizeof(unsigned)), "cudaMalloc");
	for (int i = 0; ok && i <= RtCtx::MAXS; ++i) ok = cu_ok(cudaEventCreate(&c->ev[i]), "cudaEventCreate");
	c->timer_init = ok;
	c->trace = getenv("B200_TRACE") != nullptr; // development aid: synchronise and report after every stage
	if (!ok) { delete c; return nullptr; } // (a device that cannot give a stream is unusable: nothing worth releasing piecemeal)
	return c;
}
static void rt_ctx_destroy(RtCtx* c)
{
	if (!c) return;
	cudaSetDevice(c->dev);
	if (c->stream) cudaStreamDestroy(c->stream);
	if (c->aux) cudaStreamDestroy(c->aux);
	if (c->ev_fork) cudaEventDestroy(c->ev_fork);
	if (c->ev_join) cudaEventDestroy(c->ev_join);
	if (c->scan_tmp) cudaFree(c->scan_tmp);
	if (c->counters) cudaFree(c->counters);
	if (c->timer_init) for (auto& e : c->ev) cudaEventDestroy(e);
	if (t_rt == c) t_rt = nullptr;
	delete c;
}
// bind the calling host thread to a context (and to its device) until the next rt_bind
static bool rt_bind(RtCtx* c) { t_rt = c; return c ? cu_ok(cudaSetDevice(c->dev), "cudaSetDevice") : true; }
static int rt_device_count() { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
static int rt_current_device() { int d = 0; if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); return 0; } return d; }
static cudaStream_t rt_stream() { return t_rt->override_ ? t_rt->override_ : t_rt->stream; }
static void rt_fork_aux()
{
	cu_ok(cudaEventRecord(t_rt->ev_fork, t_rt->stream), "fork"); cu_ok(cudaStreamWaitEvent(t_rt->aux, t_rt->ev_fork, 0), "fork");
	t_rt->override_ = t_rt->aux;
}
static void rt_back_to_main() { cu_ok(cudaEventRecord(t_rt->ev_join, t_rt->aux), "join"); t_rt->override_ = nullptr; }
static int rt_sm_count() { return t_rt->sm_count; }
static void* rt_malloc(size_t n) { void* p = nullptr; if (!cu_ok(cudaMalloc(&p, n ? n : 1), "cudaMalloc")) return nullptr; return p; }
static void rt_free(void* p) { cudaFree(p); }
static void rt_h2d(void* d, const void* s, size_t n) { if (n) cu_ok(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, rt_stream()), "h2d"); }
static void rt_d2h(void* d, const void* s, size_t n) { if (n) { cu_ok(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, rt_stream()), "d2h"); cu_ok(cudaStreamSynchronize(rt_stream()), "d2h sync"); } }
static void rt_memset(void* d, int v, size_t n) { cu_ok(cudaMemsetAsync(d, v, n, rt_stream()), "memset"); }
static void rt_sync() { cu_ok(cudaStreamSynchronize(rt_stream()), "sync"); }
// index replica k > 0 of a device pool: device-to-device over NVLink from the replica that took the host upload
static bool rt_peer_copy(void* dst, int dst_dev, const void* src, int src_dev, size_t n)
{
	if (!n) return true;
	return cu_ok(cudaMemcpyPeerAsync(dst, dst_dev, src, src_dev, n, rt_stream()), "peer copy");
}
// pinned host memory (staging of request / result bytes)
static void* rt_host_alloc(size_t n) { void* p = nullptr; if (cudaHostAlloc(&p, n ? n : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
static void rt_host_free(void* p) { if (p) cudaFreeHost(p); }
static bool g_last_cuda_ok() { return g_last_cuda == cudaSuccess; }

#define THREADS_PER_BLOCK 128
#define WARPS_PER_BLOCK 4
#define SEED_BLOCKS_PER_SM 8
#define WARP_BLOCKS_PER_SM 8
static int64_t rt_thread_blocks() { const int64_t d = (int64_t)rt_sm_count() * SEED_BLOCKS_PER_SM; return t_rt->thread_block_cap > 0 && t_rt->thread_block_cap < d ? t_rt->thread_block_cap : d; }
static int64_t rt_warp_blocks() { const int64_t d = (int64_t)rt_sm_count() * WARP_BLOCKS_PER_SM; return t_rt->warp_block_cap > 0 && t_rt->warp_block_cap < d ? t_rt->warp_block_cap : d; }
static int64_t rt_thread_slots() { return rt_thread_blocks() * THREADS_PER_BLOCK; }
static int64_t rt_warp_slots() { return rt_warp_blocks() * WARPS_PER_BLOCK; }
static void rt_set_slot_caps(int64_t thread_blocks, int64_t warp_blocks) { t_rt->thread_block_cap = (int)thread_blocks; t_rt->warp_block_cap = (int)warp_blocks; }

#ifdef B200_DEBUG_CKPT
__device__ volatile unsigned* g_ckpt = nullptr;
static unsigned* g_ckpt_host = nullptr;
#endif
#include "kernels.cuh"

// exclusive scan of u32 counts into u64 offsets (out[n] = total); returns the total
struct U32ToU64 { __host__ __device__ u64 operator()(const u32& v) const { return (u64)v; } };
__global__ void k_scan_close(const u32* in, u64* out, u64 n) { out[n] = n ? out[n - 1] + in[n - 1] : 0; }
// an error left behind by an earlier call of this thread (CUB returns whatever it finds in that slot as its own)
static void rt_report_pending(const char* where)
{
	const cudaError_t le = cudaGetLastError();
	if (le != cudaSuccess) { g_last_cuda = le; fprintf(stderr, "[bwamem-b200] pending CUDA error %s: %s\n", where, cudaGetErrorString(le)); }
}
static bool rt_scratch(size_t need)
{
	if (need <= t_rt->scan_tmp_bytes) return true;
	if (t_rt->scan_tmp) cudaFree(t_rt->scan_tmp);
	t_rt->scan_tmp = nullptr; t_rt->scan_tmp_bytes = 0;
	if (!cu_ok(cudaMalloc(&t_rt->scan_tmp, need + 1024), "cudaMalloc (scan scratch)")) return false;
	t_rt->scan_tmp_bytes = need + 1024;
	return true;
}
static uint64_t rt_scan_u32(const uint32_t* in, uint64_t* out, uint64_t n)
{
	rt_report_pending("before a scan / sort");
	if (n) {
		cub::TransformInputIterator<u64, U32ToU64, const u32*> it(in, U32ToU64());
		size_t need = 0;
		cu_ok(cub::DeviceScan::ExclusiveSum(nullptr, need, it, out, (int64_t)n, rt_stream()), "scan");
		if (!rt_scratch(need)) return 0; // (g_last_cuda is set: the batch fails)
		cu_ok(cub::DeviceScan::ExclusiveSum(t_rt->scan_tmp, need, it, out, (int64_t)n, rt_stream()), "scan");
	}
	k_scan_close<<<1, 1, 0, rt_stream()>>>(in, out, n);
	uint64_t total = 0;
	rt_d2h(&total, out + n, 8);
	return total;
}

static void rt_sort_u64(const uint64_t* in, uint64_t* out, uint64_t n)
{
	rt_report_pending("before a scan / sort");
	if (!n) return;
	size_t need = 0;
	cu_ok(cub::DeviceRadixSort::SortKeys(nullptr, need, in, out, (int64_t)n, 0, 64, rt_stream()), "sort");
	if (!rt_scratch(need)) return;
	cu_ok(cub::DeviceRadixSort::SortKeys(t_rt->scan_tmp, need, in, out, (int64_t)n, 0, 64, rt_stream()), "sort");
}
__global__ void k_iota_u32(u32* p, u64 n) { for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) p[i] = (u32)i; }
static void rt_iota_u32(uint32_t* p, uint64_t n) { if (n) k_iota_u32<<<(unsigned)((n + 255) / 256 < 65535 ? (n + 255) / 256 : 65535), 256, 0, rt_stream()>>>(p, n); }
// out[k] = first position of the ascending array keys[0..n) whose value is >= thr[k]  (k < n_thr <= 8)
struct Thr8 { u32 v[8]; };
__global__ void k_lower_bounds_u32(const u32* keys, u64 n, Thr8 thr, int n_thr, u32* out)
{
	const int k = (int)threadIdx.x;
	if (k >= n_thr) return;
	u64 lo = 0, hi = n;
	while (lo < hi) { const u64 mid = (lo + hi) >> 1; if (keys[mid] < thr.v[k]) lo = mid + 1; else hi = mid; }
	out[k] = (u32)lo;
}
static void rt_lower_bounds_u32(const uint32_t* keys, uint64_t n, const uint32_t* thr, int n_thr, uint32_t* out)
{
	Thr8 t; for (int k = 0; k < 8; ++k) t.v[k] = k < n_thr ? thr[k] : 0xffffffffu;
	k_lower_bounds_u32<<<1, 32, 0, rt_stream()>>>(keys, n, t, n_thr, out);
}
// order u32 payloads by u32 keys (stable); used to bucket work items so that the lanes of a warp get similar ones
static void rt_sort_pairs_u32(const uint32_t* kin, uint32_t* kout, const uint32_t* vin, uint32_t* vout, uint64_t n, int end_bit)
{
	rt_report_pending("before a scan / sort");
	if (!n) return;
	size_t need = 0;
	cu_ok(cub::DeviceRadixSort::SortPairs(nullptr, need, kin, kout, vin, vout, (int64_t)n, 0, end_bit, rt_stream()), "sort pairs");
	if (!rt_scratch(need)) return;
	cu_ok(cub::DeviceRadixSort::SortPairs(t_rt->scan_tmp, need, kin, kout, vin, vout, (int64_t)n, 0, end_bit, rt_stream()), "sort pairs");
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
		make_warp_ctx(P, slot, smem_ + (size_t)wib * ((warp_smem_bytes(P.wm_smem_cap) + 15) & ~(size_t)15), wm, ws); \
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
		make_warp_ctx(P, slot, smem_ + (size_t)wib * ((warp_smem_bytes(P.wm_smem_cap) + 15) & ~(size_t)15), wm, ws); \
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
DEF_THREAD_KERNEL(body_flt_gen)
DEF_WARP_KERNEL_DEVN(wbody_flt_sw)
DEF_THREAD_KERNEL(body_flt_apply)
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
DEF_WARP_KERNEL_DEVN_MB(wbody_rescue_sw_fb, RESCUE_SW_MIN_BLOCKS)
DEF_THREAD_QUEUE_KERNEL_MB(body_tsw_fwd_q, 2)
DEF_THREAD_QUEUE_KERNEL_MB(body_tsw_rev_q, 2)
DEF_WARP_KERNEL(wbody_extend)
DEF_THREAD_QUEUE_KERNEL_MB(body_extend_q, 2) // two blocks per SM fit the shared DP rows: let the state machine keep its state in registers
DEF_THREAD_QUEUE_KERNEL(body_extend_task_q)
// 8-bit cells: the shared rows of four blocks fit an SM; the read-level machine needs ~170 registers, so three of its blocks do
DEF_THREAD_QUEUE_KERNEL_MB(body_extend_q8, 3)
DEF_THREAD_QUEUE_KERNEL_MB(body_extend_task_q8, 4)
DEF_WARP_KERNEL_DEVN(wbody_heavy_prep)
DEF_THREAD_KERNEL(body_ext_prep)
DEF_THREAD_QUEUE_KERNEL_MB(body_ext_left_q, 6)
DEF_THREAD_QUEUE_KERNEL_MB(body_ext_right_q, 6)
DEF_THREAD_QUEUE_KERNEL_MB(body_ext_left_q8, 6)
DEF_THREAD_QUEUE_KERNEL_MB(body_ext_right_q8, 6)
DEF_THREAD_TID_KERNEL(body_ext_replay)
DEF_WARP_KERNEL_DEVN(wbody_extend_heavy)
DEF_THREAD_KERNEL(body_save_regs)
DEF_THREAD_TID_KERNEL(body_final_pe_t)
DEF_THREAD_TID_KERNEL(body_final_se_t)
DEF_THREAD_TID_KERNEL(body_final_pe_collect)
DEF_THREAD_TID_KERNEL(body_final_se_collect)
DEF_THREAD_TID_KERNEL(body_final_pe_consume)
DEF_THREAD_TID_KERNEL(body_final_se_consume)
DEF_THREAD_QUEUE_KERNEL(body_galign_task_q)
DEF_WARP_KERNEL_DEVN(wbody_galign_fb)
DEF_WARP_KERNEL_DEVN(wbody_final_pe_redo)
DEF_WARP_KERNEL_DEVN(wbody_final_se_redo)

static u32* next_counter()
{
	u32* c = t_rt->counters + (t_rt->counter_next++ & 63);
	cu_ok(cudaMemsetAsync(c, 0, sizeof(u32), rt_stream()), "memset");
	return c;
}

static void rt_timer_begin()
{
#ifdef B200_DEBUG_CKPT
	if (!g_ckpt_host) { cudaHostAlloc(&g_ckpt_host, 1024, cudaHostAllocMapped); memset(g_ckpt_host, 0, 1024); unsigned* d = nullptr; cudaHostGetDevicePointer(&d, g_ckpt_host, 0); cudaMemcpyToSymbol(g_ckpt, &d, sizeof(d)); }
#endif
	t_rt->n_marks = 0; t_rt->launches = 0;
	cudaEventRecord(t_rt->ev[0], rt_stream());
}
static void rt_mark(const char* nm)
{
	RtCtx* c = t_rt;
	{ // a launch that was refused (configuration, resources) leaves its error behind: catch it here, under the stage's name
		const cudaError_t le = cudaGetLastError();
		if (le != cudaSuccess) { g_last_cuda = le; fprintf(stderr, "[bwamem-b200] CUDA error at the launch of %s: %s\n", nm, cudaGetErrorString(le)); }
	}
	if (c->override_) return; // a kernel on the auxiliary stream: its time is inside the stages of the main stream it overlaps
	if (c->n_marks < RtCtx::MAXS) { c->name[c->n_marks] = nm; ++c->n_marks; cudaEventRecord(c->ev[c->n_marks], rt_stream()); }
	if (c->trace) {
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
// the main stream waits for the auxiliary one; the stage table shows the wait (the part of the auxiliary kernel that its
// neighbour did not cover) under the auxiliary kernel's name
static void rt_join_aux(const char* nm) { cu_ok(cudaStreamWaitEvent(rt_stream(), t_rt->ev_join, 0), "join"); rt_mark(nm); }
static int rt_launches() { return t_rt ? t_rt->launches : 0; }
// opt-in dynamic shared memory is a per-device function attribute: remembered per launch site and device
// Opt-in to more than 48 KB of dynamic shared memory: per KERNEL and device, only ever raised.  (Round 2 kept this state per
// launch SITE: a kernel launched from two sites had its attribute LOWERED by the second site's first, smaller request, and
// the first site's next launch was refused with "invalid argument" -- found by tests/fuzz_sim.py on batches of different
// read lengths through one process.  The refused launch left its error unread; rt_mark now reads it.)
template <auto Kern>
static void rt_set_smem(size_t bytes, int dev)
{
	static std::atomic<size_t> cur[RT_MAX_DEV];
	static std::mutex mu;
	const int d = dev & (RT_MAX_DEV - 1);
	if (bytes <= cur[d].load(std::memory_order_acquire)) return;
	std::lock_guard<std::mutex> g(mu);
	if (bytes <= cur[d].load(std::memory_order_relaxed)) return;
	if (cu_ok(cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), "cudaFuncSetAttribute")) cur[d].store(bytes, std::memory_order_release);
}
#define RT_SET_SMEM(kern, bytes) rt_set_smem<kern>((size_t)(bytes), t_rt->dev)

#define RT_THREADS(body, P, n) do { i64 n_ = (i64)(n); if (n_ > 0) { \
	i64 nb_ = (n_ + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK; i64 cap_ = (i64)rt_sm_count() * 16; if (nb_ > cap_) nb_ = cap_; \
	k_##body<<<(unsigned)nb_, THREADS_PER_BLOCK, 0, rt_stream()>>>(P, n_); ++t_rt->launches; } rt_mark(#body); } while (0)
#define RT_THREADS_TID(body, P, n) do { i64 n_ = (i64)(n); if (n_ > 0) { \
	i64 nb_ = (n_ + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK; i64 cap_ = rt_thread_blocks(); if (nb_ > cap_) nb_ = cap_; \
	k_##body<<<(unsigned)nb_, THREADS_PER_BLOCK, 0, rt_stream()>>>(P, n_); ++t_rt->launches; } rt_mark(#body); } while (0)
// thread kernels whose threads use the shared-memory rings (make_thread_ctx)
#define RT_THREADS_TID_RING(body, P, n) do { i64 n_ = (i64)(n); if (n_ > 0) { \
	i64 nb_ = (n_ + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK; i64 cap_ = rt_thread_blocks(); if (nb_ > cap_) nb_ = cap_; \
	size_t tsm_ = thread_block_smem_bytes(THREADS_PER_BLOCK, (P).wm_cap); \
	RT_SET_SMEM(k_##body, tsm_); \
	k_##body<<<(unsigned)nb_, THREADS_PER_BLOCK, tsm_, rt_stream()>>>(P, n_); ++t_rt->launches; } rt_mark(#body); } while (0)
// persistent queue kernels with an explicit dynamic shared-memory size (as many blocks per SM as that size allows)
#define RT_THREADS_QUEUE_SMEM(body, P, n, bytes) do { i64 n_ = (i64)(n); if (n_ > 0) { \
	size_t tsm_ = (bytes); \
	i64 nb_ = (n_ + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK; i64 per_sm_ = (i64)((227 * 1024) / (tsm_ + 1024)); if (per_sm_ < 1) per_sm_ = 1; if (per_sm_ > SEED_BLOCKS_PER_SM) per_sm_ = SEED_BLOCKS_PER_SM; \
	i64 cap_ = (i64)rt_sm_count() * per_sm_; if (cap_ > rt_thread_blocks()) cap_ = rt_thread_blocks(); if (nb_ > cap_) nb_ = cap_; \
	RT_SET_SMEM(k_##body, tsm_); \
	k_##body<<<(unsigned)nb_, THREADS_PER_BLOCK, tsm_, rt_stream()>>>(P, n_, next_counter()); ++t_rt->launches; } rt_mark(#body); } while (0)
#define RT_WARPS(wbody, P, n) do { i64 n_ = (i64)(n); if (n_ > 0) { \
	size_t sm_ = ((warp_smem_bytes((P).wm_smem_cap) + 15) & ~(size_t)15) * WARPS_PER_BLOCK; \
	i64 nb_ = (n_ + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK; i64 cap_ = rt_warp_blocks(); if (nb_ > cap_) nb_ = cap_; \
	RT_SET_SMEM(k_##wbody, sm_); \
	k_##wbody<<<(unsigned)nb_, WARPS_PER_BLOCK * 32, sm_, rt_stream()>>>(P, n_, next_counter()); ++t_rt->launches; } rt_mark(#wbody); } while (0)

#define RT_WARPS_DEVN(wbody, P, n_ptr, n_max) do { \
	size_t sm_ = ((warp_smem_bytes((P).wm_smem_cap) + 15) & ~(size_t)15) * WARPS_PER_BLOCK; \
	i64 nb_ = rt_warp_blocks(); \
	RT_SET_SMEM(k_##wbody, sm_); \
	k_##wbody<<<(unsigned)nb_, WARPS_PER_BLOCK * 32, sm_, rt_stream()>>>(P, n_ptr, next_counter()); ++t_rt->launches; rt_mark(#wbody); } while (0)

__global__ void k_relayout(const u32* src, u64 bwt_size, u32* dst, u64 n_blocks, const u64* sb, int sb_shift)
{
	for (u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (u64)gridDim.x * blockDim.x) fm_relayout_block(src, bwt_size, dst, b, sb, sb_shift);
}
// denser suffix-array samples at load time: row k (a multiple of `dense`) walks to the next row the image samples
__global__ void k_densify_sa(DevIndex x, u64 n_rows, u32 dense, u64* out)
{
	const u64 n_out = (n_rows + dense - 1) / dense;
	for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < n_out; t += (u64)gridDim.x * blockDim.x) out[t] = fm_sa(x, t * dense);
}
// one level of the prefix tables (fm.cuh kmer_build_entry): entry e of level L1 from level L1 - 1
__global__ void k_kmer_level(DevIndex x, uint4* tab, int L1, u64 n)
{
	for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (u64)gridDim.x * blockDim.x) kmer_build_entry(x, tab, L1, e);
}
#define RT_KMER_LEVEL(x, tab, L1, n) do { k_kmer_level<<<(unsigned)(rt_sm_count() * 16), 256, 0, rt_stream()>>>((x), (tab), (L1), (n)); cu_ok(cudaGetLastError(), "kmer level"); } while (0)
#define RT_DENSIFY(x, n_rows, dense, out) do { k_densify_sa<<<(unsigned)(rt_sm_count() * 16), 256, 0, rt_stream()>>>((x), (n_rows), (dense), (out)); cu_ok(cudaGetLastError(), "densify"); } while (0)
#define RT_RELAYOUT(src, size, dst, nb, sb, sbs) do { k_relayout<<<(unsigned)(rt_sm_count() * 16), 256, 0, rt_stream()>>>((src), (size), (dst), (nb), (sb), (sbs)); cu_ok(cudaGetLastError(), "relayout"); } while (0)


#include "engine_impl.inl"

using namespace b200;

// ---- C-ABI (include/bwamem_b200.h) ---------------------------------------------------------------------
// Devices of a handle: BWA_B200_DEVICES = "all" or a comma-separated list of ordinals (an in-library pool: one index
// replica per listed GPU, calls go to the least busy one); else BWA_B200_DEVICE = one ordinal; else the CUDA device
// current on the opening thread (b200_setDevice / cudaSetDevice).  The Java API has no notion of a device.
static bool devices_from_env(std::vector<int>& out, std::string& err)
{
	int ndev = rt_device_count();
	if (ndev <= 0) { err = "no CUDA device: this engine has no CPU fallback"; return false; }
	if (ndev > RT_MAX_DEV) ndev = RT_MAX_DEV;
	const char* list = getenv("BWA_B200_DEVICES");
	if (list && *list) {
		if (strcmp(list, "all") == 0) { for (int d = 0; d < ndev; ++d) out.push_back(d); return true; }
		for (const char* p = list; *p;) {
			char* e = nullptr; long v = strtol(p, &e, 10);
			if (e == p || v < 0 || v >= ndev) { err = std::string("BWA_B200_DEVICES=") + list + " names an unusable device"; return false; }
			bool dup = false; for (int d : out) dup |= d == (int)v;
			if (!dup) out.push_back((int)v);
			p = *e == ',' ? e + 1 : e;
			if (*e && *e != ',') { err = std::string("BWA_B200_DEVICES=") + list + " is malformed"; return false; }
		}
		return !out.empty();
	}
	if (const char* dv = getenv("BWA_B200_DEVICE")) {
		const int v = atoi(dv);
		if (v < 0 || v >= ndev) { err = std::string("BWA_B200_DEVICE=") + dv + " is not a usable device"; return false; }
		out.push_back(v);
		return true;
	}
	out.push_back(rt_current_device());
	return true;
}

extern "C" {

void* jnibwa_openIndex(int fd)
{
	std::vector<int> devs; std::string err;
	if (!devices_from_env(devs, err)) {
		fprintf(stderr, "[bwamem-b200] %s\n", err.c_str());
		if (fd >= 0) close(fd);
		return nullptr;
	}
	const int dev0 = rt_current_device(); // the caller's own current device is left as it was
	g_last_cuda = cudaSuccess;
	Engine* e = new Engine();
	const bool ok = e->open_from_fd(fd, devs);
	if (!ok) { fprintf(stderr, "[bwamem-b200] openIndex: %s\n", e->last_error.c_str()); e->close_all(); delete e; e = nullptr; }
	cudaSetDevice(dev0);
	return e;
}

int jnibwa_destroyIndex(void* h)
{
	if (!h) return 0;
	Engine* e = (Engine*)h;
	const int dev0 = rt_current_device();
	e->close_all(); // (Java serialises close() and refuses it while aligners are alive: BwaMemIndex.java:383-397)
	delete e;
	cudaSetDevice(dev0);
	return 0;
}

void* jnibwa_getRefContigNames(void* h, size_t* size) { return h ? contig_names_blob(((Engine*)h)->img, size) : nullptr; }

// Re-entrant: every call borrows one of the handle's contexts (its own device binding, streams and workspace), so
// threads overlap up to the number of contexts and queue beyond it.
void* jnibwa_createAlignments(void* h, void* opt, void* pes, char* seqs, size_t* size)
{
	if (!h || !opt || !seqs || !size) return nullptr;
	Engine* e = (Engine*)h;
	Batch* b = e->acquire();
	if (!b) return nullptr;
	g_last_cuda = cudaSuccess;
	void* r = b->align((const MemOpt*)opt, (const PeStat*)pes, seqs, size);
	if (r && (g_last_cuda != cudaSuccess || cudaGetLastError() != cudaSuccess)) { free(r); r = nullptr; b->last_error = "CUDA error during the batch"; }
	if (!r) fprintf(stderr, "[bwamem-b200] createAlignments: %s\n", b->last_error.c_str());
	e->release(b);
	return r;
}

void* mem_opt_init(void) { return default_options(); }
void jnibwa_free(void* p) { free(p); }
const char* jnibwa_getVersion(void) { return "bwamem-b200 0.2 (sm_100a; bwa-mem semantics of lh3/bwa cb95061)"; }

// ---- instrumentation for bench.py ------------------------------------------------------------------
// HBM-resident measurement: stage a request (host parse + upload, untimed by the caller), then run the device
// pipeline on it leaving the record stream in device memory.  Both use context 0 of the handle (the first device of
// its pool).  alignStaged returns the stream's byte size, or -1.
int b200_stageBatch(void* h, const char* seqs)
{
	Engine* e = (Engine*)h;
	if (!e || e->ctxs.empty()) return 1;
	Batch* b = e->acquire(e->ctxs[0]);
	g_last_cuda = cudaSuccess;
	const bool ok = b->stage(seqs) && g_last_cuda == cudaSuccess;
	e->release(b);
	return ok ? 0 : 1;
}
long long b200_alignStaged(void* h, void* opt, void* pes)
{
	Engine* e = (Engine*)h;
	if (!e || e->ctxs.empty()) return -1;
	Batch* b = e->acquire(e->ctxs[0]);
	size_t size = 0;
	g_last_cuda = cudaSuccess;
	void* r = b->run((const MemOpt*)opt, (const PeStat*)pes, &size, false);
	const bool ok = r && g_last_cuda == cudaSuccess;
	if (!ok) fprintf(stderr, "[bwamem-b200] alignStaged: %s\n", b->last_error.c_str());
	e->release(b);
	return ok ? (long long)size : -1;
}
// device time of the last batch of context 0: first to last stage event on its stream (ms)
float b200_lastBatchMs(void* h)
{
	Engine* e = (Engine*)h; float t = 0;
	if (!e || e->ctxs.empty()) return 0;
	RtCtx* c = e->ctxs[0]->rt;
	cudaSetDevice(c->dev);
	if (c->n_marks > 0) cudaEventElapsedTime(&t, c->ev[0], c->ev[c->n_marks]);
	return t;
}
int b200_setDevice(int dev) { return cudaSetDevice(dev) == cudaSuccess ? 0 : 1; }
// contexts of a handle (devices x batches in flight per device) and the device of context i
int b200_contextCount(void* h) { return h ? (int)((Engine*)h)->ctxs.size() : 0; }
int b200_contextDevice(void* h, int i) { Engine* e = (Engine*)h; return e && i >= 0 && i < (int)e->ctxs.size() ? e->ctxs[(size_t)i]->rep->dev : -1; }
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
// copies up to cap stage times (ms) of the last batch of context 0; names[i] points at static strings; returns the stage count
int b200_lastStageTimes(void* h, float* ms, const char** names, int cap)
{
	Engine* e = (Engine*)h;
	if (!e || e->ctxs.empty()) return 0;
	RtCtx* c = e->ctxs[0]->rt;
	cudaSetDevice(c->dev);
	int n = 0; // aggregated by stage name, in order of first appearance
	for (int i = 0; i < c->n_marks; ++i) {
		float t = 0; cudaEventElapsedTime(&t, c->ev[i], c->ev[i + 1]);
		int j = 0;
		for (; j < n; ++j) if (strcmp(names[j], c->name[i]) == 0) break;
		if (j == n) { if (n >= cap) continue; names[n] = c->name[i]; ms[n] = 0; ++n; }
		ms[j] += t;
	}
	return n;
}
// counters of the most recent call on the handle: cells_ext, cells_glb, cells_msw, n_ext, n_glb, n_msw, n_seed_lookups, n_launches
void b200_lastCounters(void* h, uint64_t* out)
{
	Engine* e = (Engine*)h;
	std::lock_guard<std::mutex> lock(e->mu);
	memcpy(out, &e->last_counters, sizeof(Counters));
	out[6] = e->last_n_seed;
	out[7] = (uint64_t)e->n_launches;
}

} // extern "C"
