// simt.cuh -- thin SIMT vocabulary used by every kernel of the engine.
//
// Product build (nvcc, sm_100a): the macros below are the CUDA built-ins; a
// "warp variable" is one register per lane and the collectives are shuffles /
// ballots / redux.  This is the only configuration that is ever shipped or
// loaded by the package.
//
// Kernel-debug build (-DB200_SIM, host g++; used ONLY by tests/cudasim to run
// the very same kernel source under gdb/ASan on the CPU-only dev box, never
// built into libbwa_b200.so and never reachable from the package): a warp
// variable is an array of 32 values, FOR_LANES is a loop, and collectives are
// loops over that array.  Warp kernels are then executed once per warp and
// thread kernels once per thread by tests/cudasim's launcher.
#pragma once
#include <cstdint>
#include <cstring>

typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;
typedef int32_t i32;
typedef uint16_t u16;
typedef int16_t i16;
typedef uint8_t u8;
typedef int8_t i8;

#ifndef B200_SIM
// ------------------------------------------------------------------ CUDA
#include <cuda_runtime.h>
#define DEV __device__ __forceinline__
#define DEVNI __device__ __noinline__
#define HOSTDEV __host__ __device__ __forceinline__
#define DEVM __device__ __forceinline__
#define KERNEL __global__
#define FULLMASK 0xffffffffu
#define WV(T, name) T name
#define L_(name) name
#define FOR_LANES
// Every lane-aware device function is a template on W, the number of lanes that cooperate on one work item:
// W = 32 (a warp per item: shuffles, ballots, __syncwarp) or W = 1 (a thread per item: collectives are identities).
template <int W> __device__ __forceinline__ int lane_id_() { return W == 32 ? (int)(threadIdx.x & 31) : 0; }
template <int W> __device__ __forceinline__ void wsync_() { if (W == 32) __syncwarp(); }
#define LANE_DECL const int lane = lane_id_<W>()
#define ONE_LANE wsync_<W>(); if (lane == 0)       /* pre-synchronised single-lane write (may overwrite data other lanes just read) */
#define ONE_LANE_NOSYNC if (lane == 0)            /* append-only single-lane write */
#define WARP_SYNC() wsync_<W>()
#define LANE0_ONLY wsync_<W>(); if (lane == 0)    /* a scalar section run by one lane of the item's lane group */

template <class T> DEV T ldg(const T* p) { return __ldg(p); }
DEV int popc64(u64 x) { return __popcll(x); }
template <int W, class T> DEV void wv_shfl_up_(T& dst, const T& src, int d) { if (W == 32) dst = __shfl_up_sync(FULLMASK, src, d); else dst = src; }
template <int W> DEV int wv_reduce_max_(const int& v) { return W == 32 ? __reduce_max_sync(FULLMASK, v) : v; }
template <int W> DEV int wv_reduce_min_(const int& v) { return W == 32 ? __reduce_min_sync(FULLMASK, v) : v; }
template <int W> DEV u32 wv_ballot_(const bool& p) { return W == 32 ? __ballot_sync(FULLMASK, p) : (p ? 1u : 0u); }
template <int W, class T> DEV T wv_get_(const T& v, int l) { return W == 32 ? __shfl_sync(FULLMASK, v, l) : v; }
#define wv_shfl_up(dst, src, d) wv_shfl_up_<W>(dst, src, d)
#define wv_reduce_max(v) wv_reduce_max_<W>(v)
#define wv_reduce_min(v) wv_reduce_min_<W>(v)
#define wv_ballot(p) wv_ballot_<W>(p)
#define wv_get(v, l) wv_get_<W>(v, l)
template <int W> DEV i64 wv_reduce_max64_(const i64& v) { i64 r = v; if (W == 32) for (int d = 16; d; d >>= 1) { i64 o = __shfl_xor_sync(FULLMASK, r, d); r = o > r ? o : r; } return r; }
template <int W> DEV i64 wv_reduce_min64_(const i64& v) { i64 r = v; if (W == 32) for (int d = 16; d; d >>= 1) { i64 o = __shfl_xor_sync(FULLMASK, r, d); r = o < r ? o : r; } return r; }
#define wv_reduce_max64(v) wv_reduce_max64_<W>(v)
#define wv_reduce_min64(v) wv_reduce_min64_<W>(v)
// Blackwell DPX: two signed 16-bit cells per 32-bit register (VIADDMNMX.S16x2[.RELU], VIMNMX3.S16x2 in SASS)
DEV u32 dpx_addmin_s16x2(u32 a, u32 b, u32 c) { return __viaddmin_s16x2(a, b, c); }            // min(a + b, c) per half
DEV u32 dpx_addmax_relu_s16x2(u32 a, u32 b, u32 c) { return __viaddmax_s16x2_relu(a, b, c); }  // max(a + b, c, 0) per half
DEV u32 dpx_max3_s16x2(u32 a, u32 b, u32 c) { return __vimax3_s16x2(a, b, c); }                // max(a, b, c) per half
DEV u32 v_add2(u32 a, u32 b) { return __vadd2(a, b); }                                         // a + b per half (wrapping)
DEV u32 v_maxs2(u32 a, u32 b) { return __vmaxs2(a, b); }                                       // max(a, b) per half (signed)
DEV u32 byte_perm(u32 a, u32 b, u32 s) { return __byte_perm(a, b, s); }                        // PRMT: 4 bytes of {a, b} picked by nibbles (bit 3: sign fill)
// full-mask vote of a thread-per-item kernel whose 32 lanes all take part (persistent queue kernels)
DEV bool lanes_all(bool p) { return __all_sync(FULLMASK, p) != 0; }
// thread-per-item kernels that hand out work to a whole warp at a time: lanes per hand-out, this thread's lane, a
// broadcast of lane 0's value and the point where the lanes meet again
#define SIMT_GROUP 32
DEV u32 lanes_ballot(bool p) { return __ballot_sync(FULLMASK, p); }
DEV u32 lanes_shfl(u32 v, int src) { return __shfl_sync(FULLMASK, v, src); }
DEV int simt_lane() { return (int)(threadIdx.x & 31); }
DEV u32 lane_bcast0(u32 v) { return __shfl_sync(FULLMASK, v, 0); }
DEV void warp_converge() { __syncwarp(); }
DEV u64 atomic_add_u64(u64* p, u64 v) { return (u64)atomicAdd((unsigned long long*)p, (unsigned long long)v); }
DEV u32 atomic_add_u32(u32* p, u32 v) { return atomicAdd(p, v); }
DEV u32 atomic_or_u32(u32* p, u32 v) { return atomicOr(p, v); }
DEV u32 atomic_max_u32(u32* p, u32 v) { return atomicMax(p, v); }
#else
// ------------------------------------------------------------------ host kernel-debug harness (tests/cudasim only)
#include <algorithm>
#define DEV static inline
#define DEVNI static
#define HOSTDEV static inline
#define DEVM inline
#define KERNEL static
#define WV(T, name) T name[32]
#define L_(name) name[lane]
#define FOR_LANES for (int lane = 0; lane < W; ++lane)
#define LANE_DECL const int lane0_unused_ = 0; (void)lane0_unused_
#define ONE_LANE
#define ONE_LANE_NOSYNC
#define WARP_SYNC()
#define LANE0_ONLY

struct uint4 { u32 x, y, z, w; };
static inline uint4 make_uint4(u32 x, u32 y, u32 z, u32 w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
template <class T> DEV T ldg(const T* p) { return *p; }
DEV int popc64(u64 x) { return __builtin_popcountll(x); }
template <int W, class T> DEV void wv_shfl_up_(T (&dst)[32], const T (&src)[32], int d) { T t[32]; for (int l = 0; l < W; ++l) t[l] = l >= d ? src[l - d] : src[l]; for (int l = 0; l < W; ++l) dst[l] = t[l]; }
template <int W> DEV int wv_reduce_max_(const int (&v)[32]) { int m = v[0]; for (int l = 1; l < W; ++l) m = v[l] > m ? v[l] : m; return m; }
template <int W> DEV int wv_reduce_min_(const int (&v)[32]) { int m = v[0]; for (int l = 1; l < W; ++l) m = v[l] < m ? v[l] : m; return m; }
template <int W> DEV u32 wv_ballot_(const bool (&p)[32]) { u32 m = 0; for (int l = 0; l < W; ++l) m |= (u32)(p[l] ? 1u : 0u) << l; return m; }
template <int W, class T> DEV T wv_get_(const T (&v)[32], int l) { return v[l]; }
template <int W> DEV i64 wv_reduce_max64_(const i64 (&v)[32]) { i64 m = v[0]; for (int l = 1; l < W; ++l) m = v[l] > m ? v[l] : m; return m; }
template <int W> DEV i64 wv_reduce_min64_(const i64 (&v)[32]) { i64 m = v[0]; for (int l = 1; l < W; ++l) m = v[l] < m ? v[l] : m; return m; }
template <int W> DEV int wv_reduce_add_(const int (&v)[32]) { int m = 0; for (int l = 0; l < W; ++l) m += v[l]; return m; }
#define wv_shfl_up(dst, src, d) wv_shfl_up_<W>(dst, src, d)
#define wv_reduce_max(v) wv_reduce_max_<W>(v)
#define wv_reduce_min(v) wv_reduce_min_<W>(v)
#define wv_reduce_max64(v) wv_reduce_max64_<W>(v)
#define wv_reduce_min64(v) wv_reduce_min64_<W>(v)
#define wv_reduce_add(v) wv_reduce_add_<W>(v)
#define wv_ballot(p) wv_ballot_<W>(p)
#define wv_get(v, l) wv_get_<W>(v, l)
static inline i32 h16_(u32 v, int k) { return (i32)(i16)(u16)(v >> (16 * k)); }
static inline u32 pk16_(i32 lo, i32 hi) { return (u32)(u16)(i16)lo | (u32)(u16)(i16)hi << 16; }
DEV u32 dpx_addmin_s16x2(u32 a, u32 b, u32 c) { i32 r[2]; for (int k = 0; k < 2; ++k) { i32 t = (i32)(i16)(h16_(a, k) + h16_(b, k)); r[k] = t < h16_(c, k) ? t : h16_(c, k); } return pk16_(r[0], r[1]); }
DEV u32 dpx_addmax_relu_s16x2(u32 a, u32 b, u32 c) { i32 r[2]; for (int k = 0; k < 2; ++k) { i32 t = (i32)(i16)(h16_(a, k) + h16_(b, k)); t = t > h16_(c, k) ? t : h16_(c, k); r[k] = t > 0 ? t : 0; } return pk16_(r[0], r[1]); }
DEV u32 dpx_max3_s16x2(u32 a, u32 b, u32 c) { i32 r[2]; for (int k = 0; k < 2; ++k) { i32 t = h16_(a, k) > h16_(b, k) ? h16_(a, k) : h16_(b, k); r[k] = t > h16_(c, k) ? t : h16_(c, k); } return pk16_(r[0], r[1]); }
DEV u32 v_add2(u32 a, u32 b) { return pk16_(h16_(a, 0) + h16_(b, 0), h16_(a, 1) + h16_(b, 1)); }
DEV u32 v_maxs2(u32 a, u32 b) { return pk16_(h16_(a, 0) > h16_(b, 0) ? h16_(a, 0) : h16_(b, 0), h16_(a, 1) > h16_(b, 1) ? h16_(a, 1) : h16_(b, 1)); }
DEV u32 byte_perm(u32 a, u32 b, u32 s)
{
	const u64 src = (u64)a | (u64)b << 32; u32 r = 0;
	for (int k = 0; k < 4; ++k) { const u32 n = s >> (4 * k) & 0xf; u32 byte = (u32)(src >> (8 * (n & 7)) & 0xff); if (n & 8) byte = (byte & 0x80) ? 0xffu : 0u; r |= byte << (8 * k); }
	return r;
}
DEV bool lanes_all(bool p) { return p; } // the harness runs one thread at a time
#define SIMT_GROUP 1
DEV u32 lanes_ballot(bool p) { return p ? 1u : 0u; }
DEV u32 lanes_shfl(u32 v, int) { return v; }
DEV int simt_lane() { return 0; }
DEV u32 lane_bcast0(u32 v) { return v; }
DEV void warp_converge() {}
DEV u64 atomic_add_u64(u64* p, u64 v) { u64 o = *p; *p += v; return o; }
DEV u32 atomic_add_u32(u32* p, u32 v) { u32 o = *p; *p += v; return o; }
DEV u32 atomic_or_u32(u32* p, u32 v) { u32 o = *p; *p |= v; return o; }
DEV u32 atomic_max_u32(u32* p, u32 v) { u32 o = *p; if (v > o) *p = v; return o; }
DEV int __ffs(u32 x) { return x ? __builtin_ctz(x) + 1 : 0; }
DEV int __popc(u32 x) { return __builtin_popcount(x); }
DEV int __clz(u32 x) { return x ? __builtin_clz(x) : 32; }
#endif

// Development aid (-DB200_DEBUG_CKPT builds only): lane 0 of every warp posts its last checkpoint into host-mapped
// memory so that a hung kernel can be located from the host (engine.cu, B200_TRACE).
#if defined(B200_DEBUG_CKPT) && !defined(B200_SIM)
extern __device__ volatile unsigned* g_ckpt;
#define DBGP(code, val) do { if ((threadIdx.x & 31) == 0 && g_ckpt) { unsigned w_ = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; if (w_ < 64) { g_ckpt[2 * w_] = (unsigned)(code); g_ckpt[2 * w_ + 1] = (unsigned)(val); __threadfence_system(); } } } while (0)
#else
#define DBGP(code, val) do {} while (0)
#endif

// read of a flag word that this thread's own atomics may have changed (bypasses the non-coherent L1)
DEV u32 err_peek(const u32* p) { return *(const volatile u32*)p; }
template <class T> DEV T imax(T a, T b) { return a > b ? a : b; }
template <class T> DEV T imin(T a, T b) { return a < b ? a : b; }
DEV int iabs(int a) { return a < 0 ? -a : a; }
