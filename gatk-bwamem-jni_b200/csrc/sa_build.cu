// sa_build.cu -- FM-index construction on the GPU (not on the timed path).
//
// Replaces, behind jnibwa_createReferenceIndex / b200_buildImageFromBases, the suffix sorting that un-vendored
// lh3/bwa performs in bwa_idx_build (called at /root/reference/src/main/c/org_broadinstitute_hellbender_utils_bwa_BwaMemIndex.c:59)
// for genomes where the host sorter of index_build.cpp would take minutes (3 Gbp = 6 G suffixes).  Output is the
// same FmIndex the host builder produces (byte-identical image; tests/test_parity_gpu.py::test_device_index_builder).
//
// Text T = fwd(ref) || revcomp(ref), 2 bits per base, '$' at the end sorts before every base.
//   1. The suffixes are partitioned by their first 16 bases into chunks of at most `cap` suffixes; chunks are
//      produced in ascending prefix order, so the rows of a chunk are final once the chunk is sorted.
//   2. A chunk is radix-sorted (CUB) by a 64-bit key: the next 28 bases (56 bits) and min(remaining length, 28) in
//      the low byte, so a suffix that runs into '$' sorts before a longer one with the same bases.
//   3. Runs of equal keys are refined, 28 bases deeper per round, by a stable two-pass radix sort of the still
//      unresolved elements on (run head, next key); resolved elements drop out.  The number of rounds is the longest
//      repeat in the chunk / 28; only members of exact repeats stay in the working set.
//   4. BWT symbols (T[SA-1]) and the suffix-array samples (every 32nd row) are emitted per chunk; the packed,
//      occurrence-interleaved BWT (64-byte blocks of 128 symbols) is assembled at the end with a device scan.
// Peak device memory: ~1 byte per suffix (BWT symbols) + 40 bytes per chunk element + the packed text.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>
#include "bwa_image.h"

namespace b200 {

typedef uint64_t u64; typedef uint32_t u32; typedef uint8_t u8; typedef int64_t i64;

#define SB_CHECK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "[bwamem-b200] index build: %s failed: %s\n", #call, cudaGetErrorString(e_)); ok = false; goto done; } } while (0)

// ---- packed text ----------------------------------------------------------------------------------------------
__device__ __forceinline__ int pac_get_d(const u8* pac, i64 l) { return pac[l >> 2] >> ((~l & 3) << 1) & 3; }

__global__ void k_pack_text(const u8* pac, i64 L, u64* w, i64 n_words)
{
	const i64 n = 2 * L;
	for (i64 wi = (i64)blockIdx.x * blockDim.x + threadIdx.x; wi < n_words; wi += (i64)gridDim.x * blockDim.x) {
		u64 v = 0;
		const i64 i0 = wi << 5;
		for (int k = 0; k < 32; ++k) {
			const i64 i = i0 + k;
			u64 c = 0;
			if (i < L) c = (u64)pac_get_d(pac, i); else if (i < n) c = (u64)(3 - pac_get_d(pac, n - 1 - i));
			v |= c << ((31 - k) << 1);
		}
		w[wi] = v;
	}
}

__device__ __forceinline__ u64 text_key32(const u64* w, i64 i) // 32 bases from position i (zero padded by the words behind the text)
{
	const i64 wi = i >> 5; const int sh = (int)(i & 31) << 1;
	const u64 a = w[wi];
	return sh ? (a << sh) | (w[wi + 1] >> (64 - sh)) : a;
}
__device__ __forceinline__ int text_base(const u64* w, i64 i) { return (int)(w[i >> 5] >> ((31 - (i & 31)) << 1) & 3); }

// sort key of the suffix at p, `depth` bases in: 28 bases | min(remaining, 28)
__device__ __forceinline__ u64 suffix_key(const u64* w, i64 n, i64 p, i64 depth)
{
	const i64 q = p + depth;
	i64 rem = n - q; rem = rem < 0 ? 0 : (rem > 28 ? 28 : rem);
	u64 k = rem > 0 ? text_key32(w, q) >> 8 : 0; // top 56 bits = 28 bases (bases past the end of T are zero words)
	if (rem < 28) k = rem > 0 ? k & (~0ull << ((28 - rem) << 1)) : 0; // (already zero padded; kept explicit)
	return k << 8 | (u64)rem;
}

// ---- chunk selection: positions whose 16-base prefix lies in [lo, hi) -----------------------------------------
__global__ void k_select(const u64* w, i64 n, u64 lo, u64 hi, u64* out_pos, unsigned long long* count, u64 cap)
{
	const int lane = threadIdx.x & 31;
	for (i64 base = ((i64)blockIdx.x * blockDim.x + threadIdx.x - lane) * 1; base < n; base += (i64)gridDim.x * blockDim.x) {
		const i64 i = base + lane;
		bool hit = false;
		if (i < n) { const u64 pre = text_key32(w, i) >> 32; hit = pre >= lo && pre < hi; } // A-padded past the end, like the keys
		const u32 m = __ballot_sync(0xffffffffu, hit);
		if (m) {
			unsigned long long start = 0;
			if (lane == 0) start = atomicAdd(count, (unsigned long long)__popc(m));
			start = __shfl_sync(0xffffffffu, start, 0);
			if (hit) { const u64 slot = start + (u64)__popc(m & ((1u << lane) - 1)); if (slot < cap) out_pos[slot] = (u64)i; }
		}
	}
}

__global__ void k_keys0(const u64* w, i64 n, const u64* pos, u64* key, u64 m)
{
	for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (u64)gridDim.x * blockDim.x) key[j] = suffix_key(w, n, (i64)pos[j], 0);
}

// after a sort: flag the elements that are still tied with a neighbour (same group, same key, key not terminal)
// head[j] = 1 when j starts a run.  unres[j] = 1 when j belongs to a run of length >= 2.
__global__ void k_mark(const u64* key, const u32* grp, u64 m, u32* unres, u32* headidx)
{
	for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (u64)gridDim.x * blockDim.x) {
		const u64 k = key[j];
		const bool live = (k & 0xff) == 28;
		const bool tie_prev = j > 0 && live && key[j - 1] == k && (!grp || grp[j - 1] == grp[j]);
		const bool tie_next = j + 1 < m && live && key[j + 1] == k && (!grp || grp[j + 1] == grp[j]);
		unres[j] = (tie_prev || tie_next) ? 1u : 0u;
		headidx[j] = tie_prev ? 0u : (u32)j; // inclusive max-scan turns this into "index of my run's head"
	}
}

struct MaxOp { __device__ __forceinline__ u32 operator()(u32 a, u32 b) const { return a > b ? a : b; } };

// compaction of the unresolved elements (slot order is preserved): slot, position, new group = slot of the run head
__global__ void k_compact(const u32* unres, const u32* off, const u32* headidx, const u32* slot_in, const u64* pos_in, u64 m,
                          u32* slot_out, u64* pos_out, u32* grp_out)
{
	for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (u64)gridDim.x * blockDim.x) {
		if (!unres[j]) continue;
		const u32 o = off[j];
		slot_out[o] = slot_in ? slot_in[j] : (u32)j;
		pos_out[o] = pos_in[j];
		const u32 h = headidx[j];
		grp_out[o] = slot_in ? slot_in[h] : h;
	}
}

__global__ void k_keys_round(const u64* w, i64 n, const u64* pos, i64 depth, u64* key, u32* idx, u64 m)
{
	for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (u64)gridDim.x * blockDim.x) { key[j] = suffix_key(w, n, (i64)pos[j], depth); idx[j] = (u32)j; }
}
__global__ void k_gather_u32(const u32* src, const u32* perm, u32* dst, u64 m)
{
	for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (u64)gridDim.x * blockDim.x) dst[j] = src[perm[j]];
}
// apply the (group, key) order: element j of the new order is old element perm[j]; it lands in slot slot[j]
__global__ void k_apply(const u32* perm, const u64* pos_old, const u64* key_old, const u32* slot, u64* pos_new, u64* key_new, u64* chunk_sa, u64 m)
{
	for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (u64)gridDim.x * blockDim.x) {
		const u32 p = perm[j];
		const u64 ps = pos_old[p];
		pos_new[j] = ps; key_new[j] = key_old[p];
		chunk_sa[slot[j]] = ps;
	}
}

// rows of a finished chunk: BWT symbol of every row, SA sample of every 32nd row, the primary row
__global__ void k_emit(const u64* w, const u64* chunk_sa, u64 m, u64 row0, u8* bwt_sym, u64* sa_samples, unsigned long long* primary)
{
	for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (u64)gridDim.x * blockDim.x) {
		const u64 p = chunk_sa[j], row = row0 + j;
		if (p == 0) { *primary = row; bwt_sym[row] = 0xff; }
		else bwt_sym[row] = (u8)text_base(w, (i64)p - 1);
		if ((row & 31) == 0) sa_samples[row >> 5] = p;
	}
}

// ---- packed, occurrence-interleaved BWT ------------------------------------------------------------------------
__device__ __forceinline__ int sym_at(const u8* bwt_sym, u64 k, u64 primary) { return (int)bwt_sym[k < primary ? k : k + 1]; }

__global__ void k_block_counts(const u8* bwt_sym, u64 n, u64 primary, u64 n_blocks, u32* c0, u32* c1, u32* c2, u32* c3)
{
	for (u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (u64)gridDim.x * blockDim.x) {
		u32 c[4] = { 0, 0, 0, 0 };
		const u64 k0 = b << 7, k1 = k0 + 128 < n ? k0 + 128 : n;
		for (u64 k = k0; k < k1; ++k) { const int s = sym_at(bwt_sym, k, primary); c[0] += s == 0; c[1] += s == 1; c[2] += s == 2; c[3] += s == 3; }
		c0[b] = c[0]; c1[b] = c[1]; c2[b] = c[2]; c3[b] = c[3];
	}
}
struct U32ToU64b { __host__ __device__ u64 operator()(const u32& v) const { return (u64)v; } };

__global__ void k_write_blocks(const u8* bwt_sym, u64 n, u64 primary, u64 n_blocks, const u64* o0, const u64* o1, const u64* o2, const u64* o3, u32* bwt)
{
	for (u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (u64)gridDim.x * blockDim.x) {
		u32* blk = bwt + (b << 4);
		const u64 occ[4] = { o0[b], o1[b], o2[b], o3[b] };
		for (int t = 0; t < 4; ++t) { blk[2 * t] = (u32)occ[t]; blk[2 * t + 1] = (u32)(occ[t] >> 32); }
		const u64 k0 = b << 7, k1 = k0 + 128 < n ? k0 + 128 : n;
		const int n_words = (int)((k1 - k0 + 15) >> 4);
		for (int wd = 0; wd < n_words; ++wd) {
			u32 v = 0;
			for (int t = 0; t < 16; ++t) { const u64 k = k0 + (u64)wd * 16 + t; if (k < k1) v |= (u32)sym_at(bwt_sym, k, primary) << ((15 - t) << 1); }
			blk[8 + wd] = v;
		}
	}
}

static inline unsigned grid_for(u64 n) { u64 g = (n + 255) / 256; if (g < 1) g = 1; if (g > 148 * 16) g = 148 * 16; return (unsigned)g; }

// Suffix-sorts T on the current CUDA device and fills fm.  Returns false (fm untouched) when no device is usable
// or anything fails; the caller then uses the host sorter.
bool build_fm_device(const RefSeqs& ref, FmIndex& fm)
{
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return false;
	const i64 L = ref.l_pac, n = 2 * L;
	if (n <= 0) return false;
	bool ok = true;
	const i64 n_words = n / 32 + 4;
	const u64 n_rows = (u64)n + 1;
	u64 cap = (u64)400 << 20; // elements per chunk
	if (const char* e = getenv("BWA_B200_BUILD_CHUNK")) { long long v = atoll(e); if (v > 1024) cap = (u64)v; }
	if (cap > (u64)n) cap = (u64)n;
	if (cap >= ((u64)1 << 32)) cap = ((u64)1 << 32) - 1;

	u8* d_pac = nullptr; u64* d_w = nullptr; u8* d_sym = nullptr; u64* d_sa_s = nullptr; unsigned long long* d_scalar = nullptr;
	u64 *d_pos = nullptr, *d_pos2 = nullptr, *d_key = nullptr, *d_key2 = nullptr, *d_csa = nullptr;
	u64 *u_pos = nullptr, *u_pos2 = nullptr, *u_key = nullptr, *u_key2 = nullptr;
	u32 *u_slot = nullptr, *u_slot2 = nullptr, *u_grp = nullptr, *u_grp2 = nullptr, *u_idx = nullptr, *u_idx2 = nullptr, *u_idx3 = nullptr, *d_unres = nullptr, *d_off = nullptr, *d_head = nullptr;
	void* d_tmp = nullptr; size_t tmp_bytes = 0;
	u32 *c0 = nullptr, *c1 = nullptr, *c2 = nullptr, *c3 = nullptr; u64 *o0 = nullptr, *o1 = nullptr, *o2 = nullptr, *o3 = nullptr; u32* d_bwt = nullptr;
	const u64 n_sa = ((u64)n + 32) / 32;
	const u64 n_blocks = ((u64)n + 127) / 128;
	const u64 bwt_size = ((u64)n + 15) / 16 + (n_blocks + 1) * 8;
	u64 primary = 0, row0 = 1, totals[4] = { 0, 0, 0, 0 };
	cudaStream_t st = 0;
	{
		SB_CHECK(cudaMalloc(&d_pac, (size_t)(L / 4 + 2)));
		SB_CHECK(cudaMalloc(&d_w, (size_t)n_words * 8));
		SB_CHECK(cudaMemcpy(d_pac, ref.pac.data(), (size_t)(L / 4 + 1), cudaMemcpyHostToDevice));
		k_pack_text<<<grid_for((u64)n_words), 256, 0, st>>>(d_pac, L, d_w, n_words);
		SB_CHECK(cudaGetLastError());
		SB_CHECK(cudaMalloc(&d_sym, (size_t)n_rows + 2));
		SB_CHECK(cudaMalloc(&d_sa_s, (size_t)n_sa * 8));
		SB_CHECK(cudaMalloc(&d_scalar, 16));
		SB_CHECK(cudaMalloc(&d_pos, cap * 8)); SB_CHECK(cudaMalloc(&d_pos2, cap * 8)); SB_CHECK(cudaMalloc(&d_key, cap * 8)); SB_CHECK(cudaMalloc(&d_key2, cap * 8));
		SB_CHECK(cudaMalloc(&d_unres, cap * 4)); SB_CHECK(cudaMalloc(&d_off, cap * 4 + 4)); SB_CHECK(cudaMalloc(&d_head, cap * 4));
		// working set of the refinement rounds (sized for the worst case: every element tied after round 0)
		SB_CHECK(cudaMalloc(&u_pos, cap * 8)); SB_CHECK(cudaMalloc(&u_pos2, cap * 8)); SB_CHECK(cudaMalloc(&u_key, cap * 8)); SB_CHECK(cudaMalloc(&u_key2, cap * 8));
		SB_CHECK(cudaMalloc(&u_slot, cap * 4)); SB_CHECK(cudaMalloc(&u_slot2, cap * 4)); SB_CHECK(cudaMalloc(&u_grp, cap * 4)); SB_CHECK(cudaMalloc(&u_grp2, cap * 4));
		SB_CHECK(cudaMalloc(&u_idx, cap * 4)); SB_CHECK(cudaMalloc(&u_idx2, cap * 4)); SB_CHECK(cudaMalloc(&u_idx3, cap * 4));
		{
			size_t a = 0, b = 0, c = 0, d = 0;
			cub::DeviceRadixSort::SortPairs(nullptr, a, d_key, d_key2, d_pos, d_pos2, (i64)cap, 0, 64, st);
			cub::DeviceRadixSort::SortPairs(nullptr, b, u_key, u_key2, u_idx, u_idx2, (i64)cap, 0, 64, st);
			cub::DeviceScan::ExclusiveSum(nullptr, c, d_unres, d_off, (i64)cap, st);
			cub::DeviceScan::InclusiveScan(nullptr, d, d_head, d_head, MaxOp(), (i64)cap, st);
			tmp_bytes = a > b ? a : b; tmp_bytes = tmp_bytes > c ? tmp_bytes : c; tmp_bytes = tmp_bytes > d ? tmp_bytes : d; tmp_bytes += 1024;
			SB_CHECK(cudaMalloc(&d_tmp, tmp_bytes));
		}
		d_csa = d_pos2; // the sorted positions of round 0 double as the chunk's suffix array
		SB_CHECK(cudaMemset(d_scalar, 0, 16));
		// row 0 is the '$' suffix: its BWT symbol is T[n-1]; rows >= 1 come from the chunks
		{
			const u64 one = (u64)n;
			SB_CHECK(cudaMemcpy(d_pos2, &one, 8, cudaMemcpyHostToDevice));
			k_emit<<<1, 32, 0, st>>>(d_w, d_pos2, 1, 0, d_sym, d_sa_s, d_scalar + 1);
		}
		// ---- chunks in ascending prefix order; a chunk that does not fit is split
		std::vector<std::pair<u64, u64>> todo; // [lo, hi) over the 32-bit prefix space, processed back to front
		{
			u64 n_chunks = ((u64)n + cap * 7 / 10 - 1) / (cap * 7 / 10);
			if (n_chunks < 1) n_chunks = 1;
			const u64 space = (u64)1 << 32;
			for (u64 c = n_chunks; c-- > 0;) todo.push_back({ space * c / n_chunks, space * (c + 1) / n_chunks });
		}
		while (!todo.empty()) {
			const u64 lo = todo.back().first, hi = todo.back().second;
			SB_CHECK(cudaMemsetAsync(d_scalar, 0, 8, st));
			k_select<<<148 * 8, 256, 0, st>>>(d_w, n, lo, hi, d_pos, d_scalar, cap);
			unsigned long long m_ = 0;
			SB_CHECK(cudaMemcpyAsync(&m_, d_scalar, 8, cudaMemcpyDeviceToHost, st));
			SB_CHECK(cudaStreamSynchronize(st));
			u64 m = (u64)m_;
			if (m > cap) { // too many suffixes share these prefixes: halve the range
				if (hi - lo < 2) { fprintf(stderr, "[bwamem-b200] index build: more than %llu suffixes share one 16-base prefix\n", (unsigned long long)cap); ok = false; goto done; }
				todo.pop_back();
				const u64 mid = lo + (hi - lo) / 2;
				todo.push_back({ mid, hi }); todo.push_back({ lo, mid });
				continue;
			}
			todo.pop_back();
			if (m == 0) continue;
			// round 0
			k_keys0<<<grid_for(m), 256, 0, st>>>(d_w, n, d_pos, d_key, m);
			{ size_t tb = tmp_bytes; SB_CHECK(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_key, d_key2, d_pos, d_pos2, (i64)m, 0, 64, st)); }
			k_mark<<<grid_for(m), 256, 0, st>>>(d_key2, nullptr, m, d_unres, d_head);
			{ size_t tb = tmp_bytes; SB_CHECK(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_unres, d_off, (i64)m, st)); }
			{ size_t tb = tmp_bytes; SB_CHECK(cub::DeviceScan::InclusiveScan(d_tmp, tb, d_head, d_head, MaxOp(), (i64)m, st)); }
			k_compact<<<grid_for(m), 256, 0, st>>>(d_unres, d_off, d_head, nullptr, d_pos2, m, u_slot, u_pos, u_grp);
			u32 last_off = 0, last_flag = 0;
			SB_CHECK(cudaMemcpyAsync(&last_off, d_off + (m - 1), 4, cudaMemcpyDeviceToHost, st));
			SB_CHECK(cudaMemcpyAsync(&last_flag, d_unres + (m - 1), 4, cudaMemcpyDeviceToHost, st));
			SB_CHECK(cudaStreamSynchronize(st));
			u64 mu = (u64)last_off + last_flag;
			for (i64 depth = 28; mu > 0; depth += 28) {
				if (depth > n + 28) { fprintf(stderr, "[bwamem-b200] index build: refinement did not terminate\n"); ok = false; goto done; }
				k_keys_round<<<grid_for(mu), 256, 0, st>>>(d_w, n, u_pos, depth, u_key, u_idx, mu);
				// stable LSD composition: by key, then by group
				{ size_t tb = tmp_bytes; SB_CHECK(cub::DeviceRadixSort::SortPairs(d_tmp, tb, u_key, u_key2, u_idx, u_idx2, (i64)mu, 0, 64, st)); }
				k_gather_u32<<<grid_for(mu), 256, 0, st>>>(u_grp, u_idx2, u_grp2, mu);
				{ size_t tb = tmp_bytes; SB_CHECK(cub::DeviceRadixSort::SortPairs(d_tmp, tb, u_grp2, u_grp, u_idx2, u_idx3, (i64)mu, 0, 32, st)); }
				// u_grp now holds the groups in sorted order (= the old array, groups are contiguous); u_idx3 the permutation
				k_apply<<<grid_for(mu), 256, 0, st>>>(u_idx3, u_pos, u_key, u_slot, u_pos2, u_key2, d_csa, mu);
				k_mark<<<grid_for(mu), 256, 0, st>>>(u_key2, u_grp, mu, d_unres, d_head);
				{ size_t tb = tmp_bytes; SB_CHECK(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_unres, d_off, (i64)mu, st)); }
				{ size_t tb = tmp_bytes; SB_CHECK(cub::DeviceScan::InclusiveScan(d_tmp, tb, d_head, d_head, MaxOp(), (i64)mu, st)); }
				k_compact<<<grid_for(mu), 256, 0, st>>>(d_unres, d_off, d_head, u_slot, u_pos2, mu, u_slot2, u_pos, u_grp2);
				SB_CHECK(cudaMemcpyAsync(&last_off, d_off + (mu - 1), 4, cudaMemcpyDeviceToHost, st));
				SB_CHECK(cudaMemcpyAsync(&last_flag, d_unres + (mu - 1), 4, cudaMemcpyDeviceToHost, st));
				SB_CHECK(cudaStreamSynchronize(st));
				mu = (u64)last_off + last_flag;
				{ u32* t = u_slot; u_slot = u_slot2; u_slot2 = t; }
				{ u32* t = u_grp; u_grp = u_grp2; u_grp2 = t; }
			}
			k_emit<<<grid_for(m), 256, 0, st>>>(d_w, d_csa, m, row0, d_sym, d_sa_s, d_scalar + 1);
			SB_CHECK(cudaGetLastError());
			row0 += m;
		}
		if (row0 != n_rows) { fprintf(stderr, "[bwamem-b200] index build: %llu rows produced, %llu expected\n", (unsigned long long)row0, (unsigned long long)n_rows); ok = false; goto done; }
		{ unsigned long long p = 0; SB_CHECK(cudaMemcpy(&p, d_scalar + 1, 8, cudaMemcpyDeviceToHost)); primary = (u64)p; }
		// ---- release the sort buffers, assemble the interleaved BWT
		cudaFree(d_pos); cudaFree(d_pos2); cudaFree(d_key); cudaFree(d_key2); cudaFree(u_pos); cudaFree(u_pos2); cudaFree(u_key); cudaFree(u_key2);
		d_pos = d_pos2 = d_key = d_key2 = u_pos = u_pos2 = u_key = u_key2 = nullptr;
		SB_CHECK(cudaMalloc(&c0, (n_blocks + 1) * 4)); SB_CHECK(cudaMalloc(&c1, (n_blocks + 1) * 4)); SB_CHECK(cudaMalloc(&c2, (n_blocks + 1) * 4)); SB_CHECK(cudaMalloc(&c3, (n_blocks + 1) * 4));
		SB_CHECK(cudaMalloc(&o0, (n_blocks + 1) * 8)); SB_CHECK(cudaMalloc(&o1, (n_blocks + 1) * 8)); SB_CHECK(cudaMalloc(&o2, (n_blocks + 1) * 8)); SB_CHECK(cudaMalloc(&o3, (n_blocks + 1) * 8));
		SB_CHECK(cudaMalloc(&d_bwt, bwt_size * 4));
		SB_CHECK(cudaMemsetAsync(d_bwt, 0, bwt_size * 4, st));
		SB_CHECK(cudaMemsetAsync(c0 + n_blocks, 0, 4, st)); SB_CHECK(cudaMemsetAsync(c1 + n_blocks, 0, 4, st)); SB_CHECK(cudaMemsetAsync(c2 + n_blocks, 0, 4, st)); SB_CHECK(cudaMemsetAsync(c3 + n_blocks, 0, 4, st));
		k_block_counts<<<grid_for(n_blocks), 256, 0, st>>>(d_sym, (u64)n, primary, n_blocks, c0, c1, c2, c3);
		{
			u32* cs[4] = { c0, c1, c2, c3 }; u64* os[4] = { o0, o1, o2, o3 };
			for (int t = 0; t < 4; ++t) {
				cub::TransformInputIterator<u64, U32ToU64b, const u32*> it(cs[t], U32ToU64b());
				size_t need = 0;
				cub::DeviceScan::ExclusiveSum(nullptr, need, it, os[t], (i64)(n_blocks + 1), st);
				if (need > tmp_bytes) { cudaFree(d_tmp); d_tmp = nullptr; SB_CHECK(cudaMalloc(&d_tmp, need)); tmp_bytes = need; }
				SB_CHECK(cub::DeviceScan::ExclusiveSum(d_tmp, need, it, os[t], (i64)(n_blocks + 1), st));
				SB_CHECK(cudaMemcpyAsync(&totals[t], os[t] + n_blocks, 8, cudaMemcpyDeviceToHost, st));
			}
		}
		k_write_blocks<<<grid_for(n_blocks), 256, 0, st>>>(d_sym, (u64)n, primary, n_blocks, o0, o1, o2, o3, d_bwt);
		SB_CHECK(cudaStreamSynchronize(st));
		SB_CHECK(cudaGetLastError());
		// ---- back to the host model
		fm.seq_len = (u64)n; fm.primary = primary; fm.bwt_size = bwt_size; fm.sa_intv = 32; fm.n_sa = n_sa;
		fm.L2[0] = 0; for (int t = 0; t < 4; ++t) fm.L2[t + 1] = fm.L2[t] + totals[t];
		fm.bwt.assign((size_t)bwt_size, 0);
		SB_CHECK(cudaMemcpy(fm.bwt.data(), d_bwt, bwt_size * 4, cudaMemcpyDeviceToHost));
		{ // the closing occurrence block sits right behind the last symbol word
			const u64 k = (u64)n, rem = k & 127;
			const u64 pos = ((k >> 7) << 4) + (rem ? 8 + (rem + 15) / 16 : 0);
			memcpy(fm.bwt.data() + pos, totals, 32);
		}
		fm.sa.assign((size_t)n_sa, 0);
		SB_CHECK(cudaMemcpy(fm.sa.data(), d_sa_s, n_sa * 8, cudaMemcpyDeviceToHost));
		fm.sa[0] = (u64)-1;
	}
done:
	cudaFree(d_pac); cudaFree(d_w); cudaFree(d_sym); cudaFree(d_sa_s); cudaFree(d_scalar);
	cudaFree(d_pos); cudaFree(d_pos2); cudaFree(d_key); cudaFree(d_key2); cudaFree(u_pos); cudaFree(u_pos2); cudaFree(u_key); cudaFree(u_key2);
	cudaFree(u_slot); cudaFree(u_slot2); cudaFree(u_grp); cudaFree(u_grp2); cudaFree(u_idx); cudaFree(u_idx2); cudaFree(u_idx3);
	cudaFree(d_unres); cudaFree(d_off); cudaFree(d_head); cudaFree(d_tmp);
	cudaFree(c0); cudaFree(c1); cudaFree(c2); cudaFree(c3); cudaFree(o0); cudaFree(o1); cudaFree(o2); cudaFree(o3); cudaFree(d_bwt);
	return ok;
}

} // namespace b200
