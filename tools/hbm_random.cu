// hbm_random.cu -- micro-benchmark: throughput of INDEPENDENT random 64-byte reads (one FM-index occurrence block =
// two 32-byte sectors of one 64-byte line) from a buffer far larger than L2, on the current device.
// This is the roof an FM-index walk could reach if it had no dependent-load latency at all: every thread issues its
// next block read without waiting for the previous one.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
// -o tools/hbm_random tools/hbm_random.cu ; run: tools/hbm_random [GiB] -> one JSON line.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t z) { z += 0x9E3779B97F4A7C15ULL; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }

// Q = 16-byte quarters per read: 2 = one 32-byte sector, 4 = a 64-byte block, 8 = a 128-byte line
template <int ILP, int Q = 4>
__global__ void k_random64(const uint4* buf, uint64_t n_blocks, int iters, uint32_t* sink)
{
	const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t acc = 0;
	for (int it = 0; it < iters; ++it) {
		uint4 v[ILP][Q];
#pragma unroll
		for (int u = 0; u < ILP; ++u) {
			const uint64_t b = mix(tid * 1315423911ULL + (uint64_t)it * ILP + u) % n_blocks;
			const uint4* p = buf + b * Q;
#pragma unroll
			for (int q = 0; q < Q; ++q) v[u][q] = __ldg(p + q);
		}
#pragma unroll
		for (int u = 0; u < ILP; ++u)
#pragma unroll
			for (int q = 0; q < Q; ++q) acc ^= v[u][q].x ^ v[u][q].w;
	}
	if (acc == 0x12345678u) *sink = acc;
}

template <int ILP, int Q = 4>
static double run(const uint4* buf, uint64_t n_blocks, uint32_t* sink, int blocks_per_sm, int sms)
{
	const int iters = 256 / ILP * 4;
	dim3 grid(sms * blocks_per_sm), block(256);
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	k_random64<ILP, Q><<<grid, block>>>(buf, n_blocks, iters, sink);
	cudaDeviceSynchronize();
	float best = 1e30f;
	for (int rep = 0; rep < 5; ++rep) {
		cudaEventRecord(a);
		k_random64<ILP, Q><<<grid, block>>>(buf, n_blocks, iters, sink);
		cudaEventRecord(b); cudaEventSynchronize(b);
		float ms = 0; cudaEventElapsedTime(&ms, a, b);
		if (ms < best) best = ms;
	}
	const double bytes = (double)grid.x * 256 * iters * ILP * (16.0 * Q);
	return bytes / (best * 1e6); // GB/s
}

int main(int argc, char** argv)
{
	const double gib = argc > 1 ? atof(argv[1]) : 4.0;
	size_t l2g = 0;
	if (argc > 2) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[2])); // L2 fetch granularity hint (bytes)
	cudaDeviceGetLimit(&l2g, cudaLimitMaxL2FetchGranularity);
	fprintf(stderr, "L2 fetch granularity: %zu bytes\n", l2g);
	const uint64_t bytes = (uint64_t)(gib * (1ULL << 30));
	uint4* buf = nullptr; uint32_t* sink = nullptr;
	if (cudaMalloc(&buf, bytes) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) { fprintf(stderr, "alloc failed\n"); return 1; }
	cudaMemset(buf, 1, bytes);
	int dev = 0, sms = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const uint64_t n_blocks = bytes / 64;
	double best = 0; int best_ilp = 0, best_bps = 0;
	for (int bps : { 4, 8 }) {
		double r1 = run<1>(buf, n_blocks, sink, bps, sms), r2 = run<2>(buf, n_blocks, sink, bps, sms), r4 = run<4>(buf, n_blocks, sink, bps, sms), r8 = run<8>(buf, n_blocks, sink, bps, sms);
		double rs[4] = { r1, r2, r4, r8 }; int il[4] = { 1, 2, 4, 8 };
		for (int k = 0; k < 4; ++k) { fprintf(stderr, "blocks/SM %d ILP %d: %.1f GB/s\n", bps, il[k], rs[k]); if (rs[k] > best) { best = rs[k]; best_ilp = il[k]; best_bps = bps; } }
	}
	double b32 = 0, b128 = 0; // other granularities, 2048 threads/SM: the 32-byte sector and the 128-byte line
	{ const double a = run<4, 2>(buf, bytes / 32, sink, 8, sms), b = run<8, 2>(buf, bytes / 32, sink, 8, sms); b32 = a > b ? a : b; }
	{ const double a = run<2, 8>(buf, bytes / 128, sink, 8, sms), b = run<4, 8>(buf, bytes / 128, sink, 8, sms); b128 = a > b ? a : b; }
	fprintf(stderr, "32 B: %.1f GB/s (%.2f G reads/s); 64 B: %.1f GB/s (%.2f G reads/s); 128 B: %.1f GB/s (%.2f G reads/s)\n", b32, b32 / 32, best, best / 64, b128, b128 / 128);
	printf("{\"random_32B_read_gbs\": %.1f, \"random_128B_read_gbs\": %.1f, \"random_64B_read_gbs\": %.1f, \"buffer_gib\": %.2f, \"threads_per_sm\": %d, \"ilp\": %d, \"sms\": %d}\n", b32, b128, best, gib, best_bps * 256, best_ilp, sms);
	return 0;
}
