// dpx_peak.cu -- micro-benchmark: sustained issue rate of the integer min/max ("DPX") instructions the DP kernels are
// built from, per SM and for the whole device.  This is the denominator of the DP stages' roofline:
//   cells/s peak = instruction rate x cells per instruction / DPX-class instructions per cell of the recurrence.
// Measured: VIMNMX3 (s32 three-input max), VIADDMNMX.RELU (s32 add+max+relu), their packed s16x2 forms
// (__vimax3_s16x2, __viaddmax_s16x2_relu: two cells per instruction).
// Every thread runs ILP independent dependency chains (sweep 1..8); the grid fills every SM with 2048 threads.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dpx_peak tools/dpx_peak.cu ; run -> one JSON line.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

enum { OP_VIMAX3_S32, OP_VIADDMAX_S32_RELU, OP_VIMAX3_S16X2, OP_VIADDMAX_S16X2_RELU, OP_N };
static const char* OP_NAME[OP_N] = { "vimax3_s32", "viaddmax_s32_relu", "vimax3_s16x2", "viaddmax_s16x2_relu" };

template <int OP>
__device__ __forceinline__ unsigned step(unsigned a, unsigned b, unsigned c)
{
	if (OP == OP_VIMAX3_S32) return (unsigned)__vimax3_s32((int)a, (int)b, (int)c);
	if (OP == OP_VIADDMAX_S32_RELU) return (unsigned)__viaddmax_s32_relu((int)a, (int)b, (int)c);
	if (OP == OP_VIMAX3_S16X2) return __vimax3_s16x2(a, b, c);
	return __viaddmax_s16x2_relu(a, b, c);
}

template <int OP, int ILP>
__global__ void __launch_bounds__(256) k_chain(int iters, unsigned seed, unsigned* sink)
{
	unsigned v[ILP];
	const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
	for (int u = 0; u < ILP; ++u) v[u] = seed * (t + 1) + u * 0x9e3779b9u;
	unsigned b = seed ^ t, c = seed + 7u * t; // loop-invariant operands the compiler cannot fold (they depend on runtime values)
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int k = 0; k < 16; ++k) {
#pragma unroll
			for (int u = 0; u < ILP; ++u) v[u] = step<OP>(v[u], b, c);
			b += 0x01010101u; // one extra ALU op per 16*ILP measured ops keeps the chain from collapsing to a fixed point the compiler could hoist
		}
	}
	unsigned acc = 0;
#pragma unroll
	for (int u = 0; u < ILP; ++u) acc ^= v[u];
	if (acc == 0x12345u) *sink = acc;
}

template <int OP, int ILP>
static double run(int sms, unsigned* sink)
{
	const int iters = 2048;
	dim3 grid(sms * 8), block(256);
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	k_chain<OP, ILP><<<grid, block>>>(iters, 12345u, sink);
	cudaDeviceSynchronize();
	float best = 1e30f;
	for (int rep = 0; rep < 5; ++rep) {
		cudaEventRecord(a);
		k_chain<OP, ILP><<<grid, block>>>(iters, 12345u + rep, sink);
		cudaEventRecord(b); cudaEventSynchronize(b);
		float ms = 0; cudaEventElapsedTime(&ms, a, b);
		if (ms < best) best = ms;
	}
	cudaEventDestroy(a); cudaEventDestroy(b);
	const double ops = (double)grid.x * 256.0 * iters * 16.0 * ILP; // thread-level instructions
	return ops / (best * 1e-3); // thread-instructions per second, whole device
}

template <int OP>
static double best_of_ilp(int sms, unsigned* sink, int* best_ilp)
{
	double r[4] = { run<OP, 1>(sms, sink), run<OP, 2>(sms, sink), run<OP, 4>(sms, sink), run<OP, 8>(sms, sink) };
	const int il[4] = { 1, 2, 4, 8 };
	double best = 0;
	for (int k = 0; k < 4; ++k) { fprintf(stderr, "%-22s ILP %d: %.2f T thread-instr/s\n", OP_NAME[OP], il[k], r[k] / 1e12); if (r[k] > best) { best = r[k]; *best_ilp = il[k]; } }
	return best;
}

int main()
{
	int dev = 0, sms = 0, khz = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
	unsigned* sink = nullptr;
	if (cudaMalloc(&sink, 4) != cudaSuccess) { fprintf(stderr, "alloc failed\n"); return 1; }
	double rate[OP_N]; int ilp[OP_N];
	rate[0] = best_of_ilp<OP_VIMAX3_S32>(sms, sink, &ilp[0]);
	rate[1] = best_of_ilp<OP_VIADDMAX_S32_RELU>(sms, sink, &ilp[1]);
	rate[2] = best_of_ilp<OP_VIMAX3_S16X2>(sms, sink, &ilp[2]);
	rate[3] = best_of_ilp<OP_VIADDMAX_S16X2_RELU>(sms, sink, &ilp[3]);
	printf("{\"sms\": %d, \"sm_clock_mhz_nominal\": %.0f, \"unit\": \"thread-level instructions per second, whole device (best ILP of 1/2/4/8, 2048 threads/SM, best of 5)\"", sms, khz / 1e3);
	for (int k = 0; k < OP_N; ++k) printf(", \"%s\": %.4e, \"%s_per_sm_clk\": %.1f, \"%s_ilp\": %d", OP_NAME[k], rate[k], OP_NAME[k], rate[k] / sms / (khz * 1e3), OP_NAME[k], ilp[k]);
	printf("}\n");
	return 0;
}
