// host_scan.cpp -- read boundaries of a createAlignments request (host side of the boundary,
// /root/reference/src/main/c/jnibwa.c:199-212: the reference walks the blob with strlen).
//
// The wire format (u32 count, then NUL-terminated reads; BwaMemAligner.java:198-209) carries no byte length, so the
// boundaries have to be found by a scan that must not look past the last NUL.  With AVX2 the scan compares 32 bytes
// per step using ALIGNED loads only: an aligned 32-byte block that contains a valid byte lies inside that byte's page,
// and the scan returns as soon as it has seen the n-th NUL, so nothing beyond the caller's buffer is ever touched
// (the same argument glibc's strlen relies on).  Compiled by the host compiler alone (no CUDA front end).
#include <atomic>
#include <cstdint>
#include <cstring>
#include <immintrin.h>

namespace b200 {

// off[0..n] = byte offsets of the reads (off[n] = total length incl. the last NUL); *max_len = longest read;
// *progress (optional) is advanced to the number of bytes known to be valid.
static void scan_reads_plain(const char* base, uint32_t n, uint64_t* off, int* max_len, std::atomic<uint64_t>* progress)
{
	const char* p = base; int mx = 0;
	for (uint32_t i = 0; i < n; ++i) {
		size_t l = strlen(p);
		off[i] = (uint64_t)(p - base);
		if ((int)l > mx) mx = (int)l;
		p += l + 1;
		if (progress && (i & 1023) == 1023) progress->store((uint64_t)(p - base), std::memory_order_release);
	}
	off[n] = (uint64_t)(p - base);
	*max_len = mx;
}

__attribute__((target("avx2")))
static void scan_reads_avx2(const char* base, uint32_t n, uint64_t* off, int* max_len, std::atomic<uint64_t>* progress)
{
	uint32_t i = 0; uint64_t start = 0; int mx = 0;
	off[0] = 0;
	if (n == 0) { *max_len = 0; return; }
	const uintptr_t b0 = (uintptr_t)base;
	uintptr_t a = b0 & ~(uintptr_t)31;
	const __m256i z = _mm256_setzero_si256();
	uint32_t mask = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_load_si256((const __m256i*)a), z));
	mask &= ~0u << (unsigned)(b0 - a); // bytes in front of the buffer
	uint32_t tick = 0;
	for (;;) {
		while (mask) {
			const uint64_t pos = (uint64_t)(a - b0) + (uint64_t)__builtin_ctz(mask); // a NUL
			const int len = (int)(pos - start);
			if (len > mx) mx = len;
			start = pos + 1;
			off[++i] = start;
			if (i == n) { *max_len = mx; if (progress) progress->store(start, std::memory_order_release); return; }
			mask &= mask - 1;
		}
		a += 32;
		mask = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_load_si256((const __m256i*)a), z));
		if (progress && (++tick & 4095) == 0) progress->store(start, std::memory_order_release);
	}
}

void scan_reads(const char* base, uint32_t n, uint64_t* off, int* max_len, std::atomic<uint64_t>* progress)
{
	static const bool avx2 = __builtin_cpu_supports("avx2");
	if (avx2) scan_reads_avx2(base, n, off, max_len, progress); else scan_reads_plain(base, n, off, max_len, progress);
}

} // namespace b200

// test hook (tests/test_host_scan.py): the boundaries of n reads starting at blob; returns the total length
extern "C" uint64_t b200_scanReads(const char* blob, uint32_t n, uint64_t* off, int* max_len, int force_plain)
{
	if (force_plain) b200::scan_reads_plain(blob, n, off, max_len, nullptr); else b200::scan_reads(blob, n, off, max_len, nullptr);
	return off[n];
}
