// pestat_host.h -- the one batch-global step of the paired-end path, on the host: insert-size statistics
// (mem_pestat of un-vendored lh3/bwa) from the per-pair candidates the device produced, and
// the per-orientation table of the pairing term .721*log(2*erfc(|z|/sqrt2))*a for every admissible distance, so that
// the device never evaluates libm (hazard H5: results must be glibc-identical and then truncated to int).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "types.cuh"

namespace b200 {

inline void host_mem_pestat(const MemOpt& opt, const int64_t* cand, int64_t n_pairs, PeStat pes[4], bool presorted = false)
{
	std::vector<uint64_t> isize[4];
	memset(pes, 0, 4 * sizeof(PeStat));
	for (int64_t i = 0; i < n_pairs; ++i) {
		if (cand[i] < 0) continue;
		int dir = (int)(cand[i] >> 48);
		isize[dir].push_back((uint64_t)(cand[i] & ((1LL << 48) - 1)));
	}
	for (int d = 0; d < 4; ++d) {
		PeStat* r = &pes[d];
		std::vector<uint64_t>& q = isize[d];
		if (q.size() < 10) { r->failed = 1; continue; }
		if (!presorted) std::sort(q.begin(), q.end()); // integers: every correct sort yields the same array
		size_t n = q.size();
		int p25 = (int)q[(int)(.25 * n + .499)];
		int p75 = (int)q[(int)(.75 * n + .499)];
		r->low = (int)(p25 - 2.0 * (p75 - p25) + .499);
		if (r->low < 1) r->low = 1;
		r->high = (int)(p75 + 2.0 * (p75 - p25) + .499);
		size_t i; int x;
		for (i = 0, x = 0, r->avg = 0; i < n; ++i)
			if (q[i] >= (uint64_t)r->low && q[i] <= (uint64_t)r->high) { r->avg += q[i]; ++x; }
		r->avg /= x;
		for (i = 0, r->std = 0; i < n; ++i)
			if (q[i] >= (uint64_t)r->low && q[i] <= (uint64_t)r->high) r->std += (q[i] - r->avg) * (q[i] - r->avg);
		r->std = sqrt(r->std / x);
		r->low = (int)(p25 - 3.0 * (p75 - p25) + .499);
		r->high = (int)(p75 + 3.0 * (p75 - p25) + .499);
		if (r->low > r->avg - 4.0 * r->std) r->low = (int)(r->avg - 4.0 * r->std + .499);
		if (r->high < r->avg + 4.0 * r->std) r->high = (int)(r->avg + 4.0 * r->std + .499);
		if (r->low < 1) r->low = 1;
	}
	size_t mx = 0;
	for (int d = 0; d < 4; ++d) mx = mx > isize[d].size() ? mx : isize[d].size();
	for (int d = 0; d < 4; ++d)
		if (pes[d].failed == 0 && isize[d].size() < mx * 0.05) pes[d].failed = 1;
}

// returns false when an orientation's [low,high] range is too wide to tabulate
inline bool host_pair_table(const MemOpt& opt, const PeStat pes[4], std::vector<double>& tab, int64_t off[4])
{
	tab.clear();
	for (int d = 0; d < 4; ++d) {
		off[d] = (int64_t)tab.size();
		if (pes[d].failed) continue;
		int64_t lo = pes[d].low, hi = pes[d].high;
		if (hi < lo) continue;
		if (hi - lo > (1 << 24)) return false;
		for (int64_t dist = lo; dist <= hi; ++dist) {
			double ns = (dist - pes[d].avg) / pes[d].std;
			tab.push_back(.721 * log(2. * erfc(fabs(ns) * M_SQRT1_2)) * opt.a);
		}
	}
	return true;
}

} // namespace b200
