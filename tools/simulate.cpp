// simulate.cpp -- seeded synthetic genome generator and wgsim-style read simulator
// (test / benchmark tooling; all data in this repo is synthetic and offline).
// Workload definitions follow SURVEY.md 8d: uniform ACGT contigs with injected
// interspersed repeat families, tandem repeats and segmental duplications; reads
// with substitutions, indels (geometric extension), uniform base errors, optional
// chimeric reads and over-sampling of repeat copies.  Output reads are written
// directly in the wire format of BwaMemAligner.java:198-209 (u32 count, then
// NUL-terminated ASCII reads; mates adjacent).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct Rng { // splitmix64 seeding + xoshiro256**
	uint64_t s[4];
	static uint64_t sm(uint64_t& x) { uint64_t z = (x += 0x9E3779B97F4A7C15ULL); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
	explicit Rng(uint64_t seed) { for (auto& v : s) v = sm(seed); }
	static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
	uint64_t next() { uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17; s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45); return r; }
	double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
	uint64_t below(uint64_t n) { return n ? next() % n : 0; }
	double normal() { double u1 = uni(), u2 = uni(); if (u1 < 1e-300) u1 = 1e-300; return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2); }
};

const char kBase[4] = {'A', 'C', 'G', 'T'};
inline char comp(char c) { switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; } }

void mutate_copy(Rng& r, const std::string& src, double div, std::string& out)
{
	out = src;
	for (auto& c : out) if (r.uni() < div) { char n; do n = kBase[r.below(4)]; while (n == c); c = n; }
}

} // namespace

extern "C" {

// Fills `out` (sum of lens bytes) with ACGT.  repeat_frac: fraction of bases covered by interspersed repeat
// families (0 = none).  Writes up to rep_cap (start,end) global intervals of repeat copies into rep_iv.
int64_t sim_genome(uint64_t seed, int n_contigs, const int64_t* lens, double repeat_frac, double tandem_frac, int n_segdup,
                   char* out, int64_t* rep_iv, int64_t rep_cap, int64_t* n_rep_out)
{
	Rng r(seed);
	int64_t total = 0;
	for (int i = 0; i < n_contigs; ++i) total += lens[i];
	for (int64_t i = 0; i < total; i += 1) out[i] = kBase[r.next() >> 62];
	int64_t n_rep = 0;
	auto record = [&](int64_t s, int64_t e) { if (rep_iv && n_rep < rep_cap) { rep_iv[2 * n_rep] = s; rep_iv[2 * n_rep + 1] = e; } ++n_rep; };
	auto place = [&](const std::string& el) -> bool { // keep a copy inside one contig
		if ((int64_t)el.size() + 2 >= total) return false;
		for (int tries = 0; tries < 16; ++tries) {
			int64_t p = (int64_t)r.below((uint64_t)(total - (int64_t)el.size())), off = 0;
			bool ok = false;
			for (int c = 0; c < n_contigs; ++c) { if (p >= off && p + (int64_t)el.size() <= off + lens[c]) { ok = true; break; } off += lens[c]; }
			if (!ok) continue;
			bool rev = r.next() & 1;
			if (!rev) memcpy(out + p, el.data(), el.size());
			else for (size_t k = 0; k < el.size(); ++k) out[p + (int64_t)k] = comp(el[el.size() - 1 - k]);
			record(p, p + (int64_t)el.size());
			return true;
		}
		return false;
	};
	if (repeat_frac > 0) {
		// two families: short (300 bp, 5-15 % divergence) takes 70 % of the repeat budget, long (6 kbp, 1-10 %) the rest
		int64_t budget_short = (int64_t)(total * repeat_frac * 0.7), budget_long = (int64_t)(total * repeat_frac * 0.3);
		int n_fam_short = (int)std::max<int64_t>(1, std::min<int64_t>(20, budget_short / (300 * 200)));
		for (int f = 0; f < n_fam_short; ++f) {
			std::string cons(300, 'A'), cp;
			for (auto& c : cons) c = kBase[r.below(4)];
			int64_t copies = budget_short / n_fam_short / 300;
			for (int64_t k = 0; k < copies; ++k) { mutate_copy(r, cons, 0.05 + 0.10 * r.uni(), cp); place(cp); }
		}
		if (total > 200000) {
			int n_fam_long = (int)std::max<int64_t>(1, std::min<int64_t>(5, budget_long / (6000 * 20)));
			for (int f = 0; f < n_fam_long; ++f) {
				std::string cons(6000, 'A'), cp;
				for (auto& c : cons) c = kBase[r.below(4)];
				int64_t copies = budget_long / n_fam_long / 6000;
				for (int64_t k = 0; k < copies; ++k) { mutate_copy(r, cons, 0.01 + 0.09 * r.uni(), cp); place(cp); }
			}
		}
	}
	if (tandem_frac > 0) {
		int64_t budget = (int64_t)(total * tandem_frac);
		while (budget > 0) {
			int period = 2 + (int)r.below(49), n_unit = 4 + (int)r.below(40);
			std::string unit(period, 'A'), el;
			for (auto& c : unit) c = kBase[r.below(4)];
			for (int k = 0; k < n_unit; ++k) el += unit;
			std::string cp; mutate_copy(r, el, 0.02 * r.uni(), cp);
			place(cp);
			budget -= (int64_t)el.size();
		}
	}
	for (int d = 0; d < n_segdup; ++d) {
		int64_t len = 10000 + (int64_t)r.below(90001);
		if (len * 4 > total) len = total / 8;
		if (len < 500) break;
		int64_t src = (int64_t)r.below((uint64_t)(total - len));
		std::string el(out + src, out + src + len), cp;
		mutate_copy(r, el, 0.005 + 0.015 * r.uni(), cp);
		place(cp);
	}
	if (n_rep_out) *n_rep_out = n_rep;
	return total;
}

// Simulates n_units reads (single-end) or pairs (paired, FR orientation) and writes the wire blob.
// Returns the blob size in bytes, or -1 if out_cap is too small.  truth (optional): per read, the 0-based
// global forward-strand start of its first segment, negated-minus-1 when the read is reverse-complemented.
int64_t sim_reads(uint64_t seed, const char* genome, int n_contigs, const int64_t* lens, int64_t n_units, int read_len, int paired,
                  double ins_mean, double ins_std, double mut_rate, double indel_frac, double indel_ext, double err_rate,
                  double chimeric_frac, double repeat_read_frac, const int64_t* rep_iv, int64_t n_rep, double n_rate,
                  char* out, int64_t out_cap, int64_t* truth)
{
	Rng r(seed);
	int64_t total = 0;
	std::vector<int64_t> offs(n_contigs + 1, 0);
	for (int i = 0; i < n_contigs; ++i) { offs[i + 1] = offs[i] + lens[i]; }
	total = offs[n_contigs];
	int64_t n_reads = paired ? 2 * n_units : n_units, p = 4, ti = 0;
	if (out_cap < 4) return -1;
	{ uint32_t n32 = (uint32_t)n_reads; memcpy(out, &n32, 4); }
	auto contig_of = [&](int64_t pos) { int c = 0; while (c + 1 < n_contigs && pos >= offs[c + 1]) ++c; return c; };
	// a fragment: `flen` bases of the haplotype starting at forward position `start`, with mutations applied
	auto fragment = [&](int64_t start, int64_t flen, std::string& frag) {
		frag.clear();
		int c = contig_of(start);
		int64_t end = offs[c + 1], q = start;
		while ((int64_t)frag.size() < flen && q < end) {
			if (r.uni() < mut_rate) {
				if (r.uni() < indel_frac) {
					int l = 1; while (r.uni() < indel_ext && l < 30) ++l;
					if (r.next() & 1) { for (int k = 0; k < l; ++k) frag.push_back(kBase[r.below(4)]); } // insertion
					else q += l; // deletion
					continue;
				} else { char n; do n = kBase[r.below(4)]; while (n == genome[q]); frag.push_back(n); ++q; continue; }
			}
			frag.push_back(genome[q]); ++q;
		}
		while ((int64_t)frag.size() < flen) frag.push_back(kBase[r.below(4)]); // ran off the contig end
		frag.resize((size_t)flen);
	};
	int64_t longest = 0; for (int i = 0; i < n_contigs; ++i) longest = std::max<int64_t>(longest, lens[i]);
	auto pick_start = [&](int64_t span) -> int64_t {
		if (span > longest) span = longest; // (a fragment longer than every contig: it runs off the end and is padded)
		for (;;) {
			int64_t s;
			if (n_rep > 0 && rep_iv && r.uni() < repeat_read_frac) {
				int64_t k = (int64_t)r.below((uint64_t)n_rep);
				int64_t rs = rep_iv[2 * k], re = rep_iv[2 * k + 1];
				s = rs - span / 2 + (int64_t)r.below((uint64_t)std::max<int64_t>(1, re - rs));
				if (s < 0) s = 0;
			} else s = (int64_t)r.below((uint64_t)std::max<int64_t>(1, total - span));
			if (s + span > total) continue;
			int c = contig_of(s);
			if (s + span <= offs[c + 1]) return s;
		}
	};
	auto emit = [&](std::string& rd) -> bool {
		for (auto& ch : rd) {
			if (err_rate > 0 && r.uni() < err_rate) { char n; do n = kBase[r.below(4)]; while (n == ch); ch = n; }
			if (n_rate > 0 && r.uni() < n_rate) ch = 'N';
		}
		if (p + (int64_t)rd.size() + 1 > out_cap) return false;
		memcpy(out + p, rd.data(), rd.size()); p += (int64_t)rd.size(); out[p++] = 0;
		return true;
	};
	auto revcomp = [&](const std::string& s, std::string& o) { o.resize(s.size()); for (size_t k = 0; k < s.size(); ++k) o[k] = comp(s[s.size() - 1 - k]); };
	std::string frag, r1, r2, tmp, seg2;
	for (int64_t u = 0; u < n_units; ++u) {
		int64_t flen = read_len;
		if (paired) { flen = (int64_t)std::llround(ins_mean + ins_std * r.normal()); if (flen < read_len) flen = read_len; }
		int64_t start = pick_start(flen + 64);
		fragment(start, flen, frag);
		bool flip = r.next() & 1; // which strand the fragment is read from
		if (flip) { revcomp(frag, tmp); frag.swap(tmp); }
		r1.assign(frag, 0, (size_t)read_len);
		if (chimeric_frac > 0 && r.uni() < chimeric_frac) { // splice a second locus behind a breakpoint
			int bp = 30 + (int)r.below((uint64_t)std::max(1, read_len - 60));
			if (bp >= read_len) bp = read_len / 2; // (reads shorter than 60 bases)
			int64_t s2 = pick_start(read_len + 64);
			fragment(s2, read_len - bp, seg2);
			if (r.next() & 1) { revcomp(seg2, tmp); seg2.swap(tmp); }
			r1.replace((size_t)bp, std::string::npos, seg2);
		}
		if (truth) truth[ti++] = flip ? -(start + flen - read_len) - 1 : start;
		if (!emit(r1)) return -1;
		if (paired) {
			tmp.assign(frag, (size_t)(flen - read_len), (size_t)read_len);
			revcomp(tmp, r2);
			if (truth) truth[ti++] = flip ? start : -(start + flen - read_len) - 1;
			if (!emit(r2)) return -1;
		}
	}
	return p;
}

} // extern "C"
