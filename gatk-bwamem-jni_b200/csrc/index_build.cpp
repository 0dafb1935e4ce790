// index_build.cpp -- from-scratch builder of the bwa index files and of the
// BwaMemIndex ".img" image (host side; not on the timed hot path).
//
// Replaces, behind the same entry points:
//   jnibwa_createIndexFile(prefix, img)        /root/reference/src/main/c/jnibwa.c:126-152
//   Java_..._createReferenceIndex -> bwa_idx_build(fasta, prefix, algo, -1)
//                                              /root/reference/src/main/c/org_broadinstitute_hellbender_utils_bwa_BwaMemIndex.c:42-63
// File/ image layouts: SURVEY.md App. A.3 / A.5 (verified against
// /root/reference/src/test/resources/ref.fa.{amb,ann,bwt,pac,sa}).
//
// Text model: T = fwd(ref) || revcomp(ref), |T| = 2*l_pac, codes A0 C1 G2 T3,
// suffix array of T$ with '$' smallest; BWT stored without the '$' row,
// occurrence counts interleaved every 128 symbols, SA sampled every 32 rows.
// The suffix sorter is a bucketed (12-mer) multi-threaded comparison sort on
// the 2-bit packed text: adequate for genomes whose repeats have bounded LCP
// (synthetic genomes here); a device-side sorter for 3 Gbp is a "next" row.
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>

#include "bwa_image.h"

namespace b200 {

static const uint8_t kNt4[256] = {
	4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
	4,4,4,4,4,4,4,4,4,4,4,4,4,5,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
	4,0,4,1,4,4,4,2,4,4,4,4,4,4,4,4, 4,4,4,4,3,4,4,4,4,4,4,4,4,4,4,4,
	4,0,4,1,4,4,4,2,4,4,4,4,4,4,4,4, 4,4,4,4,3,4,4,4,4,4,4,4,4,4,4,4,
	4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
	4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
	4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
	4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4
};

// glibc srand48/lrand48 restated (bwa replaces ambiguous reference bases by lrand48()&3 after srand48(11)).
struct Rand48 {
	uint64_t x;
	explicit Rand48(uint32_t seed) : x(((uint64_t)seed << 16) | 0x330E) {}
	long next() { x = (x * 0x5DEECE66DULL + 0xB) & ((1ULL << 48) - 1); return (long)(x >> 17); }
};

static inline void pac_set(std::vector<uint8_t>& pac, int64_t l, int c) { pac[l >> 2] |= (uint8_t)(c << ((~l & 3) << 1)); }
static inline int pac_get(const uint8_t* pac, int64_t l) { return pac[l >> 2] >> ((~l & 3) << 1) & 3; }

// ---------------------------------------------------------------- reference -> RefSeqs

void RefSeqs::add_contig(const std::string& name, const std::string& anno, const char* bases, int64_t len, Rand48Holder& rng_)
{
	Rand48& rng = *reinterpret_cast<Rand48*>(rng_.state);
	Contig c;
	c.offset = l_pac; c.len = (int32_t)len; c.n_ambs = 0; c.gi = 0; c.is_alt = 0; c.name = name; c.anno = anno;
	if ((size_t)((l_pac + len) / 4 + 2) > pac.size()) pac.resize((size_t)((l_pac + len) / 4 + 2) * 2, 0);
	int lasts = 0;
	for (int64_t i = 0; i < len; ++i) {
		int ch = kNt4[(uint8_t)bases[i]];
		if (ch >= 4) { // ambiguous base: a hole record; the packed base is pseudo-random
			if (lasts == bases[i] && !holes.empty()) ++holes.back().len;
			else { Hole h; h.offset = l_pac; h.len = 1; h.amb = bases[i]; holes.push_back(h); ++c.n_ambs; }
		}
		lasts = bases[i];
		if (ch >= 4) ch = (int)(rng.next() & 3);
		pac_set(pac, l_pac, ch);
		++l_pac;
	}
	contigs.push_back(c);
}

// ---------------------------------------------------------------- suffix sorting

struct PackedText { // 32 bases per u64 word, first base in the top bits
	std::vector<uint64_t> w;
	int64_t n = 0;
	inline uint64_t key(int64_t i) const { // 32 bases starting at i, zero (A) padded
		int64_t wi = i >> 5; int sh = (int)(i & 31) << 1;
		uint64_t a = w[wi];
		if (sh == 0) return a;
		return (a << sh) | (w[wi + 1] >> (64 - sh));
	}
};

static inline bool suffix_less(const PackedText& t, int64_t a, int64_t b)
{
	for (;;) {
		int64_t la = t.n - a, lb = t.n - b, m = la < lb ? la : lb;
		uint64_t ka = t.key(a), kb = t.key(b);
		if (m >= 32) { if (ka != kb) return ka < kb; a += 32; b += 32; continue; }
		if (m > 0) { int sh = (int)(64 - 2 * m); uint64_t xa = ka >> sh, xb = kb >> sh; if (xa != xb) return xa < xb; }
		return la < lb; // one suffix ran into '$': the shorter one is smaller
	}
}

template <typename idx_t>
static void sort_suffixes(const PackedText& t, std::vector<idx_t>& sa, int n_threads)
{
	const int K = 12; // bucket by the first K bases
	const size_t NB = (size_t)1 << (2 * K);
	int64_t n = t.n;
	std::vector<uint64_t> cnt(NB + 1, 0);
	auto bucket_of = [&](int64_t i) -> size_t { return (size_t)(t.key(i) >> (64 - 2 * K)); };
	for (int64_t i = 0; i < n; ++i) ++cnt[bucket_of(i) + 1];
	for (size_t b = 0; b < NB; ++b) cnt[b + 1] += cnt[b];
	sa.resize((size_t)n);
	{
		std::vector<uint64_t> pos(cnt.begin(), cnt.end() - 1);
		for (int64_t i = 0; i < n; ++i) sa[pos[bucket_of(i)]++] = (idx_t)i;
	}
	std::atomic<size_t> next(0);
	auto worker = [&]() {
		const size_t CH = 256;
		for (;;) {
			size_t b0 = next.fetch_add(CH);
			if (b0 >= NB) break;
			size_t b1 = b0 + CH < NB ? b0 + CH : NB;
			for (size_t b = b0; b < b1; ++b) {
				uint64_t s = cnt[b], e = cnt[b + 1];
				if (e - s > 1)
					std::sort(sa.begin() + s, sa.begin() + e, [&](idx_t x, idx_t y) { return suffix_less(t, (int64_t)x, (int64_t)y); });
			}
		}
	};
	if (n_threads < 1) n_threads = 1;
	std::vector<std::thread> th;
	for (int i = 1; i < n_threads; ++i) th.emplace_back(worker);
	worker();
	for (auto& x : th) x.join();
}

// ---------------------------------------------------------------- FM index from the packed reference

template <typename idx_t>
static void build_fm_t(const RefSeqs& ref, FmIndex& fm, int n_threads)
{
	int64_t L = ref.l_pac, n = 2 * L;
	PackedText t; t.n = n; t.w.assign((size_t)(n / 32 + 3), 0);
	auto put = [&](int64_t i, uint64_t c) { t.w[i >> 5] |= c << ((31 - (i & 31)) << 1); };
	for (int64_t i = 0; i < L; ++i) put(i, (uint64_t)pac_get(ref.pac.data(), i));
	for (int64_t i = 0; i < L; ++i) put(L + i, (uint64_t)(3 - pac_get(ref.pac.data(), L - 1 - i)));
	std::vector<idx_t> sa;
	sort_suffixes<idx_t>(t, sa, n_threads);
	auto base_at = [&](int64_t i) -> int { return (int)(t.w[i >> 5] >> ((31 - (i & 31)) << 1) & 3); };
	// rows of the full SA: row 0 = '$' suffix (position n), row r>=1 = sa[r-1]
	fm.seq_len = (uint64_t)n;
	fm.L2[0] = 0;
	{
		uint64_t c[4] = {0, 0, 0, 0};
		for (int64_t i = 0; i < n; ++i) ++c[base_at(i)];
		for (int k = 0; k < 4; ++k) fm.L2[k + 1] = fm.L2[k] + c[k];
	}
	// BWT without the '$' row: for row r (0..n), char = T[SA[r]-1]; the row with SA[r]==0 is `primary`
	uint64_t n_blocks = ((uint64_t)n + 127) / 128;
	fm.bwt_size = ((uint64_t)n + 15) / 16 + (n_blocks + 1) * 8; // words: symbols + interleaved occ blocks (+1 closing)
	// NB: bwa's interleaved size formula: bwt_size_old = (n+15)/16 ; n_occ = (n+127)/128 + 1 ; total = old + n_occ*8
	fm.bwt.assign((size_t)fm.bwt_size, 0);
	uint64_t occ[4] = {0, 0, 0, 0};
	uint64_t k = 0; // index into the packed BWT
	auto block_ptr = [&](uint64_t kk) -> uint32_t* { return fm.bwt.data() + ((kk >> 7) << 4); };
	auto emit = [&](int c) {
		if ((k & 127) == 0) memcpy(block_ptr(k), occ, 32);
		uint32_t* p = block_ptr(k) + 8 + ((k & 127) >> 4);
		*p |= (uint32_t)c << ((15 - (k & 15)) << 1);
		++occ[c]; ++k;
	};
	fm.primary = 0;
	// row 0: '$' suffix, preceded by T[n-1]
	emit(base_at(n - 1));
	for (int64_t r = 1; r <= n; ++r) {
		int64_t p = (int64_t)sa[(size_t)(r - 1)];
		if (p == 0) { fm.primary = (uint64_t)r; continue; }
		emit(base_at(p - 1));
	}
	{ // the closing occ block sits right behind the last symbol word
		uint64_t rem = k & 127;
		uint64_t pos = ((k >> 7) << 4) + (rem ? 8 + (rem + 15) / 16 : 0);
		memcpy(fm.bwt.data() + pos, occ, 32);
	}
	// sampled SA
	fm.sa_intv = 32;
	fm.n_sa = ((uint64_t)n + 32) / 32;
	fm.sa.assign((size_t)fm.n_sa, 0);
	fm.sa[0] = (uint64_t)-1;
	for (uint64_t r = 32; r <= (uint64_t)n; r += 32) fm.sa[r / 32] = (uint64_t)sa[(size_t)(r - 1)];
}

// Defined in sa_build.cu when this file is linked into libbwa_b200.so (the CUDA build); absent from the host-only
// tool library.  BWA_B200_BUILD=host|device forces one sorter; by default the device is used when one is present
// and the text is large enough to be worth the upload.
bool build_fm_device(const RefSeqs& ref, FmIndex& fm) __attribute__((weak));

void build_fm(const RefSeqs& ref, FmIndex& fm, int n_threads)
{
	const char* mode = getenv("BWA_B200_BUILD");
	const bool force_host = mode && strcmp(mode, "host") == 0, force_dev = mode && strcmp(mode, "device") == 0;
	if (build_fm_device && !force_host && (force_dev || ref.l_pac >= (int64_t)8 << 20)) {
		if (build_fm_device(ref, fm)) return;
		fprintf(stderr, "[bwamem-b200] index build: device sorter unavailable, using the host sorter\n");
	}
	if (2 * ref.l_pac < (int64_t)0xFFFFFFF0LL) build_fm_t<uint32_t>(ref, fm, n_threads);
	else build_fm_t<uint64_t>(ref, fm, n_threads);
}

// ---------------------------------------------------------------- image serialisation (App. A.5)

static void put_bytes(std::vector<uint8_t>& o, const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; o.insert(o.end(), b, b + n); }
template <typename T> static void put(std::vector<uint8_t>& o, T v) { put_bytes(o, &v, sizeof(T)); }

void serialize_image(const RefSeqs& ref, const FmIndex& fm, std::vector<uint8_t>& o)
{
	o.clear();
	o.reserve(1120 + fm.bwt.size() * 4 + fm.sa.size() * 8 + (size_t)ref.l_pac / 4 + 4096);
	// bwt_t header, 1120 bytes
	put<uint64_t>(o, fm.primary);
	for (int i = 0; i < 5; ++i) put<uint64_t>(o, fm.L2[i]);
	put<uint64_t>(o, fm.seq_len);
	put<uint64_t>(o, fm.bwt_size);
	put<uint64_t>(o, 0); // pointer slot
	for (int i = 0; i < 256; ++i) {
		uint32_t x = 0;
		for (int j = 0; j < 4; ++j)
			x |= (uint32_t)(((i & 3) == j) + ((i >> 2 & 3) == j) + ((i >> 4 & 3) == j) + ((i >> 6) == j)) << (j << 3);
		put<uint32_t>(o, x);
	}
	put<int32_t>(o, fm.sa_intv); put<int32_t>(o, 0);
	put<uint64_t>(o, fm.n_sa);
	put<uint64_t>(o, 0); // pointer slot
	put_bytes(o, fm.bwt.data(), fm.bwt.size() * 4);
	put_bytes(o, fm.sa.data(), fm.sa.size() * 8);
	// bntseq_t, 48 bytes
	put<int64_t>(o, ref.l_pac);
	put<int32_t>(o, (int32_t)ref.contigs.size());
	put<uint32_t>(o, ref.seed);
	put<uint64_t>(o, 0);
	put<int32_t>(o, (int32_t)ref.holes.size()); put<int32_t>(o, 0);
	put<uint64_t>(o, 0);
	put<uint64_t>(o, 0);
	for (const Hole& h : ref.holes) { put<int64_t>(o, h.offset); put<int32_t>(o, h.len); put<uint8_t>(o, (uint8_t)h.amb); put<uint8_t>(o, 0); put<uint8_t>(o, 0); put<uint8_t>(o, 0); }
	for (const Contig& c : ref.contigs) {
		put<int64_t>(o, c.offset); put<int32_t>(o, c.len); put<int32_t>(o, c.n_ambs); put<uint32_t>(o, c.gi); put<int32_t>(o, c.is_alt);
		put<uint64_t>(o, 0); put<uint64_t>(o, 0);
	}
	for (const Contig& c : ref.contigs) {
		put_bytes(o, c.name.c_str(), c.name.size() + 1);
		put_bytes(o, c.anno.c_str(), c.anno.size() + 1);
	}
	put_bytes(o, ref.pac.data(), (size_t)(ref.l_pac / 4 + 1));
}

// ---------------------------------------------------------------- FASTA reader

static bool read_fasta(const char* path, RefSeqs& ref)
{
	gzFile fp = gzopen(path, "r");
	if (!fp) return false;
	Rand48 rng(ref.seed);
	Rand48Holder holder; holder.state = &rng;
	std::string name, anno, seq, line;
	bool have = false;
	char buf[65536];
	auto flush = [&]() { if (have) ref.add_contig(name, anno, seq.data(), (int64_t)seq.size(), holder); seq.clear(); };
	std::string pending;
	for (;;) {
		int nr = gzread(fp, buf, sizeof buf);
		if (nr < 0) { gzclose(fp); return false; }
		size_t s = 0;
		for (int i = 0; i <= nr; ++i) {
			bool eof_flush = (nr == 0 && i == 0);
			if (i == nr && !eof_flush) { pending.append(buf + s, (size_t)nr - s); break; }
			if (eof_flush || buf[i] == '\n') {
				if (!eof_flush) pending.append(buf + s, (size_t)i - s);
				s = (size_t)i + 1;
				if (!pending.empty() && pending.back() == '\r') pending.pop_back();
				if (!pending.empty()) {
					if (pending[0] == '>') {
						flush();
						size_t sp = pending.find_first_of(" \t");
						name = pending.substr(1, sp == std::string::npos ? std::string::npos : sp - 1);
						anno = sp == std::string::npos ? "" : pending.substr(sp + 1);
						have = true;
					} else if (have) {
						for (char ch : pending) if (ch > ' ') seq.push_back(ch); // bwa's kseq keeps every graphical character
					}
				}
				pending.clear();
				if (eof_flush) break;
			}
		}
		if (nr == 0) break;
	}
	flush();
	gzclose(fp);
	return !ref.contigs.empty();
}

// ---------------------------------------------------------------- the five bwa files

static bool write_file(const std::string& path, const void* p, size_t n)
{
	FILE* f = fopen(path.c_str(), "wb");
	if (!f) return false;
	bool ok = n == 0 || fwrite(p, 1, n, f) == n;
	ok = (fclose(f) == 0) && ok;
	return ok;
}

static bool read_file(const std::string& path, std::vector<uint8_t>& out)
{
	FILE* f = fopen(path.c_str(), "rb");
	if (!f) return false;
	fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
	out.resize((size_t)sz);
	bool ok = sz == 0 || fread(out.data(), 1, (size_t)sz, f) == (size_t)sz;
	fclose(f);
	return ok;
}

bool write_index_files(const RefSeqs& ref, const FmIndex& fm, const std::string& prefix)
{
	{ // .pac
		std::vector<uint8_t> pac(ref.pac.begin(), ref.pac.begin() + (size_t)((ref.l_pac >> 2) + ((ref.l_pac & 3) == 0 ? 0 : 1)));
		if ((ref.l_pac & 3) == 0) pac.push_back(0);
		pac.push_back((uint8_t)(ref.l_pac & 3));
		if (!write_file(prefix + ".pac", pac.data(), pac.size())) return false;
	}
	{ // .ann / .amb (the annotation of a contig without a FASTA comment is printed as "(null)")
		std::string ann, amb; char tmp[4096];
		snprintf(tmp, sizeof tmp, "%lld %d %u\n", (long long)ref.l_pac, (int)ref.contigs.size(), ref.seed); ann += tmp;
		for (const Contig& c : ref.contigs) {
			ann += std::to_string(c.gi) + " " + c.name + " " + (c.anno.empty() ? std::string("(null)") : c.anno) + "\n";
			snprintf(tmp, sizeof tmp, "%lld %d %d\n", (long long)c.offset, c.len, c.n_ambs); ann += tmp;
		}
		snprintf(tmp, sizeof tmp, "%lld %d %u\n", (long long)ref.l_pac, (int)ref.contigs.size(), (unsigned)ref.holes.size()); amb += tmp;
		for (const Hole& h : ref.holes) { snprintf(tmp, sizeof tmp, "%lld %d %c\n", (long long)h.offset, h.len, h.amb); amb += tmp; }
		if (!write_file(prefix + ".ann", ann.data(), ann.size())) return false;
		if (!write_file(prefix + ".amb", amb.data(), amb.size())) return false;
	}
	{ // .bwt: primary, L2[1..4], words
		std::vector<uint8_t> o;
		put<uint64_t>(o, fm.primary);
		for (int i = 1; i < 5; ++i) put<uint64_t>(o, fm.L2[i]);
		put_bytes(o, fm.bwt.data(), fm.bwt.size() * 4);
		if (!write_file(prefix + ".bwt", o.data(), o.size())) return false;
	}
	{ // .sa: primary, L2[1..4], sa_intv, seq_len, sa[1..]
		std::vector<uint8_t> o;
		put<uint64_t>(o, fm.primary);
		for (int i = 1; i < 5; ++i) put<uint64_t>(o, fm.L2[i]);
		put<uint64_t>(o, (uint64_t)fm.sa_intv);
		put<uint64_t>(o, fm.seq_len);
		put_bytes(o, fm.sa.data() + 1, (fm.sa.size() - 1) * 8);
		if (!write_file(prefix + ".sa", o.data(), o.size())) return false;
	}
	return true;
}

static bool next_token(const std::vector<uint8_t>& b, size_t& p, std::string& tok)
{
	while (p < b.size() && (b[p] == ' ' || b[p] == '\n' || b[p] == '\t' || b[p] == '\r')) ++p;
	if (p >= b.size()) return false;
	size_t s = p;
	while (p < b.size() && !(b[p] == ' ' || b[p] == '\n' || b[p] == '\t' || b[p] == '\r')) ++p;
	tok.assign((const char*)b.data() + s, p - s);
	return true;
}

bool load_index_files(const std::string& prefix, RefSeqs& ref, FmIndex& fm)
{
	std::vector<uint8_t> b;
	std::string tok;
	// .ann
	if (!read_file(prefix + ".ann", b)) return false;
	size_t p = 0;
	int n_seqs = 0;
	if (!next_token(b, p, tok)) return false; ref.l_pac = atoll(tok.c_str());
	if (!next_token(b, p, tok)) return false; n_seqs = atoi(tok.c_str());
	if (!next_token(b, p, tok)) return false; ref.seed = (uint32_t)strtoul(tok.c_str(), 0, 10);
	ref.contigs.clear();
	for (int i = 0; i < n_seqs; ++i) {
		Contig c;
		if (!next_token(b, p, tok)) return false; c.gi = (uint32_t)strtoul(tok.c_str(), 0, 10);
		if (!next_token(b, p, tok)) return false; c.name = tok;
		// rest of the line = " anno"; " (null)" means no annotation
		size_t s = p; while (p < b.size() && b[p] != '\n') ++p;
		std::string rest((const char*)b.data() + s, p - s);
		if (rest.size() > 1 && rest != " (null)") c.anno = rest.substr(1); else c.anno = "";
		if (!next_token(b, p, tok)) return false; c.offset = atoll(tok.c_str());
		if (!next_token(b, p, tok)) return false; c.len = atoi(tok.c_str());
		if (!next_token(b, p, tok)) return false; c.n_ambs = atoi(tok.c_str());
		c.is_alt = 0;
		ref.contigs.push_back(c);
	}
	// .amb
	if (!read_file(prefix + ".amb", b)) return false;
	p = 0;
	{
		long long l_pac2; int n2, n_holes;
		if (!next_token(b, p, tok)) return false; l_pac2 = atoll(tok.c_str());
		if (!next_token(b, p, tok)) return false; n2 = atoi(tok.c_str());
		if (!next_token(b, p, tok)) return false; n_holes = atoi(tok.c_str());
		if (l_pac2 != ref.l_pac || n2 != n_seqs) return false;
		ref.holes.clear();
		for (int i = 0; i < n_holes; ++i) {
			Hole h;
			if (!next_token(b, p, tok)) return false; h.offset = atoll(tok.c_str());
			if (!next_token(b, p, tok)) return false; h.len = atoi(tok.c_str());
			if (!next_token(b, p, tok)) return false; h.amb = tok[0];
			ref.holes.push_back(h);
		}
	}
	// .pac
	if (!read_file(prefix + ".pac", b)) return false;
	if (b.size() < (size_t)(ref.l_pac / 4 + 1)) return false;
	ref.pac.assign(b.begin(), b.begin() + (size_t)(ref.l_pac / 4 + 1));
	// .bwt
	if (!read_file(prefix + ".bwt", b)) return false;
	if (b.size() < 40 || (b.size() - 40) % 4) return false;
	memcpy(&fm.primary, b.data(), 8);
	fm.L2[0] = 0; memcpy(&fm.L2[1], b.data() + 8, 32);
	fm.seq_len = fm.L2[4];
	fm.bwt_size = (b.size() - 40) / 4;
	fm.bwt.resize((size_t)fm.bwt_size);
	memcpy(fm.bwt.data(), b.data() + 40, (size_t)fm.bwt_size * 4);
	// .sa
	if (!read_file(prefix + ".sa", b)) return false;
	if (b.size() < 56) return false;
	{
		uint64_t primary, intv, seq_len;
		memcpy(&primary, b.data(), 8); memcpy(&intv, b.data() + 40, 8); memcpy(&seq_len, b.data() + 48, 8);
		if (primary != fm.primary || seq_len != fm.seq_len) return false;
		fm.sa_intv = (int32_t)intv;
		fm.n_sa = (fm.seq_len + intv) / intv;
		if (b.size() != 56 + (fm.n_sa - 1) * 8) return false;
		fm.sa.resize((size_t)fm.n_sa);
		fm.sa[0] = (uint64_t)-1;
		memcpy(fm.sa.data() + 1, b.data() + 56, (size_t)(fm.n_sa - 1) * 8);
	}
	return true;
}

} // namespace b200

// ---------------------------------------------------------------- C-ABI

using namespace b200;

static int hw_threads()
{
	unsigned n = std::thread::hardware_concurrency();
	const char* e = getenv("BWA_B200_BUILD_THREADS");
	if (e && atoi(e) > 0) return atoi(e);
	return n ? (int)n : 4;
}

// <prefix>.alt (bwa bns_restore): the first field of every line that does not start with '@' names an ALT contig
static void mark_alt_contigs(const char* prefix, RefSeqs& ref)
{
	FILE* fp = fopen((std::string(prefix) + ".alt").c_str(), "r");
	if (!fp) return;
	std::string tok; int c;
	while ((c = fgetc(fp)) != EOF) {
		if (c == '\t' || c == '\n' || c == '\r') {
			if (!tok.empty() && tok[0] != '@') for (auto& ctg : ref.contigs) if (ctg.name == tok) ctg.is_alt = 1;
			while (c != '\n' && c != EOF) c = fgetc(fp);
			tok.clear();
		} else tok.push_back((char)c);
	}
	fclose(fp);
}

extern "C" int jnibwa_createIndexFile(const char* refName, const char* imgName)
{
	RefSeqs ref; FmIndex fm;
	if (!load_index_files(refName, ref, fm)) { printf("Failed to load bwa index files with prefix %s\n", refName); return 2; }
	mark_alt_contigs(refName, ref); // bwa_idx_load (jnibwa.c:127) reads <prefix>.alt when it exists
	std::vector<uint8_t> img;
	serialize_image(ref, fm, img);
	FILE* f = fopen(imgName, "wb");
	if (!f) { printf("Failed to open %s for writing: %s\n", imgName, strerror(errno)); return 2; }
	size_t off = 0;
	while (off < img.size()) { // <= 1 GiB per write, like the reference
		size_t n = img.size() - off; if (n > (1UL << 30)) n = 1UL << 30;
		if (fwrite(img.data() + off, 1, n, f) != n) { printf("Failed to write %s: %s\n", imgName, strerror(errno)); fclose(f); return 2; }
		off += n;
	}
	if (fclose(f) != 0) { printf("Failed to close %s: %s\n", imgName, strerror(errno)); return 2; }
	return 0;
}

// algoName: "auto" | "rb2" | "is" are accepted for API compatibility; every value selects the one sorter here.
extern "C" int jnibwa_createReferenceIndex(const char* refFileName, const char* indexPrefix, const char* algoName)
{
	if (!(strcmp(algoName, "auto") == 0 || strcmp(algoName, "rb2") == 0 || strcmp(algoName, "is") == 0)) return 3;
	RefSeqs ref; FmIndex fm;
	if (!read_fasta(refFileName, ref)) return 2;
	build_fm(ref, fm, hw_threads());
	return write_index_files(ref, fm, indexPrefix) ? 0 : 2;
}

// bwa_idx_build (un-vendored lh3/bwa bwtindex.c) as the reference's own JNI glue calls it: ...BwaMemIndex.c:59
// `bwa_idx_build(reference_file_name, index_prefix, algo_type, -1)` with algo_type 0 = auto, 1 = rb2, 3 = is (:50-52).
// Exported so that the reference's unmodified glue links against this library (option A).  Returns 0
// on success like bwa's; block_size only tunes bwa's own construction and has no meaning here.
extern "C" int bwa_idx_build(const char* fa, const char* prefix, int algo_type, int block_size)
{
	(void)block_size;
	const char* name = algo_type == 0 ? "auto" : algo_type == 1 ? "rb2" : algo_type == 3 ? "is" : nullptr;
	if (!name || !fa || !prefix) return 1;
	return jnibwa_createReferenceIndex(fa, prefix, name) == 0 ? 0 : 1;
}

// Build an image straight from base codes held in memory (used by the synthetic-genome tools and bench.py).
// names: n NUL-terminated strings back to back; bases: ASCII, contigs concatenated; lens[n].
extern "C" int b200_buildImageFromBases(int n_contigs, const char* names, const int64_t* lens, const char* bases, const char* imgName, const char* prefixOrNull)
{
	RefSeqs ref; FmIndex fm;
	Rand48 rng(ref.seed);
	Rand48Holder holder; holder.state = &rng;
	const char* nm = names; const char* b = bases;
	for (int i = 0; i < n_contigs; ++i) {
		ref.add_contig(nm, "", b, lens[i], holder);
		nm += strlen(nm) + 1; b += lens[i];
	}
	build_fm(ref, fm, hw_threads());
	if (prefixOrNull && !write_index_files(ref, fm, prefixOrNull)) return 2;
	std::vector<uint8_t> img;
	serialize_image(ref, fm, img);
	FILE* f = fopen(imgName, "wb");
	if (!f) return 2;
	bool ok = fwrite(img.data(), 1, img.size(), f) == img.size();
	ok = (fclose(f) == 0) && ok;
	return ok ? 0 : 2;
}
