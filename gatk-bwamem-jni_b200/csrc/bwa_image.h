// bwa_image.h -- host-side model of the bwa index (the five files) and of the
// BwaMemIndex ".img" image (written by the reference at
// /root/reference/src/main/c/jnibwa.c:126-152, loaded at :154-165).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace b200 {

struct Contig { int64_t offset = 0; int32_t len = 0, n_ambs = 0; uint32_t gi = 0; int32_t is_alt = 0; std::string name, anno; };
struct Hole { int64_t offset = 0; int32_t len = 0; char amb = 'N'; };
struct Rand48Holder { void* state = nullptr; };

struct RefSeqs {
	int64_t l_pac = 0;
	uint32_t seed = 11; // bwa fixes srand48(11) for ambiguous-base replacement
	std::vector<Contig> contigs;
	std::vector<Hole> holes;
	std::vector<uint8_t> pac; // 2 bits per base, first base in the top bits of a byte
	void add_contig(const std::string& name, const std::string& anno, const char* bases, int64_t len, Rand48Holder& rng);
};

struct FmIndex {
	uint64_t primary = 0, L2[5] = {0, 0, 0, 0, 0}, seq_len = 0, bwt_size = 0;
	std::vector<uint32_t> bwt; // occ-interleaved, 64-byte blocks of 128 symbols
	int32_t sa_intv = 32;
	uint64_t n_sa = 0;
	std::vector<uint64_t> sa;
};

void build_fm(const RefSeqs& ref, FmIndex& fm, int n_threads);
void serialize_image(const RefSeqs& ref, const FmIndex& fm, std::vector<uint8_t>& out);
bool write_index_files(const RefSeqs& ref, const FmIndex& fm, const std::string& prefix);
bool load_index_files(const std::string& prefix, RefSeqs& ref, FmIndex& fm);

// Parsed view of an image held in host memory (pointers into the buffer).
struct ImageView {
	uint64_t primary = 0, L2[5] = {0, 0, 0, 0, 0}, seq_len = 0, bwt_size = 0;
	const uint32_t* bwt = nullptr;
	int32_t sa_intv = 0;
	uint64_t n_sa = 0;
	const uint64_t* sa = nullptr;
	int64_t l_pac = 0;
	int32_t n_seqs = 0, n_holes = 0;
	std::vector<Contig> contigs;
	const uint8_t* pac = nullptr;
	bool parse(const uint8_t* mem, size_t len);
};

} // namespace b200
