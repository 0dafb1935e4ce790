/*
 * bwamem_b200.h -- C-ABI of libbwa_b200.so, the B200-native BWA-MEM engine.
 *
 * This is the drop-in boundary for the hot path of broadinstitute/gatk-bwamem-jni: every entry point below has the
 * name and signature of the function the reference's JNI glue already binds (reference file:line given per entry),
 * with plain pointers and sizes only.  The index handle is opaque (the reference's Java side holds it as a `long`,
 * BwaMemIndex.java:328-351), so `bwaidx_t*` is spelled `void*` here; `mem_opt_t*` / `mem_pestat_t*` are the 168-byte
 * and 4x32-byte blocks whose layouts BwaMemAligner.java:46-137 and ...BwaMemIndex.c:21-40 pin (structs below).
 *
 * INTEGRATION.md shows the JNI stub a maintainer of the reference would keep (it is the reference's own
 * org_broadinstitute_hellbender_utils_bwa_BwaMemIndex.c, unchanged, linked against this library instead of libbwa.a).
 */
#ifndef BWAMEM_B200_H_
#define BWAMEM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* mem_opt_t as laid out by bwa @cb95061 on x86-64 (168 bytes; offsets hard-coded at BwaMemAligner.java:46-137) */
typedef struct {
	int32_t a, b, o_del, e_del, o_ins, e_ins, pen_unpaired, pen_clip5, pen_clip3, w, zdrop, pad0_;
	uint64_t max_mem_intv;
	int32_t T, flag, min_seed_len, min_chain_weight, max_chain_extend;
	float split_factor;
	int32_t split_width, max_occ, max_chain_gap, n_threads, chunk_size;
	float mask_level, drop_ratio, XA_drop_ratio, mask_level_redun, mapQ_coef_len;
	int32_t mapQ_coef_fac, max_ins, max_matesw, max_XA_hits, max_XA_hits_alt;
	int8_t mat[25];
	int8_t pad1_[3];
} b200_mem_opt_t;

/* mem_pestat_t (32 bytes); an array of 4 (orientations FF, FR, RF, RR) as built at ...BwaMemIndex.c:21-40 */
typedef struct { int32_t low, high, failed, pad_; double avg, std; } b200_mem_pestat_t;

/* replaces jnibwa_openIndex            jnibwa.h:13, jnibwa.c:154-165 : takes ownership of fd (closes it), parses the
 * BwaMemIndex image, uploads occ-BWT / sampled SA / 2-bit pac to HBM of the current CUDA device.  NULL on failure. */
void* jnibwa_openIndex(int fd);

/* replaces jnibwa_destroyIndex         jnibwa.h:14, jnibwa.c:167-172 : releases host mapping and device memory. */
int jnibwa_destroyIndex(void* idx);

/* replaces jnibwa_getRefContigNames    jnibwa.h:15, jnibwa.c:174-195 : malloc'd blob `i32 n; {i32 len; bytes}*n`. */
void* jnibwa_getRefContigNames(void* idx, size_t* pBufSize);

/* replaces jnibwa_createAlignments     jnibwa.h:16, jnibwa.c:197-235 (the hot path; bwa's mem_process_seqs at :214).
 * pSeq: `u32 nSeqs` then nSeqs NUL-terminated reads (BwaMemAligner.java:198-209); pOpts: live mem_opt_t; pStats:
 * NULL (infer per call) or mem_pestat_t[4].  Returns a malloc'd int32 stream (layout ...BwaMemIndex.c:115-141) and
 * its byte size; NULL on failure (never a partial or CPU-computed result).  Unlike the reference, the read bytes
 * in pSeq are left untouched (the reference overwrites them with 2-bit codes, an unobservable side effect in Java). */
void* jnibwa_createAlignments(void* idx, void* pOpts, void* pStats, char* pSeq, size_t* pBufSize);

/* replaces jnibwa_createIndexFile      jnibwa.h:12, jnibwa.c:126-152 : bwa index files <refName>.{amb,ann,bwt,pac,sa}
 * -> BwaMemIndex image.  0 on success. */
int jnibwa_createIndexFile(const char* refName, const char* imgName);

/* replaces bwa_idx_build as called by Java_..._createReferenceIndex, ...BwaMemIndex.c:42-63 (jnibwa.h:11 declares
 * jnibwa_createReferenceIndex but the reference never defines it).  algoName in {"auto","rb2","is"}; 3 = unknown. */
int jnibwa_createReferenceIndex(const char* refFileName, const char* indexPrefix, const char* algoName);

/* replaces mem_opt_init (bwa), called at ...BwaMemIndex.c:91 : malloc'd defaults. */
void* mem_opt_init(void);

/* what Java_..._destroyByteBuffer (...BwaMemIndex.c:157-160) must call for every buffer returned above. */
void jnibwa_free(void* p);

/* replaces the BWA_COMMIT string of ...BwaMemIndex.c:162-165. */
const char* jnibwa_getVersion(void);

/* ---- instrumentation (not part of the reference surface; used by bench.py) ---- */
int b200_setDevice(int dev);
/* contexts of a handle = devices of its pool (BWA_B200_DEVICES) x batches in flight per device (BWA_B200_CONTEXTS);
 * the device of context i.  The b200_*Staged / lastBatchMs / lastStageTimes calls below act on context 0. */
int b200_contextCount(void* idx);
int b200_contextDevice(void* idx, int i);
const char* b200_lastError(void* idx);
int b200_lastStageTimes(void* idx, float* ms, const char** names, int cap);
void b200_lastCounters(void* idx, uint64_t* out8);
float b200_lastBatchMs(void* idx);
/* HBM-resident measurement: stage = host parse + upload; alignStaged = device pipeline only, result stays in HBM
 * (returns its byte size, -1 on failure) */
int b200_stageBatch(void* idx, const char* pSeq);
long long b200_alignStaged(void* idx, void* pOpts, void* pStats);
/* read boundaries of a request body (n NUL-terminated reads at blob): off[0..n], longest read; returns off[n].  Test hook of
 * the host-side scan (csrc/host_scan.cpp), which never reads past the n-th NUL. */
uint64_t b200_scanReads(const char* blob, uint32_t n, uint64_t* off, int* max_len, int force_plain);
/* test hook: shared-memory layout of the thread-level extension kernels (see csrc/kernels.cuh: thread_ext_layout);
 * out[9] = row offset, row stride, query offset, query stride, ringH offset, ringE offset, ring stride, ring slots, block bytes */
void b200_extLayout(int threads, int tid, int rw, int wm_cap, int cell, int patches, uint64_t* out);
/* index construction straight from bases in memory (synthetic-genome tools) */
int b200_buildImageFromBases(int n_contigs, const char* names, const int64_t* lens, const char* bases,
                             const char* imgName, const char* prefixOrNull);

#ifdef __cplusplus
}
#endif
#endif
