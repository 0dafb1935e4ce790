/*
 * compat/bwa/bwamem.h -- what the reference's native glue needs from bwa's bwamem.h, served by libbwa_b200.so.
 *
 * The reference's jnibwa.h (src/main/c/jnibwa.h:8) includes "bwa/bwamem.h" from a bwa checkout that its Makefile
 * clones at build time (src/main/c/Makefile:17,25-26).  With this directory on the include path instead, the
 * reference's UNMODIFIED org_broadinstitute_hellbender_utils_bwa_BwaMemIndex.c and init.c compile and link against
 * the B200 engine:
 *
 *   cc -shared -fPIC -I include/compat -I $JAVA_HOME/include -I $JAVA_HOME/include/linux \
 *      src/main/c/org_broadinstitute_hellbender_utils_bwa_BwaMemIndex.c src/main/c/init.c \
 *      -L gatk-bwamem-jni_b200/lib -lbwa_b200 -o libbwa.Linux.so
 *
 * The glue uses: the three types below (bwaidx_t only as an opaque pointer; sizeof(mem_opt_t) at ...BwaMemIndex.c:91;
 * the mem_pestat_t fields at :21-40), mem_opt_init (:91), bwa_idx_build (:59), and the libc headers bwa's own header
 * happens to pull in (strdup / strcmp / memset / sprintf / malloc are used without further includes).
 */
#ifndef B200_COMPAT_BWAMEM_H_
#define B200_COMPAT_BWAMEM_H_
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bwaidx_s bwaidx_t; /* opaque: Java holds it as a long (BwaMemIndex.java:328-351) */

typedef struct { /* 168 bytes; offsets pinned by BwaMemAligner.java:46-137 */
	int a, b, o_del, e_del, o_ins, e_ins, pen_unpaired, pen_clip5, pen_clip3, w, zdrop;
	uint64_t max_mem_intv;
	int T, flag, min_seed_len, min_chain_weight, max_chain_extend;
	float split_factor;
	int split_width, max_occ, max_chain_gap, n_threads, chunk_size;
	float mask_level, drop_ratio, XA_drop_ratio, mask_level_redun, mapQ_coef_len;
	int mapQ_coef_fac, max_ins, max_matesw, max_XA_hits, max_XA_hits_alt;
	int8_t mat[25];
} mem_opt_t;

typedef struct { int low, high; int failed; double avg, std; } mem_pestat_t; /* 32 bytes */

_Static_assert(sizeof(mem_opt_t) == 168, "mem_opt_t must be the 168-byte block BwaMemAligner.java pokes by offset");
_Static_assert(sizeof(mem_pestat_t) == 32, "mem_pestat_t layout");

mem_opt_t* mem_opt_init(void);
int bwa_idx_build(const char* fa, const char* prefix, int algo_type, int block_size);

#ifdef __cplusplus
}
#endif
#endif
