/* compat/bwa/bwa_commit.h -- the reference's Makefile writes the pinned bwa commit here (src/main/c/Makefile:25-26) and
 * Java_..._getVersion returns it (...BwaMemIndex.c:162-165).  With the B200 engine it names the engine instead. */
#ifndef B200_COMPAT_BWA_COMMIT_H_
#define B200_COMPAT_BWA_COMMIT_H_
#define BWA_COMMIT "bwamem-b200 0.2 (sm_100a; bwa-mem semantics of lh3/bwa cb95061)"
#endif
