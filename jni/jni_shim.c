/*
 * jni_shim.c -- the JNI face of libbwa_b200.so: the nine static natives of BwaMemIndex
 * (/root/reference/src/main/java/org/broadinstitute/hellbender/utils/bwa/BwaMemIndex.java:479-487) and JNI_OnLoad.
 *
 * Restates, over the C-ABI of include/bwamem_b200.h, what the reference's glue does over bwa:
 *   /root/reference/src/main/c/org_broadinstitute_hellbender_utils_bwa_BwaMemIndex.c:42-165   (the natives)
 *   /root/reference/src/main/c/init.c:12-29                                                    (JNI_OnLoad, field-ID cache)
 * Same symbol names, argument meaning and error behaviour: 0 / NULL returns become the Java exceptions of
 * BwaMemIndex.java:333-340,411-414; an unknown index algorithm raises IllegalArgumentException (…BwaMemIndex.c:52-58).
 * Pure marshalling -- no alignment logic lives here.
 */
#ifdef B200_USE_SYSTEM_JNI
#include <jni.h>
#else
#include "jni_min.h"
#endif
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bwamem_b200.h"

#define NATIVE(ret, name) JNIEXPORT ret JNICALL Java_org_broadinstitute_hellbender_utils_bwa_BwaMemIndex_##name

/* BwaMemPairEndStats field IDs (failed:Z low:I high:I average:D std:D), cached once per JVM: init.c:3-7,20-25 */
static jfieldID g_fid_failed, g_fid_low, g_fid_high, g_fid_average, g_fid_std;

JNIEXPORT jint JNICALL JNI_OnLoad(JavaVM* vm, void* reserved)
{
	JNIEnv* env = NULL;
	(void)reserved;
	if ((*vm)->GetEnv(vm, (void**)&env, JNI_VERSION_1_8) != JNI_OK) return JNI_ERR;
	jclass cls = (*env)->FindClass(env, "org/broadinstitute/hellbender/utils/bwa/BwaMemPairEndStats");
	if (!cls) return JNI_ERR;
	if (!(g_fid_failed = (*env)->GetFieldID(env, cls, "failed", "Z"))) return JNI_ERR;
	if (!(g_fid_low = (*env)->GetFieldID(env, cls, "low", "I"))) return JNI_ERR;
	if (!(g_fid_high = (*env)->GetFieldID(env, cls, "high", "I"))) return JNI_ERR;
	if (!(g_fid_average = (*env)->GetFieldID(env, cls, "average", "D"))) return JNI_ERR;
	if (!(g_fid_std = (*env)->GetFieldID(env, cls, "std", "D"))) return JNI_ERR;
	(*env)->DeleteLocalRef(env, cls);
	return JNI_VERSION_1_8;
}

static char* copy_jstring(JNIEnv* env, jstring s)
{
	const char* tmp = (*env)->GetStringUTFChars(env, s, 0);
	char* res = tmp ? strdup(tmp) : NULL;
	if (tmp) (*env)->ReleaseStringUTFChars(env, s, tmp);
	return res;
}

/* Java stats describe the FR orientation only: slot 1 is filled, slots 0, 2, 3 are marked failed (…BwaMemIndex.c:21-40) */
static int pestat_from_java(JNIEnv* env, jobject in, b200_mem_pestat_t out[4])
{
	if (in == NULL) return 0;
	memset(out, 0, 4 * sizeof(b200_mem_pestat_t));
	for (int i = 0; i < 4; ++i) {
		if (i != 1) { out[i].failed = 1; continue; }
		out[i].failed = (int)(*env)->GetBooleanField(env, in, g_fid_failed);
		if (!out[i].failed) {
			out[i].low = (int)(*env)->GetIntField(env, in, g_fid_low);
			out[i].high = (int)(*env)->GetIntField(env, in, g_fid_high);
			out[i].avg = (double)(*env)->GetDoubleField(env, in, g_fid_average);
			out[i].std = (double)(*env)->GetDoubleField(env, in, g_fid_std);
		}
	}
	return 1;
}

NATIVE(jboolean, createReferenceIndex)(JNIEnv* env, jclass cls, jstring jfasta, jstring jprefix, jstring jalgo)
{
	(void)cls;
	char* fasta = copy_jstring(env, jfasta); char* prefix = copy_jstring(env, jprefix); char* algo = copy_jstring(env, jalgo);
	jboolean res = 0;
	if (fasta && prefix && algo) {
		if (strcmp(algo, "auto") && strcmp(algo, "is") && strcmp(algo, "rb2")) {
			char* msg = (char*)malloc(strlen(algo) + 100);
			sprintf(msg, "wrong algorithm name '%s'", algo);
			(*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/IllegalArgumentException"), msg);
			free(msg);
		} else res = (jboolean)!jnibwa_createReferenceIndex(fasta, prefix, algo);
	}
	free(fasta); free(prefix); free(algo);
	return res;
}

NATIVE(jboolean, createIndexImageFile)(JNIEnv* env, jclass cls, jstring jprefix, jstring jimg)
{
	(void)cls;
	char* prefix = copy_jstring(env, jprefix); char* img = copy_jstring(env, jimg);
	jboolean res = (jboolean)(prefix && img && !jnibwa_createIndexFile(prefix, img));
	free(prefix); free(img);
	return res;
}

NATIVE(jlong, openIndex)(JNIEnv* env, jclass cls, jstring jimg)
{
	(void)cls;
	char* fname = copy_jstring(env, jimg);
	int fd = fname ? open(fname, O_RDONLY) : -1;
	free(fname);
	if (fd == -1) return 0;
	return (jlong)(intptr_t)jnibwa_openIndex(fd); /* takes ownership of fd; NULL (no device / bad image) -> 0 -> CouldNotReadImageException */
}

NATIVE(jint, destroyIndex)(JNIEnv* env, jclass cls, jlong addr)
{
	(void)env; (void)cls;
	if (!addr) return 0;
	return (jint)jnibwa_destroyIndex((void*)(intptr_t)addr);
}

NATIVE(jobject, createDefaultOptions)(JNIEnv* env, jclass cls)
{
	(void)cls;
	void* opts = mem_opt_init();
	if (!opts) return NULL;
	jobject buf = (*env)->NewDirectByteBuffer(env, opts, (jlong)sizeof(b200_mem_opt_t));
	if (!buf) jnibwa_free(opts);
	return buf;
}

NATIVE(jobject, getRefContigNames)(JNIEnv* env, jclass cls, jlong addr)
{
	(void)cls;
	if (!addr) return NULL;
	size_t size = 0;
	void* mem = jnibwa_getRefContigNames((void*)(intptr_t)addr, &size);
	if (!mem) return NULL;
	jobject buf = (*env)->NewDirectByteBuffer(env, mem, (jlong)size);
	if (!buf) jnibwa_free(mem);
	return buf;
}

NATIVE(jobject, createAlignments)(JNIEnv* env, jclass cls, jobject seqsBuf, jlong addr, jobject optsBuf, jobject frPEStats)
{
	(void)cls;
	if (!addr || !seqsBuf || !optsBuf) return NULL;
	void* opts = (*env)->GetDirectBufferAddress(env, optsBuf);
	char* seqs = (char*)(*env)->GetDirectBufferAddress(env, seqsBuf);
	if (!opts || !seqs) return NULL;
	b200_mem_pestat_t stats[4];
	int provided = pestat_from_java(env, frPEStats, stats);
	size_t size = 0;
	void* mem = jnibwa_createAlignments((void*)(intptr_t)addr, opts, provided ? stats : NULL, seqs, &size);
	if (!mem) return NULL; /* -> IllegalStateException in BwaMemIndex.doAlignment (BwaMemIndex.java:411-414) */
	jobject buf = (*env)->NewDirectByteBuffer(env, mem, (jlong)size);
	if (!buf) jnibwa_free(mem);
	return buf;
}

NATIVE(void, destroyByteBuffer)(JNIEnv* env, jclass cls, jobject buf)
{
	(void)cls;
	if (buf) jnibwa_free((*env)->GetDirectBufferAddress(env, buf));
}

NATIVE(jstring, getVersion)(JNIEnv* env, jclass cls)
{
	(void)cls;
	return (*env)->NewStringUTF(env, jnibwa_getVersion());
}
