/*
 * compat/jni.h -- stands in for the JDK's <jni.h> on machines without a JDK (this image has none): the slice of the
 * JNI 1.8 interface that the reference's own glue (org_broadinstitute_hellbender_utils_bwa_BwaMemIndex.c, init.c)
 * and jni/jni_shim.c use, declared from the JNI specification (jni/jni_min.h).  With a JDK, put
 * $JAVA_HOME/include in front of this directory on the include path and the real header wins.
 */
#ifndef B200_COMPAT_JNI_H_
#define B200_COMPAT_JNI_H_
#include "../../jni/jni_min.h"
#endif
