/*
 * jni_min.h -- the slice of the JNI 1.8 C interface that jni_shim.c uses, declared from the JNI specification so
 * that the shim compiles on machines without a JDK (this image has none).  The JNIEnv / JavaVM function tables are
 * arrays of function pointers whose slot order is fixed by the specification ("JNI Functions: Interface Function
 * Table"); only the slots used here are given real prototypes, the others are padded with void*.  With a JDK
 * present, build with -DB200_USE_SYSTEM_JNI and the real <jni.h> is used instead -- the shim source is the same.
 */
#ifndef B200_JNI_MIN_H_
#define B200_JNI_MIN_H_
#include <stdint.h>

typedef int32_t jint;
typedef int64_t jlong;
typedef uint8_t jboolean;
typedef double jdouble;
typedef void* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef struct jfieldID_* jfieldID;

#define JNI_OK 0
#define JNI_ERR (-1)
#define JNI_VERSION_1_8 0x00010008
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL

struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNIInvokeInterface_;
typedef const struct JNIInvokeInterface_* JavaVM;

struct JNINativeInterface_ {
	void* reserved0_3[4];                                                     /*   0-3   */
	void* slot4_5[2];                                                         /*   4 GetVersion, 5 DefineClass */
	jclass (*FindClass)(JNIEnv*, const char*);                                /*   6     */
	void* slot7_13[7];                                                        /*   7-13  */
	jint (*ThrowNew)(JNIEnv*, jclass, const char*);                           /*  14     */
	void* slot15_22[8];                                                       /*  15-22  */
	void (*DeleteLocalRef)(JNIEnv*, jobject);                                 /*  23     */
	void* slot24_93[70];                                                      /*  24-93  */
	jfieldID (*GetFieldID)(JNIEnv*, jclass, const char*, const char*);        /*  94     */
	void* slot95;                                                             /*  95 GetObjectField */
	jboolean (*GetBooleanField)(JNIEnv*, jobject, jfieldID);                  /*  96     */
	void* slot97_99[3];                                                       /*  97-99  Byte, Char, Short */
	jint (*GetIntField)(JNIEnv*, jobject, jfieldID);                          /* 100     */
	void* slot101_102[2];                                                     /* 101 Long, 102 Float */
	jdouble (*GetDoubleField)(JNIEnv*, jobject, jfieldID);                    /* 103     */
	void* slot104_166[63];                                                    /* 104-166 */
	jstring (*NewStringUTF)(JNIEnv*, const char*);                            /* 167     */
	void* slot168;                                                            /* 168 GetStringUTFLength */
	const char* (*GetStringUTFChars)(JNIEnv*, jstring, jboolean*);            /* 169     */
	void (*ReleaseStringUTFChars)(JNIEnv*, jstring, const char*);             /* 170     */
	void* slot171_228[58];                                                    /* 171-228 */
	jobject (*NewDirectByteBuffer)(JNIEnv*, void*, jlong);                    /* 229     */
	void* (*GetDirectBufferAddress)(JNIEnv*, jobject);                        /* 230     */
	jlong (*GetDirectBufferCapacity)(JNIEnv*, jobject);                       /* 231     */
	void* slot232;                                                            /* 232 GetObjectRefType */
};

struct JNIInvokeInterface_ {
	void* reserved0_2[3];                                                     /* 0-2 */
	void* slot3_5[3];                                                         /* 3 DestroyJavaVM, 4 AttachCurrentThread, 5 DetachCurrentThread */
	jint (*GetEnv)(JavaVM*, void**, jint);                                    /* 6   */
	void* slot7;                                                              /* 7 AttachCurrentThreadAsDaemon */
};

/* slot positions, as the JNI specification numbers them */
#include <stddef.h>
#define B200_JNI_SLOT(member, idx) _Static_assert(offsetof(struct JNINativeInterface_, member) == (idx) * sizeof(void*), "JNI slot " #member)
B200_JNI_SLOT(FindClass, 6); B200_JNI_SLOT(ThrowNew, 14); B200_JNI_SLOT(DeleteLocalRef, 23); B200_JNI_SLOT(GetFieldID, 94);
B200_JNI_SLOT(GetBooleanField, 96); B200_JNI_SLOT(GetIntField, 100); B200_JNI_SLOT(GetDoubleField, 103); B200_JNI_SLOT(NewStringUTF, 167);
B200_JNI_SLOT(GetStringUTFChars, 169); B200_JNI_SLOT(ReleaseStringUTFChars, 170); B200_JNI_SLOT(NewDirectByteBuffer, 229);
B200_JNI_SLOT(GetDirectBufferAddress, 230); B200_JNI_SLOT(GetDirectBufferCapacity, 231);
_Static_assert(offsetof(struct JNIInvokeInterface_, GetEnv) == 6 * sizeof(void*), "JNI slot GetEnv");
_Static_assert(sizeof(struct JNINativeInterface_) == 233 * sizeof(void*), "JNI 1.8 table has 233 slots");
#endif
