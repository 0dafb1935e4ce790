/*
 * mock_env.c -- TEST TOOLING: a stand-in JVM for jni/jni_shim.c (no JDK exists in this image).
 * Builds a JNIEnv / JavaVM function table with the slots the shim uses; "Java objects" are small C structs.
 * tests/test_jni_shim.py compiles this together with jni/jni_shim.c and drives the Java_* exports through ctypes,
 * passing C strings as jstring and MockBuffer / MockPeStats pointers as jobject.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "jni_min.h"

typedef struct { int kind; void* addr; long long cap; } MockBuffer;          /* java.nio.DirectByteBuffer */
typedef struct { int kind; int failed, low, high; double average, std; } MockPeStats; /* BwaMemPairEndStats */
static char g_exc_class[128], g_exc_msg[512];
static int g_live_strings = 0;
static const char* FIELD_NAMES[5] = { "failed", "low", "high", "average", "std" };

static jclass m_FindClass(JNIEnv* e, const char* name) { (void)e; return (jclass)strdup(name); }
static jint m_ThrowNew(JNIEnv* e, jclass c, const char* msg) { (void)e; snprintf(g_exc_class, sizeof g_exc_class, "%s", (const char*)c); snprintf(g_exc_msg, sizeof g_exc_msg, "%s", msg); return 0; }
static void m_DeleteLocalRef(JNIEnv* e, jobject o) { (void)e; free(o); }
static jfieldID m_GetFieldID(JNIEnv* e, jclass c, const char* name, const char* sig)
{
	(void)e; (void)sig;
	if (strcmp((const char*)c, "org/broadinstitute/hellbender/utils/bwa/BwaMemPairEndStats")) return NULL;
	for (int i = 0; i < 5; ++i) if (!strcmp(name, FIELD_NAMES[i])) return (jfieldID)FIELD_NAMES[i];
	return NULL;
}
static jboolean m_GetBooleanField(JNIEnv* e, jobject o, jfieldID f) { (void)e; return (const char*)f == FIELD_NAMES[0] ? (jboolean)((MockPeStats*)o)->failed : 0; }
static jint m_GetIntField(JNIEnv* e, jobject o, jfieldID f) { (void)e; MockPeStats* p = (MockPeStats*)o; return (const char*)f == FIELD_NAMES[1] ? p->low : (const char*)f == FIELD_NAMES[2] ? p->high : -1; }
static jdouble m_GetDoubleField(JNIEnv* e, jobject o, jfieldID f) { (void)e; MockPeStats* p = (MockPeStats*)o; return (const char*)f == FIELD_NAMES[3] ? p->average : (const char*)f == FIELD_NAMES[4] ? p->std : -1.; }
static jstring m_NewStringUTF(JNIEnv* e, const char* s) { (void)e; return (jstring)strdup(s); }
static const char* m_GetStringUTFChars(JNIEnv* e, jstring s, jboolean* copy) { (void)e; if (copy) *copy = 0; ++g_live_strings; return (const char*)s; }
static void m_ReleaseStringUTFChars(JNIEnv* e, jstring s, const char* c) { (void)e; (void)s; (void)c; --g_live_strings; }
static jobject m_NewDirectByteBuffer(JNIEnv* e, void* addr, jlong cap) { (void)e; MockBuffer* b = (MockBuffer*)malloc(sizeof *b); b->kind = 1; b->addr = addr; b->cap = cap; return b; }
static void* m_GetDirectBufferAddress(JNIEnv* e, jobject o) { (void)e; return ((MockBuffer*)o)->addr; }
static jlong m_GetDirectBufferCapacity(JNIEnv* e, jobject o) { (void)e; return ((MockBuffer*)o)->cap; }

static struct JNINativeInterface_ g_table;
static JNIEnv g_env = &g_table;
static jint m_GetEnv(JavaVM* vm, void** penv, jint version) { (void)vm; if (version != JNI_VERSION_1_8) return JNI_ERR; *penv = &g_env; return JNI_OK; }
static struct JNIInvokeInterface_ g_vm_table;
static JavaVM g_vm = &g_vm_table;

JNIEnv* mock_env(void)
{
	memset(&g_table, 0, sizeof g_table);
	g_table.FindClass = m_FindClass; g_table.ThrowNew = m_ThrowNew; g_table.DeleteLocalRef = m_DeleteLocalRef; g_table.GetFieldID = m_GetFieldID;
	g_table.GetBooleanField = m_GetBooleanField; g_table.GetIntField = m_GetIntField; g_table.GetDoubleField = m_GetDoubleField;
	g_table.NewStringUTF = m_NewStringUTF; g_table.GetStringUTFChars = m_GetStringUTFChars; g_table.ReleaseStringUTFChars = m_ReleaseStringUTFChars;
	g_table.NewDirectByteBuffer = m_NewDirectByteBuffer; g_table.GetDirectBufferAddress = m_GetDirectBufferAddress; g_table.GetDirectBufferCapacity = m_GetDirectBufferCapacity;
	return &g_env;
}
JavaVM* mock_vm(void) { memset(&g_vm_table, 0, sizeof g_vm_table); g_vm_table.GetEnv = m_GetEnv; mock_env(); return &g_vm; }
void* mock_new_buffer(void* addr, long long cap) { return m_NewDirectByteBuffer(&g_env, addr, cap); }
void* mock_buffer_addr(void* b) { return ((MockBuffer*)b)->addr; }
long long mock_buffer_cap(void* b) { return ((MockBuffer*)b)->cap; }
void mock_free_buffer_object(void* b) { free(b); }
void* mock_new_pestats(int failed, int low, int high, double average, double std)
{ MockPeStats* p = (MockPeStats*)malloc(sizeof *p); p->kind = 2; p->failed = failed; p->low = low; p->high = high; p->average = average; p->std = std; return p; }
const char* mock_exception_class(void) { return g_exc_class; }
const char* mock_exception_message(void) { return g_exc_msg; }
void mock_clear_exception(void) { g_exc_class[0] = g_exc_msg[0] = 0; }
int mock_live_strings(void) { return g_live_strings; }
