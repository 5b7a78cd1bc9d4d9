/* Minimal stand-in for the JDK's <jni.h>, for SYNTAX AND SYMBOL CHECKS ONLY (tests/test_jni_shim_cpu.py): the build image
 * has no JDK.  It declares exactly the types, macros and JNIEnv functions java/jni/sieve_jni.c uses, with the signatures
 * of the JNI specification (Java SE "JNI Functions": ThrowNew, FindClass, GetArrayLength, Get/ReleasePrimitiveArrayCritical,
 * SetDoubleArrayRegion).  A real build uses the JDK's header. */
#ifndef SIEVE_TEST_JNI_STUB_H
#define SIEVE_TEST_JNI_STUB_H
#include <stdint.h>
typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef double jdouble;
typedef jint jsize;
struct _jobject;
typedef struct _jobject *jobject;
typedef jobject jclass, jthrowable, jarray, jintArray, jlongArray, jdoubleArray, jbyteArray;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ {
    jclass (*FindClass)(JNIEnv *env, const char *name);
    jint (*ThrowNew)(JNIEnv *env, jclass clazz, const char *msg);
    jsize (*GetArrayLength)(JNIEnv *env, jarray array);
    void *(*GetPrimitiveArrayCritical)(JNIEnv *env, jarray array, jboolean *isCopy);
    void (*ReleasePrimitiveArrayCritical)(JNIEnv *env, jarray array, void *carray, jint mode);
    void (*SetDoubleArrayRegion)(JNIEnv *env, jdoubleArray array, jsize start, jsize len, const jdouble *buf);
};
#endif
