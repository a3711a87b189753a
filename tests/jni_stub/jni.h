/* Stand-in for the JDK's jni.h: ONLY the declarations jni/vrec_jni.c uses, so that the shim can be type-checked
 * against include/vrec.h in an image without a JDK (tests/test_jni_shim_cpu.py).  Never shipped, never linked. */
#ifndef VREC_TEST_JNI_STUB_H
#define VREC_TEST_JNI_STUB_H
#include <stdint.h>
typedef int32_t jint;
typedef int64_t jlong;
typedef double jdouble;
typedef jint jsize;
typedef unsigned char jboolean;
typedef struct _jobject *jobject;
typedef jobject jarray;
typedef jobject jstring;
typedef jarray jlongArray;
typedef jarray jintArray;
typedef jarray jdoubleArray;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ {
    jsize (*GetArrayLength)(JNIEnv *env, jarray array);
    void *(*GetPrimitiveArrayCritical)(JNIEnv *env, jarray array, jboolean *isCopy);
    void (*ReleasePrimitiveArrayCritical)(JNIEnv *env, jarray array, void *carray, jint mode);
    jstring (*NewStringUTF)(JNIEnv *env, const char *utf);
};
#endif
