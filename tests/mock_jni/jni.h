/* MOCK jni.h -- test infrastructure only (tests/test_host_logic.py::test_jni_shim_type_checks).
 * This image has no JDK.  The mock declares just the JNI types and JNIEnv entry points bindings/jni/skeres_jni.c uses, with
 * the signatures of the JNI specification, so that `gcc -fsyntax-only` can type-check the shim's calls into include/skeres.h.
 * It is NOT a JNI implementation and nothing links against it. */
#ifndef MOCK_JNI_H
#define MOCK_JNI_H
#include <stdint.h>
typedef int32_t jint; typedef int64_t jlong; typedef int8_t jbyte; typedef uint8_t jboolean; typedef double jdouble; typedef jint jsize;
typedef struct _jobject* jobject; typedef jobject jclass; typedef jobject jstring; typedef jobject jarray; typedef jarray jdoubleArray;
typedef jarray jlongArray; typedef jarray jbyteArray; typedef jarray jintArray;
#define JNI_FALSE 0
#define JNI_TRUE 1
#define JNI_ABORT 2
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv*, const char*);
  jint (*ThrowNew)(JNIEnv*, jclass, const char*);
  jsize (*GetArrayLength)(JNIEnv*, jarray);
  void* (*GetPrimitiveArrayCritical)(JNIEnv*, jarray, jboolean*);
  void (*ReleasePrimitiveArrayCritical)(JNIEnv*, jarray, void*, jint);
  void (*GetDoubleArrayRegion)(JNIEnv*, jdoubleArray, jsize, jsize, jdouble*);
  void (*GetLongArrayRegion)(JNIEnv*, jlongArray, jsize, jsize, jlong*);
  void (*GetIntArrayRegion)(JNIEnv*, jintArray, jsize, jsize, jint*);
  void (*GetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, jbyte*);
  void (*SetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, const jbyte*);
  jbyteArray (*NewByteArray)(JNIEnv*, jsize);
  jstring (*NewStringUTF)(JNIEnv*, const char*);
  const char* (*GetStringUTFChars)(JNIEnv*, jstring, jboolean*);
  void (*ReleaseStringUTFChars)(JNIEnv*, jstring, const char*);
};
#endif
