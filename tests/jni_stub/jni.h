/* tests/jni_stub/jni.h -- NOT the JDK's header.  Minimal declarations of the JNI types and of the JNIEnv function-table entries
 * that integration/jni/mlease_b200_jni.c uses, with the JNI specification's signatures, so that the shim can be type-checked and
 * linked against libmlease_b200.so on an image without a JDK (tests/test_abi.py).  The table layout is NOT the real one: nothing
 * compiled against this file may ever be loaded into a JVM. */
#ifndef MLEASE_TEST_JNI_STUB_H
#define MLEASE_TEST_JNI_STUB_H
#include <stdint.h>
typedef int32_t jint;
typedef int64_t jlong;
typedef uint8_t jboolean;
typedef float jfloat;
typedef double jdouble;
typedef jint jsize;
struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jfloatArray;
typedef jarray jdoubleArray;
typedef jarray jlongArray;
typedef jobject jstring;
struct _jfieldID;
typedef struct _jfieldID* jfieldID;
#define JNI_FALSE 0
#define JNI_TRUE 1
#define JNI_ABORT 2
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv* env, const char* name);
  jint (*ThrowNew)(JNIEnv* env, jclass clazz, const char* msg);
  jclass (*GetObjectClass)(JNIEnv* env, jobject obj);
  jfieldID (*GetFieldID)(JNIEnv* env, jclass clazz, const char* name, const char* sig);
  jlong (*GetLongField)(JNIEnv* env, jobject obj, jfieldID fieldID);
  void (*SetLongField)(JNIEnv* env, jobject obj, jfieldID fieldID, jlong val);
  jsize (*GetArrayLength)(JNIEnv* env, jarray array);
  jint* (*GetIntArrayElements)(JNIEnv* env, jintArray array, jboolean* isCopy);
  jfloat* (*GetFloatArrayElements)(JNIEnv* env, jfloatArray array, jboolean* isCopy);
  jdouble* (*GetDoubleArrayElements)(JNIEnv* env, jdoubleArray array, jboolean* isCopy);
  void (*ReleaseIntArrayElements)(JNIEnv* env, jintArray array, jint* elems, jint mode);
  void (*ReleaseFloatArrayElements)(JNIEnv* env, jfloatArray array, jfloat* elems, jint mode);
  void (*ReleaseDoubleArrayElements)(JNIEnv* env, jdoubleArray array, jdouble* elems, jint mode);
  void (*SetDoubleArrayRegion)(JNIEnv* env, jdoubleArray array, jsize start, jsize len, const jdouble* buf);
  void* (*GetDirectBufferAddress)(JNIEnv* env, jobject buf);
  const char* (*GetStringUTFChars)(JNIEnv* env, jstring str, jboolean* isCopy);
  void (*ReleaseStringUTFChars)(JNIEnv* env, jstring str, const char* chars);
  jstring (*NewStringUTF)(JNIEnv* env, const char* utf);
  jlongArray (*NewLongArray)(JNIEnv* env, jsize len);
  void (*SetLongArrayRegion)(JNIEnv* env, jlongArray array, jsize start, jsize len, const jlong* buf);
};
#endif
