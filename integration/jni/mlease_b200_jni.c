/* mlease_b200_jni.c -- JNI shim of com.linkedin.mlease.regression.gpu.NativeAdmm over the mlease_world_* entry points of
 * include/mlease_b200.h.  Reference-side source: built where a JDK exists,
 *     cc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude integration/jni/mlease_b200_jni.c \
 *        -Lml-ease_b200/lib -lmlease_b200 -o libmlease_b200_jni.so
 * This image has no JDK; tests/test_abi.py type-checks and links this file against a stub jni.h (tests/jni_stub/) so that it
 * cannot drift from the header.  Error mapping: MLEASE_ERR_STATE -> RuntimeException, everything else -> IOException with the
 * library's message ("Model fitting error!", "Some models failed!", "Only L1 and L2 regularization supported!" ...), as the
 * reference's jobs throw them (jobs/RegressionAdmmTrain.java:144-147,713-716; utils/LinearModelUtils.java:80-83). */
#include <jni.h>
#include <stdint.h>
#include <string.h>

#include "mlease_b200.h"

#define NATIVE(ret, name) JNIEXPORT ret JNICALL Java_com_linkedin_mlease_regression_gpu_NativeAdmm_##name

static void throw_for(JNIEnv* env, int rc) {
  const char* cls = rc == MLEASE_ERR_STATE ? "java/lang/RuntimeException" : "java/io/IOException";
  jclass c = (*env)->FindClass(env, cls);
  if (c) (*env)->ThrowNew(env, c, mlease_last_error());
}
static mlease_world* world_of(JNIEnv* env, jobject self) {
  jclass c = (*env)->GetObjectClass(env, self);
  jfieldID f = (*env)->GetFieldID(env, c, "handle", "J");
  return f ? (mlease_world*)(intptr_t)(*env)->GetLongField(env, self, f) : NULL;
}
static void* direct(JNIEnv* env, jobject buf) { return buf ? (*env)->GetDirectBufferAddress(env, buf) : NULL; }

NATIVE(jlong, create)(JNIEnv* env, jclass cls, jintArray devices, jint numBlocks, jint numFeatures, jfloatArray lambdas, jfloatArray rhos,
                      jfloatArray lambdaMap, jint regularizer, jboolean penalize, jdouble epsilon, jfloat rhoAdapt, jboolean aggressive, jboolean binary) {
  (void)cls;
  mlease_admm_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.num_blocks = numBlocks; cfg.num_features = numFeatures; cfg.regularizer = regularizer;   /* 1 or 2, else the library raises the job's own message */
  cfg.num_lambdas = (*env)->GetArrayLength(env, lambdas);
  jfloat* l = (*env)->GetFloatArrayElements(env, lambdas, NULL);
  jfloat* r = rhos ? (*env)->GetFloatArrayElements(env, rhos, NULL) : NULL;
  jfloat* m = lambdaMap ? (*env)->GetFloatArrayElements(env, lambdaMap, NULL) : NULL;
  jint* d = devices ? (*env)->GetIntArrayElements(env, devices, NULL) : NULL;
  cfg.lambdas = l; cfg.rhos = r; cfg.lambda_map = m;
  cfg.penalize_intercept = penalize; cfg.epsilon = epsilon; cfg.rho_adapt_coefficient = rhoAdapt;
  cfg.aggressive_decay = aggressive; cfg.binary_feature = binary;
  mlease_world* w = NULL;
  const int rc = mlease_world_create(&cfg, (const int32_t*)d, d ? (*env)->GetArrayLength(env, devices) : 0, &w);   /* the library copies the arrays */
  if (d) (*env)->ReleaseIntArrayElements(env, devices, d, JNI_ABORT);
  (*env)->ReleaseFloatArrayElements(env, lambdas, l, JNI_ABORT);
  if (r) (*env)->ReleaseFloatArrayElements(env, rhos, r, JNI_ABORT);
  if (m) (*env)->ReleaseFloatArrayElements(env, lambdaMap, m, JNI_ABORT);
  if (rc) { throw_for(env, rc); return 0; }
  return (jlong)(intptr_t)w;
}

NATIVE(void, addPartitionCsr)(JNIEnv* env, jobject self, jint pid, jlong nrows, jobject rowptr, jobject colidx, jobject vals, jobject resp,
                              jobject weight, jobject offset) {
  const int rc = mlease_world_add_partition_csr(world_of(env, self), pid, nrows, (const int64_t*)direct(env, rowptr), (const int32_t*)direct(env, colidx),
                                                (const float*)direct(env, vals), (const int32_t*)direct(env, resp), (const float*)direct(env, weight),
                                                (const float*)direct(env, offset));
  if (rc) throw_for(env, rc);
}

NATIVE(void, begin)(JNIEnv* env, jobject self) {
  const int rc = mlease_world_begin(world_of(env, self));
  if (rc) throw_for(env, rc);
}

NATIVE(void, beginInitialized)(JNIEnv* env, jobject self, jdoubleArray z0, jfloat boost) {
  jdouble* z = (*env)->GetDoubleArrayElements(env, z0, NULL);
  const int rc = mlease_world_begin_initialized(world_of(env, self), z, boost);
  (*env)->ReleaseDoubleArrayElements(env, z0, z, JNI_ABORT);
  if (rc) throw_for(env, rc);
}

NATIVE(jint, fitPartition)(JNIEnv* env, jobject self, jint pid, jdoubleArray x, jdoubleArray mean, jdoubleArray prec) {
  jdouble* xv = (*env)->GetDoubleArrayElements(env, x, NULL);
  jdouble* mv = (*env)->GetDoubleArrayElements(env, mean, NULL);
  jdouble* qv = (*env)->GetDoubleArrayElements(env, prec, NULL);
  int32_t steps = 0;
  const int rc = mlease_world_fit_partition(world_of(env, self), pid, xv, mv, qv, &steps);
  (*env)->ReleaseDoubleArrayElements(env, x, xv, rc ? JNI_ABORT : 0);      /* 0 = copy the fitted model back */
  (*env)->ReleaseDoubleArrayElements(env, mean, mv, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, prec, qv, JNI_ABORT);
  if (rc) throw_for(env, rc);
  return steps;
}

NATIVE(jboolean, iterate)(JNIEnv* env, jobject self, jdoubleArray maxdiffOut) {
  double md = 0;
  int32_t stop = 0;
  const int rc = mlease_world_iterate(world_of(env, self), &md, &stop);   /* MLEASE_ERR_NUMERIC = a reducer's fit failed: IOException("Model fitting error!") */
  if (rc) { throw_for(env, rc); return JNI_FALSE; }
  if (maxdiffOut) (*env)->SetDoubleArrayRegion(env, maxdiffOut, 0, 1, &md);
  return stop ? JNI_TRUE : JNI_FALSE;
}

NATIVE(jint, run)(JNIEnv* env, jobject self, jint numIters) {
  int32_t done = 0;
  const int rc = mlease_world_run(world_of(env, self), numIters, &done);
  if (rc) throw_for(env, rc);
  return done;
}

NATIVE(void, getZ)(JNIEnv* env, jobject self, jint l, jdoubleArray out) {
  jdouble* o = (*env)->GetDoubleArrayElements(env, out, NULL);
  const int rc = mlease_world_get_z(world_of(env, self), l, o);
  (*env)->ReleaseDoubleArrayElements(env, out, o, rc ? JNI_ABORT : 0);
  if (rc) throw_for(env, rc);
}
NATIVE(void, getFinalModel)(JNIEnv* env, jobject self, jint l, jfloatArray out) {
  jfloat* o = (*env)->GetFloatArrayElements(env, out, NULL);
  const int rc = mlease_world_get_final_model(world_of(env, self), l, o);
  (*env)->ReleaseFloatArrayElements(env, out, o, rc ? JNI_ABORT : 0);
  if (rc) throw_for(env, rc);
}
NATIVE(void, getX)(JNIEnv* env, jobject self, jint p, jint l, jdoubleArray out) {
  jdouble* o = (*env)->GetDoubleArrayElements(env, out, NULL);
  const int rc = mlease_world_get_x(world_of(env, self), p, l, o);
  (*env)->ReleaseDoubleArrayElements(env, out, o, rc ? JNI_ABORT : 0);
  if (rc) throw_for(env, rc);
}
NATIVE(void, getU)(JNIEnv* env, jobject self, jint p, jint l, jfloatArray out) {
  jfloat* o = (*env)->GetFloatArrayElements(env, out, NULL);
  const int rc = mlease_world_get_u(world_of(env, self), p, l, o);
  (*env)->ReleaseFloatArrayElements(env, out, o, rc ? JNI_ABORT : 0);
  if (rc) throw_for(env, rc);
}
NATIVE(void, getUplusx)(JNIEnv* env, jobject self, jint p, jint l, jfloatArray out) {
  jfloat* o = (*env)->GetFloatArrayElements(env, out, NULL);
  const int rc = mlease_world_get_uplusx(world_of(env, self), p, l, o);
  (*env)->ReleaseFloatArrayElements(env, out, o, rc ? JNI_ABORT : 0);
  if (rc) throw_for(env, rc);
}

NATIVE(void, close)(JNIEnv* env, jobject self) {
  jclass c = (*env)->GetObjectClass(env, self);
  jfieldID f = (*env)->GetFieldID(env, c, "handle", "J");
  if (!f) return;
  mlease_world* w = (mlease_world*)(intptr_t)(*env)->GetLongField(env, self, f);
  if (w) mlease_world_destroy(w);
  (*env)->SetLongField(env, self, f, 0);
}

/* ---------------------------------------------------------------------------------------------------------------------------
 * com.linkedin.mlease.regression.gpu.NativeOps: the stateless entry points (NaiveTrain reducer, scoring, test log-likelihood)
 * ------------------------------------------------------------------------------------------------------------------------- */
#define OPS(ret, name) JNIEXPORT ret JNICALL Java_com_linkedin_mlease_regression_gpu_NativeOps_##name

OPS(void, naiveTrain)(JNIEnv* env, jclass cls, jint device, jint numKeys, jint numFeatures, jobject keyRowStart, jobject rowptr, jobject colidx,
                      jobject vals, jobject response, jobject weight, jobject offset, jfloatArray lambdas, jfloatArray lambdaMap, jfloat priorMean,
                      jboolean penalize, jboolean hasIntercept, jint threshold, jboolean binary, jdoubleArray outModel, jintArray skipped) {
  (void)cls;
  jfloat* l = (*env)->GetFloatArrayElements(env, lambdas, NULL);
  jfloat* m = lambdaMap ? (*env)->GetFloatArrayElements(env, lambdaMap, NULL) : NULL;
  jdouble* out = (*env)->GetDoubleArrayElements(env, outModel, NULL);
  jint* sk = (*env)->GetIntArrayElements(env, skipped, NULL);
  const int rc = mlease_naive_train(device, NULL, numKeys, numFeatures, (const int64_t*)direct(env, keyRowStart), (const int64_t*)direct(env, rowptr),
                                    (const int32_t*)direct(env, colidx), (const float*)direct(env, vals), 0, (const int32_t*)direct(env, response),
                                    (const float*)direct(env, weight), (const float*)direct(env, offset), (*env)->GetArrayLength(env, lambdas), l, m, priorMean,
                                    penalize, hasIntercept, threshold, binary, out, (int32_t*)sk);
  (*env)->ReleaseIntArrayElements(env, skipped, sk, rc ? JNI_ABORT : 0);
  (*env)->ReleaseDoubleArrayElements(env, outModel, out, rc ? JNI_ABORT : 0);
  if (m) (*env)->ReleaseFloatArrayElements(env, lambdaMap, m, JNI_ABORT);
  (*env)->ReleaseFloatArrayElements(env, lambdas, l, JNI_ABORT);
  if (rc) throw_for(env, rc);
}

OPS(void, score)(JNIEnv* env, jclass cls, jint device, jint numFeatures, jlong nrows, jobject rowptr, jobject colidx, jobject vals, jobject offset,
                 jdoubleArray model, jint reps, jboolean binary, jfloatArray predOut) {
  (void)cls;
  jdouble* b = (*env)->GetDoubleArrayElements(env, model, NULL);
  jfloat* p = (*env)->GetFloatArrayElements(env, predOut, NULL);
  const int rc = mlease_score(device, NULL, numFeatures, nrows, (const int64_t*)direct(env, rowptr), (const int32_t*)direct(env, colidx),
                              (const float*)direct(env, vals), 0, (const float*)direct(env, offset), b, reps, binary, p);
  (*env)->ReleaseFloatArrayElements(env, predOut, p, rc ? JNI_ABORT : 0);
  (*env)->ReleaseDoubleArrayElements(env, model, b, JNI_ABORT);
  if (rc) throw_for(env, rc);
}

OPS(jfloat, testLoglik)(JNIEnv* env, jclass cls, jint device, jintArray response, jfloatArray pred, jfloatArray weight, jlong combinerBlock,
                        jdoubleArray countOut) {
  (void)cls;
  const jsize n = (*env)->GetArrayLength(env, response);
  jint* r = (*env)->GetIntArrayElements(env, response, NULL);
  jfloat* p = (*env)->GetFloatArrayElements(env, pred, NULL);
  jfloat* w = (*env)->GetFloatArrayElements(env, weight, NULL);
  float ll = 0.f;
  double cnt = 0.0;
  const int rc = mlease_test_loglik(device, NULL, n, (const int32_t*)r, p, w, combinerBlock, &ll, &cnt);
  (*env)->ReleaseFloatArrayElements(env, weight, w, JNI_ABORT);
  (*env)->ReleaseFloatArrayElements(env, pred, p, JNI_ABORT);
  (*env)->ReleaseIntArrayElements(env, response, r, JNI_ABORT);
  if (rc) { throw_for(env, rc); return 0.f; }
  if (countOut) (*env)->SetDoubleArrayRegion(env, countOut, 0, 1, &cnt);
  return ll;
}

/* ---------------------------------------------------------------------------------------------------------------------------
 * com.linkedin.mlease.regression.gpu.NativeIngest: the job layer's avro ingest (include/mlease_host.h; link libmlease_host.so too)
 * ------------------------------------------------------------------------------------------------------------------------- */
#include "mlease_host.h"
#define ING(ret, name) JNIEXPORT ret JNICALL Java_com_linkedin_mlease_regression_gpu_NativeIngest_##name

static mlease_rows* rows_of(JNIEnv* env, jobject self) {
  jclass c = (*env)->GetObjectClass(env, self);
  jfieldID f = (*env)->GetFieldID(env, c, "handle", "J");
  return f ? (mlease_rows*)(intptr_t)(*env)->GetLongField(env, self, f) : NULL;
}

ING(jlong, read)(JNIEnv* env, jclass cls, jstring path, jboolean raw, jboolean binary) {
  (void)cls;
  const char* p = (*env)->GetStringUTFChars(env, path, NULL);
  mlease_rows* r = NULL;
  const int rc = mlease_rows_read(p, raw, binary, 0, &r);
  (*env)->ReleaseStringUTFChars(env, path, p);
  if (rc) {
    jclass c = (*env)->FindClass(env, "java/io/IOException");
    if (c) (*env)->ThrowNew(env, c, mlease_job_last_error());   /* the reference's texts: "features is null", "name is null", ... */
    return 0;
  }
  return (jlong)(intptr_t)r;
}
ING(jlongArray, counts)(JNIEnv* env, jobject self) {
  int64_t nnz = 0;
  int32_t nf = 0;
  const int64_t n = mlease_rows_count(rows_of(env, self), &nnz, &nf);
  jlong v[3];
  v[0] = n; v[1] = nnz; v[2] = nf;
  jlongArray out = (*env)->NewLongArray(env, 3);
  if (out) (*env)->SetLongArrayRegion(env, out, 0, 3, v);
  return out;
}
ING(void, get)(JNIEnv* env, jobject self, jobject rowptr, jobject colidx, jobject vals, jobject response, jobject weight, jobject offset) {
  mlease_rows_get(rows_of(env, self), (int64_t*)direct(env, rowptr), (int32_t*)direct(env, colidx), (float*)direct(env, vals), (int32_t*)direct(env, response),
                  (float*)direct(env, weight), (float*)direct(env, offset));
}
ING(jstring, feature)(JNIEnv* env, jobject self, jint k) {
  const char* s = mlease_rows_feature(rows_of(env, self), k);
  return s ? (*env)->NewStringUTF(env, s) : NULL;   /* U+0001 is a one-byte sequence in modified UTF-8 as well */
}
ING(jstring, key)(JNIEnv* env, jobject self, jlong i) {
  const char* s = mlease_rows_key(rows_of(env, self), i);
  return s ? (*env)->NewStringUTF(env, s) : NULL;
}
ING(void, close)(JNIEnv* env, jobject self) {
  jclass c = (*env)->GetObjectClass(env, self);
  jfieldID f = (*env)->GetFieldID(env, c, "handle", "J");
  if (!f) return;
  mlease_rows* r = (mlease_rows*)(intptr_t)(*env)->GetLongField(env, self, f);
  if (r) mlease_rows_free(r);
  (*env)->SetLongField(env, self, f, 0);
}
