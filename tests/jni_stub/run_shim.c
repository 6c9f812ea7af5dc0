/* tests/jni_stub/run_shim.c -- executes integration/jni/mlease_b200_jni.c on a machine without a JVM or a GPU: a fake JNIEnv
 * (arrays are handed out as COPIES, so results only reach the "Java" side when the shim releases them with mode 0; direct buffers are
 * plain memory; objects carry one long field) drives the NativeAdmm / NativeOps natives against the test double of the device
 * library (tests/fake_device/) and compares what comes back with direct calls of the C ABI.  Built and run by tests/test_abi.py. */
#define _POSIX_C_SOURCE 200809L
#include <jni.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mlease_b200.h"

struct _jobject { int kind; jsize len; size_t elem; void* data; jlong field; };   /* kind: 0 object, 1 array, 2 direct buffer, 3 class */
struct _jfieldID { int unused; };
static struct _jfieldID g_field;
static char g_thrown[256];
static int g_outstanding = 0;   /* Get*ArrayElements without Release */

static jclass f_FindClass(JNIEnv* e, const char* n) { (void)e; (void)n; static struct _jobject c = {3, 0, 0, NULL, 0}; return &c; }
static jint f_ThrowNew(JNIEnv* e, jclass c, const char* m) { (void)e; (void)c; snprintf(g_thrown, sizeof g_thrown, "%s", m); return 0; }
static jclass f_GetObjectClass(JNIEnv* e, jobject o) { (void)o; return f_FindClass(e, ""); }
static jfieldID f_GetFieldID(JNIEnv* e, jclass c, const char* n, const char* s) { (void)e; (void)c; return (strcmp(n, "handle") == 0 && strcmp(s, "J") == 0) ? &g_field : NULL; }
static jlong f_GetLongField(JNIEnv* e, jobject o, jfieldID f) { (void)e; (void)f; return o->field; }
static void f_SetLongField(JNIEnv* e, jobject o, jfieldID f, jlong v) { (void)e; (void)f; o->field = v; }
static jsize f_GetArrayLength(JNIEnv* e, jarray a) { (void)e; return a->len; }
static void* get_elems(jarray a) { void* c = malloc(a->len * a->elem + 1); memcpy(c, a->data, a->len * a->elem); g_outstanding++; return c; }
static void rel_elems(jarray a, void* c, jint mode) { if (mode != JNI_ABORT) memcpy(a->data, c, a->len * a->elem); free(c); g_outstanding--; }
static jint* f_GetInt(JNIEnv* e, jintArray a, jboolean* ic) { (void)e; if (ic) *ic = 1; return (jint*)get_elems(a); }
static jfloat* f_GetFloat(JNIEnv* e, jfloatArray a, jboolean* ic) { (void)e; if (ic) *ic = 1; return (jfloat*)get_elems(a); }
static jdouble* f_GetDouble(JNIEnv* e, jdoubleArray a, jboolean* ic) { (void)e; if (ic) *ic = 1; return (jdouble*)get_elems(a); }
static void f_RelInt(JNIEnv* e, jintArray a, jint* c, jint m) { (void)e; rel_elems(a, c, m); }
static void f_RelFloat(JNIEnv* e, jfloatArray a, jfloat* c, jint m) { (void)e; rel_elems(a, c, m); }
static void f_RelDouble(JNIEnv* e, jdoubleArray a, jdouble* c, jint m) { (void)e; rel_elems(a, c, m); }
static void f_SetDoubleRegion(JNIEnv* e, jdoubleArray a, jsize s, jsize l, const jdouble* b) { (void)e; memcpy((jdouble*)a->data + s, b, l * sizeof(jdouble)); }
static void* f_Direct(JNIEnv* e, jobject b) { (void)e; return b->data; }
static const char* f_GetUTF(JNIEnv* e, jstring s, jboolean* ic) { (void)e; if (ic) *ic = 0; return (const char*)s->data; }
static void f_RelUTF(JNIEnv* e, jstring s, const char* c) { (void)e; (void)s; (void)c; }
static jstring f_NewUTF(JNIEnv* e, const char* u) { (void)e; jstring s = calloc(1, sizeof(*s)); s->data = strdup(u); return s; }
static jlongArray f_NewLongArray(JNIEnv* e, jsize n) { (void)e; jarray a = calloc(1, sizeof(*a)); a->kind = 1; a->len = n; a->elem = 8; a->data = calloc(n, 8); return a; }
static void f_SetLongRegion(JNIEnv* e, jlongArray a, jsize s, jsize l, const jlong* b) { (void)e; memcpy((jlong*)a->data + s, b, l * sizeof(jlong)); }

static const struct JNINativeInterface_ g_iface = {f_FindClass, f_ThrowNew, f_GetObjectClass, f_GetFieldID, f_GetLongField, f_SetLongField, f_GetArrayLength,
  f_GetInt, f_GetFloat, f_GetDouble, f_RelInt, f_RelFloat, f_RelDouble, f_SetDoubleRegion, f_Direct, f_GetUTF, f_RelUTF, f_NewUTF, f_NewLongArray, f_SetLongRegion};

static struct _jobject arr(void* data, jsize len, size_t elem) { struct _jobject a = {1, len, elem, data, 0}; return a; }
static struct _jobject buf(void* data) { struct _jobject b = {2, 0, 0, data, 0}; return b; }
#define CHECK(c) do { if (!(c)) { printf("FAILED line %d: %s (thrown: %s)\n", __LINE__, #c, g_thrown); return 1; } } while (0)

/* the natives under test */
#define NA(name) Java_com_linkedin_mlease_regression_gpu_NativeAdmm_##name
jlong NA(create)(JNIEnv*, jclass, jintArray, jint, jint, jfloatArray, jfloatArray, jfloatArray, jint, jboolean, jdouble, jfloat, jboolean, jboolean);
void NA(addPartitionCsr)(JNIEnv*, jobject, jint, jlong, jobject, jobject, jobject, jobject, jobject, jobject);
void NA(begin)(JNIEnv*, jobject);
jboolean NA(iterate)(JNIEnv*, jobject, jdoubleArray);
jint NA(fitPartition)(JNIEnv*, jobject, jint, jdoubleArray, jdoubleArray, jdoubleArray);
void NA(getZ)(JNIEnv*, jobject, jint, jdoubleArray);
void NA(getX)(JNIEnv*, jobject, jint, jint, jdoubleArray);
void NA(getU)(JNIEnv*, jobject, jint, jint, jfloatArray);
void NA(getUplusx)(JNIEnv*, jobject, jint, jint, jfloatArray);
void NA(close)(JNIEnv*, jobject);
jfloat Java_com_linkedin_mlease_regression_gpu_NativeOps_testLoglik(JNIEnv*, jclass, jint, jintArray, jfloatArray, jfloatArray, jlong, jdoubleArray);
#define NI(name) Java_com_linkedin_mlease_regression_gpu_NativeIngest_##name
jlong NI(read)(JNIEnv*, jclass, jstring, jboolean, jboolean);
jlongArray NI(counts)(JNIEnv*, jobject);
void NI(get)(JNIEnv*, jobject, jobject, jobject, jobject, jobject, jobject, jobject);
jstring NI(feature)(JNIEnv*, jobject, jint);
jstring NI(key)(JNIEnv*, jobject, jlong);
void NI(close)(JNIEnv*, jobject);

int main(int argc, char** argv) {
  JNIEnv env = &g_iface;
  enum { D = 5, P = 2, L = 2, N = 4 };
  float lambdas[L] = {1.f, 10.f};
  int64_t rowptr[N + 1] = {0, 2, 3, 5, 6};
  int32_t colidx[6] = {0, 3, 1, 2, 4, 0};
  float vals[6] = {1.f, -2.f, .5f, 3.f, 1.5f, -1.f};
  int32_t resp[N] = {1, 0, 1, 0};
  struct _jobject jl = arr(lambdas, L, 4), self = {0, 0, 0, NULL, 0};
  struct _jobject brp = buf(rowptr), bci = buf(colidx), bv = buf(vals), br = buf(resp);
  /* the same world through the C ABI directly */
  mlease_admm_config cfg; memset(&cfg, 0, sizeof cfg);
  cfg.num_blocks = P; cfg.num_features = D; cfg.num_lambdas = L; cfg.lambdas = lambdas; cfg.regularizer = 2;
  mlease_world* ref = NULL;
  CHECK(mlease_world_create(&cfg, NULL, 0, &ref) == 0);
  for (int p = 0; p < P; p++) CHECK(mlease_world_add_partition_csr(ref, p, N, rowptr, colidx, vals, resp, NULL, NULL) == 0);
  CHECK(mlease_world_begin(ref) == 0);
  /* through the shim */
  self.field = NA(create)(&env, NULL, NULL, P, D, &jl, NULL, NULL, 2, 0, 1e-4, 0.f, 0, 0);
  CHECK(self.field != 0 && g_thrown[0] == 0);
  for (int p = 0; p < P; p++) NA(addPartitionCsr)(&env, &self, p, N, &brp, &bci, &bv, &br, NULL, NULL);
  NA(begin)(&env, &self);
  CHECK(g_thrown[0] == 0);
  for (int it = 0; it < 3; it++) {
    double md_ref = 0, md = -1; int32_t stop_ref = 0;
    struct _jobject jmd = arr(&md, 1, 8);
    CHECK(mlease_world_iterate(ref, &md_ref, &stop_ref) == 0);
    const jboolean stop = NA(iterate)(&env, &self, &jmd);
    CHECK(md == md_ref && (stop != 0) == (stop_ref != 0));
  }
  double z[D + 1], zr[D + 1], x[D + 1], xr[D + 1];
  float u[D + 1], ur[D + 1];
  for (int l = 0; l < L; l++) {
    memset(z, 0, sizeof z);
    struct _jobject jz = arr(z, D + 1, 8);
    NA(getZ)(&env, &self, l, &jz);
    CHECK(mlease_world_get_z(ref, l, zr) == 0 && memcmp(z, zr, sizeof z) == 0);     /* copied back: released with mode 0 */
    for (int p = 0; p < P; p++) {
      struct _jobject jx = arr(x, D + 1, 8), ju = arr(u, D + 1, 4);
      NA(getX)(&env, &self, p, l, &jx);
      CHECK(mlease_world_get_x(ref, p, l, xr) == 0 && memcmp(x, xr, sizeof x) == 0);
      NA(getU)(&env, &self, p, l, &ju);
      CHECK(mlease_world_get_u(ref, p, l, ur) == 0 && memcmp(u, ur, sizeof u) == 0);
      NA(getUplusx)(&env, &self, p, l, &ju);
      CHECK(mlease_world_get_uplusx(ref, p, l, ur) == 0 && memcmp(u, ur, sizeof u) == 0);
    }
  }
  {   /* fitPartition: x is in/out, mean and precision must come back untouched */
    double xin[D + 1] = {0}, m[D + 1] = {.1, .2, .3, .4, .5, .6}, q[D + 1] = {1, 1, 1, 1, 1, 1}, m0[D + 1], xr2[D + 1] = {0};
    memcpy(m0, m, sizeof m);
    struct _jobject jx = arr(xin, D + 1, 8), jm = arr(m, D + 1, 8), jq = arr(q, D + 1, 8);
    const jint steps = NA(fitPartition)(&env, &self, 1, &jx, &jm, &jq);
    int32_t sr = 0;
    CHECK(mlease_world_fit_partition(ref, 1, xr2, m0, q, &sr) == 0 && steps == sr && memcmp(xin, xr2, sizeof xin) == 0 && memcmp(m, m0, sizeof m) == 0);
  }
  {   /* NativeOps.testLoglik */
    float pred[N] = {.5f, -1.f, 2.f, .25f}, w[N] = {1.f, 2.f, 1.f, .5f}, ll_ref; double cnt_ref, cnt = -1;
    struct _jobject jr = arr(resp, N, 4), jp = arr(pred, N, 4), jw = arr(w, N, 4), jc = arr(&cnt, 1, 8);
    const jfloat ll = Java_com_linkedin_mlease_regression_gpu_NativeOps_testLoglik(&env, NULL, 0, &jr, &jp, &jw, 0, &jc);
    CHECK(mlease_test_loglik(0, NULL, N, resp, pred, w, 0, &ll_ref, &cnt_ref) == 0 && ll == ll_ref && cnt == cnt_ref);
  }
  NA(close)(&env, &self);
  CHECK(self.field == 0 && g_outstanding == 0 && g_thrown[0] == 0);
  if (argc > 1) {   /* NativeIngest on a prepared avro file written by the test: argv[2] records, argv[3] stored values, argv[4] features, argv[5] first feature */
    struct _jobject path = {0, 0, 0, argv[1], 0}, ing = {0, 0, 0, NULL, 0};
    ing.field = NI(read)(&env, NULL, &path, 0, 0);
    CHECK(ing.field != 0 && g_thrown[0] == 0);
    jlongArray c = NI(counts)(&env, &ing);
    const jlong* cv = (const jlong*)c->data;
    CHECK(cv[0] == atol(argv[2]) && cv[1] == atol(argv[3]) && cv[2] == atol(argv[4]));
    int64_t* rp2 = calloc((size_t)cv[0] + 1, 8); int32_t* ci2 = calloc((size_t)cv[1] + 1, 4); float* vv2 = calloc((size_t)cv[1] + 1, 4);
    struct _jobject b1 = buf(rp2), b2 = buf(ci2), b3 = buf(vv2);
    NI(get)(&env, &ing, &b1, &b2, &b3, NULL, NULL, NULL);
    CHECK(rp2[0] == 0 && rp2[cv[0]] == cv[1]);
    jstring f0 = NI(feature)(&env, &ing, 0), k0 = NI(key)(&env, &ing, 0);
    CHECK(f0 && strcmp((const char*)f0->data, argv[5]) == 0 && k0 && strlen((const char*)k0->data) > 0);
    CHECK(NI(feature)(&env, &ing, (jint)cv[2]) == NULL);
    NI(close)(&env, &ing);
    CHECK(ing.field == 0);
    path.data = "/nonexistent/file.avro";
    CHECK(NI(read)(&env, NULL, &path, 0, 0) == 0 && g_thrown[0] != 0);   /* IOException with the library's message */
    g_thrown[0] = 0;
  }
  mlease_world_destroy(ref);
  printf("JNI shim OK\n");
  return 0;
}
