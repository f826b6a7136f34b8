/* skeres_jni.c -- JNI shim between the Scala facade (bindings/scala) and libskeres.so.
 *
 * Replaces the SWIG-generated ceres_wrap.cc of the reference (build.sh:13-24, ceres.i).  UNTESTED on a JVM: this image
 * has no JDK; tests/test_host_logic.py only type-checks it against include/skeres.h with a mock jni.h.  The tested
 * equivalent of the same call sequence is the ctypes mirror skeres_b200/api.py.
 *
 *   cc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude bindings/jni/skeres_jni.c \
 *      -Lskeres_b200 -lskeres -o libskeres_jni.so
 *
 * Conventions: handles travel as jlong; every failure throws java.lang.RuntimeException(sk_last_error()) -- the
 * reference never lets an exception cross JNI and reports through Summary only; here errors are
 * explicit.  Bulk data crosses once per array (GetPrimitiveArrayCritical), never once per element (the reference's
 * per-element crossings are the bottleneck this library removes, AutodiffCostFunction.scala:74-134). */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "skeres.h"

#define H(type, h) ((type*)(intptr_t)(h))
#define J(p) ((jlong)(intptr_t)(p))
#define FN(name) Java_com_google_ceres_Native_##name

static void throw_last(JNIEnv* env) {
  jclass c = (*env)->FindClass(env, "java/lang/RuntimeException");
  if (c != NULL) (*env)->ThrowNew(env, c, sk_last_error());
}
static int ok(JNIEnv* env, int status) {
  if (status != SK_OK) { throw_last(env); return 0; }
  return 1;
}

/* ---- DoubleArray (ceres.i:95-107, RichDoubleArray.scala:14-75) ------------------------------------------------ */
JNIEXPORT jlong JNICALL FN(doubleArrayCreate)(JNIEnv* env, jclass cls, jlong n) {
  sk_double_array* a = NULL;
  (void)cls;
  return ok(env, sk_double_array_create((int64_t)n, &a)) ? J(a) : 0;
}
JNIEXPORT void JNICALL FN(doubleArrayDestroy)(JNIEnv* env, jclass cls, jlong h) { (void)cls; ok(env, sk_double_array_destroy(H(sk_double_array, h))); }
JNIEXPORT jlong JNICALL FN(doubleArraySize)(JNIEnv* env, jclass cls, jlong h) { (void)env; (void)cls; return (jlong)sk_double_array_size(H(sk_double_array, h)); }
JNIEXPORT jdouble JNICALL FN(doubleArrayGet)(JNIEnv* env, jclass cls, jlong h, jlong i) {
  double v = 0.0;
  (void)cls;
  ok(env, sk_double_array_get(H(sk_double_array, h), (int64_t)i, &v));
  return v;
}
JNIEXPORT void JNICALL FN(doubleArraySet)(JNIEnv* env, jclass cls, jlong h, jlong i, jdouble v) {
  (void)cls;
  ok(env, sk_double_array_set(H(sk_double_array, h), (int64_t)i, v));
}
/* copyFrom / toArray: ONE crossing per array */
JNIEXPORT void JNICALL FN(doubleArrayUpload)(JNIEnv* env, jclass cls, jlong h, jlong offset, jdoubleArray src) {
  jsize n = (*env)->GetArrayLength(env, src);
  jdouble* p = (jdouble*)(*env)->GetPrimitiveArrayCritical(env, src, NULL);
  int st;
  (void)cls;
  if (p == NULL) return;                                   /* OutOfMemoryError already pending */
  st = sk_double_array_upload(H(sk_double_array, h), (int64_t)offset, p, (int64_t)n);
  (*env)->ReleasePrimitiveArrayCritical(env, src, p, JNI_ABORT);
  ok(env, st);
}
JNIEXPORT void JNICALL FN(doubleArrayDownload)(JNIEnv* env, jclass cls, jlong h, jlong offset, jdoubleArray dst) {
  jsize n = (*env)->GetArrayLength(env, dst);
  jdouble* p = (jdouble*)(*env)->GetPrimitiveArrayCritical(env, dst, NULL);
  int st;
  (void)cls;
  if (p == NULL) return;
  st = sk_double_array_download(H(sk_double_array, h), (int64_t)offset, p, (int64_t)n);
  (*env)->ReleasePrimitiveArrayCritical(env, dst, p, 0);
  ok(env, st);
}

/* ---- LossFunction (ceres.i:160-184) ------------------------------------------------------------------------------- */
JNIEXPORT jlong JNICALL FN(lossTrivial)(JNIEnv* env, jclass cls) {
  sk_loss_function* l = NULL;
  (void)cls;
  return ok(env, sk_loss_trivial(&l)) ? J(l) : 0;
}
JNIEXPORT jlong JNICALL FN(lossHuber)(JNIEnv* env, jclass cls, jdouble a) {
  sk_loss_function* l = NULL;
  (void)cls;
  return ok(env, sk_loss_huber(a, &l)) ? J(l) : 0;
}
JNIEXPORT jlong JNICALL FN(lossCauchy)(JNIEnv* env, jclass cls, jdouble a) {
  sk_loss_function* l = NULL;
  (void)cls;
  return ok(env, sk_loss_cauchy(a, &l)) ? J(l) : 0;
}
JNIEXPORT jlong JNICALL FN(lossTolerant)(JNIEnv* env, jclass cls, jdouble a, jdouble b) {
  sk_loss_function* l = NULL;
  (void)cls;
  return ok(env, sk_loss_tolerant(a, b, &l)) ? J(l) : 0;
}
JNIEXPORT void JNICALL FN(lossDestroy)(JNIEnv* env, jclass cls, jlong h) { (void)cls; ok(env, sk_loss_destroy(H(sk_loss_function, h))); }

/* ---- CostFunction (CostFunctor.scala:31-51, AutodiffCostFunction.scala:68-134) ------------------------------------ */
JNIEXPORT jlong JNICALL FN(costFunctionCreate)(JNIEnv* env, jclass cls, jint functor, jdoubleArray consts) {
  sk_cost_function* f = NULL;
  double c[SK_MAX_CONSTS] = {0};
  jsize n = consts ? (*env)->GetArrayLength(env, consts) : 0;
  (void)cls;
  if (n > SK_MAX_CONSTS) n = SK_MAX_CONSTS + 1;             /* let the library report the mismatch */
  if (n > 0 && n <= SK_MAX_CONSTS) (*env)->GetDoubleArrayRegion(env, consts, 0, n, c);
  return ok(env, sk_cost_function_create((int)functor, c, (int)n, &f)) ? J(f) : 0;
}
JNIEXPORT void JNICALL FN(costFunctionDestroy)(JNIEnv* env, jclass cls, jlong h) { (void)cls; ok(env, sk_cost_function_destroy(H(sk_cost_function, h))); }
/* A functor given as CUDA source over T = double / Jet<N> (sk_functor_register_source): the device counterpart of subclassing
 * CostFunctor on the JVM.  Returns the functor id to pass to costFunctionCreate. */
JNIEXPORT jint JNICALL FN(functorRegisterSource)(JNIEnv* env, jclass cls, jstring name, jstring source, jint numResiduals, jintArray blockSizes,
                                                 jint numConsts) {
  int id = -1, sizes[SK_MAX_PARAMETER_BLOCKS] = {0};
  const char* cname; const char* csrc;
  jsize nblk = blockSizes ? (*env)->GetArrayLength(env, blockSizes) : 0;
  (void)cls;
  if (name == NULL || source == NULL || nblk < 1 || nblk > SK_MAX_PARAMETER_BLOCKS) { ok(env, SK_ERR_INVALID_ARGUMENT); return -1; }
  (*env)->GetIntArrayRegion(env, blockSizes, 0, nblk, (jint*)sizes);
  cname = (*env)->GetStringUTFChars(env, name, NULL);
  csrc = (*env)->GetStringUTFChars(env, source, NULL);
  ok(env, (cname && csrc) ? sk_functor_register_source(cname, csrc, (int)numResiduals, (int)nblk, sizes, (int)numConsts, &id) : SK_ERR_INTERNAL);
  if (csrc) (*env)->ReleaseStringUTFChars(env, source, csrc);
  if (cname) (*env)->ReleaseStringUTFChars(env, name, cname);
  return (jint)id;
}
/* bool evaluate(parameters, residuals, jacobians) on device-resident blocks: arrays/offsets describe the DoublePointers;
 * jacobianArrays == null <=> jacobians.isNull, a 0 handle inside it <=> jacobians.getRow(i).isNull (:80, :118) */
JNIEXPORT jboolean JNICALL FN(costFunctionEvaluate)(JNIEnv* env, jclass cls, jlong h, jlongArray paramArrays, jlongArray paramOffsets,
                                                    jlong residualArray, jlong residualOffset, jlongArray jacobianArrays,
                                                    jlongArray jacobianOffsets) {
  sk_double_pointer params[SK_MAX_PARAMETER_BLOCKS], jacs[SK_MAX_PARAMETER_BLOCKS], res;
  jlong pa[SK_MAX_PARAMETER_BLOCKS], po[SK_MAX_PARAMETER_BLOCKS], ja[SK_MAX_PARAMETER_BLOCKS], jo[SK_MAX_PARAMETER_BLOCKS];
  jsize n = (*env)->GetArrayLength(env, paramArrays), i;
  int success = 0;
  (void)cls;
  if (n > SK_MAX_PARAMETER_BLOCKS) n = SK_MAX_PARAMETER_BLOCKS;
  (*env)->GetLongArrayRegion(env, paramArrays, 0, n, pa);
  (*env)->GetLongArrayRegion(env, paramOffsets, 0, n, po);
  for (i = 0; i < n; ++i) { params[i].array = H(sk_double_array, pa[i]); params[i].offset = (int64_t)po[i]; }
  res.array = H(sk_double_array, residualArray); res.offset = (int64_t)residualOffset;
  if (jacobianArrays != NULL) {
    (*env)->GetLongArrayRegion(env, jacobianArrays, 0, n, ja);
    (*env)->GetLongArrayRegion(env, jacobianOffsets, 0, n, jo);
    for (i = 0; i < n; ++i) { jacs[i].array = H(sk_double_array, ja[i]); jacs[i].offset = (int64_t)jo[i]; }
  }
  if (!ok(env, sk_cost_function_evaluate(H(sk_cost_function, h), params, res, jacobianArrays != NULL ? jacs : NULL, &success))) return JNI_FALSE;
  return success ? JNI_TRUE : JNI_FALSE;
}

/* ---- Problem (Problem.scala:16-33) --------------------------------------------------------------------------------- */
JNIEXPORT jlong JNICALL FN(problemCreate)(JNIEnv* env, jclass cls) {
  sk_problem* p = NULL;
  (void)cls;
  return ok(env, sk_problem_create(&p)) ? J(p) : 0;
}
JNIEXPORT void JNICALL FN(problemDestroy)(JNIEnv* env, jclass cls, jlong h) { (void)cls; ok(env, sk_problem_destroy(H(sk_problem, h))); }
JNIEXPORT jlong JNICALL FN(addResidualBlock)(JNIEnv* env, jclass cls, jlong problem, jlong cost, jlong loss, jlongArray arrays,
                                             jlongArray offsets) {
  sk_double_pointer blocks[SK_MAX_PARAMETER_BLOCKS];
  jlong a[SK_MAX_PARAMETER_BLOCKS], o[SK_MAX_PARAMETER_BLOCKS];
  jsize n = (*env)->GetArrayLength(env, arrays), i;
  sk_residual_block_id id = -1;
  (void)cls;
  if (n > SK_MAX_PARAMETER_BLOCKS) n = SK_MAX_PARAMETER_BLOCKS;
  (*env)->GetLongArrayRegion(env, arrays, 0, n, a);
  (*env)->GetLongArrayRegion(env, offsets, 0, n, o);
  for (i = 0; i < n; ++i) { blocks[i].array = H(sk_double_array, a[i]); blocks[i].offset = (int64_t)o[i]; }
  ok(env, sk_problem_add_residual_block(H(sk_problem, problem), H(sk_cost_function, cost), H(sk_loss_function, loss), blocks, (int)n, &id));
  return (jlong)id;
}
/* the bulk form of the loop at SimpleBundleAdjuster.scala:139-145 */
JNIEXPORT jlong JNICALL FN(addResidualBlocks)(JNIEnv* env, jclass cls, jlong problem, jint functor, jdoubleArray consts, jlong loss,
                                              jlong array, jlongArray offsets) {
  int nres = 0, nblk = 0, sizes[SK_MAX_PARAMETER_BLOCKS], nconst = 0, st;
  jsize n_off = (*env)->GetArrayLength(env, offsets);
  jdouble* cp;
  jlong* op;
  sk_residual_block_id first = -1;
  (void)cls;
  if (!ok(env, sk_functor_info((int)functor, &nres, &nblk, sizes, &nconst))) return -1;
  cp = consts ? (jdouble*)(*env)->GetPrimitiveArrayCritical(env, consts, NULL) : NULL;
  op = (jlong*)(*env)->GetPrimitiveArrayCritical(env, offsets, NULL);
  st = (op == NULL) ? SK_ERR_INVALID_ARGUMENT
                    : sk_problem_add_residual_blocks(H(sk_problem, problem), (int)functor, (int64_t)(n_off / nblk), cp,
                                                     H(sk_loss_function, loss), H(sk_double_array, array), (const int64_t*)op, &first);
  if (op) (*env)->ReleasePrimitiveArrayCritical(env, offsets, op, JNI_ABORT);
  if (cp) (*env)->ReleasePrimitiveArrayCritical(env, consts, cp, JNI_ABORT);
  ok(env, st);
  return (jlong)first;
}
/* Problem::AddParameterBlock in bulk (the rank-local multi-GPU mode declares every camera with it) */
JNIEXPORT void JNICALL FN(addParameterBlocks)(JNIEnv* env, jclass cls, jlong problem, jlong array, jlongArray offsets, jint size) {
  jsize n = (*env)->GetArrayLength(env, offsets);
  jlong* op = (jlong*)(*env)->GetPrimitiveArrayCritical(env, offsets, NULL);
  int st;
  (void)cls;
  if (op == NULL) return;
  st = sk_problem_add_parameter_blocks(H(sk_problem, problem), H(sk_double_array, array), (int64_t)n, (const int64_t*)op, (int32_t)size);
  (*env)->ReleasePrimitiveArrayCritical(env, offsets, op, JNI_ABORT);
  ok(env, st);
}
JNIEXPORT jlong JNICALL FN(problemNumResidualBlocks)(JNIEnv* env, jclass cls, jlong h) { (void)env; (void)cls; return (jlong)sk_problem_num_residual_blocks(H(sk_problem, h)); }
JNIEXPORT jlong JNICALL FN(problemNumParameterBlocks)(JNIEnv* env, jclass cls, jlong h) { (void)env; (void)cls; return (jlong)sk_problem_num_parameter_blocks(H(sk_problem, h)); }

/* ---- Solver.Options / Solver.Summary / ceres.solve (ceres.i:151) ----------------------------------------------------- */
/* Options live in a malloc'ed POD owned by the Scala object; the setters the reference uses map to one call each. */
JNIEXPORT jlong JNICALL FN(optionsCreate)(JNIEnv* env, jclass cls) {
  sk_solver_options* o = (sk_solver_options*)malloc(sizeof *o);
  (void)env; (void)cls;
  if (o) sk_solver_options_init(o);
  return J(o);
}
JNIEXPORT void JNICALL FN(optionsDestroy)(JNIEnv* env, jclass cls, jlong h) { (void)env; (void)cls; free(H(sk_solver_options, h)); }
JNIEXPORT void JNICALL FN(optionsSetLinearSolverType)(JNIEnv* env, jclass cls, jlong h, jint v) { (void)env; (void)cls; H(sk_solver_options, h)->linear_solver_type = v; }
JNIEXPORT void JNICALL FN(optionsSetPreconditionerType)(JNIEnv* env, jclass cls, jlong h, jint v) { (void)env; (void)cls; H(sk_solver_options, h)->preconditioner_type = v; }
JNIEXPORT void JNICALL FN(optionsSetMinimizerType)(JNIEnv* env, jclass cls, jlong h, jint v) { (void)env; (void)cls; H(sk_solver_options, h)->minimizer_type = v; }
JNIEXPORT void JNICALL FN(optionsSetMaxNumIterations)(JNIEnv* env, jclass cls, jlong h, jint v) { (void)env; (void)cls; H(sk_solver_options, h)->max_num_iterations = v; }
JNIEXPORT void JNICALL FN(optionsSetMinimizerProgressToStdout)(JNIEnv* env, jclass cls, jlong h, jboolean v) { (void)env; (void)cls; H(sk_solver_options, h)->minimizer_progress_to_stdout = v ? 1 : 0; }
JNIEXPORT void JNICALL FN(optionsSetFunctionTolerance)(JNIEnv* env, jclass cls, jlong h, jdouble v) { (void)env; (void)cls; H(sk_solver_options, h)->function_tolerance = v; }
JNIEXPORT void JNICALL FN(optionsSetComm)(JNIEnv* env, jclass cls, jlong h, jlong comm) { (void)env; (void)cls; H(sk_solver_options, h)->comm = H(sk_comm, comm); }
JNIEXPORT void JNICALL FN(optionsSetResidualBlocksAreLocal)(JNIEnv* env, jclass cls, jlong h, jboolean v) { (void)env; (void)cls; H(sk_solver_options, h)->residual_blocks_are_local = v ? 1 : 0; }

JNIEXPORT jlong JNICALL FN(summaryCreate)(JNIEnv* env, jclass cls) {
  sk_solver_summary* s = NULL;
  (void)cls;
  return ok(env, sk_solver_summary_create(&s)) ? J(s) : 0;
}
JNIEXPORT void JNICALL FN(summaryDestroy)(JNIEnv* env, jclass cls, jlong h) { (void)cls; ok(env, sk_solver_summary_destroy(H(sk_solver_summary, h))); }
JNIEXPORT jstring JNICALL FN(summaryBriefReport)(JNIEnv* env, jclass cls, jlong h) { (void)cls; return (*env)->NewStringUTF(env, sk_solver_summary_brief_report(H(sk_solver_summary, h))); }
JNIEXPORT jstring JNICALL FN(summaryFullReport)(JNIEnv* env, jclass cls, jlong h) { (void)cls; return (*env)->NewStringUTF(env, sk_solver_summary_full_report(H(sk_solver_summary, h))); }
JNIEXPORT jstring JNICALL FN(summaryMessage)(JNIEnv* env, jclass cls, jlong h) { (void)cls; return (*env)->NewStringUTF(env, sk_solver_summary_message(H(sk_solver_summary, h))); }
JNIEXPORT jdouble JNICALL FN(summaryFinalCost)(JNIEnv* env, jclass cls, jlong h) {
  sk_solver_summary_data d;
  (void)cls;
  memset(&d, 0, sizeof d);
  ok(env, sk_solver_summary_get(H(sk_solver_summary, h), &d));
  return d.final_cost;
}
JNIEXPORT jint JNICALL FN(summaryTerminationType)(JNIEnv* env, jclass cls, jlong h) {
  sk_solver_summary_data d;
  (void)cls;
  memset(&d, 0, sizeof d);
  ok(env, sk_solver_summary_get(H(sk_solver_summary, h), &d));
  return (jint)d.termination_type;
}
JNIEXPORT void JNICALL FN(solve)(JNIEnv* env, jclass cls, jlong options, jlong problem, jlong summary) {
  (void)cls;
  ok(env, sk_solve(H(sk_solver_options, options), H(sk_problem, problem), H(sk_solver_summary, summary)));
}
JNIEXPORT void JNICALL FN(initGoogleLogging)(JNIEnv* env, jclass cls, jstring name) {
  const char* s = (*env)->GetStringUTFChars(env, name, NULL);
  (void)cls;
  if (s) { sk_init_google_logging(s); (*env)->ReleaseStringUTFChars(env, name, s); }
}

/* ---- multi-GPU ------------------------------------------------------------------------------------------------------- */
JNIEXPORT void JNICALL FN(setDevice)(JNIEnv* env, jclass cls, jint ordinal) { (void)cls; ok(env, sk_set_device((int)ordinal)); }
JNIEXPORT jbyteArray JNICALL FN(commUniqueId)(JNIEnv* env, jclass cls) {
  char id[SK_COMM_UNIQUE_ID_BYTES];
  jbyteArray out;
  (void)cls;
  if (!ok(env, sk_comm_get_unique_id(id))) return NULL;
  out = (*env)->NewByteArray(env, SK_COMM_UNIQUE_ID_BYTES);
  if (out) (*env)->SetByteArrayRegion(env, out, 0, SK_COMM_UNIQUE_ID_BYTES, (const jbyte*)id);
  return out;
}
JNIEXPORT jlong JNICALL FN(commCreate)(JNIEnv* env, jclass cls, jbyteArray id, jint rank, jint world) {
  char buf[SK_COMM_UNIQUE_ID_BYTES];
  sk_comm* c = NULL;
  (void)cls;
  (*env)->GetByteArrayRegion(env, id, 0, SK_COMM_UNIQUE_ID_BYTES, (jbyte*)buf);
  return ok(env, sk_comm_create(buf, (int)rank, (int)world, &c)) ? J(c) : 0;
}
JNIEXPORT void JNICALL FN(commDestroy)(JNIEnv* env, jclass cls, jlong h) { (void)cls; ok(env, sk_comm_destroy(H(sk_comm, h))); }
