/*
 * vrec_jni.c -- JNI binding of libvrec.so for the JVM side of the Visit Recommender.
 *
 * One stub per native method of com.github.tashoyan.recommender.gpu.Vrec (jni/scala/.../gpu/Vrec.scala); every
 * stub pins its arrays with GetPrimitiveArrayCritical, calls the C ABI of include/vrec.h and releases the arrays
 * (JNI_ABORT for inputs: nothing to copy back).  Status codes are returned to Scala unchanged; the Scala side
 * turns VREC_ENOENT / VREC_EINVAL into the IllegalArgumentException the reference throws
 * (knn/KnnRecommender.scala:17-20,83; stochastic/StochasticRecommender.scala:33-34,70).
 *
 * Build on a machine with a JDK (this repository's image has none, so CI only checks that the file compiles
 * against include/vrec.h with a stand-in jni.h -- tests/test_jni_shim_cpu.py):
 *   gcc -O2 -fPIC -shared -I"$JAVA_HOME/include" -I"$JAVA_HOME/include/linux" -I../include \
 *       -o libvrec_jni.so vrec_jni.c -L../locations-recommender_b200 -lvrec -Wl,-rpath,'$ORIGIN'
 */
#include <jni.h>
#include <stddef.h>
#include <stdint.h>

#include "vrec.h"

#define VREC_JNI(ret, name) JNIEXPORT ret JNICALL Java_com_github_tashoyan_recommender_gpu_Vrec_00024_##name

/* pin / unpin helpers; a NULL Java array is a NULL pointer */
static void *pin(JNIEnv *env, jarray a) { return a ? (*env)->GetPrimitiveArrayCritical(env, a, NULL) : NULL; }
static void unpin_in(JNIEnv *env, jarray a, void *p) { if (a) (*env)->ReleasePrimitiveArrayCritical(env, a, p, JNI_ABORT); }
static void unpin_out(JNIEnv *env, jarray a, void *p) { if (a) (*env)->ReleasePrimitiveArrayCritical(env, a, p, 0); }
static jsize len(JNIEnv *env, jarray a) { return a ? (*env)->GetArrayLength(env, a) : 0; }

/* ------------------------------------------------------------------ context */
VREC_JNI(jlong, init)(JNIEnv *env, jobject self, jint device) {
    vrec_ctx *ctx = NULL;
    (void)env; (void)self;
    return vrec_init((int)device, &ctx) == VREC_OK ? (jlong)(intptr_t)ctx : 0;
}

VREC_JNI(void, shutdown)(JNIEnv *env, jobject self, jlong ctx) {
    (void)env; (void)self;
    vrec_shutdown((vrec_ctx *)(intptr_t)ctx);
}

VREC_JNI(jstring, lastError)(JNIEnv *env, jobject self) {
    (void)self;
    return (*env)->NewStringUTF(env, vrec_last_error());
}

/* ------------------------------------------------------------------ KNN path */
VREC_JNI(jlong, knnLoad)(JNIEnv *env, jobject self, jlong ctx, jlongArray personId, jlongArray placeRowPtr,
                         jintArray placeCol, jdoubleArray placeVal, jint placeDim, jlongArray catRowPtr,
                         jintArray catCol, jdoubleArray catVal, jint catDim, jlongArray ratingPerson,
                         jlongArray ratingPlace, jlongArray ratingValue) {
    (void)self;
    const jsize P = len(env, personId), nr = len(env, ratingPerson);
    jlong *pid = pin(env, personId), *prp = pin(env, placeRowPtr), *crp = pin(env, catRowPtr);
    jint *pc = pin(env, placeCol), *cc = pin(env, catCol);
    jdouble *pv = pin(env, placeVal), *cv = pin(env, catVal);
    jlong *rp = pin(env, ratingPerson), *rl = pin(env, ratingPlace), *rv = pin(env, ratingValue);
    vrec_knn *knn = NULL;
    int rc = vrec_knn_load((vrec_ctx *)(intptr_t)ctx, (int64_t)P, (const int64_t *)pid, (const int64_t *)prp,
                           (const int32_t *)pc, pv, (int32_t)placeDim, (const int64_t *)crp, (const int32_t *)cc, cv,
                           (int32_t)catDim, (int64_t)nr, (const int64_t *)rp, (const int64_t *)rl,
                           (const int64_t *)rv, &knn);
    unpin_in(env, ratingValue, rv); unpin_in(env, ratingPlace, rl); unpin_in(env, ratingPerson, rp);
    unpin_in(env, catVal, cv); unpin_in(env, placeVal, pv); unpin_in(env, catCol, cc); unpin_in(env, placeCol, pc);
    unpin_in(env, catRowPtr, crp); unpin_in(env, placeRowPtr, prp); unpin_in(env, personId, pid);
    return rc == VREC_OK ? (jlong)(intptr_t)knn : (jlong)rc;          /* handles are > 0, codes < 0 */
}

VREC_JNI(void, knnFree)(JNIEnv *env, jobject self, jlong knn) {
    (void)env; (void)self;
    vrec_knn_free((vrec_knn *)(intptr_t)knn);
}

VREC_JNI(jint, knnQuery)(JNIEnv *env, jobject self, jlong knn, jlongArray targets, jdouble placeWeight,
                         jdouble categoryWeight, jint kNearest, jlongArray placeFilter, jint maxRecs,
                         jlongArray outPlace, jdoubleArray outRating, jintArray outCount, jintArray outStatus) {
    (void)self;
    const jsize n = len(env, targets), nf = len(env, placeFilter);
    jlong *t = pin(env, targets), *f = pin(env, placeFilter), *op = pin(env, outPlace);
    jdouble *orr = pin(env, outRating);
    jint *oc = pin(env, outCount), *os = pin(env, outStatus);
    int rc = vrec_knn_query((vrec_knn *)(intptr_t)knn, (const int64_t *)t, (int32_t)n, placeWeight, categoryWeight,
                            (int32_t)kNearest, (const int64_t *)f, (int64_t)nf, (int32_t)maxRecs, (int64_t *)op, orr,
                            (int32_t *)oc, (int32_t *)os);
    unpin_out(env, outStatus, os); unpin_out(env, outCount, oc); unpin_out(env, outRating, orr);
    unpin_out(env, outPlace, op); unpin_in(env, placeFilter, f); unpin_in(env, targets, t);
    return (jint)rc;
}

/* returns the number of rows (>= 0; == capacity means "call again with larger arrays") or a VREC_E* code */
VREC_JNI(jlong, knnEstimates)(JNIEnv *env, jobject self, jlong knn, jlong target, jdouble placeWeight,
                              jdouble categoryWeight, jint kNearest, jlongArray outPlace, jdoubleArray outRating) {
    (void)self;
    const jsize cap = len(env, outPlace);
    jlong *op = pin(env, outPlace);
    jdouble *orr = pin(env, outRating);
    int64_t n = 0;
    int rc = vrec_knn_estimates((vrec_knn *)(intptr_t)knn, (int64_t)target, placeWeight, categoryWeight,
                                (int32_t)kNearest, (int64_t *)op, orr, (int64_t)cap, &n);
    unpin_out(env, outRating, orr); unpin_out(env, outPlace, op);
    return rc == VREC_OK ? (jlong)n : (jlong)rc;
}

VREC_JNI(jint, knnNeighbours)(JNIEnv *env, jobject self, jlong knn, jlong target, jdouble placeWeight,
                              jdouble categoryWeight, jint kNearest, jlongArray outPerson, jdoubleArray outSimilarity) {
    (void)self;
    const jsize cap = len(env, outPerson);
    jlong *op = pin(env, outPerson);
    jdouble *os = pin(env, outSimilarity);
    int32_t n = 0;
    int rc = vrec_knn_neighbours((vrec_knn *)(intptr_t)knn, (int64_t)target, placeWeight, categoryWeight,
                                 (int32_t)kNearest, (int64_t *)op, os, (int32_t)cap, &n);
    unpin_out(env, outSimilarity, os); unpin_out(env, outPerson, op);
    return rc == VREC_OK ? (jint)n : (jint)rc;
}

/* ------------------------------------------------------------------ SG path */
VREC_JNI(jlong, sgLoad)(JNIEnv *env, jobject self, jlong ctx, jlongArray source, jlongArray target,
                        jdoubleArray weight) {
    (void)self;
    const jsize nnz = len(env, source);
    jlong *s = pin(env, source), *t = pin(env, target);
    jdouble *w = pin(env, weight);
    vrec_sg *sg = NULL;
    int rc = vrec_sg_load((vrec_ctx *)(intptr_t)ctx, (int64_t)nnz, (const int64_t *)s, (const int64_t *)t, w, &sg);
    unpin_in(env, weight, w); unpin_in(env, target, t); unpin_in(env, source, s);
    return rc == VREC_OK ? (jlong)(intptr_t)sg : (jlong)rc;
}

VREC_JNI(void, sgFree)(JNIEnv *env, jobject self, jlong sg) {
    (void)env; (void)self;
    vrec_sg_free((vrec_sg *)(intptr_t)sg);
}

VREC_JNI(jlong, sgVertexCount)(JNIEnv *env, jobject self, jlong sg) {
    (void)env; (void)self;
    return (jlong)vrec_sg_vertex_count((vrec_sg *)(intptr_t)sg);
}

VREC_JNI(jint, sgVertexIds)(JNIEnv *env, jobject self, jlong sg, jlongArray outIds) {
    (void)self;
    jlong *o = pin(env, outIds);
    int rc = vrec_sg_vertex_ids((vrec_sg *)(intptr_t)sg, (int64_t *)o);
    unpin_out(env, outIds, o);
    return (jint)rc;
}

/* outInfo = {iterations, converged}; outX[vertex_count] in vertex-id order */
VREC_JNI(jint, sgStationary)(JNIEnv *env, jobject self, jlong sg, jlong vertex, jdouble epsilon, jint maxIterations,
                             jdoubleArray outX, jintArray outInfo) {
    (void)self;
    jdouble *x = pin(env, outX);
    jint *info = pin(env, outInfo);
    int32_t it = 0, conv = 0;
    int rc = vrec_sg_stationary((vrec_sg *)(intptr_t)sg, (int64_t)vertex, epsilon, (int32_t)maxIterations, x, &it,
                                &conv, NULL);
    if (info) { info[0] = it; info[1] = conv; }
    unpin_out(env, outInfo, info); unpin_out(env, outX, x);
    return (jint)rc;
}

VREC_JNI(jint, sgQuery)(JNIEnv *env, jobject self, jlong sg, jlongArray vertices, jdouble epsilon,
                        jint maxIterations, jlongArray placeFilter, jint maxRecs, jlongArray outId,
                        jdoubleArray outProb, jintArray outCount, jintArray outIterations, jintArray outConverged,
                        jintArray outStatus) {
    (void)self;
    const jsize n = len(env, vertices), nf = len(env, placeFilter);
    jlong *v = pin(env, vertices), *f = pin(env, placeFilter), *oi = pin(env, outId);
    jdouble *opr = pin(env, outProb);
    jint *oc = pin(env, outCount), *oit = pin(env, outIterations), *ocv = pin(env, outConverged);
    jint *os = pin(env, outStatus);
    int rc = vrec_sg_query((vrec_sg *)(intptr_t)sg, (const int64_t *)v, (int32_t)n, epsilon, (int32_t)maxIterations,
                           (const int64_t *)f, (int64_t)nf, (int32_t)maxRecs, (int64_t *)oi, opr, (int32_t *)oc,
                           (int32_t *)oit, (int32_t *)ocv, (int32_t *)os);
    unpin_out(env, outStatus, os); unpin_out(env, outConverged, ocv); unpin_out(env, outIterations, oit);
    unpin_out(env, outCount, oc); unpin_out(env, outProb, opr); unpin_out(env, outId, oi);
    unpin_in(env, placeFilter, f); unpin_in(env, vertices, v);
    return (jint)rc;
}

/* ------------------------------------------------------------------ builders */
/* all three: return the number of output rows, or a VREC_E* code; VREC_ENOMEM = arrays too small, the needed
 * size is then in outN[0] */
VREC_JNI(jlong, buildPlaceVisits)(JNIEnv *env, jobject self, jlong ctx, jlongArray person, jdoubleArray lat,
                                  jdoubleArray lon, jlongArray timestampMs, jlongArray region, jlongArray placeId,
                                  jdoubleArray placeLat, jdoubleArray placeLon, jlongArray placeCategory,
                                  jlongArray placeRegion, jint lastDaysCount, jdouble accuracyMeters,
                                  jlongArray outPerson, jlongArray outTimestampMs, jlongArray outPlace,
                                  jlongArray outRegion, jlongArray outCategory, jlongArray outN) {
    (void)self;
    const jsize nv = len(env, person), np = len(env, placeId), cap = len(env, outPerson);
    jlong *pe = pin(env, person), *ts = pin(env, timestampMs), *rg = pin(env, region), *pid = pin(env, placeId);
    jlong *pcat = pin(env, placeCategory), *preg = pin(env, placeRegion);
    jdouble *la = pin(env, lat), *lo = pin(env, lon), *pla = pin(env, placeLat), *plo = pin(env, placeLon);
    jlong *o1 = pin(env, outPerson), *o2 = pin(env, outTimestampMs), *o3 = pin(env, outPlace);
    jlong *o4 = pin(env, outRegion), *o5 = pin(env, outCategory), *on = pin(env, outN);
    int64_t n = 0;
    int rc = vrec_build_place_visits((vrec_ctx *)(intptr_t)ctx, (int64_t)nv, (const int64_t *)pe, la, lo,
                                     (const int64_t *)ts, (const int64_t *)rg, (int64_t)np, (const int64_t *)pid,
                                     pla, plo, (const int64_t *)pcat, (const int64_t *)preg, (int32_t)lastDaysCount,
                                     accuracyMeters, (int64_t)cap, &n, (int64_t *)o1, (int64_t *)o2, (int64_t *)o3,
                                     (int64_t *)o4, (int64_t *)o5);
    if (on) on[0] = (jlong)n;
    unpin_out(env, outN, on); unpin_out(env, outCategory, o5); unpin_out(env, outRegion, o4);
    unpin_out(env, outPlace, o3); unpin_out(env, outTimestampMs, o2); unpin_out(env, outPerson, o1);
    unpin_in(env, placeLon, plo); unpin_in(env, placeLat, pla); unpin_in(env, lon, lo); unpin_in(env, lat, la);
    unpin_in(env, placeRegion, preg); unpin_in(env, placeCategory, pcat); unpin_in(env, placeId, pid);
    unpin_in(env, region, rg); unpin_in(env, timestampMs, ts); unpin_in(env, person, pe);
    return rc == VREC_OK ? (jlong)n : (jlong)rc;
}

/* outInfo = {persons, non-zeros, vector size (max entity id + 1)} */
VREC_JNI(jint, buildRatingVectors)(JNIEnv *env, jobject self, jlong ctx, jlongArray person, jlongArray entity,
                                   jlongArray weight, jint topN, jlongArray outPerson, jlongArray outRowPtr,
                                   jintArray outCol, jdoubleArray outVal, jlongArray outInfo) {
    (void)self;
    const jsize n = len(env, person);
    jlong *pe = pin(env, person), *en = pin(env, entity), *we = pin(env, weight);
    jlong *op = pin(env, outPerson), *orp = pin(env, outRowPtr), *info = pin(env, outInfo);
    jint *oc = pin(env, outCol);
    jdouble *ov = pin(env, outVal);
    int64_t n_persons = 0, nnz = 0;
    int32_t dim = 0;
    int rc = vrec_build_rating_vectors((vrec_ctx *)(intptr_t)ctx, (int64_t)n, (const int64_t *)pe, (const int64_t *)en,
                                       (const int64_t *)we, (int32_t)topN, &n_persons, &nnz, (int64_t *)op,
                                       (int64_t *)orp, (int32_t *)oc, ov, &dim);
    if (info) { info[0] = (jlong)n_persons; info[1] = (jlong)nnz; info[2] = (jlong)dim; }
    unpin_out(env, outVal, ov); unpin_out(env, outCol, oc); unpin_out(env, outInfo, info);
    unpin_out(env, outRowPtr, orp); unpin_out(env, outPerson, op);
    unpin_in(env, weight, we); unpin_in(env, entity, en); unpin_in(env, person, pe);
    return (jint)rc;
}

VREC_JNI(jlong, buildStochasticGraph)(JNIEnv *env, jobject self, jlong ctx, jlongArray person, jlongArray place,
                                      jlongArray category, jlongArray timestampMs, jdouble betaPersonPlace,
                                      jdouble betaPersonCategory, jlongArray outSource, jlongArray outTarget,
                                      jdoubleArray outWeight, jlongArray outN) {
    (void)self;
    const jsize n = len(env, person), cap = len(env, outSource);
    jlong *pe = pin(env, person), *pl = pin(env, place), *ca = pin(env, category), *ts = pin(env, timestampMs);
    jlong *os = pin(env, outSource), *ot = pin(env, outTarget), *on = pin(env, outN);
    jdouble *ow = pin(env, outWeight);
    int64_t ne = 0;
    int rc = vrec_build_stochastic_graph((vrec_ctx *)(intptr_t)ctx, (int64_t)n, (const int64_t *)pe,
                                         (const int64_t *)pl, (const int64_t *)ca, (const int64_t *)ts,
                                         betaPersonPlace, betaPersonCategory, (int64_t)cap, &ne, (int64_t *)os,
                                         (int64_t *)ot, ow);
    if (on) on[0] = (jlong)ne;
    unpin_out(env, outN, on); unpin_out(env, outWeight, ow); unpin_out(env, outTarget, ot);
    unpin_out(env, outSource, os);
    unpin_in(env, timestampMs, ts); unpin_in(env, category, ca); unpin_in(env, place, pl); unpin_in(env, person, pe);
    return rc == VREC_OK ? (jlong)ne : (jlong)rc;
}
