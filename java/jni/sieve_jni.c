/* JNI glue for java/beast/evolution/likelihood/SieveB200.java: pins the arrays, forwards to include/sieve_b200.h,
 * turns a non-zero status into a RuntimeException carrying sieve_last_error().  No logic lives here.
 * tests/test_jni_shim_cpu.py compiles this file against a minimal jni.h (tests/jni_stub) with -Wall -Wextra -Werror and checks
 * the exports against SieveB200.java's native methods.
 * Build (needs a JDK, absent from this repository's image):
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include sieve_jni.c \
 *       -L../../sieve_b200 -lsieve_b200 -o libsieve_b200_jni.so
 */
#ifdef SIEVE_HAVE_JNI
#include <jni.h>
#include <stddef.h>
#include <stdint.h>
#include "sieve_b200.h"

#define H(h) ((sieve_core_t *)(intptr_t)(h))
static void chk(JNIEnv *e, int rc) {
    if (rc != SIEVE_OK) (*e)->ThrowNew(e, (*e)->FindClass(e, "java/lang/RuntimeException"), sieve_last_error());
}
#define FN(name) Java_beast_evolution_likelihood_SieveB200_##name

JNIEXPORT jlong JNICALL FN(create)(JNIEnv *e, jclass c, jint nl, jint ns, jint k, jint g, jint flags, jint dev) {
    sieve_config_t cfg = {nl, ns, k, g, (uint32_t)flags, dev};
    sieve_core_t *h = NULL;
    chk(e, sieve_create(&cfg, &h));
    return (jlong)(intptr_t)h;
}
JNIEXPORT void JNICALL FN(destroy)(JNIEnv *e, jclass c, jlong h) { chk(e, sieve_destroy(H(h))); }
JNIEXPORT void JNICALL FN(setCounts)(JNIEnv *e, jclass c, jlong h, jintArray a, jint tl) {
    jint *p = (*e)->GetPrimitiveArrayCritical(e, a, NULL);
    int rc = sieve_set_counts(H(h), (const int32_t *)p, tl);
    (*e)->ReleasePrimitiveArrayCritical(e, a, p, JNI_ABORT);
    chk(e, rc);
}
JNIEXPORT void JNICALL FN(setPatternWeights)(JNIEnv *e, jclass c, jlong h, jintArray a) {
    jint *p = a ? (*e)->GetPrimitiveArrayCritical(e, a, NULL) : NULL;
    int rc = sieve_set_pattern_weights(H(h), (const int32_t *)p);
    if (a) (*e)->ReleasePrimitiveArrayCritical(e, a, p, JNI_ABORT);
    chk(e, rc);
}
JNIEXPORT void JNICALL FN(computeSizeFactors)(JNIEnv *e, jclass c, jlong h, jint z, jdoubleArray sf, jdoubleArray st) {
    jdouble *p = (*e)->GetPrimitiveArrayCritical(e, sf, NULL);
    jdouble *q = st ? (*e)->GetPrimitiveArrayCritical(e, st, NULL) : NULL;
    int rc = sieve_compute_size_factors(H(h), z, p, q);
    if (st) (*e)->ReleasePrimitiveArrayCritical(e, st, q, 0);
    (*e)->ReleasePrimitiveArrayCritical(e, sf, p, 0);
    chk(e, rc);
}
JNIEXPORT void JNICALL FN(setSizeFactors)(JNIEnv *e, jclass c, jlong h, jdoubleArray a) {
    jdouble *p = (*e)->GetPrimitiveArrayCritical(e, a, NULL);
    int rc = sieve_set_size_factors(H(h), p);
    (*e)->ReleasePrimitiveArrayCritical(e, a, p, JNI_ABORT);
    chk(e, rc);
}
JNIEXPORT void JNICALL FN(setLeafParams)(JNIEnv *e, jclass c, jlong h, jdouble th, jdouble t, jdouble v, jdouble eps,
                                         jdouble w1, jdouble w2) {
    sieve_leaf_params_t p = {th, t, v, eps, w1, w2};
    chk(e, sieve_set_leaf_params(H(h), &p));
}
JNIEXPORT void JNICALL FN(updateLeaves)(JNIEnv *e, jclass c, jlong h, jboolean f) { chk(e, sieve_update_leaves(H(h), f)); }
JNIEXPORT void JNICALL FN(setNodeMatrixForUpdate)(JNIEnv *e, jclass c, jlong h, jint n) {
    chk(e, sieve_set_node_matrix_for_update(H(h), n));
}
JNIEXPORT void JNICALL FN(setNodeMatrix)(JNIEnv *e, jclass c, jlong h, jint n, jint k, jdoubleArray a) {
    jdouble *p = (*e)->GetPrimitiveArrayCritical(e, a, NULL);
    int rc = sieve_set_node_matrix(H(h), n, k, p);
    (*e)->ReleasePrimitiveArrayCritical(e, a, p, JNI_ABORT);
    chk(e, rc);
}
JNIEXPORT void JNICALL FN(setDistances)(JNIEnv *e, jclass c, jlong h, jintArray nodes, jdoubleArray d, jboolean f) {
    const jsize n = (*e)->GetArrayLength(e, nodes);
    jint *pn = (*e)->GetPrimitiveArrayCritical(e, nodes, NULL);
    jdouble *pd = (*e)->GetPrimitiveArrayCritical(e, d, NULL);
    int rc = sieve_set_distances(H(h), n, (const int32_t *)pn, pd, f);
    (*e)->ReleasePrimitiveArrayCritical(e, d, pd, JNI_ABORT);
    (*e)->ReleasePrimitiveArrayCritical(e, nodes, pn, JNI_ABORT);
    chk(e, rc);
}
JNIEXPORT void JNICALL FN(setPrivateSeqCov)(JNIEnv *e, jclass c, jlong h, jdoubleArray t, jdoubleArray v) {
    jdouble *pt = (*e)->GetPrimitiveArrayCritical(e, t, NULL), *pv = (*e)->GetPrimitiveArrayCritical(e, v, NULL);
    int rc = sieve_set_private_seq_cov(H(h), pt, pv);
    (*e)->ReleasePrimitiveArrayCritical(e, v, pv, JNI_ABORT);
    (*e)->ReleasePrimitiveArrayCritical(e, t, pt, JNI_ABORT);
    chk(e, rc);
}
JNIEXPORT void JNICALL FN(getPrivateSeqCov)(JNIEnv *e, jclass c, jlong h, jdoubleArray t, jdoubleArray v) {
    jdouble *pt = (*e)->GetPrimitiveArrayCritical(e, t, NULL), *pv = (*e)->GetPrimitiveArrayCritical(e, v, NULL);
    int rc = sieve_get_private_seq_cov(H(h), pt, pv);
    (*e)->ReleasePrimitiveArrayCritical(e, v, pv, 0);
    (*e)->ReleasePrimitiveArrayCritical(e, t, pt, 0);
    chk(e, rc);
}
JNIEXPORT jint JNICALL FN(privateUpdateFromMl)(JNIEnv *e, jclass c, jlong h, jint zeroCovMode, jbyteArray tied) {
    jbyte *pt = (*e)->GetPrimitiveArrayCritical(e, tied, NULL);
    int32_t changed = 0;
    int rc = sieve_private_update_from_ml(H(h), zeroCovMode, &changed, (uint8_t *)pt);
    (*e)->ReleasePrimitiveArrayCritical(e, tied, pt, 0);
    chk(e, rc);
    return changed;
}
JNIEXPORT jint JNICALL FN(privateRetraverseChanged)(JNIEnv *e, jclass c, jlong h, jintArray ops) {
    const jsize n = (*e)->GetArrayLength(e, ops) / 3;
    jint *po = (*e)->GetPrimitiveArrayCritical(e, ops, NULL);
    int32_t sites = 0;
    int rc = sieve_private_retraverse_changed(H(h), (int32_t)n, (const int32_t *)po, &sites);
    (*e)->ReleasePrimitiveArrayCritical(e, ops, po, JNI_ABORT);
    chk(e, rc);
    return sites;
}
JNIEXPORT void JNICALL FN(mlTrace)(JNIEnv *e, jclass c, jlong h) { chk(e, sieve_ml_trace(H(h))); }
JNIEXPORT void JNICALL FN(mlCall)(JNIEnv *e, jclass c, jlong h, jbyteArray g, jbyteArray a, jbyteArray k, jbyteArray m) {
    jbyte *pg = (*e)->GetPrimitiveArrayCritical(e, g, NULL), *pa = (*e)->GetPrimitiveArrayCritical(e, a, NULL);
    jbyte *pk = (*e)->GetPrimitiveArrayCritical(e, k, NULL), *pm = (*e)->GetPrimitiveArrayCritical(e, m, NULL);
    int rc = sieve_ml_call(H(h), (int8_t *)pg, (int8_t *)pa, (uint8_t *)pk, (uint8_t *)pm);
    (*e)->ReleasePrimitiveArrayCritical(e, m, pm, 0);
    (*e)->ReleasePrimitiveArrayCritical(e, k, pk, 0);
    (*e)->ReleasePrimitiveArrayCritical(e, a, pa, 0);
    (*e)->ReleasePrimitiveArrayCritical(e, g, pg, 0);
    chk(e, rc);
}
JNIEXPORT void JNICALL FN(mlGetSite)(JNIEnv *e, jclass c, jlong h, jint site, jbyteArray b, jdoubleArray p) {
    jbyte *pb = (*e)->GetPrimitiveArrayCritical(e, b, NULL);
    jdouble *pp = (*e)->GetPrimitiveArrayCritical(e, p, NULL);
    int rc = sieve_ml_get_site(H(h), site, (uint8_t *)pb, pp);
    (*e)->ReleasePrimitiveArrayCritical(e, p, pp, 0);
    (*e)->ReleasePrimitiveArrayCritical(e, b, pb, 0);
    chk(e, rc);
}
JNIEXPORT void JNICALL FN(setNodePartialsForUpdate)(JNIEnv *e, jclass c, jlong h, jint n) {
    chk(e, sieve_set_node_partials_for_update(H(h), n));
}
JNIEXPORT void JNICALL FN(calculatePartials)(JNIEnv *e, jclass c, jlong h, jint c1, jboolean l1, jint c2, jboolean l2,
                                             jint p, jint cg) {
    chk(e, sieve_calculate_partials(H(h), c1, l1, c2, l2, p, cg));
}
JNIEXPORT void JNICALL FN(setRoot)(JNIEnv *e, jclass c, jlong h, jint r, jdoubleArray a, jint rg) {
    jdouble *p = (*e)->GetPrimitiveArrayCritical(e, a, NULL);
    int rc = sieve_set_root(H(h), r, p, rg);
    (*e)->ReleasePrimitiveArrayCritical(e, a, p, JNI_ABORT);
    chk(e, rc);
}
JNIEXPORT void JNICALL FN(evaluate)(JNIEnv *e, jclass c, jlong h, jdoubleArray a) {
    sieve_shard_result_t r;
    int rc = sieve_evaluate(H(h), &r);
    if (rc == SIEVE_OK) (*e)->SetDoubleArrayRegion(e, a, 0, 5, (const jdouble *)&r);
    chk(e, rc);
}
JNIEXPORT void JNICALL FN(getSiteLogLikelihoods)(JNIEnv *e, jclass c, jlong h, jdoubleArray a, jdoubleArray b) {
    jdouble *p = (*e)->GetPrimitiveArrayCritical(e, a, NULL);
    jdouble *q = b ? (*e)->GetPrimitiveArrayCritical(e, b, NULL) : NULL;
    int rc = sieve_get_site_log_likelihoods(H(h), p, q);
    if (b) (*e)->ReleasePrimitiveArrayCritical(e, b, q, 0);
    (*e)->ReleasePrimitiveArrayCritical(e, a, p, 0);
    chk(e, rc);
}
JNIEXPORT void JNICALL FN(getNodePartials)(JNIEnv *e, jclass c, jlong h, jint n, jint w, jdoubleArray a) {
    jdouble *p = (*e)->GetPrimitiveArrayCritical(e, a, NULL);
    int rc = sieve_get_node_partials(H(h), n, w, p);
    (*e)->ReleasePrimitiveArrayCritical(e, a, p, 0);
    chk(e, rc);
}
JNIEXPORT void JNICALL FN(store)(JNIEnv *e, jclass c, jlong h) { chk(e, sieve_store(H(h))); }
JNIEXPORT void JNICALL FN(unstore)(JNIEnv *e, jclass c, jlong h) { chk(e, sieve_unstore(H(h))); }
JNIEXPORT void JNICALL FN(restore)(JNIEnv *e, jclass c, jlong h) { chk(e, sieve_restore(H(h))); }
JNIEXPORT jdouble JNICALL FN(combineShards)(JNIEnv *e, jclass c, jdoubleArray a, jint n, jint mode, jdouble M) {
    jdouble *p = (*e)->GetPrimitiveArrayCritical(e, a, NULL);
    double r = sieve_combine_shards((const sieve_shard_result_t *)p, n, mode, M);
    (*e)->ReleasePrimitiveArrayCritical(e, a, p, JNI_ABORT);
    return r;
}
JNIEXPORT void JNICALL FN(stepDistances)(JNIEnv *e, jclass c, jlong h, jboolean leaves, jintArray nodes, jdoubleArray d,
                                         jintArray ops, jdoubleArray out5) {
    const jsize nn = (*e)->GetArrayLength(e, nodes), no = (*e)->GetArrayLength(e, ops) / 3;
    jint *pn = (*e)->GetPrimitiveArrayCritical(e, nodes, NULL);
    jdouble *pd = (*e)->GetPrimitiveArrayCritical(e, d, NULL);
    jint *po = (*e)->GetPrimitiveArrayCritical(e, ops, NULL);
    sieve_shard_result_t r;
    int rc = sieve_step_distances(H(h), leaves, nn, (const int32_t *)pn, pd, no, (const int32_t *)po, &r);
    (*e)->ReleasePrimitiveArrayCritical(e, ops, po, JNI_ABORT);
    (*e)->ReleasePrimitiveArrayCritical(e, d, pd, JNI_ABORT);
    (*e)->ReleasePrimitiveArrayCritical(e, nodes, pn, JNI_ABORT);
    if (rc == SIEVE_OK) (*e)->SetDoubleArrayRegion(e, out5, 0, 5, (const jdouble *)&r);
    chk(e, rc);
}
/* size factors over the site shards of one JVM (one part per worker / GPU): sieve_sf_* */
JNIEXPORT jlong JNICALL FN(sfOpenCore)(JNIEnv *e, jclass c, jlong h, jint zeroCovMode) {
    sieve_sf_t *p = NULL;
    chk(e, sieve_sf_open_core(H(h), zeroCovMode, &p));
    return (jlong)(intptr_t)p;
}
JNIEXPORT void JNICALL FN(sfClose)(JNIEnv *e, jclass c, jlong p) { chk(e, sieve_sf_close((sieve_sf_t *)(intptr_t)p)); }
JNIEXPORT void JNICALL FN(sfSolveParts)(JNIEnv *e, jclass c, jlongArray parts, jdoubleArray sf, jdoubleArray st) {
    enum { MAX_PARTS = 64 };
    sieve_sf_t *pp[MAX_PARTS];
    const jsize n = (*e)->GetArrayLength(e, parts);
    if (n < 1 || n > MAX_PARTS) {
        (*e)->ThrowNew(e, (*e)->FindClass(e, "java/lang/IllegalArgumentException"), "sfSolveParts: 1..64 parts");
        return;
    }
    jlong *pl = (*e)->GetPrimitiveArrayCritical(e, parts, NULL);
    for (jsize i = 0; i < n; i++) pp[i] = (sieve_sf_t *)(intptr_t)pl[i];
    (*e)->ReleasePrimitiveArrayCritical(e, parts, pl, JNI_ABORT);
    jdouble *p = (*e)->GetPrimitiveArrayCritical(e, sf, NULL);
    jdouble *q = st ? (*e)->GetPrimitiveArrayCritical(e, st, NULL) : NULL;
    int rc = sieve_sf_solve_parts(n, pp, NULL, NULL, p, q);
    if (st) (*e)->ReleasePrimitiveArrayCritical(e, st, q, 0);
    (*e)->ReleasePrimitiveArrayCritical(e, sf, p, 0);
    chk(e, rc);
}
#endif /* SIEVE_HAVE_JNI */
