package beast.evolution.likelihood;

/**
 * Native entry points of libsieve_b200.so (include/sieve_b200.h), one static method per C function.
 * The JNI glue (java/jni/sieve_jni.c) pins the Java arrays and forwards; it contains no logic.
 * The build image has no JDK (SURVEY.md section 0): tests/test_jni_shim_cpu.py compiles the glue against a minimal jni.h and
 * checks that every native method below has its Java_..._SieveB200_* export (and vice versa).
 */
public final class SieveB200 {
    static {
        System.loadLibrary("sieve_b200_jni"); // links against libsieve_b200.so
    }

    private SieveB200() {}

    public static final int FLAG_LOG_PARTIALS = 1, FLAG_CONST_PARTIALS = 2, FLAG_SINGLE_ADO = 4, FLAG_EPHEMERAL = 8, FLAG_SPARSE = 16,
            FLAG_SINGLE_LEAF_BUFFER = 32, FLAG_EXTERNAL_COUNTS = 64, FLAG_PRIVATE_SEQ_COV = 128;
    public static final int UPDATE_FLIP = 1, UPDATE_PER_LEVEL = 2;

    /** sieve_create; returns the opaque handle; throws RuntimeException(sieve_last_error()) on failure */
    public static native long create(int nLeaves, int nSites, int nCategories, int nStates, int flags, int device);
    public static native void destroy(long h);
    /** counts: [site][cell][5] flattened, exactly ScsAlignment.sitePatterns (tupleLen = 5) */
    public static native void setCounts(long h, int[] counts, int tupleLen);
    public static native void setPatternWeights(long h, int[] weights);
    public static native void computeSizeFactors(long h, int zeroCovMode, double[] outSizeFactors, double[] outStats);
    public static native void setSizeFactors(long h, double[] sizeFactors);
    public static native void setLeafParams(long h, double adoRate, double allelicSeqCov, double allelicSeqCovRawVar,
                                            double effSeqErrRate, double shapeCtrl1, double shapeCtrl2);
    /** PrivateAllelicSeqCovModel (FLAG_PRIVATE_SEQ_COV): allelicSeqCovPerPattern / allelicSeqCovRawVarPerPattern, [K * nSites] each,
     *  index matrixIndex * nrOfPatterns + patternIndex (PrivateAllelicSeqCovModel.java:16-17) */
    public static native void setPrivateSeqCov(long h, double[] allelicSeqCov, double[] allelicSeqCovRawVar);
    public static native void getPrivateSeqCov(long h, double[] allelicSeqCov, double[] allelicSeqCovRawVar);
    /** updateAllelicSeqCovAndRawVar from the max-sum trace of the current state (ScsTreeLikelihood.postProcess); returns the
     *  number of (category, pattern) pairs whose values moved; outTied[K * nSites] flags the pairs left to the host */
    public static native int privateUpdateFromMl(long h, int zeroCovMode, byte[] outTied);
    /** ScsTreeLikelihood.postProcess step 3 (:2493-2498): the patterns whose (t, v) moved are re-evaluated in place (leaves, the
     *  whole tree -- ops = (parent, child1, child2) triples in post order --, site log-likelihoods); the accepted state becomes
     *  the stored state; returns the number of re-evaluated sites.  Dense resident cores. */
    public static native int privateRetraverseChanged(long h, int[] ops);
    public static native void updateLeaves(long h, boolean forUpdate);
    public static native void setNodeMatrixForUpdate(long h, int node);
    public static native void setNodeMatrix(long h, int node, int category, double[] p16);
    /** fast path for the fixed rate matrix: distances[i * K + k] = branch length * branch rate * site rate_k of nodes[i] */
    public static native void setDistances(long h, int[] nodes, double[] distances, boolean forUpdate);
    /** max-sum pass on the current state (variant calling); see include/sieve_b200.h "max-sum pass" */
    public static native void mlTrace(long h);
    public static native void mlCall(long h, byte[] genotypes, byte[] ado, byte[] category, byte[] ambiguous);
    public static native void mlGetSite(long h, int site, byte[] back, double[] logPartials);
    public static native void setNodePartialsForUpdate(long h, int node);
    public static native void calculatePartials(long h, int child1, boolean isLeaf1, int child2, boolean isLeaf2,
                                                int parent, int constGenotype);
    public static native void setRoot(long h, int root, double[] proportions, int rootGenotype);
    /** out[5] = {sumLogL, constLse, constSum, lewisSum, weightSum} */
    public static native void evaluate(long h, double[] out5);
    public static native void getSiteLogLikelihoods(long h, double[] outLogL, double[] outLogConst);
    public static native void getNodePartials(long h, int node, int which, double[] out);
    public static native void store(long h);
    public static native void unstore(long h);
    public static native void restore(long h);
    /** sieve_step_distances: one call per MCMC state (distances of the changed branches + dirty ops [n][3] flattened) */
    public static native void stepDistances(long h, boolean updateLeaves, int[] nodes, double[] distances, int[] ops, double[] out5);
    /** global size factors over the workers' site ranges without a device that holds the full alignment (sieve_sf_*) */
    public static native long sfOpenCore(long h, int zeroCovMode);
    public static native void sfClose(long part);
    public static native void sfSolveParts(long[] parts, double[] outSizeFactors, double[] outStats);
    /** shards: nShards x 5 doubles in worker order; ascMode 0 none, 1 felsenstein/mean, 2 felsenstein/geometric, 3 lewis */
    public static native double combineShards(double[] shards, int nShards, int ascMode, double nBackgroundSites);
}
