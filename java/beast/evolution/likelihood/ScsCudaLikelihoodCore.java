package beast.evolution.likelihood;

/**
 * Drop-in for ScsBeerLikelihoodCore (likelihood/ScsBeerLikelihoodCore.java) backed by libsieve_b200.so.
 *
 * It overrides exactly the methods ScsTreeLikelihood.traverse / calcLogP call on the core during MCMC
 * (likelihood/ScsTreeLikelihood.java:430-468, 545-935, 2827-2880), so the traverse, the dirty tracking and the
 * store / restore logic of the reference run unchanged.  Two differences a maintainer has to wire (INTEGRATION.md):
 *   - leaves are computed on the device: setNodePartials(leaf, ...) is replaced by updateLeaves();
 *   - calculateLogPartials only queues; integratePartials + calculateLogLikelihoods trigger the single launch.
 * Not compiled in this repository's CI (no JDK in the build image); the JNI glue below it is syntax- and symbol-checked.
 */
public class ScsCudaLikelihoodCore extends ScsBeerLikelihoodCore {
    private long h;
    private int patternCount;
    private final double[] result = new double[5];
    private boolean evaluated;
    private boolean singleADO = true, ascBias = true, singleLeafBuffer = false;

    /**
     * The model switches the device needs, read where ScsTreeLikelihood.initAndValidate has them (before initCore):
     * singleADO from the raw-read-counts model (rawreadcountsmodel/RawReadCountsModelInterface.java:50-54, 178-181; an XML
     * input, default true) and whether the alignment asks for acquisition-bias correction
     * (alignment/ScsAlignment.java:697-699 isAscBiasCorrection): without it no constant-site work is done at all.
     */
    public void configure(boolean singleADO, boolean ascBiasCorrection) {
        this.singleADO = singleADO;
        this.ascBias = ascBiasCorrection;
        this.singleLeafBuffer = Boolean.getBoolean("sieve.b200.singleLeafBuffer");
    }

    public ScsCudaLikelihoodCore(int nrOfStates) {
        super(nrOfStates);
    }

    /** ScsBeerLikelihoodCore.initialize (:91-128); the device index is read from -Dsieve.b200.device (default 0) */
    @Override
    public void initialize(int nodeCount, int leafNodeCount, int internalNodeCount, int patternCount, int matrixCount,
                           int stateCount, boolean integrateCategories, boolean useLogPartials) {
        this.patternCount = patternCount;
        int flags = (useLogPartials ? SieveB200.FLAG_LOG_PARTIALS : 0) | (ascBias ? SieveB200.FLAG_CONST_PARTIALS : 0)
                | (singleADO ? SieveB200.FLAG_SINGLE_ADO : 0)
                | (singleLeafBuffer ? SieveB200.FLAG_SINGLE_LEAF_BUFFER : 0)
                | (Boolean.getBoolean("sieve.b200.sparse") ? SieveB200.FLAG_SPARSE : 0)        // 10 000-cell data sets
                | (Boolean.getBoolean("sieve.b200.streamed") ? SieveB200.FLAG_EPHEMERAL : 0);  // beyond that
        h = SieveB200.create(leafNodeCount, patternCount, matrixCount, stateCount, flags,
                Integer.getInteger("sieve.b200.device", 0));
    }

    /** called once from initCore instead of the per-leaf initializeLeafLikelihood loop (:436-455) */
    public void uploadData(int[] sitePatternsFlat, int[] patternWeights, double[] sizeFactors) {
        SieveB200.setCounts(h, sitePatternsFlat, 5);
        SieveB200.setPatternWeights(h, patternWeights);
        SieveB200.setSizeFactors(h, sizeFactors);
    }

    /** replaces computeLeafLikelihood + setNodePartials for every leaf (:583-601) */
    public void updateLeaves(double adoRate, double allelicSeqCov, double allelicSeqCovRawVar, double effSeqErrRate,
                             double shapeCtrl1, double shapeCtrl2, boolean forUpdate) {
        SieveB200.setLeafParams(h, adoRate, allelicSeqCov, allelicSeqCovRawVar, effSeqErrRate, shapeCtrl1, shapeCtrl2);
        SieveB200.updateLeaves(h, forUpdate);
    }

    @Override public void createNodePartials(int nodeIndex) { /* pools are allocated in initialize */ }
    @Override public void setNodeMatrixForUpdate(int nodeIndex) { SieveB200.setNodeMatrixForUpdate(h, nodeIndex); }
    @Override public void setNodeMatrix(int nodeIndex, int matrixIndex, double[] matrix) {
        SieveB200.setNodeMatrix(h, nodeIndex, matrixIndex, matrix);
    }
    @Override public void setNodePartialsForUpdate(int nodeIndex) {
        if (nodeIndex >= nrOfLeafNodes) SieveB200.setNodePartialsForUpdate(h, nodeIndex); // leaves flip in updateLeaves
    }
    @Override public void calculateLogPartials(int c1, boolean leaf1, int c2, boolean leaf2, int parent, int constGenotype) {
        SieveB200.calculatePartials(h, c1, leaf1, c2, leaf2, parent, constGenotype);
        evaluated = false;
    }
    @Override public void calculatePartials(int c1, boolean leaf1, int c2, boolean leaf2, int parent, int constGenotype) {
        calculateLogPartials(c1, leaf1, c2, leaf2, parent, constGenotype);
    }

    /** integratePartials (:153-162): remembers the root; the integration runs as the epilogue of the root's op */
    @Override
    public void integratePartials(int nodeIndex, double[] proportions, int rootGenotype, double[] outPartials, double[] constRoot) {
        SieveB200.setRoot(h, nodeIndex, proportions, rootGenotype);
    }

    /** calculateLogLikelihoods (:255-269): one launch for everything queued, then the per-site arrays */
    @Override
    public void calculateLogLikelihoods(double[] partials, double[] outLogLikelihoods, double[] constRoot, double[] logConstRoot) {
        SieveB200.evaluate(h, result);
        evaluated = true;
        SieveB200.getSiteLogLikelihoods(h, outLogLikelihoods, logConstRoot); // optional: calcLogP can use getShardResult()
    }

    /** {sum w*logL, log sum w*exp(logConst), sum w*logConst, lewis, sum w}: what ScsTreeLikelihoodCaller returns (:466-471) */
    public double[] getShardResult() { return result; }

    @Override public void getNodePartials(int nodeIndex, double[] out) { SieveB200.getNodePartials(h, nodeIndex, 0, out); }
    @Override public void store() { SieveB200.store(h); }
    @Override public void unstore() { SieveB200.unstore(h); }
    @Override public void restore() { SieveB200.restore(h); }
    @Override public void setUseScaling(double scale) { /* the device arithmetic is always scaled */ }
    @Override public void finalize() throws Throwable { SieveB200.destroy(h); h = 0; super.finalize(); }
}
