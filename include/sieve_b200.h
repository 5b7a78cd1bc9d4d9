/*
 * sieve_b200.h -- C-ABI of the B200-native tree-likelihood core for SIEVE.
 *
 * This is the drop-in boundary: the entry points are what a JNI shim behind SIEVE's
 * `LikelihoodCore` / `RawReadCountsModel` seam binds (see INTEGRATION.md for the Java side).  Every function
 * cites the reference interface it replaces; paths are relative to /root/reference/src/beast/.
 *
 * Conventions
 *   - every function returns an int status (SIEVE_OK == 0); no exception crosses the boundary;
 *     sieve_last_error() gives the text of the last failure on the calling thread.
 *   - the caller owns all host arrays; the callee has consumed them when the call returns.
 *   - one handle <=> one GPU <=> one contiguous site shard (the analogue of one worker of
 *     likelihood/ThreadedScsTreeLikelihood.java:176-219).  Calls on one handle are serialised by the caller;
 *     different handles may be driven from different host threads.
 *   - node numbering is the reference's: leaves 0..n-1, internal nodes n..2n-2, unary root 2n-1
 *     (tree/ScsTree.java:220-260).  Matrices are row = parent state, column = child state
 *     (likelihood/ScsBeerLikelihoodCore.java:107-108).  Genotype states: 0/0, 0/1, 1/1, 1/1'
 *     (substitutionmodel/ScsFiniteMuExtendedModel.java:62-66).
 *   - NaN / -inf results are returned faithfully in the outputs (e.g. eff_seq_err == 0 with alt reads);
 *     CUDA failures are status codes.
 *   - there is no CPU fallback: every compute entry point fails with SIEVE_ERR_CUDA when no device is usable.
 */
#ifndef SIEVE_B200_H
#define SIEVE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sieve_core sieve_core_t;

enum {
    SIEVE_OK = 0,
    SIEVE_ERR_INVALID = 1, /* bad argument */
    SIEVE_ERR_CUDA = 2,    /* CUDA runtime failure (incl. no device) */
    SIEVE_ERR_OOM = 3,     /* the requested residency does not fit the device */
    SIEVE_ERR_STATE = 4    /* call sequence error (e.g. compute before counts were set) */
};

/* sieve_config_t.flags */
#define SIEVE_FLAG_LOG_PARTIALS 1u   /* useLogPartials == true semantics (likelihood/ScsTreeLikelihood.java:297-300):      \
                                        transition probabilities are clamped to Double.MIN_VALUE as                        \
                                        ScsBeerLikelihoodCore.java:1430 does, and sieve_get_node_partials returns logs */
#define SIEVE_FLAG_CONST_PARTIALS 2u /* maintain the constant-site partials (acquisition-bias correction on) */
#define SIEVE_FLAG_SINGLE_ADO 4u     /* singleADO == true (rawreadcountsmodel/RawReadCountsModelInterface.java:178-181) */
#define SIEVE_FLAG_EPHEMERAL 8u      /* do not keep internal partials resident: every update is a streamed full-tree    \
                                        evaluation (needed when 2 * n * S * 136 B exceeds HBM) */
#define SIEVE_FLAG_SPARSE 16u        /* keep the leaves resident but only the partials of the larger subtrees (as many as  \
                                        fit); small subtrees are recomputed on the way.  For shapes whose full residency       \
                                        exceeds HBM but whose MCMC moves should still cost O(depth), e.g. 10^4 cells. */
#define SIEVE_FLAG_SINGLE_LEAF_BUFFER 32u /* one leaf buffer instead of two (32 B per cell and site saved): a rejected         \
                                        leaf-parameter proposal is undone by recomputing the leaves in sieve_restore */

#define SIEVE_FLAG_EXTERNAL_COUNTS 64u /* the count tensor stays in caller-owned device memory: sieve_create allocates none, \
                                        sieve_set_counts_device adopts the pointer instead of copying (it must stay valid   \
                                        until sieve_destroy).  Saves 16 B per cell and site when the producer of the counts \
                                        (an ingest kernel, a generator) already wrote the core's layout */

#define SIEVE_FLAG_PRIVATE_SEQ_COV 128u /* the coverage model is PrivateAllelicSeqCovModel (rawreadcountsmodel/seqcovmodel/        \
                                        PrivateAllelicSeqCovModel.java): one (allelicSeqCov, allelicSeqCovRawVar) pair per (rate  \
                                        category, pattern), set with sieve_set_private_seq_cov; the leaf partials then differ      \
                                        between the categories (nrOfMatrices = K, 128 B per cell and site).  Dense resident cores. */

typedef struct {
    int32_t n_leaves;     /* cells (taxa) */
    int32_t n_sites;      /* patterns of this shard */
    int32_t n_categories; /* K, site-rate categories == matrixCount (1..4) */
    int32_t n_states;     /* must be 4 */
    uint32_t flags;
    int32_t device; /* CUDA ordinal; -1 = current device */
} sieve_config_t;

/* The six RealParameter inputs of the raw-read-counts model:
 * rawreadcountsmodel/RawReadCountsModelInterface.java:44-48 (adoRate),
 * seqcovmodel/SeqCovModelInterface.java:30-67 (allelicSeqCov, allelicSeqCovRawVar),
 * nucreadcountsmodel/NucReadCountsModelInterface.java:18-34 (effSeqErrRate, shapeCtrl1, shapeCtrl2). */
typedef struct {
    double ado_rate;
    double allelic_seq_cov;
    double allelic_seq_cov_raw_var;
    double eff_seq_err_rate;
    double shape_ctrl1;
    double shape_ctrl2;
} sieve_leaf_params_t;

/* acquisition-bias modes, alignment/ScsAlignment.java:93-113 (ascertained, meanAscBiasCorrection) */
enum { SIEVE_ASC_NONE = 0, SIEVE_ASC_FELSENSTEIN_MEAN = 1, SIEVE_ASC_FELSENSTEIN_GEOMETRIC = 2, SIEVE_ASC_LEWIS = 3 };

/* What one shard reports per evaluation (the two doubles a worker returns,
 * likelihood/ThreadedScsTreeLikelihood.java:466-471, plus the sums the other bias modes need). */
typedef struct {
    double sum_log_l;     /* sum_s w_s * logL_s                      (likelihood/ScsTreeLikelihood.java:2625-2629) */
    double const_lse;     /* log sum_s w_s * exp(logConst_s)         (alignment/ScsAlignment.java:768-769) */
    double const_sum;     /* sum_s w_s * logConst_s                  (alignment/ScsAlignment.java:773-774) */
    double lewis_sum;     /* -sum_s w_s * log(1 - exp(logConst_s))   (alignment/ScsAlignment.java:746-752) */
    double weight_sum;    /* sum_s w_s (loci of this shard) */
} sieve_shard_result_t;

const char *sieve_last_error(void);
int sieve_version(void);
/* number of usable CUDA devices (0 when none); never fails */
int sieve_device_count(void);

/* ScsBeerLikelihoodCore.initialize (likelihood/ScsBeerLikelihoodCore.java:91-128) + createNodePartials (:134-143):
 * allocates the device pools.  Fails with SIEVE_ERR_OOM when the residency does not fit. */
int sieve_create(const sieve_config_t *cfg, sieve_core_t **out);
/* ScsBeerLikelihoodCore.finalize (:399-409) */
int sieve_destroy(sieve_core_t *h);
/* bytes of device memory held by the handle */
int sieve_device_bytes(sieve_core_t *h, uint64_t *out);

/* The read-count tensor: ScsAlignment.sitePatterns / getPattern (alignment/ScsAlignment.java:188, 400-446).
 * counts is int32 [site][cell][tuple_len], tuple_len 4 = (alt1, alt2, alt3, cov) or 5 = the reference's
 * (alt1, alt2, alt3, ref, cov) of datatype/FullSupsCov.java:74-77. */
int sieve_set_counts(sieve_core_t *h, const int32_t *counts, int32_t tuple_len);
/* the tensor of the ingest format (sieve_b200/ingest.py): HOST memory already in the core's own
 * order [cell][site][4] (alt1, alt2, alt3, cov), int32 (dtype_code 0) or uint16 (dtype_code 1); streamed in chunks, so
 * `counts` may be a memory-mapped file */
int sieve_set_counts_cellmajor(sieve_core_t *h, const void *counts, int32_t dtype_code);
/* as above but the tensor already lives on the device in the core's own layout int32 [cell][site][4] */
int sieve_set_counts_device(sieve_core_t *h, const void *d_counts_cell_site_4);
/* ScsAlignment.getPatternWeight (likelihood/ScsTreeLikelihood.java:2626); NULL = all 1 */
int sieve_set_pattern_weights(sieve_core_t *h, const int32_t *weights);

/* SeqCovModelInterface.Base.computeSizeFactors (seqcovmodel/SeqCovModelInterface.java:291-351) over THIS handle's sites
 * (use it on an unsharded handle, or compute once and hand the result to every shard, as
 * likelihood/ThreadedScsTreeLikelihood.java:119-131 + ExploredSharedAllelicSeqCovModel.java:103-109 do).
 * out_size_factors[n_leaves]; out_stats[2] = {mean(cov)/2, var(cov)/4} (the start values of :322-325), may be NULL. */
int sieve_compute_size_factors(sieve_core_t *h, int32_t zero_cov_mode, double *out_size_factors, double *out_stats);
int sieve_set_size_factors(sieve_core_t *h, const double *size_factors);

/* ---- size factors over a SITE-SHARDED alignment --------------------------------------------------------------------
 * likelihood/ThreadedScsTreeLikelihood.java:119-131 initialises the coverage model on the FULL alignment before it splits
 * the sites, so that every worker uses the same, global size factors (ExploredSharedAllelicSeqCovModel.java:103-109).
 * With one GPU per site range no device holds the full alignment; the per-cell median of seqcovmodel/
 * SeqCovModelInterface.java:291-351 is then found by exact distributed selection (csrc/sizefactors_global.cuh): per
 * round every part counts, per cell, the weight of its coverage ratios below a pivot, the counts are summed over the
 * parts (integers: the split cannot change the result), 63 rounds of bisection on the ratio's bit pattern.  The result is
 * bit-identical to sieve_compute_size_factors on one handle that holds all sites.
 *
 * A part borrows a device-resident count tensor int32 [cell][site][4] (the core's own layout); it needs no core and no
 * pools.  weights: host int32 [n_sites] or NULL (all 1).  */
typedef struct sieve_sf sieve_sf_t;
int sieve_sf_open(int32_t device, int32_t n_cells, int32_t n_sites, const void *d_counts_cell_site_4,
                  const int32_t *weights, int32_t zero_cov_mode, sieve_sf_t **out);
/* a part over the counts and pattern weights a handle already holds */
int sieve_sf_open_core(sieve_core_t *h, int32_t zero_cov_mode, sieve_sf_t **out);
int sieve_sf_close(sieve_sf_t *p);
/* the two local primitives of the protocol: {sum cov, sum cov^2, number of entries} of this part; and, per cell and
 * pivot (n_piv = 1 or 2 pivots per cell, bit patterns of doubles), the weight of the kept ratios <= the pivot */
int sieve_sf_moments(sieve_sf_t *p, uint64_t *out3);
int sieve_sf_weight_le(sieve_sf_t *p, int32_t n_piv, const uint64_t *pivots, int64_t *out);
/* sum over all processes, in place; 0 = ok */
typedef int (*sieve_sf_allreduce_fn)(void *user, int64_t *buf, int32_t count);
/* The protocol itself is host logic over callbacks (no device needed: CPU tests drive it with stand-in primitives).
 * moments / weight_le answer for everything THIS process holds; allreduce_sum_i64 sums over processes (NULL: one). */
typedef struct {
    void *user;
    int32_t n_cells;
    int (*moments)(void *user, uint64_t *out3);
    int (*weight_le)(void *user, int32_t n_piv, const uint64_t *pivots, int64_t *out);
    sieve_sf_allreduce_fn allreduce_sum_i64;
} sieve_sf_ops_t;
/* out_size_factors [n_cells]; out_stats[2] = {mean(cov) / 2, var(cov) / 4} (may be NULL) */
int sieve_sf_solve(const sieve_sf_ops_t *ops, double *out_size_factors, double *out_stats);
/* the usual wiring: the parts this process drives (one per GPU; the JVM drives all of them) + an optional all-reduce
 * across processes (one process per GPU: torch.distributed / MPI / NCCL behind the callback; NULL for one process) */
int sieve_sf_solve_parts(int32_t n_parts, sieve_sf_t *const *parts, sieve_sf_allreduce_fn allreduce, void *user,
                         double *out_size_factors, double *out_stats);

/* ---- PrivateAllelicSeqCovModel (SIEVE_FLAG_PRIVATE_SEQ_COV) ------------------------------------------------------------
 * allelicSeqCovPerPattern / allelicSeqCovRawVarPerPattern (PrivateAllelicSeqCovModel.java:16-17), both [K][n_sites] in the
 * reference's index order matrixIndex * nrOfPatterns + patternIndex; they replace allelic_seq_cov / allelic_seq_cov_raw_var
 * of sieve_leaf_params_t in the coverage term (computeSeqCovLikelihoodPerAllele, :499-511).  Takes effect at the next
 * sieve_update_leaves.  The model initialises every pair with the shared start values (:159-161). */
int sieve_set_private_seq_cov(sieve_core_t *h, const double *allelic_seq_cov_ks, const double *allelic_seq_cov_raw_var_ks);
int sieve_get_private_seq_cov(sieve_core_t *h, double *allelic_seq_cov_ks, double *allelic_seq_cov_raw_var_ks);
/* The accept-time re-estimation, PrivateAllelicSeqCovModel.updateAllelicSeqCovAndRawVar (:312-432) driven by
 * ScsTreeLikelihood.postProcess (likelihood/ScsTreeLikelihood.java:2454-2520): for every (category, pattern) pair the
 * maximum-likelihood numbers of sequenced alleles of the cells (from the max-sum trace of the current state, which this
 * call runs) give new (t, v) by the grouped method of moments.  Pairs whose ML combination is not unique are left to the
 * host (out_tied, [K][n_sites], may be NULL: 1 = tied, resolve with sieve_ml_get_site as for variant calling and write
 * the pair back with sieve_set_private_seq_cov).  out_changed: number of pairs whose (t, v) moved.  The caller then
 * recomputes the state (sieve_update_leaves + a whole-tree update), as postProcess does for the changed patterns. */
int sieve_private_update_from_ml(sieve_core_t *h, int32_t zero_cov_mode, int32_t *out_changed, uint8_t *out_tied);
/* ScsTreeLikelihood.postProcess step 3 (likelihood/ScsTreeLikelihood.java:2493-2498: "traverse the entire tree for changed
 * patterns"; the changedPatterns overloads of likelihood/ScsBeerLikelihoodCore.java:584-820, 1070-1266, 1516-1757): the sites
 * at which any category's (t, v) differs from the pair the current leaves were computed with are re-evaluated IN PLACE --
 * their leaves, every internal node (ops: the whole tree in post order, triples as for sieve_update_nodes), their site
 * log-likelihoods and column sums; no buffer is flipped and every other site keeps its value, so the state equals what a
 * whole-shard pass (sieve_update_leaves + sieve_update_nodes) produces, bit for bit.  The accepted state becomes the stored
 * state (as by sieve_store).  out_sites: number of re-evaluated sites (0: nothing moved, nothing done).  Follow with
 * sieve_evaluate for the totals.  Dense resident cores only (SIEVE_ERR_STATE otherwise: use the whole-shard pass). */
int sieve_private_retraverse_changed(sieve_core_t *h, int32_t n_ops, const int32_t *ops, int32_t *out_sites);

/* new values of the six leaf-model parameters; takes effect at the next sieve_update_leaves */
int sieve_set_leaf_params(sieve_core_t *h, const sieve_leaf_params_t *p);
/* rawReadCountsModel.computeLeafLikelihood + core.setNodePartials for every leaf
 * (likelihood/ScsTreeLikelihood.java:583-601, rawreadcountsmodel/RawReadCountsModelInterface.java:458-506).
 * for_update != 0 first flips the leaves' write buffer (setNodePartialsForUpdate, :586);
 * for_update == 0 is the initialisation path (initializeLeafLikelihood + storeLeafPartials, :430-468). */
int sieve_update_leaves(sieve_core_t *h, int32_t for_update);

/* BeerLikelihoodCore.setNodeMatrixForUpdate / setNodeMatrix (external; call sites likelihood/ScsTreeLikelihood.java:566-578) */
int sieve_set_node_matrix_for_update(sieve_core_t *h, int32_t node);
int sieve_set_node_matrix(sieve_core_t *h, int32_t node, int32_t category, const double *p16);
/* batched form: for every listed node (flip if for_update) all K matrices, P is [count][K][16] */
int sieve_set_matrices(sieve_core_t *h, int32_t count, const int32_t *nodes, const double *P, int32_t for_update);

/* Fast path for SIEVE's fixed rate matrix (substitutionmodel/ScsFiniteMuExtendedModel.java:62-86): instead of K
 * matrices per node the host passes K branch distances, distance[k] = branch length * branch rate * site rate_k, i.e.
 * the arguments of getTransitionProbabilities at likelihood/ScsTreeLikelihood.java:567-571.  The device forms
 * P = exp(Q d) in closed form (Q has eigenvalues 0, -0.2, -0.4, -0.4) and, when every matrix of an update came
 * in this way, walks the tree with two scalars per (node, category) instead of a matrix in shared memory.
 * distances is [count][K].  Mixing with sieve_set_matrices is allowed (the generic kernel is used then). */
int sieve_set_distances(sieve_core_t *h, int32_t count, const int32_t *nodes, const double *distances, int32_t for_update);

/* BeerLikelihoodCore.setNodePartialsForUpdate (external; call site ScsTreeLikelihood.java:2951-2956) */
int sieve_set_node_partials_for_update(sieve_core_t *h, int32_t node);
/* ScsBeerLikelihoodCore.calculateLogPartials / calculatePartials (:1283-1369 / :857-943).  child2 == -1 <=> unary root.
 * The op is queued; queued ops are launched together (one kernel) by the next call that needs their result. */
int sieve_calculate_partials(sieve_core_t *h, int32_t child1, int32_t is_leaf1, int32_t child2, int32_t is_leaf2,
                             int32_t parent, int32_t const_genotype);
/* batched form of the two calls above: ops is [n_ops][3] = (parent, child1, child2 | -1) in post-order
 * (children before parents).  level_offsets (n_levels + 1 entries, may be NULL) marks groups of mutually
 * independent ops; with it and SIEVE_UPDATE_PER_LEVEL the update is issued as one launch per tree level. */
#define SIEVE_UPDATE_FLIP 1u      /* setNodePartialsForUpdate on every parent first */
#define SIEVE_UPDATE_PER_LEVEL 2u /* one launch per level instead of one launch per update */
int sieve_update_nodes(sieve_core_t *h, int32_t n_ops, const int32_t *ops, const int32_t *level_offsets,
                       int32_t n_levels, uint32_t mode);

/* which node is the root and how categories are integrated there:
 * integratePartials(root, proportions, rootGenotype, ...) (:153-218) + calculateLogLikelihoods (:255-269).
 * The integration itself runs as the epilogue of the root's op. */
int sieve_set_root(sieve_core_t *h, int32_t root, const double *proportions, int32_t root_genotype);
/* launches everything queued, then the per-site log + reduction; fills *out (ScsTreeLikelihood.calcLogP, :2621-2651) */
int sieve_evaluate(sieve_core_t *h, sieve_shard_result_t *out);

/* One MCMC-step call: matrices of the changed nodes, the dirty ops, and the result, with one host->device copy,
 * one device->host copy.  Equivalent to sieve_set_matrices(for_update) + sieve_update_nodes(FLIP) + sieve_evaluate.
 * update_leaves != 0 runs sieve_update_leaves(for_update) first (parameters set by sieve_set_leaf_params). */
int sieve_step(sieve_core_t *h, int32_t update_leaves, int32_t n_matrices, const int32_t *matrix_nodes, const double *P,
               int32_t n_ops, const int32_t *ops, sieve_shard_result_t *out);

/* sieve_step with branch distances [n_nodes][K] in place of matrices (see sieve_set_distances) */
int sieve_step_distances(sieve_core_t *h, int32_t update_leaves, int32_t n_nodes, const int32_t *nodes,
                         const double *distances, int32_t n_ops, const int32_t *ops, sieve_shard_result_t *out);
/* bytes the last step copied host -> device (ops + matrices / distances) */
int sieve_last_stage_bytes(sieve_core_t *h, uint64_t *out);

/* Re-launches, n_times, exactly the kernels of the last sieve_step / sieve_evaluate from what is already staged in HBM
 * (matrices, ops): no host<->device copy and no synchronisation.  update_leaves != 0 also rebuilds the tables and the
 * leaf partials each time.  Used to time the device path in isolation; results are read with sieve_evaluate-style
 * getters after the caller synchronises the handle's stream. */
int sieve_replay(sieve_core_t *h, int32_t update_leaves, int32_t n_times);

/* getPatternLogLikelihoods (likelihood/ScsTreeLikelihood.java:2958-2970); out_const may be NULL */
int sieve_get_site_log_likelihoods(sieve_core_t *h, double *out_site_log_l, double *out_site_log_const);
/* device pointers to the same two arrays (valid until destroy), for callers that combine shards on the device */
int sieve_get_result_device_ptr(sieve_core_t *h, void **d_result /* sieve_shard_result_t */);

/* BeerLikelihoodCore.getNodePartials in the reference layout [K][S][G] (:1417-1499 index order), logs when
 * SIEVE_FLAG_LOG_PARTIALS.  which: 0 = partials, 1 = constant-site partials (internal nodes only). */
int sieve_get_node_partials(sieve_core_t *h, int32_t node, int32_t which, double *out);

/* Change tracking (resident cores).  The core versions every input (matrix slot contents, leaf buffers, node values); a
 * node the caller marks dirty whose inputs are the very ones a value it still holds was computed from is not recomputed
 * -- the kernels are deterministic, the bits would be the same.  This matters for callers that dirty more than changed:
 * SIEVE recomputes the whole tree for every branch-rate proposal under a relaxed clock with Felsenstein's correction
 * (likelihood/ScsTreeLikelihood.java:2814-2821, alignment/ScsAlignment.java:707-709); here the walk then covers only the
 * path from the changed branch to the root.  On by default; `on = 0` makes every dirty node recompute (A/B, tests). */
int sieve_set_elision(sieve_core_t *h, int32_t on);
/* SIEVE_FLAG_SPARSE bookkeeping: slots in the partial pool, the current keep threshold (nodes with more leaves keep their
 * partials resident; 0 for a dense core) and the slots that are free right now.  Outputs may be NULL. */
int sieve_sparse_info(sieve_core_t *h, int32_t *pool_slots, int32_t *keep_leaves, int32_t *free_slots);
int sieve_elided_count(sieve_core_t *h, uint64_t *out); /* dirty nodes not recomputed so far */
/* Diagnostic: resident blocks per SM (6, 7 or 8) that the single-launch walk over n_sites sites is compiled for -- chosen
 * per launch from how its blocks of 64 sites divide into waves on the 148 SMs.  No device needed. */
int sieve_walk_blocks_per_sm(int64_t n_sites);

/* ---- max-sum pass: maximum-likelihood genotypes and dropout states (variant calling) --------------------------------
 * Replaces the traceMLGenotypes arm of the reference for ONE fixed state (tree, matrices, leaf parameters), i.e. what
 * VariantCaller needs after the MCMC: the max-sum twin of the pruning with its arg-max lists
 * (likelihood/ScsBeerLikelihoodCore.java:2918-3486), the leaves' arg-max dropout component
 * (rawreadcountsmodel/RawReadCountsModelFiniteMuDM.java:54-77), the ML rate category per pattern
 * (ScsBeerLikelihoodCore.java:278-296) and the top-down collection of the genotypes every node takes
 * (likelihood/ScsTreeLikelihood.java:947-1460, MLGenotypesCollection).  A tie list of the reference is a bit mask here
 * (bit g set <=> genotype / component g is listed).  The pass is recomputed by every call of sieve_ml_trace; any later
 * change of the state invalidates the results (the getters then fail with SIEVE_ERR_STATE).
 * A core with linear partials (no SIEVE_FLAG_LOG_PARTIALS) runs the linear-space twin (ScsBeerLikelihoodCore.java:1759-2335):
 * a LEAF child then contributes log(P[p][c] * leaf[c]) with the unclamped P, an internal child as in the log-space twin.
 * Dense cores trace the whole shard and serve every getter.  SIEVE_FLAG_SPARSE cores (no room for ~190 B per cell and site of
 * max-sum state) trace in site chunks inside sieve_ml_call / sieve_ml_get_categories and keep the calls only:
 * sieve_ml_get_node / sieve_ml_get_site fail with SIEVE_ERR_STATE there -- resolve the (rare) tied sites on a small dense
 * core over just those sites.  SIEVE_FLAG_EPHEMERAL cores hold no leaves or root to trace from: SIEVE_ERR_STATE. */
int sieve_ml_trace(sieve_core_t *h);
/* The common case of ScsTreeLikelihood.collectMLGenotypesAndLogLikelihoods (:2033-2081), for the lowest ML category of
 * every site: genotypes [node][S] and ado [leaf][S] hold the unique ML genotype / number of dropped alleles, or -1 where
 * several are tied; category [S]; ambiguous [S]: bit 0 = several ML categories, bit 1 = several ML combinations. */
int sieve_ml_call(sieve_core_t *h, int8_t *genotypes, int8_t *ado, uint8_t *category, uint8_t *ambiguous);
/* getPatternCategories: bit mask over the categories per site */
int sieve_ml_get_categories(sieve_core_t *h, uint8_t *out_S);
/* Reference layouts, any of the outputs may be NULL: ml_partials [K][S][4] (MLPartials; a leaf: its log partials),
 * back [K][S][4] one byte per (category, site, parent genotype): low nibble = arg-max list of child 0, high nibble = of
 * child 1 (MLGenotypesNodes[..][node][..][0|1]); for a leaf the byte is the arg-max list of its dropout components;
 * collection [K][S] = MLGenotypesCollection of the node. */
int sieve_ml_get_node(sieve_core_t *h, int32_t node, double *ml_partials, uint8_t *back, uint8_t *collection);
/* One site, all nodes: back [node][K][4] (bytes as above) and the nodes' sum-product log partials [node][K][4]
 * (ScsBeerLikelihoodCore.getLogLikelihood :338-349): the inputs of the tie-break :2083-2160, done by the host. */
int sieve_ml_get_site(sieve_core_t *h, int32_t site, uint8_t *back, double *log_partials);

/* ScsBeerLikelihoodCore.store / unstore (:4053-4065), BeerLikelihoodCore.restore (external) -- index flips only */
int sieve_store(sieve_core_t *h);
int sieve_unstore(sieve_core_t *h);
int sieve_restore(sieve_core_t *h);

/* ThreadedScsTreeLikelihood.calculateLogPByBeagle (likelihood/ThreadedScsTreeLikelihood.java:366-389) with
 * ScsAlignment.getAscBiasCorrection (alignment/ScsAlignment.java:764-782 for one shard, :811-820 for several):
 * logP = sum_r a_r + bias(asc_mode, b_r, M).  Pure host arithmetic on n_shards small structs. */
double sieve_combine_shards(const sieve_shard_result_t *shards, int32_t n_shards, int32_t asc_mode,
                            double n_background_sites);
/* One process per GPU: the join of the reference's workers (likelihood/ThreadedScsTreeLikelihood.java:366-379) becomes an
 * NCCL all-gather of the 5-double results on the handle's stream, issued by every sieve_evaluate / sieve_step once a
 * communicator is bound.  Rank 0 draws an id (128 bytes), distributes it by any means, every rank calls sieve_comm_init.
 * sieve_last_global_results returns the results of all ranks, rank order, of the last evaluation (out_world[comm size]). */
int sieve_comm_unique_id(uint8_t *id_out128);
int sieve_comm_init(sieve_core_t *h, int32_t rank, int32_t world, const uint8_t *id128);
int sieve_comm_size(sieve_core_t *h);
/* how the per-evaluation exchange runs on this handle: 0 = not bound to a communicator, 1 = ncclAllGather behind the
 * reduction, 2 = fused into the reduction kernel over peer memory (every rank maps its peers' mailboxes through CUDA IPC at
 * sieve_comm_init; chosen when all ranks could, SIEVE_B200_P2P=0 forces 1) */
int sieve_comm_exchange_mode(sieve_core_t *h);
int sieve_last_global_results(sieve_core_t *h, sieve_shard_result_t *out_world);

/* calcPatternPoints (likelihood/ThreadedScsTreeLikelihood.java:227-238): points[n_shards + 1] */
void sieve_pattern_points(int32_t n_sites, int32_t n_shards, int32_t *points);

/* Stage-1 background likelihood, ScsBackgroundLikelihood.calculateLogPDirichletMultinomial
 * (likelihood/ScsBackgroundLikelihood.java:169-268): the five histograms (variant1, variant2, variant3, normal, coverage;
 * alignment/ScsBackgroundInfo.java:184-186) are given flattened: entry i has read count values[i] with multiplicity
 * counts[i]; histogram h owns entries offsets6[h] .. offsets6[h+1].  Needs no handle. */
int sieve_background_log_likelihood(int32_t device, const int64_t *values, const int64_t *counts, const int32_t *offsets6,
                                    int64_t n_background_sites, int32_t n_taxa, double eff_seq_err_rate,
                                    double shape_ctrl1, double *out_log_p);

/* kernel-time breakdown of the last sieve_evaluate / sieve_step in milliseconds (CUDA events on the handle's stream):
 * out[0] leaf tables+kernel, out[1] pruning launches, out[2] reduce, out[3] total; and how many kernels were launched. */
int sieve_last_timing(sieve_core_t *h, float *out_ms4, int32_t *out_launches);
int sieve_enable_timing(sieve_core_t *h, int32_t on);
/* total number of kernels this handle has launched so far */
int sieve_launch_count(sieve_core_t *h, uint64_t *out);
/* the CUDA stream the handle launches on (cudaStream_t) */
int sieve_get_stream(sieve_core_t *h, void **out_stream);

#ifdef __cplusplus
}
#endif
#endif /* SIEVE_B200_H */
