/*
 * sieve_b200_mcmc.h -- a native MCMC replay driver above the C-ABI of sieve_b200.h.
 *
 * In production BEAST 2 (Java) owns the chain: state, operators, priors, acceptance.  There is no JVM in this project's
 * image, so to measure "MCMC states per second" the way the plugin would load the device, this driver replays the
 * reference's operator schedule natively: it holds a trunk tree (tree/ScsTree.java:220-260) and the model parameters,
 * draws operators with the weights of the templates (examples/templates/smc_stage_1.xml:139-160, smc_stage_2.xml:85-95),
 * applies the move, derives the dirty set exactly as ScsTreeLikelihood.requiresRecalculation / traverse do
 * (likelihood/ScsTreeLikelihood.java:2786-2824, 545-935), calls sieve_store / sieve_step_distances / sieve_restore,
 * and accepts with the Metropolis-Hastings ratio on the tree likelihood (flat priors; the Hastings terms of the scale
 * operators are included).  It uses ONLY the public entry points of sieve_b200.h.
 */
#ifndef SIEVE_B200_MCMC_H
#define SIEVE_B200_MCMC_H

#include "sieve_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sieve_mcmc sieve_mcmc_t;

enum {
    SIEVE_MIX_STAGE1 = 1, /* smc_stage_1.xml: tree moves 81, coalescent parameters 5, leaf-model parameters 11, gammaShape 2 */
    SIEVE_MIX_STAGE2 = 2, /* smc_stage_2.xml: tree moves only (leaf-model parameters fixed) */
    SIEVE_MIX_RELAXED2 = 3 /* rmc_stage_2.xml: tree moves + branch-rate moves, which dirty the whole tree when the
                              Felsenstein correction is on (isForcingTreeDirtyRMC, ScsTreeLikelihood.java:2814-2821) */
};

/* operator classes reported in the statistics */
enum {
    SIEVE_OP_TREE_SCALE = 0,   /* ScsTreeScaleOperator: every height, full tree */
    SIEVE_OP_ROOT_SCALE = 1,   /* ScsTreeScaleOperator rootOnly */
    SIEVE_OP_UNIFORM = 2,      /* ScsUniform */
    SIEVE_OP_SUBTREE_SLIDE = 3,
    SIEVE_OP_NARROW = 4,       /* ScsExchange isNarrow */
    SIEVE_OP_WIDE = 5,         /* ScsExchange wide */
    SIEVE_OP_WILSON_BALDING = 6,
    SIEVE_OP_PRIOR_ONLY = 7,   /* ePopSize / growthRate: no likelihood work */
    SIEVE_OP_COVERAGE = 8,     /* allelicSeqCov, allelicSeqCovRawVar: NB tables + leaves + full tree */
    SIEVE_OP_NUC = 9,          /* effSeqErrRate, shapeCtrl1, shapeCtrl2: DM tables + leaves + full tree */
    SIEVE_OP_ADO = 10,         /* adoRate: remix + leaves + full tree */
    SIEVE_OP_GAMMA_SHAPE = 11, /* all matrices, full tree */
    SIEVE_OP_BRANCH_RATE = 12, /* relaxed clock rate move */
    SIEVE_OP_CLASSES = 13
};

typedef struct {
    int32_t n_leaves;
    int32_t n_categories;
    int32_t asc_mode;            /* SIEVE_ASC_* */
    double n_background_sites;
    double gamma_shape;
    double clock_rate;
    sieve_leaf_params_t leaf;
    int32_t mix;                 /* SIEVE_MIX_* */
    uint64_t seed;
} sieve_mcmc_config_t;

typedef struct {
    int64_t states;
    int64_t accepted;
    int64_t likelihood_evaluations; /* states that launched device work */
    int64_t ops_total;              /* internal nodes recomputed, summed over states */
    int64_t matrices_total;         /* branch matrices refreshed */
    int64_t leaf_updates;           /* states that recomputed the leaves */
    double seconds;                 /* wall time of the run (host + device, synchronous per state) */
    double log_p;                   /* current log-likelihood */
    int64_t proposed[SIEVE_OP_CLASSES];
    int64_t accepted_by_op[SIEVE_OP_CLASSES];
    double seconds_by_op[SIEVE_OP_CLASSES];
} sieve_mcmc_stats_t;

/* `core` must already hold counts, size factors and pattern weights; the driver sets the leaf parameters, the root and
 * does the initial full evaluation.  left/right/height describe the start tree (node numbering of sieve_b200.h;
 * right[root] = -1).  branch_rates may be NULL (strict clock). */
int sieve_mcmc_create(const sieve_mcmc_config_t *cfg, sieve_core_t *core, const int32_t *left, const int32_t *right,
                      const double *height, const double *branch_rates, sieve_mcmc_t **out);
int sieve_mcmc_destroy(sieve_mcmc_t *m);
/* Multi-shard chains: `combine` turns this shard's result into the log-likelihood of the WHOLE data set (the all-gather
 * + fixed-order combine of sieve_combine_shards over ranks).  Every rank runs the same chain with the same seed and takes
 * identical decisions because the combined value is bit-identical everywhere.  NULL restores the single-shard combine. */
typedef double (*sieve_combine_fn)(const sieve_shard_result_t *local, void *user);
int sieve_mcmc_set_combine(sieve_mcmc_t *m, sieve_combine_fn combine, void *user);
/* runs n_states proposals; statistics accumulate over calls unless reset != 0 */
int sieve_mcmc_run(sieve_mcmc_t *m, int64_t n_states, int32_t reset, sieve_mcmc_stats_t *out);
/* the current tree (arrays of 2 n entries) and log-likelihood, e.g. to re-evaluate it from scratch as a check */
int sieve_mcmc_get_state(sieve_mcmc_t *m, int32_t *left, int32_t *right, double *height, double *branch_rates,
                         sieve_leaf_params_t *leaf, double *gamma_shape, double *log_p);
/* recomputes the current state from scratch (all leaves, all matrices, all nodes) and returns that log-likelihood:
 * BEAST's robustlyCalcPosterior check (core/ScsState.java:15-27) */
int sieve_mcmc_full_recompute(sieve_mcmc_t *m, double *log_p);
/* discrete gamma category rates as ScsSiteModel yields them (BEAST1-style medians, mean 1); rates_out[k] */
int sieve_gamma_category_rates(double shape, int32_t k, double *rates_out);

/* The reference's own traverse sequence for a whole-tree update, issued call by call through the C-ABI the way the JNI
 * shim forwards it when only the core behind the LikelihoodCore seam is swapped (likelihood/ScsTreeLikelihood.java:563-601,
 * 621-628, 745-752, 854-861, 888-920, 2621-2651): per non-root node setNodeMatrixForUpdate + K x setNodeMatrix, (leaves),
 * per internal node setNodePartialsForUpdate + calculateLogPartials, evaluation, getPatternLogLikelihoods.
 * ops_post_order [n_ops][3] = (parent, child1, child2 | -1); P [n_nodes][K][16]; site_log_l [S] or NULL. */
int sieve_traverse_full(sieve_core_t *h, int32_t n_nodes, int32_t n_categories, int32_t root, int32_t n_ops,
                        const int32_t *ops_post_order, const double *P, int32_t update_leaves, double *site_log_l,
                        sieve_shard_result_t *out);

#ifdef __cplusplus
}
#endif
#endif
