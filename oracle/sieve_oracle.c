/*
 * sieve_oracle.c -- CPU restatement of SIEVE's tree-likelihood hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sieve_b200/ may include, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference (szczurek-lab/SIEVE v0.15.9, 100 % Java on top of BEAST 2.6.7)
 * cannot be compiled or run in the build container (no JDK, no BEAST jar) and its own tests hold
 * no numeric golden vector for this path.  The oracle is therefore pinned
 * by independent derivations instead (tests/test_oracle_*.py): brute-force enumeration of internal
 * genotypes, log-mode == linear-mode, mpmath 40-digit densities, thread-split invariance.
 *
 * Every function cites the reference file:line (relative to /root/reference/src/beast/) whose
 * loop order and arithmetic it follows.  Third-party arithmetic that is not in /root/reference:
 *   org.apache.commons.math.special.Gamma.logGamma / Beta.logBeta  (commons-math 2.x bundled in
 *   BEAST 2.6.7, build.xml:31-32): Lanczos g = 607/128, 15 coefficients, restated from the
 *   published algorithm; call sites math/distributions/DirichletMultinomial.java:49 and
 *   math/distributions/NegativeBinomial.java:78.
 *   beast.math.statistic.DiscreteStatistics.{max,mean,variance,median} (BEAST 2.6.7): trivial.
 *   beast.evolution.likelihood.BeerLikelihoodCore buffer-index semantics (BEAST 2.6.7).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <pthread.h>
#include <unistd.h>

#define SO_MIN_VALUE 4.9e-324 /* java.lang.Double.MIN_VALUE */

/* ------------------------------------------------------------------------------------------ */
/* commons-math 2.x Gamma.logGamma (external; published Lanczos form)                          */
/* ------------------------------------------------------------------------------------------ */
static const double SO_LANCZOS[15] = {
    0.99999999999999709182,
    57.156235665862923517,
    -59.597960355475491248,
    14.136097974741747174,
    -0.49191381609762019978,
    .33994649984811888699e-4,
    .46523628927048575665e-4,
    -.98374475304879564677e-4,
    .15808870322491248884e-3,
    -.21026444172410488319e-3,
    .21743961811521264320e-3,
    -.16431810653676389022e-3,
    .84418223983852743293e-4,
    -.26190838401581408670e-4,
    .36899182659531622704e-5,
};
#define SO_HALF_LOG_2_PI 0.91893853320467274178 /* 0.5 * log(2 pi) */

double so_log_gamma(double x)
{
    if (isnan(x) || x <= 0.0) return NAN;
    const double g = 607.0 / 128.0;
    double sum = 0.0;
    for (int i = 14; i > 0; --i) sum = sum + (SO_LANCZOS[i] / (x + i));
    sum = sum + SO_LANCZOS[0];
    double tmp = x + g + .5;
    return ((x + .5) * log(tmp)) - tmp + SO_HALF_LOG_2_PI + log(sum / x);
}

/* commons-math 2.x Beta.logBeta(a, b) */
double so_log_beta(double a, double b)
{
    if (isnan(a) || isnan(b) || a <= 0.0 || b <= 0.0) return NAN;
    return so_log_gamma(a) + so_log_gamma(b) - so_log_gamma(a + b);
}

/* math/util/MathFunctions.java:101-112 -- max + log(sum_i exp(x_i - max)), left to right */
double so_log_sum_exp(const double *a, int n)
{
    double result = 0.0;
    double max = a[0]; /* DiscreteStatistics.max */
    for (int i = 1; i < n; i++)
        if (a[i] > max) max = a[i];
    for (int i = 0; i < n; i++) result += exp(a[i] - max);
    return log(result) + max;
}

/* math/util/MathFunctions.java:73-92 */
double so_geometric_mean(const int *arr, int start, int end)
{
    double sum = 0.0;
    int len = 0;
    for (int i = start; i < end; i++) {
        if (arr[i] > 0) {
            ++len;
            sum += log((double)arr[i]);
        }
    }
    return len == 0 ? 0.0 : exp(sum / len);
}

/* math/distributions/DirichletMultinomial.java:48-50 */
static double so_dm_log_partial(double count, double fraction)
{
    return count <= 0 ? 0 : log(count) + so_log_beta(fraction, count);
}

/* math/distributions/DirichletMultinomial.java:25-39 (counts[4] is the total; returns 1.0 when it is 0) */
double so_dm_log_density(const int *counts5, const double *alpha5)
{
    if (counts5[4] == 0) return 1.0;
    double logValue = 0;
    int i = 0;
    for (; i < 4; i++) logValue += so_dm_log_partial(counts5[i], alpha5[i]);
    return so_dm_log_partial(counts5[i], alpha5[i]) - logValue;
}

/* math/distributions/NegativeBinomial.java:76-79 */
double so_nb_log_density(double x, double r, double p)
{
    double beta = 1.0 / p - 1.0;
    return so_log_gamma(r + x) + (r + x) * log(1.0 - p) - so_log_gamma(x + 1.0) - so_log_gamma(r) - r * log(beta);
}

/* ------------------------------------------------------------------------------------------ */
/* Size factors: rawreadcountsmodel/seqcovmodel/SeqCovModelInterface.java:291-351              */
/* counts: [pattern][taxon][5] (alt1, alt2, alt3, ref, cov); weights may be NULL (all 1).      */
/* stats_out[0] = mean(cov)/2, stats_out[1] = variance(cov)/4 (the start values of :322-325).  */
/* ------------------------------------------------------------------------------------------ */
static int so_cmp_double(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

void so_size_factors(const int *counts, const int *weights, int n_patterns, int n_taxa, int zero_cov_mode,
                     double *size_factors, double *stats_out)
{
    int *processed = (int *)malloc(sizeof(int) * (size_t)n_patterns * n_taxa);
    double *geo = (double *)malloc(sizeof(double) * (size_t)n_patterns);
    long loci = 0;
    int id = 0;
    for (int p = 0; p < n_patterns; p++) {
        loci += weights ? weights[p] : 1;
        for (int t = 0; t < n_taxa; t++) {
            int cov = counts[((size_t)p * n_taxa + t) * 5 + 4];
            processed[id + t] = zero_cov_mode == 0 ? cov : cov + 1;
        }
        geo[p] = so_geometric_mean(processed, id, id + n_taxa);
        id += n_taxa;
    }
    if (stats_out) {
        /* DiscreteStatistics.mean / variance (n - 1 denominator) over processedSeqCov */
        size_t N = (size_t)n_patterns * n_taxa;
        double m = 0;
        for (size_t i = 0; i < N; i++) m += processed[i];
        m /= (double)N;
        double var = 0;
        for (size_t i = 0; i < N; i++) {
            double d = processed[i] - m;
            var += d * d;
        }
        var /= (double)(N - 1);
        stats_out[0] = m / 2;
        stats_out[1] = var / 4;
    }
    double *inter = (double *)malloc(sizeof(double) * (size_t)loci);
    for (int t = 0; t < n_taxa; t++) {
        long index = 0;
        for (int p = 0; p < n_patterns; p++) {
            int w = weights ? weights[p] : 1;
            int c = processed[(size_t)p * n_taxa + t];
            if (zero_cov_mode == 1 || (c > 0 && geo[p] > 0)) {
                for (int k = 0; k < w; k++) inter[index + k] = c / geo[p];
                index += w;
            }
        }
        /* DiscreteStatistics.median: sort, odd -> middle, even -> mean of the two middles */
        qsort(inter, (size_t)index, sizeof(double), so_cmp_double);
        if (index == 0)
            size_factors[t] = NAN;
        else if (index % 2 == 1)
            size_factors[t] = inter[index / 2];
        else
            size_factors[t] = (inter[index / 2 - 1] + inter[index / 2]) / 2.0;
    }
    free(inter);
    free(geo);
    free(processed);
}

/* ------------------------------------------------------------------------------------------ */
/* Leaf model                                                                                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double ado_rate;        /* theta, RawReadCountsModelInterface.java:44-48 */
    double allelic_cov;     /* t,     SeqCovModelInterface.java:30-67 */
    double allelic_raw_var; /* v */
    double eff_seq_err;     /* eps,   NucReadCountsModelInterface.java:18-34 */
    double shape1;          /* w1 */
    double shape2;          /* w2 */
    int single_ado;         /* RawReadCountsModelInterface.java:178-181 (default true) */
    int use_log;            /* useLogPartials */
} so_leaf_params;

/* seqcovmodel/ExploredSharedAllelicSeqCovModel.java:266-277 */
double so_seq_cov_likelihood(int cov, int n_alleles, double size_factor, const so_leaf_params *P)
{
    const double stab = 1e-6; /* SeqCovModelInterface.java:218 */
    const double mean = (n_alleles + stab) * P->allelic_cov * size_factor;
    const double var = mean + pow(size_factor, 2) * pow(n_alleles + stab, 2) * P->allelic_raw_var;
    const double p = mean / var;
    const double r = pow(mean, 2) / (var - mean);
    double l = so_nb_log_density(cov, r, p);
    return P->use_log ? l : exp(l);
}

/* nucreadcountsmodel/NucReadCountsModelFiniteMuDM.java:33-96 -- out[0..3] = 0/0, 1/1, 1/1', 0/1 */
void so_nuc_likelihoods(const int *reads5, const so_leaf_params *P, double *out4)
{
    const double w1 = P->shape1, w2 = P->shape2;
    const double oneThird = P->eff_seq_err / 3;
    const double oneMinus = 1 - P->eff_seq_err;
    const double halfMinus = 0.5 - P->eff_seq_err / 3;
    const double f1 = oneThird * w1, m1 = oneMinus * w1;
    const double f2 = oneThird * w2, h2 = halfMinus * w2;
    const double a0[5] = {f1, f1, f1, m1, w1};
    const double a1[5] = {m1, f1, f1, f1, w1};
    const double a2[5] = {h2, h2, f2, f2, w2};
    const double a3[5] = {h2, f2, f2, h2, w2};
    out4[0] = so_dm_log_density(reads5, a0);
    out4[1] = so_dm_log_density(reads5, a1);
    out4[2] = so_dm_log_density(reads5, a2);
    out4[3] = so_dm_log_density(reads5, a3);
    if (!P->use_log)
        for (int i = 0; i < 4; i++) out4[i] = exp(out4[i]);
}

/* rawreadcountsmodel/RawReadCountsModelFiniteMuDM.java:121-260.
 * N = nuc likelihoods (index 0..3 as above), C = coverage likelihoods indexed by position in
 * modeledAlleles ({1,2} for single ADO, {0,1,2} for locus ADO); genotype 0..3 = 0/0,0/1,1/1,1/1'. */
static double so_mixed_core(int genotype, const double *N, const double *C, const so_leaf_params *P, double *comp)
{
    const double theta = P->ado_rate;
    double tmp_local[3], in2[2];
    double *tmp = comp ? comp : tmp_local; /* :130 "comp == null ? new double[...] : comp" */
    /* which nuc category carries the "no dropout" term, which the "one allele left" term */
    if (P->single_ado) {
        if (P->use_log) {
            switch (genotype) {
            case 0:
                tmp[0] = log(1 - theta) + N[0] + C[1];
                tmp[1] = log(theta) + N[0] + C[0];
                return so_log_sum_exp(tmp, 2);
            case 1:
                tmp[0] = log(1 - theta) + N[3] + C[1];
                in2[0] = N[0];
                in2[1] = N[1];
                tmp[1] = log(theta) - log(2) + so_log_sum_exp(in2, 2) + C[0];
                return so_log_sum_exp(tmp, 2);
            case 2:
                tmp[0] = log(1 - theta) + N[1] + C[1];
                tmp[1] = log(theta) + N[1] + C[0];
                return so_log_sum_exp(tmp, 2);
            default:
                tmp[0] = log(1 - theta) + N[2] + C[1];
                tmp[1] = log(theta) + N[1] + C[0];
                return so_log_sum_exp(tmp, 2);
            }
        } else {
            switch (genotype) {
            case 0:
                tmp[0] = (1 - theta) * N[0] * C[1];
                tmp[1] = theta * N[0] * C[0];
                return tmp[0] + tmp[1];
            case 1:
                tmp[0] = (1 - theta) * N[3] * C[1];
                tmp[1] = (theta / 2) * (N[0] + N[1]) * C[0];
                return tmp[0] + tmp[1];
            case 2:
                tmp[0] = (1 - theta) * N[1] * C[1];
                tmp[1] = theta * N[1] * C[0];
                return tmp[0] + tmp[1];
            default:
                tmp[0] = (1 - theta) * N[2] * C[1];
                tmp[1] = theta * N[1] * C[0];
                return tmp[0] + tmp[1];
            }
        }
    } else { /* locus ADO */
        if (P->use_log) {
            switch (genotype) {
            case 0:
                tmp[0] = 2 * log(1 - theta) + N[0] + C[2];
                tmp[1] = log(2) + log(theta) + log(1 - theta) + N[0] + C[1];
                tmp[2] = 2 * log(theta) + C[0];
                return so_log_sum_exp(tmp, 3);
            case 1:
                tmp[0] = 2 * log(1 - theta) + N[3] + C[2];
                in2[0] = N[0];
                in2[1] = N[1];
                tmp[1] = log(theta) + log(1 - theta) + so_log_sum_exp(in2, 2) + C[1];
                tmp[2] = 2 * log(theta) + C[0];
                return so_log_sum_exp(tmp, 3);
            case 2:
                tmp[0] = 2 * log(1 - theta) + N[1] + C[2];
                tmp[1] = log(2) + log(theta) + log(1 - theta) + N[1] + C[1];
                tmp[2] = 2 * log(theta) + C[0];
                return so_log_sum_exp(tmp, 3);
            default:
                tmp[0] = 2 * log(1 - theta) + N[2] + C[2];
                tmp[1] = log(2) + log(theta) + log(1 - theta) + N[1] + C[1];
                tmp[2] = 2 * log(theta) + C[0];
                return so_log_sum_exp(tmp, 3);
            }
        } else {
            switch (genotype) {
            case 0:
                tmp[0] = pow(1 - theta, 2) * N[0] * C[2];
                tmp[1] = 2 * theta * (1 - theta) * N[0] * C[1];
                tmp[2] = pow(theta, 2) * C[0];
                return tmp[0] + tmp[1] + tmp[2];
            case 1:
                tmp[0] = pow(1 - theta, 2) * N[3] * C[2];
                tmp[1] = theta * (1 - theta) * (N[0] + N[1]) * C[1];
                tmp[2] = pow(theta, 2) * C[0];
                return tmp[0] + tmp[1] + tmp[2];
            case 2:
                tmp[0] = pow(1 - theta, 2) * N[1] * C[2];
                tmp[1] = 2 * theta * (1 - theta) * N[1] * C[1];
                tmp[2] = pow(theta, 2) * C[0];
                return tmp[0] + tmp[1] + tmp[2];
            default:
                tmp[0] = pow(1 - theta, 2) * N[2] * C[2];
                tmp[1] = 2 * theta * (1 - theta) * N[1] * C[1];
                tmp[2] = pow(theta, 2) * C[0];
                return tmp[0] + tmp[1] + tmp[2];
            }
        }
    }
}

double so_mixed_likelihood(int genotype, const double *N, const double *C, const so_leaf_params *P)
{
    return so_mixed_core(genotype, N, C, P, NULL);
}

/* math/util/MathFunctions.java:168-180 maxIndices, as a bit mask (bit i set <=> index i is listed) */
static unsigned so_max_indices_mask(const double *v, int n)
{
    double max = v[0];
    unsigned mask = 1u;
    for (int i = 1; i < n; i++) {
        if (v[i] > max) {
            max = v[i];
            mask = 1u << i;
        } else if (v[i] == max)
            mask |= 1u << i;
    }
    return mask;
}

/* RawReadCountsModelFiniteMuDM.java:54-77 computeMixedLikelihood(..., MLCompIndex): the mixed likelihood plus the
 * arg-max set of its dropout components (index = number of dropped alleles), used when traceMLGenotypes is on
 * (RawReadCountsModelInterface.java:515-567).  out: [pattern][4]; ado: [pattern][4] tie masks. */
void so_leaf_likelihood_ml(const int *counts, int n_taxa, int p0, int p1, int taxon, double size_factor,
                           const so_leaf_params *P, double *out, unsigned char *ado)
{
    const int n_all = P->single_ado ? 2 : 3;
    const int first = P->single_ado ? 1 : 0;
    for (int p = p0; p < p1; p++) {
        const int *reads = counts + ((size_t)p * n_taxa + taxon) * 5;
        double C[3], N[4], comp[3];
        for (int i = 0; i < n_all; i++) C[i] = so_seq_cov_likelihood(reads[4], first + i, size_factor, P);
        so_nuc_likelihoods(reads, P, N);
        for (int g = 0; g < 4; g++) {
            out[(size_t)(p - p0) * 4 + g] = so_mixed_core(g, N, C, P, comp);
            ado[(size_t)(p - p0) * 4 + g] = (unsigned char)so_max_indices_mask(comp, n_all);
        }
    }
}

/* rawreadcountsmodel/RawReadCountsModelInterface.java:458-506 (all sub-models dirty) == :316-359.
 * counts: [pattern][taxon][5]; out: [pattern][4] for one taxon (nrOfMatrices == 1,
 * ExploredSharedAllelicSeqCovModel.java:94).  p0..p1 restricts the pattern range. */
void so_leaf_likelihood(const int *counts, int n_taxa, int p0, int p1, int taxon, double size_factor,
                        const so_leaf_params *P, double *out)
{
    /* modeledAlleles = {2} from ScsFiniteMuExtendedModel.java:31, widened by ADO,
     * RawReadCountsModelInterface.java:222-233 */
    const int n_all = P->single_ado ? 2 : 3;
    const int first = P->single_ado ? 1 : 0;
    for (int p = p0; p < p1; p++) {
        const int *reads = counts + ((size_t)p * n_taxa + taxon) * 5;
        double C[3], N[4];
        for (int i = 0; i < n_all; i++) C[i] = so_seq_cov_likelihood(reads[4], first + i, size_factor, P);
        so_nuc_likelihoods(reads, P, N);
        for (int g = 0; g < 4; g++) out[(size_t)(p - p0) * 4 + g] = so_mixed_likelihood(g, N, C, P);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Likelihood core: likelihood/ScsBeerLikelihoodCore.java (+ external BeerLikelihoodCore)      */
/* ------------------------------------------------------------------------------------------ */
typedef struct so_core {
    int n_nodes, n_leaves, n_internal, n_patterns, n_matrices, n_states;
    int use_log, use_scaling;
    size_t partials_size, matrix_size;
    double *partials[2];       /* [buf][node][K*S*G] */
    double *const_partials[2]; /* [buf][internal][K*S*G] */
    double *matrices[2];       /* [buf][node][K*G*G] */
    double *scaling[2];        /* [buf][node][S] */
    double *const_scaling[2];  /* [buf][internal][S] */
    int *cur_matrix, *stored_matrix, *cur_partials, *stored_partials;
} so_core;

/* ScsBeerLikelihoodCore.java:91-128 (integrateCategories == true) */
so_core *so_core_create(int node_count, int leaf_count, int pattern_count, int matrix_count, int state_count,
                        int use_log)
{
    so_core *c = (so_core *)calloc(1, sizeof(so_core));
    c->n_nodes = node_count;
    c->n_leaves = leaf_count;
    c->n_internal = node_count - leaf_count;
    c->n_patterns = pattern_count;
    c->n_matrices = matrix_count;
    c->n_states = state_count;
    c->use_log = use_log;
    c->partials_size = (size_t)matrix_count * pattern_count * state_count;
    c->matrix_size = (size_t)state_count * state_count;
    for (int b = 0; b < 2; b++) {
        c->partials[b] = (double *)calloc(c->partials_size * node_count, sizeof(double));
        c->const_partials[b] = (double *)calloc(c->partials_size * c->n_internal, sizeof(double));
        c->matrices[b] = (double *)calloc(c->matrix_size * matrix_count * node_count, sizeof(double));
        c->scaling[b] = NULL;
        c->const_scaling[b] = NULL;
    }
    c->cur_matrix = (int *)calloc(node_count, sizeof(int));
    c->stored_matrix = (int *)calloc(node_count, sizeof(int));
    c->cur_partials = (int *)calloc(node_count, sizeof(int));
    c->stored_partials = (int *)calloc(node_count, sizeof(int));
    return c;
}

void so_core_free(so_core *c)
{
    if (!c) return;
    for (int b = 0; b < 2; b++) {
        free(c->partials[b]);
        free(c->const_partials[b]);
        free(c->matrices[b]);
        free(c->scaling[b]);
        free(c->const_scaling[b]);
    }
    free(c->cur_matrix);
    free(c->stored_matrix);
    free(c->cur_partials);
    free(c->stored_partials);
    free(c);
}

static double *so_part(so_core *c, int node) { return c->partials[c->cur_partials[node]] + c->partials_size * node; }
static double *so_cpart(so_core *c, int node)
{
    return c->const_partials[c->cur_partials[node]] + c->partials_size * (node - c->n_leaves);
}
static double *so_mat(so_core *c, int node)
{
    return c->matrices[c->cur_matrix[node]] + c->matrix_size * c->n_matrices * node;
}

/* ScsBeerLikelihoodCore.java:383-393 */
void so_core_set_use_scaling(so_core *c, double scale)
{
    c->use_scaling = (scale != 1.0);
    if (c->use_scaling) {
        for (int b = 0; b < 2; b++) {
            free(c->scaling[b]);
            free(c->const_scaling[b]);
            c->scaling[b] = (double *)calloc((size_t)c->n_nodes * c->n_patterns, sizeof(double));
            c->const_scaling[b] = (double *)calloc((size_t)c->n_internal * c->n_patterns, sizeof(double));
        }
    }
}

/* ScsBeerLikelihoodCore.java:476-573 (full-pattern overload) */
static void so_scale_partials(so_core *c, int node)
{
    const double threshold = 1.0E-100; /* :66 */
    const int S = c->n_patterns, G = c->n_states, K = c->n_matrices;
    double *part = so_part(c, node);
    double *cpart = node >= c->n_leaves ? so_cpart(c, node) : NULL;
    double *sf = c->scaling[c->cur_partials[node]] + (size_t)node * S;
    double *csf = cpart ? c->const_scaling[c->cur_partials[node]] + (size_t)(node - c->n_leaves) * S : NULL;
    size_t u = 0;
    for (int i = 0; i < S; i++) {
        double scale = 0.0, cscale = 0.0;
        size_t v = u;
        for (int k = 0; k < K; k++) {
            for (int j = 0; j < G; j++) {
                if (part[v] > scale) scale = part[v];
                if (cpart && cpart[v] > cscale) cscale = cpart[v];
                v++;
            }
            v += (size_t)(S - 1) * G;
        }
        if (scale < threshold) {
            v = u;
            for (int k = 0; k < K; k++) {
                for (int j = 0; j < G; j++) {
                    if (c->use_log)
                        part[v] -= scale;
                    else
                        part[v] /= scale;
                    v++;
                }
                v += (size_t)(S - 1) * G;
            }
            sf[i] = c->use_log ? scale : log(scale);
        } else
            sf[i] = 0.0;
        if (cpart) {
            if (cscale < threshold) {
                v = u;
                for (int k = 0; k < K; k++) {
                    for (int j = 0; j < G; j++) {
                        if (c->use_log)
                            cpart[v] -= cscale;
                        else
                            cpart[v] /= cscale;
                        v++;
                    }
                    v += (size_t)(S - 1) * G;
                }
                csf[i] = c->use_log ? cscale : log(cscale);
            } else
                csf[i] = 0.0;
        }
        u += G;
    }
}

/* ScsBeerLikelihoodCore.java:414-429: a [S*G] input is replicated over the K matrices */
void so_core_set_node_partials(so_core *c, int node, const double *partials, size_t len)
{
    double *dst = so_part(c, node);
    if (len < c->partials_size) {
        size_t k = 0;
        for (int i = 0; i < c->n_matrices; i++) {
            memcpy(dst + k, partials, len * sizeof(double));
            k += len;
        }
    } else
        memcpy(dst, partials, len * sizeof(double));
    if (c->use_scaling) so_scale_partials(c, node);
}

/* ScsBeerLikelihoodCore.java:437-445 */
void so_core_store_leaf_partials(so_core *c, int node)
{
    memcpy(c->partials[1 - c->cur_partials[node]] + c->partials_size * node, so_part(c, node),
           c->partials_size * sizeof(double));
}

/* external BeerLikelihoodCore: flip the write buffer of one node */
void so_core_set_node_matrix_for_update(so_core *c, int node) { c->cur_matrix[node] = 1 - c->cur_matrix[node]; }
void so_core_set_node_partials_for_update(so_core *c, int node)
{
    c->cur_partials[node] = 1 - c->cur_partials[node];
}
void so_core_set_node_matrix(so_core *c, int node, int cat, const double *P)
{
    memcpy(so_mat(c, node) + c->matrix_size * cat, P, c->matrix_size * sizeof(double));
}
void so_core_get_node_partials(so_core *c, int node, double *out)
{
    memcpy(out, so_part(c, node), c->partials_size * sizeof(double));
}
void so_core_get_const_partials(so_core *c, int node, double *out)
{
    memcpy(out, so_cpart(c, node), c->partials_size * sizeof(double));
}

/* ScsBeerLikelihoodCore.java:4053-4065 and external restore() (index-array swap) */
void so_core_store(so_core *c)
{
    memcpy(c->stored_matrix, c->cur_matrix, sizeof(int) * (c->n_nodes - 1));
    memcpy(c->stored_partials, c->cur_partials, sizeof(int) * c->n_nodes);
}
void so_core_unstore(so_core *c)
{
    memcpy(c->cur_matrix, c->stored_matrix, sizeof(int) * (c->n_nodes - 1));
    memcpy(c->cur_partials, c->stored_partials, sizeof(int) * c->n_nodes);
}
void so_core_restore(so_core *c)
{
    int *t = c->cur_matrix;
    c->cur_matrix = c->stored_matrix;
    c->stored_matrix = t;
    t = c->cur_partials;
    c->cur_partials = c->stored_partials;
    c->stored_partials = t;
}

/* ScsBeerLikelihoodCore.java:1384-1501 (log space).  ch*c == NULL <=> that child is a leaf. */
static void so_log_pruning(const so_core *c, const double *ch1, const double *ch1c, const double *m1,
                           const double *ch2, const double *ch2c, const double *m2, double *par, double *parc,
                           int const_genotype)
{
    const int K = c->n_matrices, S = c->n_patterns, G = c->n_states;
    const int has2 = (ch2 != NULL && m2 != NULL);
    double cst1, cst2;
    double cstArr1[8], cstArr2[8], sp1[8], sp2[8];
    double tmp1, tmp2;
    size_t pIndex = 0, cIndex = 0, mIndex;
    for (int k = 0; k < K; k++) {
        for (int s = 0; s < S; s++) {
            mIndex = (size_t)k * c->matrix_size;
            for (int pg = 0; pg < G; pg++) {
                cst1 = cst2 = 0.0;
                tmp1 = tmp2 = 0.0;
                for (int cg = 0; cg < G; cg++) {
                    tmp1 = log(m1[mIndex] > 0.0 ? m1[mIndex] : SO_MIN_VALUE) + ch1[cIndex + cg];
                    if (ch1c == NULL) {
                        if (cg == const_genotype) cst1 = tmp1;
                    } else
                        cstArr1[cg] = log(m1[mIndex] > 0.0 ? m1[mIndex] : SO_MIN_VALUE) + ch1c[cIndex + cg];
                    sp1[cg] = tmp1;
                    if (has2) {
                        tmp2 = log(m2[mIndex] > 0.0 ? m2[mIndex] : SO_MIN_VALUE) + ch2[cIndex + cg];
                        if (ch2c == NULL) {
                            if (cg == const_genotype) cst2 = tmp2;
                        } else
                            cstArr2[cg] = log(m2[mIndex] > 0.0 ? m2[mIndex] : SO_MIN_VALUE) + ch2c[cIndex + cg];
                        sp2[cg] = tmp2;
                    }
                    mIndex++;
                }
                if (has2) {
                    if (ch1c == NULL) {
                        if (ch2c == NULL)
                            parc[pIndex] = cst1 + cst2;
                        else
                            parc[pIndex] = cst1 + so_log_sum_exp(cstArr2, G);
                    } else {
                        if (ch2c == NULL)
                            parc[pIndex] = so_log_sum_exp(cstArr1, G) + cst2;
                        else
                            parc[pIndex] = so_log_sum_exp(cstArr1, G) + so_log_sum_exp(cstArr2, G);
                    }
                    par[pIndex] = so_log_sum_exp(sp1, G) + so_log_sum_exp(sp2, G);
                } else {
                    if (ch1c == NULL)
                        parc[pIndex] = cst1;
                    else
                        parc[pIndex] = so_log_sum_exp(cstArr1, G);
                    par[pIndex] = so_log_sum_exp(sp1, G);
                }
                pIndex++;
            }
            cIndex += G;
        }
    }
}

/* ScsBeerLikelihoodCore.java:958-1055 (linear space) */
static void so_lin_pruning(const so_core *c, const double *ch1, const double *ch1c, const double *m1,
                           const double *ch2, const double *ch2c, const double *m2, double *par, double *parc,
                           int const_genotype)
{
    const int K = c->n_matrices, S = c->n_patterns, G = c->n_states;
    const int has2 = (ch2 != NULL && m2 != NULL);
    double cst1, cst2, sum1, sum2, tmp1, tmp2;
    size_t pIndex = 0, cIndex = 0, mIndex;
    for (int k = 0; k < K; k++) {
        for (int s = 0; s < S; s++) {
            mIndex = (size_t)k * c->matrix_size;
            for (int pg = 0; pg < G; pg++) {
                cst1 = cst2 = 0.0;
                sum1 = sum2 = 0.0;
                for (int cg = 0; cg < G; cg++) {
                    tmp1 = m1[mIndex] * ch1[cIndex + cg];
                    if (ch1c == NULL) {
                        if (cg == const_genotype) cst1 = tmp1;
                    } else
                        cst1 += m1[mIndex] * ch1c[cIndex + cg];
                    sum1 += tmp1;
                    if (has2) {
                        tmp2 = m2[mIndex] * ch2[cIndex + cg];
                        if (ch2c == NULL) {
                            if (cg == const_genotype) cst2 = tmp2;
                        } else
                            cst2 += m2[mIndex] * ch2c[cIndex + cg];
                        sum2 += tmp2;
                    }
                    mIndex++;
                }
                if (has2) {
                    parc[pIndex] = cst1 * cst2 > 0.0 ? cst1 * cst2 : SO_MIN_VALUE;
                    par[pIndex] = sum1 * sum2;
                } else {
                    parc[pIndex] = cst1 > 0.0 ? cst1 : SO_MIN_VALUE;
                    par[pIndex] = sum1;
                }
                pIndex++;
            }
            cIndex += G;
        }
    }
}

/* ScsBeerLikelihoodCore.java:1283-1369 (log) / :857-943 (linear); child2 == -1 <=> unary root */
void so_core_calculate_partials(so_core *c, int child1, int is_leaf1, int child2, int is_leaf2, int parent,
                                int const_genotype)
{
    const double *ch1 = so_part(c, child1);
    const double *ch1c = is_leaf1 ? NULL : so_cpart(c, child1);
    const double *m1 = so_mat(c, child1);
    const double *ch2 = NULL, *ch2c = NULL, *m2 = NULL;
    if (child2 >= 0) {
        ch2 = so_part(c, child2);
        ch2c = is_leaf2 ? NULL : so_cpart(c, child2);
        m2 = so_mat(c, child2);
    }
    if (c->use_log)
        so_log_pruning(c, ch1, ch1c, m1, ch2, ch2c, m2, so_part(c, parent), so_cpart(c, parent), const_genotype);
    else
        so_lin_pruning(c, ch1, ch1c, m1, ch2, ch2c, m2, so_part(c, parent), so_cpart(c, parent), const_genotype);
    if (c->use_scaling) so_scale_partials(c, parent);
}

/* Max-sum twin of the log-space pruning, ScsBeerLikelihoodCore.java:3054-3173 (leaf, leaf), :3175-3322 (leaf,
 * internal), :3324-3486 (internal, internal / unary root).  The three reference functions differ only in which array
 * a child contributes: a leaf its (sum-product) partials, an internal node its MLPartials -- the caller passes the
 * right one as ch1 / ch2.  ch2 == NULL <=> unary root.  Layouts as the core: ch*, par_ml [K][S][G]; m* [K][G*G].
 * back[K][S][G]: low nibble = arg-max set over the genotypes of child 1 (parentMLGenotypes[pIndex][0]), high nibble =
 * child 2, bit c set <=> genotype c is listed; ties by exact equality, first index seeds the list (:3107-3131). */
void so_ml_log_pruning(int K, int S, const double *ch1, const double *m1, const double *ch2, const double *m2,
                       double *par_ml, unsigned char *back)
{
    const int G = 4;
    size_t pIndex = 0, cIndex = 0;
    for (int k = 0; k < K; k++) {
        for (int s = 0; s < S; s++) {
            size_t mIndex = (size_t)k * G * G;
            for (int pg = 0; pg < G; pg++) {
                double max1 = 0.0, max2 = 0.0;
                unsigned l1 = 0, l2 = 0;
                for (int cg = 0; cg < G; cg++) {
                    const double tmp1 = log(m1[mIndex] > 0.0 ? m1[mIndex] : SO_MIN_VALUE) + ch1[cIndex + cg];
                    double tmp2 = 0.0;
                    if (ch2) tmp2 = log(m2[mIndex] > 0.0 ? m2[mIndex] : SO_MIN_VALUE) + ch2[cIndex + cg];
                    if (cg == 0) {
                        max1 = tmp1;
                        l1 = 1u;
                        if (ch2) {
                            max2 = tmp2;
                            l2 = 1u;
                        }
                    } else {
                        if (max1 < tmp1) {
                            max1 = tmp1;
                            l1 = 1u << cg;
                        } else if (max1 == tmp1)
                            l1 |= 1u << cg;
                        if (ch2) {
                            if (max2 < tmp2) {
                                max2 = tmp2;
                                l2 = 1u << cg;
                            } else if (max2 == tmp2)
                                l2 |= 1u << cg;
                        }
                    }
                    mIndex++;
                }
                par_ml[pIndex] = ch2 ? max1 + max2 : max1;
                back[pIndex] = (unsigned char)(l1 | (l2 << 4));
                pIndex++;
            }
            cIndex += G;
        }
    }
}

/* ScsBeerLikelihoodCore.java:278-296 getPatternCategories: arg-max set over the categories of the root's sum-product
 * partial at the root genotype, as a bit mask per pattern. */
void so_pattern_categories(int K, int S, const double *root_partials, int root_genotype, unsigned char *out)
{
    for (int i = 0; i < S; i++) {
        double max = -1.7976931348623157e308;
        unsigned mask = 0;
        for (int j = 0; j < K; j++) {
            const double tmp = root_partials[((size_t)j * S + i) * 4 + root_genotype];
            if (max < tmp) {
                max = tmp;
                mask = 1u << j;
            } else if (max == tmp)
                mask |= 1u << j;
        }
        out[i] = (unsigned char)mask;
    }
}

/* ScsBeerLikelihoodCore.java:153-218 */
void so_core_integrate_partials(so_core *c, int node, const double *proportions, int root_genotype, double *out,
                                double *const_root)
{
    const int K = c->n_matrices, S = c->n_patterns, G = c->n_states;
    const double *in = so_part(c, node);
    const double *cin = so_cpart(c, node);
    if (c->use_log) {
        double o[16], cs[16];
        for (int s = 0; s < S; s++) {
            size_t inIndex = (size_t)s * G + root_genotype;
            for (int k = 0; k < K; k++) {
                o[k] = log(proportions[k]) + in[inIndex];
                cs[k] = log(proportions[k]) + cin[inIndex];
                inIndex += (size_t)S * G;
            }
            out[s] = so_log_sum_exp(o, K);
            const_root[s] = so_log_sum_exp(cs, K);
        }
    } else {
        size_t inIndex = root_genotype;
        for (int s = 0; s < S; s++) {
            out[s] = in[inIndex] * proportions[0];
            const_root[s] = cin[inIndex] * proportions[0];
            inIndex += G;
        }
        for (int k = 1; k < K; k++) {
            for (int s = 0; s < S; s++) {
                out[s] += in[inIndex] * proportions[k];
                const_root[s] += cin[inIndex] * proportions[k];
                inIndex += G;
            }
        }
    }
}

/* ScsBeerLikelihoodCore.java:255-269 with external getLogScalingFactor / :830-839 */
void so_core_calculate_log_likelihoods(so_core *c, const double *in, double *out, const double *const_root,
                                       double *log_const_root)
{
    const int S = c->n_patterns;
    for (int s = 0; s < S; s++) {
        double ls = 0.0, cls = 0.0;
        if (c->use_scaling) {
            for (int i = 0; i < c->n_nodes; i++) ls += c->scaling[c->cur_partials[i]][(size_t)i * S + s];
            for (int i = c->n_leaves; i < c->n_nodes; i++)
                cls += c->const_scaling[c->cur_partials[i]][(size_t)(i - c->n_leaves) * S + s];
        }
        out[s] = ls;
        log_const_root[s] = cls;
        if (c->use_log) {
            out[s] += in[s];
            log_const_root[s] += const_root[s];
        } else {
            out[s] += log(in[s]);
            log_const_root[s] += log(const_root[s]);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* logP and the acquisition-bias terms                                                         */
/* ------------------------------------------------------------------------------------------ */

/* likelihood/ScsTreeLikelihood.java:2621-2630 */
double so_sum_pattern_logl(const double *pattern_logl, const int *weights, int S)
{
    double logP = 0.0;
    for (int i = 0; i < S; i++) logP += pattern_logl[i] * (weights ? weights[i] : 1);
    return logP;
}

/* alignment/ScsAlignment.java:746-752 */
double so_asc_bias_lewis(const double *log_const_root, const int *weights, int S)
{
    double corr = 0.0;
    for (int i = 0; i < S; i++) corr -= log(1 - exp(log_const_root[i])) * (weights ? weights[i] : 1);
    return corr;
}

/* alignment/ScsAlignment.java:764-782; the input is expanded by pattern weight as
 * ScsTreeLikelihood.java:2624-2630 does.  mode: 0 = mean, 1 = geometric. */
double so_asc_bias_felsenstein(const double *log_const_root, const int *weights, int S, int mode, double M,
                               int return_const_sum)
{
    long len = 0;
    for (int i = 0; i < S; i++) len += weights ? weights[i] : 1;
    double *all = (double *)malloc(sizeof(double) * (size_t)len);
    long j = 0;
    for (int i = 0; i < S; i++)
        for (int w = 0; w < (weights ? weights[i] : 1); w++) all[j++] = log_const_root[i];
    double corr, corrected;
    if (mode == 0) {
        corr = so_log_sum_exp(all, (int)len);
        corrected = (corr - log((double)len)) * M;
    } else {
        corr = 0.0;
        for (long i = 0; i < len; i++) corr += all[i];
        corrected = corr * M / (double)len;
    }
    free(all);
    return return_const_sum ? corr : corrected;
}

/* alignment/ScsAlignment.java:811-820 (combine of per-thread constant sums) */
double so_asc_bias_threaded(const double *const_sum_by_thread, int T, int mode, long loci_nr, double M)
{
    if (mode == 0) return (so_log_sum_exp(const_sum_by_thread, T) - log((double)loci_nr)) * M;
    double s = 0.0;
    for (int i = 0; i < T; i++) s += const_sum_by_thread[i];
    return s * M / (double)T; /* logConstRoot.length is the thread count there (:816) */
}

/* likelihood/ThreadedScsTreeLikelihood.java:227-238 (no 'proportions' input) */
void so_pattern_points(int n_sites, int T, int *points)
{
    points[0] = 0;
    for (int i = 0; i < T; i++) {
        int cnt = n_sites / T + (i < n_sites % T ? 1 : 0);
        points[i + 1] = points[i] + cnt;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Whole evaluation, parallelised like ThreadedScsTreeLikelihood.java:176-219, 366-389:        */
/* contiguous site ranges, one private core per worker, per-worker (logP, constSum) combined    */
/* in worker order.  Used as the timed CPU baseline and as a one-call checker.                  */
/*   counts    [S][n][5]                                                                        */
/*   ops       [n_ops][3] = (parent, child1, child2 | -1), post-order; leaves are 0..n-1        */
/*   matrices  [2n][K][16]  (row = parent state, col = child state), entry of the root unused   */
/*   asc_mode  0 none, 1 felsenstein/mean, 2 felsenstein/geometric, 3 lewis                     */
/* outputs: site_logl[S], site_const[S] (may be NULL), returns logP                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double logp;
    double const_sum;
} so_thread_out;

static void so_eval_range(const int *counts, const int *weights, int n, int p0, int p1, int K,
                          const double *size_factors, const so_leaf_params *P, const int *ops, int n_ops,
                          const double *matrices, const double *proportions, int root, int asc_mode, double M,
                          int threaded, double *site_logl, double *site_const, so_thread_out *out)
{
    const int S = p1 - p0;
    const int G = 4;
    so_core *c = so_core_create(2 * n, n, S, K, G, P->use_log);
    double *leaf = (double *)malloc(sizeof(double) * (size_t)S * G);
    /* ScsTreeLikelihood.java:583-601: every leaf refreshed, then copied into the core */
    for (int t = 0; t < n; t++) {
        so_leaf_likelihood(counts, n, p0, p1, t, size_factors[t], P, leaf);
        so_core_set_node_partials(c, t, leaf, (size_t)S * G);
    }
    /* ScsTreeLikelihood.java:563-580: matrices of every non-root node */
    for (int node = 0; node < 2 * n; node++) {
        if (node == root) continue;
        for (int k = 0; k < K; k++) so_core_set_node_matrix(c, node, k, matrices + ((size_t)node * K + k) * 16);
    }
    for (int i = 0; i < n_ops; i++) {
        int parent = ops[3 * i], c1 = ops[3 * i + 1], c2 = ops[3 * i + 2];
        so_core_calculate_partials(c, c1, c1 < n, c2, c2 >= 0 && c2 < n, parent, 0);
    }
    double *rootp = (double *)malloc(sizeof(double) * S);
    double *croot = (double *)malloc(sizeof(double) * S);
    double *ll = (double *)malloc(sizeof(double) * S);
    double *cll = (double *)malloc(sizeof(double) * S);
    so_core_integrate_partials(c, root, proportions, 0, rootp, croot);
    so_core_calculate_log_likelihoods(c, rootp, ll, croot, cll);
    const int *w = weights ? weights + p0 : NULL;
    out->logp = so_sum_pattern_logl(ll, w, S);
    out->const_sum = 0.0;
    if (asc_mode == 1 || asc_mode == 2) {
        double v = so_asc_bias_felsenstein(cll, w, S, asc_mode - 1, M, threaded);
        if (threaded)
            out->const_sum = v;
        else
            out->logp += v;
    } else if (asc_mode == 3)
        out->logp += so_asc_bias_lewis(cll, w, S);
    if (site_logl) memcpy(site_logl + p0, ll, sizeof(double) * S);
    if (site_const) memcpy(site_const + p0, cll, sizeof(double) * S);
    free(rootp);
    free(croot);
    free(ll);
    free(cll);
    free(leaf);
    so_core_free(c);
}

typedef struct {
    const int *counts;
    const int *weights;
    int n, p0, p1, K;
    const double *size_factors;
    const so_leaf_params *P;
    const int *ops;
    int n_ops;
    const double *matrices;
    const double *proportions;
    int root, asc_mode;
    double M;
    double *site_logl, *site_const;
    so_thread_out *out;
} so_job;

/* one worker == one ScsTreeLikelihoodCaller (ThreadedScsTreeLikelihood.java:457-480) */
static void *so_job_run(void *arg)
{
    so_job *j = (so_job *)arg;
    so_eval_range(j->counts, j->weights, j->n, j->p0, j->p1, j->K, j->size_factors, j->P, j->ops, j->n_ops,
                  j->matrices, j->proportions, j->root, j->asc_mode, j->M, 1, j->site_logl, j->site_const, j->out);
    return NULL;
}

double so_full_eval(const int *counts, const int *weights, int n, int S, int K, const double *size_factors,
                    const so_leaf_params *P, const int *ops, int n_ops, const double *matrices,
                    const double *proportions, int root, int asc_mode, double M, int n_threads, double *site_logl,
                    double *site_const)
{
    if (n_threads <= 1) {
        so_thread_out o;
        so_eval_range(counts, weights, n, 0, S, K, size_factors, P, ops, n_ops, matrices, proportions, root,
                      asc_mode, M, 0, site_logl, site_const, &o);
        return o.logp;
    }
    int *points = (int *)malloc(sizeof(int) * (n_threads + 1));
    so_thread_out *outs = (so_thread_out *)malloc(sizeof(so_thread_out) * n_threads);
    so_pattern_points(S, n_threads, points);
    so_job *jobs = (so_job *)malloc(sizeof(so_job) * n_threads);
    pthread_t *tids = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    for (int t = 0; t < n_threads; t++) {
        so_job j = {counts, weights, n, points[t], points[t + 1], K, size_factors, P, ops, n_ops, matrices,
                    proportions, root, asc_mode, M, site_logl, site_const, &outs[t]};
        jobs[t] = j;
        pthread_create(&tids[t], NULL, so_job_run, &jobs[t]);
    }
    for (int t = 0; t < n_threads; t++) pthread_join(tids[t], NULL);
    free(tids);
    free(jobs);
    /* ThreadedScsTreeLikelihood.java:371-379 */
    double logP = 0;
    for (int t = 0; t < n_threads; t++) logP += outs[t].logp;
    if (asc_mode == 1 || asc_mode == 2) {
        double *cs = (double *)malloc(sizeof(double) * n_threads);
        long loci = 0;
        for (int i = 0; i < S; i++) loci += weights ? weights[i] : 1;
        for (int t = 0; t < n_threads; t++) cs[t] = outs[t].const_sum;
        logP += so_asc_bias_threaded(cs, n_threads, asc_mode - 1, loci, M);
        free(cs);
    }
    free(outs);
    free(points);
    return logP;
}

int so_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* ------------------------------------------------------------------------------------------ */
/* Stage-1 background likelihood: likelihood/ScsBackgroundLikelihood.java:169-268              */
/* (calculateLogPDirichletMultinomial).  hist h = 0..4 is (variant1, variant2, variant3,        */
/* normal, coverage) in the order of ScsBackgroundInfo.getOrderedNamesFullSupsCov (:184-186);   */
/* its entries are values[off[h]..off[h+1]) with multiplicities counts[...].                    */
/* ------------------------------------------------------------------------------------------ */
double so_background_log_likelihood(const long long *values, const long long *counts, const int *off,
                                    long long n_background_sites, int n_taxa, double eff_seq_err, double shape1)
{
    /* NucReadCountsModelFiniteMuDM.getWildTypeNucReadCountModelParams (:99-112) */
    double a[5];
    a[0] = a[1] = a[2] = (eff_seq_err / 3) * shape1;
    a[3] = (1 - eff_seq_err) * shape1;
    a[4] = shape1;
    double logP = (double)n_background_sites * n_taxa *
                  (so_log_gamma(a[4]) - so_log_gamma(a[0]) - so_log_gamma(a[1]) - so_log_gamma(a[2]) - so_log_gamma(a[3]));
    for (int i = off[4]; i < off[5]; i++) logP -= counts[i] * so_log_gamma(a[4] + values[i]);
    for (int h = 0; h < 4; h++)
        for (int i = off[h]; i < off[h + 1]; i++) logP += counts[i] * so_log_gamma(a[h] + values[i]);
    return logP;
}
