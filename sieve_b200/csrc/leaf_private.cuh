// leaf_private.cuh -- leaf genotype likelihoods under the per-pattern coverage model.
//
// Replaces RawReadCountsModelInterface.Base.computeLeafLikelihood (rawreadcountsmodel/RawReadCountsModelInterface.java:
// 458-506) when the sequencing-coverage model is PrivateAllelicSeqCovModel (rawreadcountsmodel/seqcovmodel/
// PrivateAllelicSeqCovModel.java): every (rate category, pattern) pair has its OWN allelic coverage t and raw variance v
// (allelicSeqCovPerPattern / allelicSeqCovRawVarPerPattern, :16-17), so nrOfMatrices = K and the leaf partials differ
// between the categories (:272-296, computeSeqCovLikelihoodPerAllele :499-511).  The nucleotide (Dirichlet-multinomial)
// terms are those of the shared model and come from the same two look-up tables (leaf.cuh); the negative-binomial
// coverage terms cannot be tabulated per cell any more -- (t, v) change with the site -- and are evaluated directly, with
// the reference's operation order (sv_seq_cov_cell_terms / sv_seq_cov_log_terms of leaf.cuh): 1 + 2 K A logGamma per
// (cell, site) for A modelled allele counts.  The kernel is bound by the fp64 pipe, not by memory.
//
// Output per (cell, site): K x 4 scaled linear genotype likelihoods [cell][site][K][4] (the layout of an internal
// partial, 128 bytes per site) under ONE natural-log scale per (cell, site); as in leaf.cuh only the per-site SUM of the
// scales over the cells is kept, and -- for the algebraic constant-site path -- per category the sum of the constant
// genotype's log-likelihoods.
#pragma once
#include "leaf.cuh"

struct SvLeafPrivArgs {
    const int4* counts;       // [cell][site]
    const double* dmA;        // [C][4]
    const double* dmB;        // [C][2]
    const double* size_factors;
    const int* cell_list;     // NULL: entry e is cell e
    const int* site_list;     // NULL: every site of the shard; else the n_list sites to evaluate (in place: the patterns whose
    int n_list;               // (t, v) moved at an accepted state, sieve_private_retraverse_changed)
    const double2* tv;        // [K][tv_stride] (t, v) of every (category, site) of the shard
    double* leafXK;           // [cell][out_stride][KMAX][4] (already offset to the write buffer) or NULL: sums only
    double* sumS;             // [cell group][sum_stride]
    double* sum0;             // [cell group][KMAX][sum_stride] or NULL
    int const_genotype;
    int n_cells, S, C, K;
    long long counts_stride, site_begin, out_stride, sum_stride, tv_stride;
    SvDmAlphas al;
    double theta;
    int single_ado;
};

#define SV_LEAFP_THREADS 128
__global__ void __launch_bounds__(SV_LEAFP_THREADS) k_leaf_private(SvLeafPrivArgs a) {
    extern __shared__ __align__(16) double sm[];
    const int C = a.C;
    const int tid = threadIdx.x;
    // the two Dirichlet-multinomial tables, staged once per block when they fit (else read through L1)
    const double* dmA = a.dmA;
    const double* dmB = a.dmB;
    if ((size_t)C * 6 * sizeof(double) <= 96 * 1024) {
        double* sA = sm;
        double* sB = sm + 4 * (size_t)C;
        for (int i = tid; i < 4 * C; i += SV_LEAFP_THREADS) sA[i] = a.dmA[i];
        for (int i = tid; i < 2 * C; i += SV_LEAFP_THREADS) sB[i] = a.dmB[i];
        __syncthreads();
        dmA = sA;
        dmB = sB;
    }
    int s = blockIdx.x * SV_LEAFP_THREADS + tid;
    if (a.site_list) {
        if (s >= a.n_list) return;
        s = a.site_list[s];
    } else if (s >= a.S)
        return;
    const int e0 = blockIdx.y * SV_LEAF_CELLS;
    const int e1 = min(a.n_cells, e0 + SV_LEAF_CELLS);
    const int K = a.K;
    double t[SV_KMAX], v[SV_KMAX];
#pragma unroll
    for (int k = 0; k < SV_KMAX; k++) {
        const double2 p = k < K ? a.tv[(size_t)k * a.tv_stride + a.site_begin + s] : make_double2(1.0, 1.0);
        t[k] = p.x;
        v[k] = p.y;
    }
    const double lt = log(a.theta), l1t = log(1 - a.theta);
    const bool want0 = a.sum0 != nullptr;
    double accS = 0.0, acc0[SV_KMAX] = {0.0, 0.0, 0.0, 0.0};

    for (int e = e0; e < e1; e++) {
        const int j = a.cell_list ? a.cell_list[e] : e;
        const double sf = __ldg(a.size_factors + j);
        const int4 c = __ldg(a.counts + (size_t)j * a.counts_stride + a.site_begin + s);
        double N0, N1, N2, N3;
        sv_leaf_nuc_terms<false>(c, dmA, dmB, C, a.al, N0, N1, N2, N3);
        // coverage terms of every category: log NB(cov; alleles, t_k, v_k) for 2 and 1 (and, locus dropout, 0) alleles
        const int cov = c.w;
        const double lg_x1 = sv_log_gamma(__dadd_rn((double)cov, 1.0));
        double ca[SV_KMAX], cb[SV_KMAX], cz[SV_KMAX];
        double M = __longlong_as_double(0xfff0000000000000LL);
#pragma unroll
        for (int k = 0; k < SV_KMAX; k++) {
            if (k < K) {
                const double c2 = sv_seq_cov_log_terms(cov, sv_seq_cov_cell_terms(2, sf, t[k], v[k]), lg_x1);
                const double c1 = sv_seq_cov_log_terms(cov, sv_seq_cov_cell_terms(1, sf, t[k], v[k]), lg_x1);
                if (a.single_ado) {
                    ca[k] = l1t + c2;
                    cb[k] = lt + c1;
                    cz[k] = __longlong_as_double(0xfff0000000000000LL);
                } else {
                    const double c0 = sv_seq_cov_log_terms(cov, sv_seq_cov_cell_terms(0, sf, t[k], v[k]), lg_x1);
                    ca[k] = 2 * l1t + c2;
                    cb[k] = lt + l1t + c1;
                    cz[k] = 2 * lt + c0;
                }
                M = fmax(M, fmax(fmax(ca[k], cb[k]), cz[k]));
            } else
                ca[k] = cb[k] = cz[k] = __longlong_as_double(0xfff0000000000000LL);
        }
        // RawReadCountsModelFiniteMuDM.java:121-260 per category, in linear space under the common scale exp(m + M)
        double m = fmax(fmax(N0, N1), fmax(N2, N3));
        if (!a.single_ado) {
#pragma unroll
            for (int k = 0; k < SV_KMAX; k++)
                if (k < K) m = fmax(m, cz[k] - M);
        }
        const double n0 = exp(N0 - m), n1 = exp(N1 - m), n2 = exp(N2 - m), n3 = exp(N3 - m);
        double x[SV_KMAX][4];
        int hi = 0;
#pragma unroll
        for (int k = 0; k < SV_KMAX; k++) {
            if (k < K) {
                const double q0 = exp(ca[k] - M), q1 = exp(cb[k] - M);
                if (a.single_ado) {
                    const double both = q0 + q1;
                    x[k][0] = n0 * both;
                    x[k][1] = n3 * q0 + 0.5 * (n0 + n1) * q1;
                    x[k][2] = n1 * both;
                    x[k][3] = n2 * q0 + n1 * q1;
                    if (want0) {
                        const int cg = a.const_genotype;
                        acc0[k] += cg == 0 ? N0 + (log(both) + M)
                                           : log(cg == 1 ? x[k][1] : (cg == 2 ? x[k][2] : x[k][3])) + (m + M);
                    }
                } else {
                    const double z = exp((cz[k] - M) - m);
                    x[k][0] = n0 * q0 + 2.0 * n0 * q1 + z;
                    x[k][1] = n3 * q0 + (n0 + n1) * q1 + z;
                    x[k][2] = n1 * q0 + 2.0 * n1 * q1 + z;
                    x[k][3] = n2 * q0 + 2.0 * n1 * q1 + z;
                    if (want0) {
                        const int cg = a.const_genotype;
                        acc0[k] += log(cg == 0 ? x[k][0] : (cg == 1 ? x[k][1] : (cg == 2 ? x[k][2] : x[k][3]))) + (m + M);
                    }
                }
                hi = max(hi, sv_maxhi4(x[k]));
            } else
                x[k][0] = x[k][1] = x[k][2] = x[k][3] = 0.0;
        }
        // the largest of the K x 4 values into [1, 2): the walk's magnitude bounds start from a normalised leaf
        int e2;
        const double f = sv_pow2_rescale_hi_int(hi, &e2);
        accS += fma((double)e2, SV_LN2, m + M);
        if (a.leafXK) {
            double* out = a.leafXK + ((size_t)j * a.out_stride + s) * (SV_KMAX * 4);
#pragma unroll
            for (int k = 0; k < SV_KMAX; k++) {
                double y[4] = {x[k][0] * f, x[k][1] * f, x[k][2] * f, x[k][3] * f};
                sv_st256(out + 4 * k, y);
            }
        }
    }
    a.sumS[(size_t)blockIdx.y * a.sum_stride + s] = accS;
    if (want0)
        for (int k = 0; k < K; k++) a.sum0[((size_t)blockIdx.y * SV_KMAX + k) * a.sum_stride + s] = acc0[k];
}

// adds the per-cell-group partial sums of the private leaf kernel in group order: lsum[s] and lambdaK[k][s]
__global__ void __launch_bounds__(256) k_colsum_private(const double* __restrict__ sumS, const double* __restrict__ sum0,
                                                        int n_groups, int S, int K, long long stride,
                                                        double* __restrict__ lsum, double* __restrict__ lambdaK,
                                                        long long lambda_stride, const int* __restrict__ site_list = nullptr,
                                                        int n_list = 0) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (site_list) {
        if (s >= n_list) return;
        s = site_list[s];
    } else if (s >= S)
        return;
    double a = 0.0;
    for (int g = 0; g < n_groups; g++) a += __ldg(sumS + (size_t)g * stride + s);
    lsum[s] = a;
    if (sum0)
        for (int k = 0; k < K; k++) {
            double b = 0.0;
            for (int g = 0; g < n_groups; g++) b += __ldg(sum0 + ((size_t)g * SV_KMAX + k) * stride + s);
            lambdaK[(size_t)k * lambda_stride + s] = b;
        }
}

// logConstRoot of every site under the private model: the constant-site partial of the root factorises per category,
// constRoot[s] = sum_k prop_k G_root[k][root genotype] exp(Lambda_k[s])  (core_planner.cuh), so
// logConstRoot[s] = logSumExp_k(logw[k] + Lambda_k[s]) with logw[k] = log(prop_k G_root[k][rg]) + log-scale of G_root.
struct SvConstPrivW {
    double logw[SV_KMAX];
};
__global__ void __launch_bounds__(256) k_const_private(const double* __restrict__ lambdaK, long long lambda_stride, int S,
                                                       int K, SvConstPrivW w, double* __restrict__ site_const) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double v[SV_KMAX];
    double mx = __longlong_as_double(0xfff0000000000000LL);
    for (int k = 0; k < K; k++) {
        v[k] = w.logw[k] + lambdaK[(size_t)k * lambda_stride + s];
        mx = fmax(mx, v[k]);
    }
    double sum = 0.0;
    for (int k = 0; k < K; k++) sum += exp(v[k] - mx);
    site_const[s] = log(sum) + mx;
}

// flag[s] = 1 where the (t, v) pair of any category of site s differs (bitwise) from the pair the current leaves were
// computed with: the patterns ScsTreeLikelihood.postProcess re-traverses (likelihood/ScsTreeLikelihood.java:2493-2498)
__global__ void __launch_bounds__(256) k_tv_changed(const double2* __restrict__ tv, const double2* __restrict__ tv_leaf, int S,
                                                    int K, unsigned char* __restrict__ flag) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    unsigned char f = 0;
    for (int k = 0; k < K; k++) {
        const double2 a = tv[(size_t)k * S + s], b = tv_leaf[(size_t)k * S + s];
        f |= (__double_as_longlong(a.x) != __double_as_longlong(b.x)) || (__double_as_longlong(a.y) != __double_as_longlong(b.y));
    }
    flag[s] = f;
}

// ---- accept-time re-estimation of (t, v) --------------------------------------------------------------------------
// PrivateAllelicSeqCovModel.updateAllelicSeqCovAndRawVar (:312-432) for every (category, site) pair whose maximum-
// likelihood combination is unique (one combination of weight 1; tied pairs are flagged and left to the host, which has
// the enumeration of sieve_b200/variants.py).  The ML number of sequenced alleles of a cell is
// getNrOfAlleles(genotype) - dropout state = 2 - (index of the arg-max dropout component)
// (ScsTreeLikelihood.java:1080, ScsFiniteMuExtendedModel.java:245-247); modeledAlleles = {1, 2} (single dropout) or
// {0, 1, 2} (RawReadCountsModelInterface.java:223-233).  Per group of cells with the same number of alleles: q = mean of
// cov / sizeFactor, w = sample variance, z = w - q mean(1 / sizeFactor); t = mean over the non-empty groups of
// q / (alleles + 1e-6), v = mean over the groups with z > 0 of z / (alleles + 1e-6)^2, v kept when there is none.
// Sums run over the cells in index order, like the reference's `indices` lists.  One thread per (site, category).
struct SvPrivUpdArgs {
    const int4* counts;             // [cell][S]
    const double* size_factors;
    const unsigned char* ado_sel;   // [cell][S][K] arg-max dropout components of the leaf over its ML genotypes
    const unsigned char* ambiguous; // [S][K]
    double2* tv;                    // [KMAX][S], updated in place
    unsigned char* tied;            // [K][S] out
    int* changed;                   // out: pairs whose (t, v) moved
    int n_cells, S, K, zero_cov_mode, single_ado;
};
__global__ void __launch_bounds__(128) k_private_update(SvPrivUpdArgs a) {
    const long long lane = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long lanes = (long long)a.S * a.K;
    if (lane >= lanes) return;
    const int s = (int)(lane / a.K), k = (int)(lane - (long long)s * a.K);
    if (a.ambiguous[lane]) {
        a.tied[(size_t)k * a.S + s] = 1;
        return;
    }
    a.tied[(size_t)k * a.S + s] = 0;
    const int add = a.zero_cov_mode == 1 ? 1 : 0;
    // pass 1: group sizes and means (index = 2 - dropout component: 2, 1, 0 sequenced alleles)
    int h[3] = {0, 0, 0};
    double sq[3] = {0.0, 0.0, 0.0};
    for (int j = 0; j < a.n_cells; j++) {
        const unsigned u = a.ado_sel[(size_t)j * lanes + lane];
        const int comp = __ffs(u) - 1;  // unique by construction (the pair is not ambiguous)
        const double x = __ddiv_rn((double)(__ldg(&a.counts[(size_t)j * a.S + s].w) + add), __ldg(a.size_factors + j));
#pragma unroll
        for (int c = 0; c < 3; c++)
            if (c == comp) {
                h[c]++;
                sq[c] = __dadd_rn(sq[c], x);
            }
    }
    double q[3] = {0.0, 0.0, 0.0}, z[3] = {0.0, 0.0, 0.0};
    int N0 = 0, N1 = 0;
#pragma unroll
    for (int c = 0; c < 3; c++)
        if (h[c] > 0) {
            N0++;
            q[c] = __ddiv_rn(sq[c], (double)h[c]);
        }
    // pass 2: sample variances and the 1 / sizeFactor sums of the groups with at least two cells
    double sw[3] = {0.0, 0.0, 0.0}, sz[3] = {0.0, 0.0, 0.0};
    for (int j = 0; j < a.n_cells; j++) {
        const unsigned u = a.ado_sel[(size_t)j * lanes + lane];
        const int comp = __ffs(u) - 1;
        const double sf = __ldg(a.size_factors + j);
        const double x = __ddiv_rn((double)(__ldg(&a.counts[(size_t)j * a.S + s].w) + add), sf);
#pragma unroll
        for (int c = 0; c < 3; c++)
            if (c == comp) {
                const double d = __dadd_rn(x, -q[c]);
                sw[c] = __dadd_rn(sw[c], __dmul_rn(d, d));
                sz[c] = __dadd_rn(sz[c], __ddiv_rn(1.0, sf));
            }
    }
#pragma unroll
    for (int c = 0; c < 3; c++)
        if (h[c] > 1) {
            N1++;
            const double w = __ddiv_rn(sw[c], (double)(h[c] - 1));
            z[c] = __dadd_rn(w, -__ddiv_rn(__dmul_rn(q[c], sz[c]), (double)h[c]));
            if (z[c] <= 0.0) {
                z[c] = 0.0;
                N1--;
            }
        }
    // sums over the modelled allele numbers in ascending order (modeledAlleles is sorted): 0 (locus dropout only), 1, 2
    double sumT = 0.0, sumV = 0.0;
    for (int al = a.single_ado ? 1 : 0; al <= 2; al++) {
        const int c = 2 - al;
        const double ka = __dadd_rn((double)al, 1e-6);
        sumT = __dadd_rn(sumT, __ddiv_rn(q[c], ka));
        sumV = __dadd_rn(sumV, __ddiv_rn(z[c], __dmul_rn(ka, ka)));
    }
    const double2 old = a.tv[(size_t)k * a.S + s];
    double2 nw = old;
    // weight 1, one combination: sumSeqCov / sumWeights[0] = sumT / N[0]; the variance likewise when N[1] > 0 and positive
    nw.x = __ddiv_rn(__ddiv_rn(sumT, (double)N0), 1.0);
    if (N1 > 0) {
        const double sv = __ddiv_rn(sumV, (double)N1);
        if (sv > 0.0) nw.y = sv;
    }
    if (nw.x != old.x || nw.y != old.y) {
        a.tv[(size_t)k * a.S + s] = nw;
        atomicAdd(a.changed, 1);
    }
}
