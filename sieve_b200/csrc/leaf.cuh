// leaf.cuh -- leaf genotype likelihoods from the read-count tensor.
//
// Replaces RawReadCountsModelInterface.Base.computeLeafLikelihood (rawreadcountsmodel/RawReadCountsModelInterface.java:458-506)
// with its callees NucReadCountsModelFiniteMuDM.computeNucReadCountsLikelihoods (nucreadcountsmodel/...DM.java:33-96),
// ExploredSharedAllelicSeqCovModel.computeSeqCovLikelihoodPerAllele (seqcovmodel/...java:266-277) and
// RawReadCountsModelFiniteMuDM.computeMixedLikelihoodCore (rawreadcountsmodel/RawReadCountsModelFiniteMuDM.java:121-260).
//
// Every logGamma argument of the reference is (alpha + c), c, (r + x) or (x + 1) with INTEGER c, x, only six distinct
// alpha per parameter set and one (r, p) per (cell, allele count).  So per parameter proposal two small tables are built
// on the device with the reference's own arithmetic (Lanczos logGamma, same operation order):
//   dmA[c] = { LP(c, f*w1), LP(c, (1-eps)*w1), LP(c, (1/2-f)*w2), LP(c, f*w2) },  dmB[c] = { LP(c, w1), LP(c, w2) }
//       with LP(c, a) = c <= 0 ? 0 : log c + logBeta(a, c)   (math/distributions/DirichletMultinomial.java:48-50)
//   cell[j][x] = the allelic-dropout mixture weights of the two (three) negative-binomial coverage terms at coverage x
// and the per-(cell, site) work becomes 4 table sums + 4 exp: HBM-bound instead of lgamma-bound.
// Counts beyond the table are evaluated directly (same formulas), so any input is handled.
//
// Output per (cell, site): 4 linear genotype likelihoods, jointly scaled: true likelihood_g = x[g] * exp(lscale).
// lscale itself is not stored per leaf, only its sum over the cells per site (see the kernel below).
#pragma once
#include "common.cuh"

// ---- commons-math 2.x Gamma.logGamma, restated (external to the reference; see oracle/sieve_oracle.c header) ----
__device__ __forceinline__ double sv_log_gamma(double x) {
    if (!(x > 0.0)) return __longlong_as_double(0x7ff8000000000000LL);
    const double L[15] = {0.99999999999999709182,     57.156235665862923517,      -59.597960355475491248,
                          14.136097974741747174,      -0.49191381609762019978,    .33994649984811888699e-4,
                          .46523628927048575665e-4,   -.98374475304879564677e-4,  .15808870322491248884e-3,
                          -.21026444172410488319e-3,  .21743961811521264320e-3,   -.16431810653676389022e-3,
                          .84418223983852743293e-4,   -.26190838401581408670e-4,  .36899182659531622704e-5};
    double sum = 0.0;
#pragma unroll
    for (int i = 14; i > 0; --i) sum = __dadd_rn(sum, __ddiv_rn(L[i], __dadd_rn(x, (double)i)));
    sum = __dadd_rn(sum, L[0]);
    const double tmp = __dadd_rn(__dadd_rn(x, 607.0 / 128.0), .5);
    // ((x + .5) * log(tmp)) - tmp + HALF_LOG_2_PI + log(sum / x), left to right, no contraction
    double r = __dmul_rn(__dadd_rn(x, .5), log(tmp));
    r = __dadd_rn(r, -tmp);
    r = __dadd_rn(r, 0.91893853320467274178);
    r = __dadd_rn(r, log(__ddiv_rn(sum, x)));
    return r;
}

// LP(c, a) of DirichletMultinomial.logPartial
__device__ __forceinline__ double sv_dm_log_partial(int c, double a) {
    if (c <= 0) return 0.0;
    const double cd = (double)c;
    if (!(a > 0.0)) return __longlong_as_double(0x7ff8000000000000LL);  // logBeta is NaN for a <= 0
    const double lb = __dadd_rn(__dadd_rn(sv_log_gamma(a), sv_log_gamma(cd)), -sv_log_gamma(__dadd_rn(a, cd)));
    return __dadd_rn(log(cd), lb);
}

struct SvDmAlphas {
    double fw1, omw1, hw2, fw2, w1, w2;
};

// NucReadCountsModelFiniteMuDM.java:40-51
__host__ __device__ inline SvDmAlphas sv_dm_alphas(double eps, double w1, double w2) {
    SvDmAlphas a;
    const double oneThird = eps / 3;
    const double oneMinus = 1 - eps;
    const double halfMinus = 0.5 - eps / 3;
    a.fw1 = oneThird * w1;
    a.omw1 = oneMinus * w1;
    a.fw2 = oneThird * w2;
    a.hw2 = halfMinus * w2;
    a.w1 = w1;
    a.w2 = w2;
    return a;
}

// exp(y) = m * 2^e with m in [1, 2): Cody-Waite reduction, so that |y| up to a few thousand keeps full precision
__device__ __forceinline__ void sv_split_exp(double y, double* m_out, int* e_out) {
    const double kf = rint(y * 1.4426950408889634074);
    double r = fma(-kf, 6.93147180369123816490e-01, y);   // ln 2, high part (fdlibm)
    r = fma(-kf, 1.90821492927058770002e-10, r);           // ln 2, low part
    double m = exp(r);                                     // [0.70, 1.42]
    int k = (int)kf;
    if (m < 1.0) {
        m *= 2.0;
        k -= 1;
    }
    *m_out = m;
    *e_out = k;
}
// The same tables in LINEAR space for the default leaf kernel (k_leaf_lin): the reciprocals exp(-LP(c, alpha)) of the four
// denominators' columns and exp(+LP(c, alpha)) of the two numerators, each as a mantissa in [1, 2) and a binary exponent,
// so that a Dirichlet-multinomial likelihood is five multiplications per genotype instead of four sums and an exp.
// The kernel is bound by L1TEX wavefronts (ncu: 97 % of the LSU data pipe, more than half of them these look-ups), so the
// rows are as narrow as they can be: the exponent travels INSIDE the double -- a mantissa in [1, 2) has the constant sign /
// exponent field 0x3FF, and those 12 bits hold the biased exponent instead (sv_lin_pack / sv_lin_unpack).  Three arrays of
// 16-byte entries, so that a random row falls into a random one of the 8 groups of 4 banks:
//   linA = [ (col 0, col 1) x C | (col 2, col 3) x C ],  linB = [ (w1, w2) x C ];  linB[0] = e^1: the cov == 0 quirk.
// 12 bits: |exponent| < 2048, i.e. |LP| <= SV_LIN_MAX_LOG (host check sv_lin_tables_ok: otherwise, and for NaN entries,
// the log-space kernel runs).
#define SV_LIN_BIAS 2048
#define SV_LIN_MAX_LOG 1400.0
#define SV_LINA_STRIDE 4      // doubles per count in linA / linB (sizes only: see the layout above)
#define SV_LINB_STRIDE 2

__device__ __forceinline__ double sv_lin_pack(double m /* in [1, 2) */, int e) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(m) & 0x000FFFFFFFFFFFFFull;
    return __longlong_as_double((long long)(b | ((unsigned long long)(unsigned)(e + SV_LIN_BIAS) << 52)));
}
// -> the mantissa as a double in [1, 2); E = the biased exponent
__device__ __forceinline__ double sv_lin_unpack(double d, unsigned& E) {
    const int hi = __double2hiint(d);
    E = (unsigned)hi >> 20;
    return __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(d));
}

__global__ void k_build_dm_tables(double* __restrict__ dmA, double* __restrict__ dmB, double* __restrict__ linA,
                                  double* __restrict__ linB, int C, SvDmAlphas al) {
    // one thread per (count, column): six short logGamma chains per count instead of one long one
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = idx / 6, col = idx % 6;
    if (c >= C) return;
    const double alpha = col == 0 ? al.fw1 : col == 1 ? al.omw1 : col == 2 ? al.hw2 : col == 3 ? al.fw2 : col == 4 ? al.w1 : al.w2;
    const double v = sv_dm_log_partial(c, alpha);
    double m;
    int e;
    if (col < 4) {
        dmA[4 * c + col] = v;
        sv_split_exp(-v, &m, &e);
        linA[(size_t)(col >> 1) * 2 * C + 2 * c + (col & 1)] = sv_lin_pack(m, e);
    } else {
        dmB[2 * c + (col - 4)] = v;
        // DirichletMultinomial.logDensity returns 1.0 for an empty site (:29-30): row 0 is only ever read when cov == 0
        sv_split_exp(c == 0 ? 1.0 : v, &m, &e);
        linB[2 * c + (col - 4)] = sv_lin_pack(m, e);
    }
}

struct SvCovParams {
    double t, v;        // allelicSeqCov, allelicSeqCovRawVar
    double theta;       // adoRate
    int single_ado;
};

// log NB(x; r, p) for `alleles` sequenced alleles of a cell with size factor s:
// ExploredSharedAllelicSeqCovModel.java:266-277 + math/distributions/NegativeBinomial.java:76-79, same order.
__device__ __forceinline__ double sv_seq_cov_log(int x, int alleles, double s, double t, double v) {
    const double ka = __dadd_rn((double)alleles, 1e-6);
    const double mean = __dmul_rn(__dmul_rn(ka, t), s);
    const double var = __dadd_rn(mean, __dmul_rn(__dmul_rn(__dmul_rn(s, s), __dmul_rn(ka, ka)), v));
    const double p = __ddiv_rn(mean, var);
    const double r = __ddiv_rn(__dmul_rn(mean, mean), __dadd_rn(var, -mean));
    const double beta = __dadd_rn(__ddiv_rn(1.0, p), -1.0);
    const double rx = __dadd_rn(r, (double)x);
    double l = __dadd_rn(sv_log_gamma(rx), __dmul_rn(rx, log(__dadd_rn(1.0, -p))));
    l = __dadd_rn(l, -sv_log_gamma(__dadd_rn((double)x, 1.0)));
    l = __dadd_rn(l, -sv_log_gamma(r));
    l = __dadd_rn(l, -__dmul_rn(r, log(beta)));
    return l;
}

// The coverage part of the dropout mixture at coverage x, scaled by its own maximum:
//   single ADO: e[0] = (1-theta) * NB_2 / M, e[1] = theta * NB_1 / M, e[2] = log M, e[3] = log(e[0] + e[1]) + log M
//   locus  ADO: e[0] = (1-theta)^2 * NB_2 / M, e[1] = theta (1-theta) * NB_1 / M, e[2] = log M, e[3] = log(theta^2 NB_0 / M)
__device__ __forceinline__ void sv_cov_entry(int x, double s, const SvCovParams& cp, double (&e)[4]) {
    const double c2 = sv_seq_cov_log(x, 2, s, cp.t, cp.v);
    const double c1 = sv_seq_cov_log(x, 1, s, cp.t, cp.v);
    const double lt = log(cp.theta), l1t = log(1 - cp.theta);
    if (cp.single_ado) {
        const double a = l1t + c2, b = lt + c1;
        const double m = fmax(a, b);
        e[0] = exp(a - m);
        e[1] = exp(b - m);
        e[2] = m;
        e[3] = log(e[0] + e[1]) + m;  // log of the coverage mixture: the constant genotype's leaf log-likelihood is N0 + this
    } else {
        const double c0 = sv_seq_cov_log(x, 0, s, cp.t, cp.v);
        const double a = 2 * l1t + c2, b = lt + l1t + c1, z = 2 * lt + c0;
        const double m = fmax(fmax(a, b), z);
        e[0] = exp(a - m);
        e[1] = exp(b - m);
        e[2] = m;
        e[3] = z - m;
    }
}

// The x-independent terms of sv_seq_cov_log for one (cell, allele count), computed once per block: the table kernel
// then evaluates one logGamma per allele count and entry (plus logGamma(x + 1), shared by the allele counts) instead of
// three.  Same functions of the same arguments in the same order as sv_seq_cov_log: the entries keep their bits.
struct SvCovCellTerms {
    double r, log1mp, lg_r, r_log_beta;
};
__device__ __forceinline__ SvCovCellTerms sv_seq_cov_cell_terms(int alleles, double s, double t, double v) {
    const double ka = __dadd_rn((double)alleles, 1e-6);
    const double mean = __dmul_rn(__dmul_rn(ka, t), s);
    const double var = __dadd_rn(mean, __dmul_rn(__dmul_rn(__dmul_rn(s, s), __dmul_rn(ka, ka)), v));
    const double p = __ddiv_rn(mean, var);
    SvCovCellTerms o;
    o.r = __ddiv_rn(__dmul_rn(mean, mean), __dadd_rn(var, -mean));
    const double beta = __dadd_rn(__ddiv_rn(1.0, p), -1.0);
    o.log1mp = log(__dadd_rn(1.0, -p));
    o.lg_r = sv_log_gamma(o.r);
    o.r_log_beta = __dmul_rn(o.r, log(beta));
    return o;
}
__device__ __forceinline__ double sv_seq_cov_log_terms(int x, const SvCovCellTerms& c, double lg_x1) {
    const double rx = __dadd_rn(c.r, (double)x);
    double l = __dadd_rn(sv_log_gamma(rx), __dmul_rn(rx, c.log1mp));
    l = __dadd_rn(l, -lg_x1);
    l = __dadd_rn(l, -c.lg_r);
    l = __dadd_rn(l, -c.r_log_beta);
    return l;
}

// logGamma(x + 1) for every table row: depends on nothing but x, so it is tabulated once per count tensor (sv_alloc_tables)
// instead of once per (cell, row) of every coverage-parameter proposal -- a third of k_build_cell_tables' logGamma chains
__global__ void k_build_lgx1(double* __restrict__ lgx1, int C) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x < C) lgx1[x] = sv_log_gamma(__dadd_rn((double)x, 1.0));
}

// cellT2 (or NULL): the compact twin of the rows for k_leaf_lin, single dropout only: (+-rho, m) in 16 bytes.  Of the two
// mixture weights exp(a - m), exp(b - m) with m = max(a, b) the larger one is exp(0) = 1 exactly, so the row only needs the
// smaller one, rho, and which of the two it is (the sign bit: negative <=> the FIRST weight is rho) -- the same two
// doubles come back bit for bit, from half the bytes of L1 / L2 traffic per look-up.
__global__ void k_build_cell_tables(double* __restrict__ cellT, const double* __restrict__ size_factors, int n_cells,
                                    int C, SvCovParams cp, const double* __restrict__ lgx1, double2* __restrict__ cellT2) {
    __shared__ SvCovCellTerms terms[3];  // 2, 1 and (locus dropout) 0 sequenced alleles
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (j >= n_cells) return;
    if (threadIdx.x < 3) terms[threadIdx.x] = sv_seq_cov_cell_terms(2 - (int)threadIdx.x, size_factors[j], cp.t, cp.v);
    __syncthreads();
    if (x >= C) return;
    const double lg_x1 = __ldg(lgx1 + x);  // the same function of the same argument as before: same bits
    const double c2 = sv_seq_cov_log_terms(x, terms[0], lg_x1);
    const double c1 = sv_seq_cov_log_terms(x, terms[1], lg_x1);
    const double lt = log(cp.theta), l1t = log(1 - cp.theta);
    double e[4];
    if (cp.single_ado) {
        const double a = l1t + c2, b = lt + c1;
        const double m = fmax(a, b);
        e[0] = exp(a - m);
        e[1] = exp(b - m);
        e[2] = m;
        e[3] = log(e[0] + e[1]) + m;
        if (cellT2) {
            // (NaN parameters give NaN in both forms; a == b gives (1, 1) in both)
            const bool first_is_one = a >= b;
            cellT2[(size_t)j * C + x] = make_double2(first_is_one ? e[1] : -e[0], m);
        }
    } else {
        const double c0 = sv_seq_cov_log_terms(x, terms[2], lg_x1);
        const double a = 2 * l1t + c2, b = lt + l1t + c1, z = 2 * lt + c0;
        const double m = fmax(fmax(a, b), z);
        e[0] = exp(a - m);
        e[1] = exp(b - m);
        e[2] = m;
        e[3] = z - m;
    }
    sv_st256(cellT + ((size_t)j * C + x) * 4, e);
}

// Slow paths for counts beyond the tables: out of line and returning by value, so that the hot loop keeps its
// operands in registers and stays small.
struct SvD4 {
    double a, b, c, d;
};
struct SvD2 {
    double a, b;
};
__device__ __noinline__ SvD4 sv_dm_partials_direct(int c, SvDmAlphas al) {
    SvD4 o;
    o.a = sv_dm_log_partial(c, al.fw1);
    o.b = sv_dm_log_partial(c, al.omw1);
    o.c = sv_dm_log_partial(c, al.hw2);
    o.d = sv_dm_log_partial(c, al.fw2);
    return o;
}
__device__ __noinline__ SvD2 sv_dm_totals_direct(int c, SvDmAlphas al) {
    SvD2 o;
    o.a = sv_dm_log_partial(c, al.w1);
    o.b = sv_dm_log_partial(c, al.w2);
    return o;
}
__device__ __noinline__ SvD4 sv_cov_entry_direct(int x, double s, SvCovParams cp) {
    double e[4];
    sv_cov_entry(x, s, cp, e);
    SvD4 o;
    o.a = e[0]; o.b = e[1]; o.c = e[2]; o.d = e[3];
    return o;
}

// ---- the leaf kernel -------------------------------------------------------------------------------------------
// One block = SV_LEAF_SITES sites x one GROUP of SV_LEAF_CELLS cells, cells taken one after the other.  Per (cell, site)
// the kernel writes only the four scaled genotype likelihoods (32 B).  The natural-log scale of a leaf is NOT stored
// per leaf: scales are additive along the pruning, so whatever the tree looks like, the scales of all leaves arrive at the
// root as their plain sum over the cells.  The block therefore accumulates, per site and in registers, the sum of the
// scales of its cells (and of the constant genotype's log-likelihoods, the column sum Lambda of the algebraic
// constant-site path) and writes one partial sum per (cell group, site); k_colsum adds the groups in index order.  The
// walk adds the sum once, at the root (prune.cuh).  Groups are a fixed number of consecutive cells, independent of the
// site range, so resident, sparse and streamed evaluations produce the same bits.
struct SvLeafArgs {
    const int4* counts;     // [cell][site] (alt1, alt2, alt3, cov)
    const double* dmA;      // [C][4]
    const double* dmB;      // [C][2]
    const double* cellT;    // [cell][C][4]
    const double2* cellT2;  // [cell][C] compact rows (+-rho, m) for k_leaf_lin, or NULL
    const double *linA, *linB;  // linear-space tables of k_leaf_lin (k_build_dm_tables)
    const double* size_factors;
    const int* cell_list;   // NULL: entry e is cell e; else the cells to evaluate (read-back of a subtree's sums)
    double* leafX;          // [cell][site][4] (already offset to the write buffer), or NULL: sums only
    double* sumS;           // [cell group][sum_stride] sum over the group's cells of the leaves' log scales
    double* sum0;           // [cell group][sum_stride] ... of the constant genotype's log-likelihoods, or NULL
    int const_genotype;
    int n_cells, S, C;     // n_cells = entries evaluated; S = sites evaluated by this launch
    long long counts_stride;  // sites per cell in `counts` (the whole shard)
    long long site_begin;     // first site of this launch inside the shard (streamed evaluation walks site chunks)
    long long out_stride;     // sites per cell in leafX (== S when resident, chunk capacity when streamed)
    long long sum_stride;     // sites per cell group in sumS / sum0
    SvDmAlphas al;
    SvCovParams cp;
};

#define SV_LEAF_CELLS 32  // cells per group (fixed: the grouping of the sums must not depend on the launch)

__device__ __forceinline__ void sv_leaf_cp16(double* dst_smem, const double* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void sv_leaf_cp_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// one (cell, site): x[4] scaled linear genotype likelihoods, ls their natural-log scale, l0 the constant genotype's
// log-likelihood (only when want0)
// COVERED: every count is known to index the tables (checked once at ingest, k_counts_outside_tables): no bounds
// checks, no out-of-line calls -- the site evaluation is straight-line code that the compiler can interleave across sites.
// one 32-byte table row with one 256-bit load (global memory only; not volatile: the compiler may schedule it freely)
__device__ __forceinline__ void sv_leaf_row256(const double* p, double (&v)[4]) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}

// CT_GLOBAL: the cell table is in global memory (read through L1 with one 256-bit load per row: one sector per lane).
// The four Dirichlet-multinomial log-densities N0..N3 of one (cell, site) (genotypes 0/0, 1/1, 1/1', 0/1;
// NucReadCountsModelFiniteMuDM.java:33-96) from the two LP tables; counts beyond the tables take the direct path.
template <bool COVERED>
__device__ __forceinline__ void sv_leaf_nuc_terms(const int4 c, const double* __restrict__ dmA, const double* __restrict__ dmB,
                                                  const int C, const SvDmAlphas& al, double& N0, double& N1, double& N2,
                                                  double& N3) {
    const int cov = c.w;
    const int ref = cov - (c.x + c.y + c.z);
    if (cov == 0) {
        // DirichletMultinomial.logDensity returns 1.0 for an empty site, log mode included (:29-30)
        N0 = N1 = N2 = N3 = 1.0;
    } else {
        // LP terms: table when the count is inside it, else the same formula evaluated directly
        auto lpA = [&](int cval, double (&o)[4]) {
            const int ci = cval < 0 ? 0 : cval;
            if (ci == 0) {
                // LP(0, .) == 0 (DirichletMultinomial.java:49): most alt counts are 0, and skipping the look-up keeps
                // the shared-memory pipe for the entries that carry information
                o[0] = o[1] = o[2] = o[3] = 0.0;
            } else if (COVERED || ci < C) {
                SV_BOUND(ci < C);
                const double2 u = *reinterpret_cast<const double2*>(dmA + 4 * ci);
                const double2 w = *reinterpret_cast<const double2*>(dmA + 4 * ci + 2);
                o[0] = u.x; o[1] = u.y; o[2] = w.x; o[3] = w.y;
            } else {
                const SvD4 r = sv_dm_partials_direct(ci, al);
                o[0] = r.a; o[1] = r.b; o[2] = r.c; o[3] = r.d;
            }
        };
        // logDensity = LP(cov, a0) - (((LP(c1,a1) + LP(c2,a2)) + LP(c3,a3)) + LP(ref,a4)), summed left to right;
        // table columns: [0] f*w1, [1] (1-eps)*w1, [2] (1/2-f)*w2, [3] f*w2
        double T[4];
        lpA(c.x, T);
        double s0 = T[0], s1 = T[1], s2 = T[2], s3 = T[2];  // alt1: 0/0 f | 1/1 (1-eps) | 1/1' h | 0/1 h
        lpA(c.y, T);
        s0 += T[0]; s1 += T[0]; s2 += T[2]; s3 += T[3];      // alt2: f | f | h | f
        lpA(c.z, T);
        s0 += T[0]; s1 += T[0]; s2 += T[3]; s3 += T[3];      // alt3: f | f | f | f
        lpA(ref, T);
        s0 += T[1]; s1 += T[0]; s2 += T[3]; s3 += T[2];      // ref : (1-eps) | f | f | h
        double B0, B1;
        if (COVERED || (cov > 0 && cov < C)) {
            SV_BOUND(cov > 0 && cov < C);
            const double2 u = *reinterpret_cast<const double2*>(dmB + 2 * cov);
            B0 = u.x; B1 = u.y;
        } else {
            const SvD2 r = sv_dm_totals_direct(cov, al);
            B0 = r.a; B1 = r.b;
        }
        N0 = B0 - s0;
        N1 = B0 - s1;
        N2 = B1 - s2;
        N3 = B1 - s3;
    }
}

template <bool COVERED, bool CT_GLOBAL>
__device__ __forceinline__ void sv_leaf_site(const int4 c, const double* __restrict__ dmA, const double* __restrict__ dmB,
                                             const double* __restrict__ cT, const int C, const double sf,
                                             const SvLeafArgs& a, const bool want0, double (&x)[4], double& ls,
                                             double& l0) {
    const int cov = c.w;
    double N0, N1, N2, N3;
    sv_leaf_nuc_terms<COVERED>(c, dmA, dmB, C, a.al, N0, N1, N2, N3);
    double e[4];
    if (COVERED || (cov >= 0 && cov < C)) {
        SV_BOUND(cov >= 0 && cov < C);
        if (CT_GLOBAL) {
            sv_leaf_row256(cT + 4 * cov, e);
        } else {
            const double2 u = *reinterpret_cast<const double2*>(cT + 4 * cov);
            const double2 w = *reinterpret_cast<const double2*>(cT + 4 * cov + 2);
            e[0] = u.x; e[1] = u.y; e[2] = w.x; e[3] = w.y;
        }
    } else {
        const SvD4 r = sv_cov_entry_direct(cov, sf, a.cp);
        e[0] = r.a; e[1] = r.b; e[2] = r.c; e[3] = r.d;
    }
    if (a.cp.single_ado) {
        // RawReadCountsModelFiniteMuDM.java:133-225, single-ADO arms, in linear space with a common scale
        const double m = fmax(fmax(N0, N1), fmax(N2, N3));
        const double n0 = exp(N0 - m), n1 = exp(N1 - m), n2 = exp(N2 - m), n3 = exp(N3 - m);
        const double both = e[0] + e[1];
        x[0] = n0 * both;                              // 0/0
        x[1] = n3 * e[0] + 0.5 * (n0 + n1) * e[1];     // 0/1
        x[2] = n1 * both;                              // 1/1
        x[3] = n2 * e[0] + n1 * e[1];                  // 1/1'
        ls = m + e[2];
    } else {
        // locus-dropout arms (:147-160 etc.); the theta^2 * NB_0 term carries no nucleotide factor
        const double m = fmax(fmax(fmax(N0, N1), fmax(N2, N3)), e[3]);
        const double n0 = exp(N0 - m), n1 = exp(N1 - m), n2 = exp(N2 - m), n3 = exp(N3 - m);
        const double z = exp(e[3] - m);
        x[0] = n0 * e[0] + 2.0 * n0 * e[1] + z;
        x[1] = n3 * e[0] + (n0 + n1) * e[1] + z;
        x[2] = n1 * e[0] + 2.0 * n1 * e[1] + z;
        x[3] = n2 * e[0] + 2.0 * n1 * e[1] + z;
        ls = m + e[2];
    }
    {
        // the largest of the four values into [1, 2) (exact power of two): the walk's magnitude bounds (core_planner.cuh,
        // sv_bound_op) start from a normalised leaf
        int e2;
        const double f = sv_pow2_rescale_hi_int(sv_maxhi4(x), &e2);
        x[0] *= f; x[1] *= f; x[2] *= f; x[3] *= f;
        ls = fma((double)e2, SV_LN2, ls);
    }
    l0 = 0.0;
    if (want0) {
        const int cg = a.const_genotype;
        if (a.cp.single_ado && cg == 0)
            l0 = N0 + e[3];  // log(e^{N0} (e0 + e1) M): no transcendental on the common path
        else
            l0 = log(cg == 0 ? x[0] : (cg == 1 ? x[1] : (cg == 2 ? x[2] : x[3]))) + ls;
    }
}

// TAB: where the look-up tables are read from.
//   1: the two Dirichlet-multinomial tables staged once per block (6 C doubles of shared memory), the per-cell coverage
//      table through L1: no barrier inside the cell loop, the warps of a block drift apart               [default]
//   2: the cell table too in shared memory, double-buffered with cp.async (the next cell's table arrives while this
//      cell's sites are evaluated; one barrier per cell): 14 C doubles.  Measured slower (profiles/README.md)
//   0: everything through L1/L2 (tables too large for shared memory)
// 256 threads x 2 blocks per SM: 112-128 registers hold the four unrolled site evaluations and the eight accumulators.
// Measured alternatives, all within +2..+16 % (profiles/README.md): 3 blocks per SM with spilled accumulators, with the
// accumulators in shared memory, with 2 sites per thread (80 registers, no spill), and 6 sites per thread at 2 blocks:
// the kernel is bound by the L1TEX data pipe (the random table look-ups: 2.5-way bank conflicts in shared memory, 32
// sectors per look-up through L1), not by occupancy.
#define SV_LEAF_THREADS 256
#define SV_LEAF_IT 4  // sites per thread and cell: four count tuples in flight per lane
#define SV_LEAF_SITES (SV_LEAF_THREADS * SV_LEAF_IT)
template <int TAB, bool COVERED = false>
__global__ void __launch_bounds__(SV_LEAF_THREADS, 2) k_leaf(SvLeafArgs a) {
    extern __shared__ __align__(16) double sm[];
    const int C = a.C;
    const int e0 = blockIdx.y * SV_LEAF_CELLS;
    const int e1 = min(a.n_cells, e0 + SV_LEAF_CELLS);
    const double* dmA = a.dmA;
    const double* dmB = a.dmB;
    double* sT = sm + 6 * (size_t)C;
    const int tid = threadIdx.x;
    if (TAB >= 1) {
        double* sA = sm;
        double* sB = sm + 4 * (size_t)C;
        if (TAB == 2) {
            const int j0 = a.cell_list ? a.cell_list[e0] : e0;
            const double* src = a.cellT + (size_t)j0 * C * 4;
            for (int i = tid; i < 2 * C; i += SV_LEAF_THREADS) sv_leaf_cp16(sT + 2 * i, src + 2 * i);
        }
        for (int i = tid; i < 4 * C; i += SV_LEAF_THREADS) sA[i] = dmA[i];
        for (int i = tid; i < 2 * C; i += SV_LEAF_THREADS) sB[i] = dmB[i];
        if (TAB == 2) sv_leaf_cp_wait_all();
        __syncthreads();
        dmA = sA;
        dmB = sB;
    }
    const int sbase = blockIdx.x * SV_LEAF_SITES + tid;
    const bool want0 = a.sum0 != nullptr;
    double accS[SV_LEAF_IT], acc0[SV_LEAF_IT];
#pragma unroll
    for (int it = 0; it < SV_LEAF_IT; it++) accS[it] = acc0[it] = 0.0;

    for (int e = e0; e < e1; e++) {
        const int j = a.cell_list ? a.cell_list[e] : e;
        const double* cT = a.cellT + (size_t)j * C * 4;
        if (TAB == 2) {
            cT = sT + (size_t)((e - e0) & 1) * 4 * C;
            if (e + 1 < e1) {
                const int jn = a.cell_list ? a.cell_list[e + 1] : e + 1;
                const double* src = a.cellT + (size_t)jn * C * 4;
                double* dst = sT + (size_t)((e + 1 - e0) & 1) * 4 * C;
                for (int i = tid; i < 2 * C; i += SV_LEAF_THREADS) sv_leaf_cp16(dst + 2 * i, src + 2 * i);
            }
        }
        const double sf = __ldg(a.size_factors + j);
        double* outX = a.leafX ? a.leafX + (size_t)j * a.out_stride * 4 : nullptr;
        // the four count tuples of this cell: four independent 16-byte loads in flight per lane
        const int4* cnt = a.counts + (size_t)j * a.counts_stride + a.site_begin;
        int4 c[SV_LEAF_IT];
#pragma unroll
        for (int it = 0; it < SV_LEAF_IT; it++) {
            const int s = sbase + it * SV_LEAF_THREADS;
            c[it] = s < a.S ? __ldg(cnt + s) : make_int4(0, 0, 0, 0);
        }
#pragma unroll
        for (int it = 0; it < SV_LEAF_IT; it++) {
            const int s = sbase + it * SV_LEAF_THREADS;
            if (s >= a.S) continue;
            double x[4], ls, l0;
            sv_leaf_site<COVERED, TAB != 2>(c[it], dmA, dmB, cT, C, sf, a, want0, x, ls, l0);
            if (outX) sv_st256(outX + (size_t)s * 4, x);
            accS[it] += ls;
            acc0[it] += l0;
        }
        if (TAB == 2) {
            sv_leaf_cp_wait_all();
            __syncthreads();  // the next cell's table is complete; every warp is done with the buffer it will replace
        }
    }
#pragma unroll
    for (int it = 0; it < SV_LEAF_IT; it++) {
        const int s = sbase + it * SV_LEAF_THREADS;
        if (s >= a.S) continue;
        a.sumS[(size_t)blockIdx.y * a.sum_stride + s] = accS[it];
        if (want0) a.sum0[(size_t)blockIdx.y * a.sum_stride + s] = acc0[it];
    }
}


// ---- the default leaf kernel: linear-space Dirichlet-multinomial terms --------------------------------------------------
// Same model, same outputs and the same block / cell-group structure as k_leaf<1, true>; what differs is how the four
// nucleotide likelihoods are formed: as products of tabulated (mantissa, exponent) pairs (k_build_dm_tables) instead of
// exp(sum of tabulated logs - max).  No transcendental is evaluated per (cell, site) -- ncu had the FP64 pipe and the
// issue slots of k_leaf<1, true> as its limiters, four exp being two thirds of its FP64 instructions -- and the scale of
// a leaf is an exact power of two times the coverage table's maximum, summed as an integer.  The constant genotype's
// log-likelihood (the column sum of the algebraic constant-site path) is accumulated as a running product over the
// group's cells and logged once per (group, site).
// Preconditions (host, sv_leaf_lin_ok): every count inside the tables, single dropout, constant genotype 0, tables within
// the exponent range and small enough for shared memory; otherwise k_leaf<...> runs.
// 512 threads x 2 sites per thread, two blocks per SM (60 registers): 32 resident warps.  Measured against 256 x 4 at two
// and three blocks per SM, 512 x 1, 1024 x 1 / 2, and again after the row compactions 256 x 2 x 4, 384 x 2 x 3, 512 x 3
// (profiles/README.md): the kernel is bound by the L1TEX wavefronts of its look-ups (ncu: 97 % of the LSU data pipe before
// the rows were compacted), so what counts is bytes and look-ups per row -- the 16-byte coverage row (cellT2) and the
// exponents packed into the mantissas (sv_lin_unpack).
#ifndef SV_LEAFL_THREADS
#define SV_LEAFL_THREADS 512
#endif
#ifndef SV_LEAFL_IT
#define SV_LEAFL_IT 2
#endif
#ifndef SV_LEAFL_MINB
#define SV_LEAFL_MINB 2
#endif
#define SV_LEAFL_SITES (SV_LEAFL_THREADS * SV_LEAFL_IT)
__global__ void __launch_bounds__(SV_LEAFL_THREADS, SV_LEAFL_MINB) k_leaf_lin(SvLeafArgs a) {
    extern __shared__ __align__(16) double sm[];
    const int C = a.C;
    const int e0c = blockIdx.y * SV_LEAF_CELLS;
    const int e1c = min(a.n_cells, e0c + SV_LEAF_CELLS);
    const int tid = threadIdx.x;
    double2* sA01 = reinterpret_cast<double2*>(sm);  // (col 0, col 1) of count c at [c], (col 2, col 3) at [C + c], numerators at [2 C + c]
    {
        const double2* gA = reinterpret_cast<const double2*>(a.linA);
        const double2* gB = reinterpret_cast<const double2*>(a.linB);
        for (int i = tid; i < 2 * C; i += SV_LEAFL_THREADS) sA01[i] = gA[i];
        for (int i = tid; i < C; i += SV_LEAFL_THREADS) sA01[2 * C + i] = gB[i];
    }
    const double2* sA23 = sA01 + C;
    const double2* sB = sA01 + 2 * C;
    __syncthreads();
    const int sbase = blockIdx.x * SV_LEAFL_SITES + tid;
    const bool want0 = a.sum0 != nullptr;
    constexpr unsigned BIAS2 = ((unsigned)SV_LIN_BIAS << 16) | (unsigned)SV_LIN_BIAS;
    double accS[SV_LEAFL_IT], acc0m[SV_LEAFL_IT];
    int accE[SV_LEAFL_IT], acc0e[SV_LEAFL_IT];
#pragma unroll
    for (int it = 0; it < SV_LEAFL_IT; it++) {
        accS[it] = 0.0;
        acc0m[it] = 1.0;
        accE[it] = acc0e[it] = 0;
    }
    // the count tuples of the NEXT cell are requested while this cell's sites are evaluated: with no transcendental left,
    // the DRAM latency of that load is what a warp would otherwise wait for once per cell
    int4 cn[SV_LEAFL_IT];
    {
        const int j0 = a.cell_list ? a.cell_list[e0c] : e0c;
        const int4* cnt = a.counts + (size_t)j0 * a.counts_stride + a.site_begin;
#pragma unroll
        for (int it = 0; it < SV_LEAFL_IT; it++) {
            const int s = sbase + it * SV_LEAFL_THREADS;
            cn[it] = s < a.S ? __ldg(cnt + s) : make_int4(0, 0, 0, 0);
        }
    }
    for (int e = e0c; e < e1c; e++) {
        const int j = a.cell_list ? a.cell_list[e] : e;
#ifdef SV_LEAF_CT32
        const double* cT = a.cellT + (size_t)j * C * 4;
#else
        const double2* cT2 = a.cellT2 + (size_t)j * C;
#endif
        double* outX = a.leafX ? a.leafX + (size_t)j * a.out_stride * 4 : nullptr;
        int4 c[SV_LEAFL_IT];
#pragma unroll
        for (int it = 0; it < SV_LEAFL_IT; it++) c[it] = cn[it];
        if (e + 1 < e1c) {
            const int jn = a.cell_list ? a.cell_list[e + 1] : e + 1;
            const int4* cnt = a.counts + (size_t)jn * a.counts_stride + a.site_begin;
#pragma unroll
            for (int it = 0; it < SV_LEAFL_IT; it++) {
                const int s = sbase + it * SV_LEAFL_THREADS;
                if (s < a.S) cn[it] = __ldg(cnt + s);
            }
        }
#pragma unroll
        for (int it = 0; it < SV_LEAFL_IT; it++) {
            const int s = sbase + it * SV_LEAFL_THREADS;
            if (s >= a.S) continue;
            const int cov = c[it].w;
            const int ref = max(cov - (c[it].x + c[it].y + c[it].z), 0);
            SV_BOUND(cov >= 0 && cov < C && c[it].x >= 0 && c[it].x < C && c[it].y >= 0 && c[it].y < C && c[it].z >= 0 &&
                     c[it].z < C && ref < C);
            // the coverage mixture of this (cell, coverage): one 256-bit row through L1 (k_build_cell_tables)
            double ce[4];
#ifdef SV_LEAF_CT32
            sv_leaf_row256(cT + 4 * cov, ce);
#else
            {
                // the compact row: 16 bytes through L1 (k_build_cell_tables); the larger weight is exactly 1
                const double2 t = __ldg(cT2 + cov);
                const bool first_is_rho = __double2hiint(t.x) < 0;
                const double rho = fabs(t.x);
                ce[0] = first_is_rho ? rho : 1.0;
                ce[1] = first_is_rho ? 1.0 : rho;
                ce[2] = t.y;
            }
#endif
            // numerators: genotypes 0/0 and 1/1 take alpha0 = w1, 1/1' and 0/1 take w2
            unsigned E0, E1, E2, E3;  // scratch: biased exponents of the row just read
            const double2 bp = sB[cov];
            const double bm0 = sv_lin_unpack(bp.x, E0), bm1 = sv_lin_unpack(bp.y, E1);
            double v0 = bm0, v1 = bm0, v2 = bm1, v3 = bm1;
            unsigned elo = E0 * 0x10001u, ehi = E1 * 0x10001u;  // packed exponents of (g0, g1), (g2, g3)
            // denominators (reciprocals); table columns: [0] f*w1, [1] (1-eps)*w1, [2] (1/2-f)*w2, [3] f*w2
            //   alt1: 0/0 f | 1/1 (1-eps) | 1/1' h | 0/1 h     alt2: f | f | h | f     alt3: f | f | f | f     ref: (1-eps) | f | f | h
            // (a zero count contributes the factor 1: only the exponent bias is added)
            if (c[it].x > 0) {
                const double2 u = sA01[c[it].x], w = sA23[c[it].x];
                const double m0 = sv_lin_unpack(u.x, E0), m1 = sv_lin_unpack(u.y, E1), m2 = sv_lin_unpack(w.x, E2);
                v0 *= m0; v1 *= m1; v2 *= m2; v3 *= m2;
                elo += E0 | (E1 << 16);
                ehi += E2 * 0x10001u;
            } else {
                elo += BIAS2;
                ehi += BIAS2;
            }
            if (c[it].y > 0) {
                const double2 u = sA01[c[it].y], w = sA23[c[it].y];
                const double m0 = sv_lin_unpack(u.x, E0), m2 = sv_lin_unpack(w.x, E2), m3 = sv_lin_unpack(w.y, E3);
                v0 *= m0; v1 *= m0; v2 *= m2; v3 *= m3;
                elo += E0 * 0x10001u;
                ehi += E2 | (E3 << 16);
            } else {
                elo += BIAS2;
                ehi += BIAS2;
            }
            if (c[it].z > 0) {
                const double2 u = sA01[c[it].z], w = sA23[c[it].z];
                const double m0 = sv_lin_unpack(u.x, E0), m3 = sv_lin_unpack(w.y, E3);
                v0 *= m0; v1 *= m0; v2 *= m3; v3 *= m3;
                elo += E0 * 0x10001u;
                ehi += E3 * 0x10001u;
            } else {
                elo += BIAS2;
                ehi += BIAS2;
            }
            if (ref > 0) {
                const double2 u = sA01[ref], w = sA23[ref];
                const double m0 = sv_lin_unpack(u.x, E0), m1 = sv_lin_unpack(u.y, E1);
                const double m2 = sv_lin_unpack(w.x, E2), m3 = sv_lin_unpack(w.y, E3);
                v0 *= m1; v1 *= m0; v2 *= m3; v3 *= m2;
                elo += E1 | (E0 << 16);
                ehi += E3 | (E2 << 16);
            } else {
                elo += BIAS2;
                ehi += BIAS2;
            }
            // common binary scale of the four likelihoods: the largest exponent; a genotype more than 2^-1000 below it is
            // flushed to that floor (the documented limit of the scaled representation)
            const int S0 = (int)(elo & 0xffffu), S1 = (int)(elo >> 16), S2 = (int)(ehi & 0xffffu), S3 = (int)(ehi >> 16);
            const int Emax = max(max(S0, S1), max(S2, S3));
            const double n0 = v0 * __hiloint2double((1023 + max(S0 - Emax, -1000)) << 20, 0);
            const double n1 = v1 * __hiloint2double((1023 + max(S1 - Emax, -1000)) << 20, 0);
            const double n2 = v2 * __hiloint2double((1023 + max(S2 - Emax, -1000)) << 20, 0);
            const double n3 = v3 * __hiloint2double((1023 + max(S3 - Emax, -1000)) << 20, 0);
            // RawReadCountsModelFiniteMuDM.java:133-225, single-dropout arms, as in sv_leaf_site
            const double both = ce[0] + ce[1];
            double x[4];
            x[0] = n0 * both;
            x[1] = n3 * ce[0] + 0.5 * (n0 + n1) * ce[1];
            x[2] = n1 * both;
            x[3] = n2 * ce[0] + n1 * ce[1];
            int e2;
            const double f = sv_pow2_rescale_hi_int(sv_maxhi4(x), &e2);
            x[0] *= f; x[1] *= f; x[2] *= f; x[3] *= f;
            if (outX) sv_st256(outX + (size_t)s * 4, x);
            // log scale of this leaf = ce[2] + (Emax - 5 bias + e2) ln 2: the table part as a double, the rest as an integer
            accS[it] += ce[2];
            accE[it] += Emax - 5 * SV_LIN_BIAS + e2;
            if (want0) {
                // constant genotype 0/0: log-likelihood = log(v0 * both) + (E0 - 5 bias) ln 2 + ce[2], kept as a running
                // product (mantissa renormalised every cell) so that the log is taken once per (cell group, site)
                double p = acc0m[it] * (v0 * both);
                const int hi = __double2hiint(p);
                const int ex = ((hi >> 20) & 0x7ff) - 1023;
                if ((unsigned)ex < 256u) {  // finite, >= 1: always, unless a table entry is NaN (which then propagates)
                    p = __hiloint2double(hi - (ex << 20), __double2loint(p));
                    acc0e[it] += ex;
                }
                acc0m[it] = p;
                acc0e[it] += S0 - 5 * SV_LIN_BIAS;
            }
        }
    }
#pragma unroll
    for (int it = 0; it < SV_LEAFL_IT; it++) {
        const int s = sbase + it * SV_LEAFL_THREADS;
        if (s >= a.S) continue;
        a.sumS[(size_t)blockIdx.y * a.sum_stride + s] = fma((double)accE[it], SV_LN2, accS[it]);
        if (want0) a.sum0[(size_t)blockIdx.y * a.sum_stride + s] = fma((double)acc0e[it], SV_LN2, accS[it]) + log(acc0m[it]);
    }
}
