// maxsum.cuh -- the max-sum ("maximum-likelihood genotypes") half of the likelihood core, for variant calling.
//
// Replaces, for one fixed tree and parameter set (traceMLGenotypes / VariantCaller, not the MCMC inner loop):
//   * RawReadCountsModelFiniteMuDM.computeMixedLikelihood(..., MLCompIndex)  rawreadcountsmodel/...DM.java:54-77, 121-260
//   * ScsBeerLikelihoodCore.calculateLogPartials(..., parentMLGenotypes, ...) likelihood/ScsBeerLikelihoodCore.java:2918-3486
//       (the max-sum part and the arg-max lists; the sum-product / constant-site parts are the walk in prune.cuh)
//   * ScsBeerLikelihoodCore.getPatternCategories                              :278-296
//   * ScsTreeLikelihood.getMLGenotypes (first run)                            likelihood/ScsTreeLikelihood.java:947-1460
//       as a top-down pass that yields, per (category, site, node), the SET of genotypes the node takes in any
//       maximum-likelihood combination (MLGenotypesCollection) and whether the combination is unique.
//
// The reference keeps java.util.List<Integer> tie lists per (node, category, pattern, parent genotype, child); here a
// tie list is a 4-bit mask and the two children's masks share one byte: 4 bytes per (node, site, category).
//
// Arithmetic: log space like the reference (max-sum only adds and compares, so nothing underflows); log P is taken on
// the host in double precision, the device adds and compares -- ties are decided by exact equality (:3107-3131).
//
// Thread mapping as in prune.cuh: lane = site * K + category owns one (site, category) for the whole walk over the
// ops; ops are dependent, lanes are not, so there is no barrier and one launch per pass.  HBM traffic per
// (internal node, site): read 2 x K x 32 B children, write K x 32 B max-sum partials + K x 4 B back-pointers.
#pragma once
#include "common.cuh"
#include "leaf.cuh"

struct SvMlOp {
    int dst;      // internal index of the parent (node - n_leaves)
    int c1, c2;   // children: leaf index, or internal index; c2 < 0 for the unary root
    int n1, n2;   // node numbers of the children (rows of the log-matrix table)
    int flags;    // SV_OP_LEAF1 | SV_OP_LEAF2 | SV_OP_UNARY
    int pad0, pad1;
};

// ---- leaves: log-space mixture, component arg-max ---------------------------------------------------------------
// log NB(x; r, p) of a cell at coverage x for 0, 1 and 2 sequenced alleles, tabulated like the linear-space tables of
// leaf.cuh (same formula, same operation order): the max-sum leaf kernel then needs no logGamma per (cell, site)
__global__ void k_build_cell_log_tables(double* __restrict__ cellL, const double* __restrict__ size_factors, int n_cells,
                                        int C, SvCovParams cp) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (x >= C || j >= n_cells) return;
    double e[4];
    e[0] = cp.single_ado ? 0.0 : sv_seq_cov_log(x, 0, size_factors[j], cp.t, cp.v);
    e[1] = sv_seq_cov_log(x, 1, size_factors[j], cp.t, cp.v);
    e[2] = sv_seq_cov_log(x, 2, size_factors[j], cp.t, cp.v);
    e[3] = 0.0;
    sv_st256(cellL + ((size_t)j * C + x) * 4, e);
}

struct SvMlLeafArgs {
    const int4* counts;  // [cell][site]
    const double *dmA, *dmB;
    const double* cellL;     // [cell][C][4] log NB for 0 / 1 / 2 sequenced alleles
    const double* size_factors;
    double* leafL;           // [cell][site][4] log genotype likelihoods (0/0, 0/1, 1/1, 1/1')
    unsigned short* ado;     // [cell][site] four nibbles: arg-max set of the dropout components of genotype g
    int n_cells, S, C;      // S = sites of this launch (a site chunk of a sparse core, else the shard)
    long long counts_stride, site_begin;  // sites per cell in `counts`, first site of this launch
    SvDmAlphas al;
    SvCovParams cp;
};

// MathFunctions.logSumExp (math/util/MathFunctions.java:101-112): max + log(sum exp(x_i - max)), left to right
__device__ __forceinline__ double sv_lse2(double a, double b) {
    const double m = fmax(a, b);
    return __dadd_rn(m, log(__dadd_rn(exp(__dadd_rn(a, -m)), exp(__dadd_rn(b, -m)))));
}
__device__ __forceinline__ double sv_lse3(double a, double b, double c) {
    const double m = fmax(fmax(a, b), c);
    double s = __dadd_rn(exp(__dadd_rn(a, -m)), exp(__dadd_rn(b, -m)));
    s = __dadd_rn(s, exp(__dadd_rn(c, -m)));
    return __dadd_rn(m, log(s));
}
// MathFunctions.maxIndices (:168-180) over up to three components, as a bit mask
__device__ __forceinline__ unsigned sv_max_indices(const double* v, int n) {
    double mx = v[0];
    unsigned mask = 1u;
    for (int i = 1; i < n; i++) {
        if (v[i] > mx) {
            mx = v[i];
            mask = 1u << i;
        } else if (v[i] == mx)
            mask |= 1u << i;
    }
    return mask;
}

// The dropout mixture of one (cell, site[, category]) in log space with the arg-max set of its components per genotype
// (RawReadCountsModelFiniteMuDM.java:54-77, 121-260).  N0..N3: nucleotide log-densities of 0/0, 1/1, 1/1', 0/1;
// nb0..nb2: log NB of the coverage for 0 / 1 / 2 sequenced alleles.  -> out[g], four nibbles of component masks.
__device__ __forceinline__ unsigned sv_ml_leaf_mixture(double N0, double N1, double N2, double N3, double nb0, double nb1,
                                                       double nb2, double lt, double l1t, int single_ado, double (&out)[4]) {
    const double n01 = sv_lse2(N0, N1);  // logSumExp{nuc[+0], nuc[+1]} of the heterozygous one-allele term
    unsigned masks = 0;
    double comp[3];
    if (single_ado) {
        // :137-141, 166-170, 195-199, 224-228: tmp[0] no dropout, tmp[1] one allele lost
        const double nn[4] = {N0, N3, N1, N2};      // "no dropout" nucleotide term of genotype g
        const double na[4] = {N0, 0.0, N1, N1};     // "one allele left" term (g == 1 handled apart)
#pragma unroll
        for (int g = 0; g < 4; g++) {
            comp[0] = __dadd_rn(__dadd_rn(l1t, nn[g]), nb2);
            comp[1] = g == 1 ? __dadd_rn(__dadd_rn(__dadd_rn(lt, -log(2.0)), n01), nb1)
                             : __dadd_rn(__dadd_rn(lt, na[g]), nb1);
            out[g] = sv_lse2(comp[0], comp[1]);
            masks |= sv_max_indices(comp, 2) << (4 * g);
        }
    } else {
        // locus dropout (:147-153 etc.): tmp[0] both alleles, tmp[1] one lost, tmp[2] both lost (no nucleotide factor)
        const double nn[4] = {N0, N3, N1, N2};
        const double na[4] = {N0, 0.0, N1, N1};
        const double two_l1t = __dmul_rn(2.0, l1t), two_lt = __dmul_rn(2.0, lt);
#pragma unroll
        for (int g = 0; g < 4; g++) {
            comp[0] = __dadd_rn(__dadd_rn(two_l1t, nn[g]), nb2);
            comp[1] = g == 1 ? __dadd_rn(__dadd_rn(__dadd_rn(lt, l1t), n01), nb1)
                             : __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(log(2.0), lt), l1t), na[g]), nb1);
            comp[2] = __dadd_rn(two_lt, nb0);
            out[g] = sv_lse3(comp[0], comp[1], comp[2]);
            masks |= sv_max_indices(comp, 3) << (4 * g);
        }
    }
    return masks;
}

__global__ void __launch_bounds__(128) k_leaf_ml(SvMlLeafArgs a) {
    const int j = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.S) return;
    const int C = a.C;
    const int4 c = __ldg(a.counts + (size_t)j * a.counts_stride + a.site_begin + s);
    const int cov = c.w;
    double N0, N1, N2, N3;  // nucleotide log-densities of 0/0, 1/1, 1/1', 0/1 (NucReadCountsModelFiniteMuDM.java:33-96)
    sv_leaf_nuc_terms<false>(c, a.dmA, a.dmB, C, a.al, N0, N1, N2, N3);
    const double sf = a.size_factors[j];
    const double lt = log(a.cp.theta), l1t = log(1 - a.cp.theta);
    double nb0, nb1, nb2;
    if (cov >= 0 && cov < C) {
        double e[4];
        sv_ld256_nc(a.cellL + ((size_t)j * C + cov) * 4, e);
        nb0 = e[0]; nb1 = e[1]; nb2 = e[2];
    } else {
        nb2 = sv_seq_cov_log(cov, 2, sf, a.cp.t, a.cp.v);
        nb1 = sv_seq_cov_log(cov, 1, sf, a.cp.t, a.cp.v);
        nb0 = a.cp.single_ado ? 0.0 : sv_seq_cov_log(cov, 0, sf, a.cp.t, a.cp.v);
    }
    double out[4];
    const unsigned masks = sv_ml_leaf_mixture(N0, N1, N2, N3, nb0, nb1, nb2, lt, l1t, a.cp.single_ado, out);
    sv_st256(a.leafL + ((size_t)j * a.S + s) * 4, out);
    a.ado[(size_t)j * a.S + s] = (unsigned short)masks;
}

// The same under the per-pattern coverage model (SIEVE_FLAG_PRIVATE_SEQ_COV, leaf_private.cuh): the coverage terms of
// category k come from the (category, site) pair's own (t, v), evaluated directly; the leaves then differ between the
// categories: leafL [cell][site][K][4], ado [cell][site][K].
struct SvMlLeafPrivArgs {
    const int4* counts;
    const double *dmA, *dmB;
    const double* size_factors;
    const double2* tv;       // [K][S]
    double* leafL;           // [cell][S][K][4]
    unsigned short* ado;     // [cell][S][K]
    int n_cells, S, C, K;
    SvDmAlphas al;
    double theta;
    int single_ado;
};
__global__ void __launch_bounds__(128) k_leaf_ml_private(SvMlLeafPrivArgs a) {
    const int j = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.S) return;
    const int4 c = __ldg(a.counts + (size_t)j * a.S + s);
    const int cov = c.w;
    double N0, N1, N2, N3;
    sv_leaf_nuc_terms<false>(c, a.dmA, a.dmB, a.C, a.al, N0, N1, N2, N3);
    const double sf = a.size_factors[j];
    const double lt = log(a.theta), l1t = log(1 - a.theta);
    for (int k = 0; k < a.K; k++) {
        const double2 p = a.tv[(size_t)k * a.S + s];
        const double nb2 = sv_seq_cov_log(cov, 2, sf, p.x, p.y);
        const double nb1 = sv_seq_cov_log(cov, 1, sf, p.x, p.y);
        const double nb0 = a.single_ado ? 0.0 : sv_seq_cov_log(cov, 0, sf, p.x, p.y);
        double out[4];
        const unsigned masks = sv_ml_leaf_mixture(N0, N1, N2, N3, nb0, nb1, nb2, lt, l1t, a.single_ado, out);
        const size_t at = ((size_t)j * a.S + s) * a.K + k;
        sv_st256(a.leafL + at * 4, out);
        a.ado[at] = (unsigned short)masks;
    }
}

// ---- bottom-up: max-sum partials and back-pointers -------------------------------------------------------------
struct SvMlArgs {
    const SvMlOp* ops;
    int n_ops;
    const double* leafL;    // [cell][S][4]
    const unsigned short* ado;  // [cell][S]
    const double* logm;     // [node][K][16] log(max(P, Double.MIN_VALUE)), row = parent genotype
    double* ml;             // [internal][S][K][4]
    unsigned* back;         // [internal][S][K] byte p: low nibble = arg-max set over child 1's genotypes, high = child 2
    unsigned char* coll;    // [node][S][K] genotype set of the node over all ML combinations
    unsigned char* ado_sel; // [cell][S][K] union of the arg-max dropout components over the leaf's genotype set
    unsigned char* ambiguous;  // [S][K] 1 <=> more than one ML combination
    int S, K, n_leaves, root_genotype;
    const double* rawm;     // [node][K][16] P itself, for the leaf children of the linear-space twin (linear == 1)
    int linear;             // 1: ScsBeerLikelihoodCore.java:1759-2335 (useLogPartials = false): a LEAF child contributes
                            //    log(P[p][c] * leaf[c]) with the linear leaf likelihood and the unclamped P (:1948-1949, 1964-1978)
    int leaf_k;             // 1: the leaves are per category (k_leaf_ml_private): leafL [cell][S][K][4], ado [cell][S][K]
};

// one child's side of ScsBeerLikelihoodCore.java:3100-3131: max over child genotypes of log P[p][c] + value[c], with
// the reference's sequential tie rule (first index seeds, '<' replaces, '==' appends)
__device__ __forceinline__ void sv_ml_side(const double* lm /* shared */, const double (&v)[4], double (&mx)[4],
                                           unsigned (&mask)[4]) {
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const double2 ma = *reinterpret_cast<const double2*>(lm + 4 * p);
        const double2 mb = *reinterpret_cast<const double2*>(lm + 4 * p + 2);
        const double m[4] = {ma.x, ma.y, mb.x, mb.y};
        double best = __dadd_rn(m[0], v[0]);
        unsigned bm = 1u;
#pragma unroll
        for (int cgt = 1; cgt < 4; cgt++) {
            const double t = __dadd_rn(m[cgt], v[cgt]);
            if (best < t) {
                best = t;
                bm = 1u << cgt;
            } else if (best == t)
                bm |= 1u << cgt;
        }
        mx[p] = best;
        mask[p] = bm;
    }
}

// a LEAF child in the linear-space twin (calculateLeafLeafPruning / calculateLeafPartialPruning, :1895-2166): the candidates
// are log(P[p][c] * leaf[c]) -- product first, no Double.MIN_VALUE clamp, so a zero entry of P gives -inf and ties among
// -inf candidates are ties (the comparison is on the logs, :1964-1978).  lm holds P, v the LOG leaf likelihoods.
__device__ __forceinline__ void sv_ml_side_lin_leaf(const double* lm /* shared */, const double (&v)[4], double (&mx)[4],
                                                    unsigned (&mask)[4]) {
    const double L[4] = {exp(v[0]), exp(v[1]), exp(v[2]), exp(v[3])};
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const double2 ma = *reinterpret_cast<const double2*>(lm + 4 * p);
        const double2 mb = *reinterpret_cast<const double2*>(lm + 4 * p + 2);
        const double m[4] = {ma.x, ma.y, mb.x, mb.y};
        double best = log(__dmul_rn(m[0], L[0]));
        unsigned bm = 1u;
#pragma unroll
        for (int cgt = 1; cgt < 4; cgt++) {
            const double t = log(__dmul_rn(m[cgt], L[cgt]));
            if (best < t) {
                best = t;
                bm = 1u << cgt;
            } else if (best == t)
                bm |= 1u << cgt;
        }
        mx[p] = best;
        mask[p] = bm;
    }
}

__global__ void __launch_bounds__(128) k_ml_up(SvMlArgs a) {
    const long long lane_raw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long lanes = (long long)a.S * a.K;
    // lanes past the end stay (every lane of a warp stages a matrix row) as duplicates of the last lane; they never store
    const bool lane_ok = lane_raw < lanes;
    const long long lane = lane_ok ? lane_raw : lanes - 1;
    const int K = a.K;
    const long long s = lane / K;
    const int k = (int)(lane - s * K);
    // warp-private copies of the op's two log matrices (all K categories): one 256-bit load per lane stages them, the
    // 32 reads per op then are shared-memory broadcasts instead of L1 transactions (ncu: the global-load queue was the
    // limiter -- lg_throttle -- when every lane fetched its rows itself)
    __shared__ __align__(32) double sm[(128 / 32) * 2 * SV_KMAX * SV_MAT_STRIDE];
    const int wl = threadIdx.x & 31;
    double* wm = sm + (threadIdx.x >> 5) * (2 * SV_KMAX * SV_MAT_STRIDE);
    const int st_which = wl >> 4, st_kk = (wl >> 2) & 3, st_q = wl & 3;
    double* st_dst = wm + st_which * (SV_KMAX * SV_MAT_STRIDE) + st_kk * SV_MAT_STRIDE + st_q * 4;
    const double* M1 = wm + k * SV_MAT_STRIDE;
    const double* M2 = wm + SV_KMAX * SV_MAT_STRIDE + k * SV_MAT_STRIDE;
    for (int i = 0; i < a.n_ops; i++) {
        const SvMlOp op = a.ops[i];
        double v1[4], v2[4], m1[4], m2[4];
        unsigned b1[4], b2[4] = {0u, 0u, 0u, 0u};
        {
            double m[4] = {0.0, 0.0, 0.0, 0.0};
            const bool lin_leaf = a.linear && (op.flags & (st_which ? SV_OP_LEAF2 : SV_OP_LEAF1));
            if (st_kk < K && !(st_which == 1 && (op.flags & SV_OP_UNARY)))
                sv_ld256_nc((lin_leaf ? a.rawm : a.logm) + ((size_t)(st_which ? op.n2 : op.n1) * K + st_kk) * 16 + st_q * 4, m);
            __syncwarp();  // every lane is done with the previous op's matrices
            *reinterpret_cast<double2*>(st_dst) = make_double2(m[0], m[1]);
            *reinterpret_cast<double2*>(st_dst + 2) = make_double2(m[2], m[3]);
            __syncwarp();
        }
        if (op.flags & SV_OP_LEAF1)
            sv_ld256_nc(a.leaf_k ? a.leafL + ((size_t)op.c1 * lanes + lane) * 4 : a.leafL + ((size_t)op.c1 * a.S + s) * 4, v1);
        else
            sv_ld256(a.ml + ((size_t)op.c1 * lanes + lane) * 4, v1);
        if (a.linear && (op.flags & SV_OP_LEAF1))
            sv_ml_side_lin_leaf(M1, v1, m1, b1);
        else
            sv_ml_side(M1, v1, m1, b1);
        if (!(op.flags & SV_OP_UNARY)) {
            if (op.flags & SV_OP_LEAF2)
                sv_ld256_nc(a.leaf_k ? a.leafL + ((size_t)op.c2 * lanes + lane) * 4 : a.leafL + ((size_t)op.c2 * a.S + s) * 4, v2);
            else
                sv_ld256(a.ml + ((size_t)op.c2 * lanes + lane) * 4, v2);
            if (a.linear && (op.flags & SV_OP_LEAF2))
                sv_ml_side_lin_leaf(M2, v2, m2, b2);
            else
                sv_ml_side(M2, v2, m2, b2);
#pragma unroll
            for (int p = 0; p < 4; p++) m1[p] = __dadd_rn(m1[p], m2[p]);  // parentMLPartials = max1 + max2 (:3139)
        }
        unsigned w = 0;
#pragma unroll
        for (int p = 0; p < 4; p++) w |= (b1[p] | (b2[p] << 4)) << (8 * p);
        if (lane_ok) {
            sv_st256(a.ml + ((size_t)op.dst * lanes + lane) * 4, m1);
            a.back[(size_t)op.dst * lanes + lane] = w;
        }
    }
}

// ---- top-down: which genotypes does every node take in the ML combinations? ------------------------------------
__global__ void __launch_bounds__(128) k_ml_down(SvMlArgs a) {
    const long long lane = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long lanes = (long long)a.S * a.K;
    if (lane >= lanes) return;
    const int n = a.n_leaves;
    const long long s = lane / a.K;
    unsigned amb = 0;
    for (int i = a.n_ops - 1; i >= 0; i--) {
        const SvMlOp op = a.ops[i];
        unsigned pc;
        if (i == a.n_ops - 1) {
            pc = 1u << a.root_genotype;  // the root's genotype is fixed (ScsTreeLikelihood.java:960, 995-1001)
            a.coll[(size_t)(n + op.dst) * lanes + lane] = (unsigned char)pc;
        } else
            pc = a.coll[(size_t)(n + op.dst) * lanes + lane];
        const unsigned w = a.back[(size_t)op.dst * lanes + lane];
        unsigned g1 = 0, g2 = 0;
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const unsigned b = (pc >> p) & 1u ? (w >> (8 * p)) & 0xffu : 0u;
            g1 |= b & 15u;
            g2 |= b >> 4;
        }
        amb |= (g1 & (g1 - 1)) | (g2 & (g2 - 1));
        for (int side = 0; side < 2; side++) {
            if (side == 1 && (op.flags & SV_OP_UNARY)) break;
            const bool is_leaf = op.flags & (side == 0 ? SV_OP_LEAF1 : SV_OP_LEAF2);
            const int c = side == 0 ? op.c1 : op.c2;
            const unsigned g = side == 0 ? g1 : g2;
            if (is_leaf) {
                a.coll[(size_t)c * lanes + lane] = (unsigned char)g;
                const unsigned ad = a.leaf_k ? a.ado[(size_t)c * lanes + lane] : a.ado[(size_t)c * a.S + s];
                unsigned u = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const unsigned m = (g >> q) & 1u ? (ad >> (4 * q)) & 15u : 0u;
                    u |= m;
                    amb |= m & (m - 1);
                }
                a.ado_sel[(size_t)c * lanes + lane] = (unsigned char)u;
            } else
                a.coll[(size_t)(n + c) * lanes + lane] = (unsigned char)g;
        }
    }
    a.ambiguous[lane] = amb ? 1 : 0;
}

// getPatternCategories (ScsBeerLikelihoodCore.java:278-296): arg-max set over the categories of the root's partial at
// the root genotype.  The partial is scaled-linear with ONE scale per (node, site), so the linear values order like
// the logs the reference compares.
__global__ void k_ml_categories(const double* __restrict__ root_part /*[S][SV_KMAX][4]*/, int S, int K, int root_genotype,
                                unsigned char* __restrict__ cats) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double mx = -1.7976931348623157e308;
    unsigned mask = 0;
    for (int k = 0; k < K; k++) {
        const double t = root_part[((size_t)s * SV_KMAX + k) * 4 + root_genotype];
        if (mx < t) {
            mx = t;
            mask = 1u << k;
        } else if (mx == t)
            mask |= 1u << k;
    }
    cats[s] = (unsigned char)mask;
}

// Variant call per (node, site) for the lowest ML category: the genotype when it is unique, else -1; same for the
// dropout state of the leaves (collectMLGenotypesAndLogLikelihoods' common case, ScsTreeLikelihood.java:2033-2081).
__global__ void k_ml_call(const unsigned char* __restrict__ coll, const unsigned char* __restrict__ ado_sel,
                          const unsigned char* __restrict__ cats, const unsigned char* __restrict__ ambiguous, int S,
                          int K, int n_nodes, int n_leaves, signed char* __restrict__ genotypes /*[node][S]*/,
                          signed char* __restrict__ ados /*[leaf][S]*/, unsigned char* __restrict__ category,
                          unsigned char* __restrict__ amb_site) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int node = blockIdx.y;
    if (s >= S) return;
    const unsigned cm = cats[s];
    const int k = cm ? __ffs(cm) - 1 : 0;
    const size_t lane = (size_t)s * K + k;
    const size_t lanes = (size_t)S * K;
    const unsigned g = coll[(size_t)node * lanes + lane];
    genotypes[(size_t)node * S + s] = (g && !(g & (g - 1))) ? (signed char)(__ffs(g) - 1) : (signed char)-1;
    if (node < n_leaves) {
        const unsigned u = ado_sel[(size_t)node * lanes + lane];
        ados[(size_t)node * S + s] = (u && !(u & (u - 1))) ? (signed char)(__ffs(u) - 1) : (signed char)-1;
    }
    if (node == 0) {
        category[s] = (unsigned char)k;
        amb_site[s] = (unsigned char)(((cm & (cm - 1)) ? 1 : 0) | (ambiguous[lane] ? 2 : 0));
    }
}

// Everything the host needs to enumerate the tied combinations of ONE site (rare): per node and category the
// back-pointer bytes (leaves: the dropout arg-max nibbles spread to bytes) and the node's sum-product log partial.
__global__ void k_ml_site_gather(const unsigned* __restrict__ back, const unsigned short* __restrict__ ado,
                                 const double* __restrict__ leafL, const double* __restrict__ part,
                                 const int* __restrict__ scale, const unsigned* __restrict__ part_row /*[node]*/,
                                 int S, int K, int n_leaves, int n_nodes, int site, int leaf_k,
                                 unsigned char* __restrict__ out_back /*[node][K][4]*/,
                                 double* __restrict__ out_logp /*[node][K][4]*/) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_nodes * K) return;
    const int node = t / K, k = t - node * K;
    const size_t lanes = (size_t)S * K, lane = (size_t)site * K + k;
    for (int g = 0; g < 4; g++) {
        if (node < n_leaves) {
            const size_t at = leaf_k ? (size_t)node * lanes + lane : (size_t)node * S + site;
            out_back[(size_t)t * 4 + g] = (unsigned char)((ado[at] >> (4 * g)) & 15u);
            out_logp[(size_t)t * 4 + g] = leafL[at * 4 + g];
        } else {
            out_back[(size_t)t * 4 + g] = (unsigned char)((back[(size_t)(node - n_leaves) * lanes + lane] >> (8 * g)) & 0xffu);
            const size_t row = (size_t)part_row[node] + site;
            out_logp[(size_t)t * 4 + g] = log(part[(row * SV_KMAX + k) * 4 + g]) + (double)scale[row] * SV_LN2;
        }
    }
}
