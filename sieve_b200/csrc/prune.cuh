// prune.cuh -- Felsenstein pruning over dirty nodes, root integration and per-site log-likelihood.
//
// Replaces ScsBeerLikelihoodCore.calculateLogPartials / calculatePartials -> calculate(Log)PartialPartialPruning
// (likelihood/ScsBeerLikelihoodCore.java:1283-1501 / :857-1055), integratePartials (:153-218) and
// calculateLogLikelihoods (:255-269).
//
// Arithmetic: scaled linear space.  A node's partial for (site, category k) is 4 doubles x[g] and ONE binary exponent
// per (node, site) shared by all categories and states: true value = x[g] * 2^e * exp(sum of the leaves' scales below).
// The 16 values of a site are renormalised by an exact power of two (sv_pow2_rescale_hi_int) only where the host's bound
// on their magnitude asks for it (SV_OP_RENORM, core_planner.cuh sv_bound_op): scaling by 2^k is exact and the exponents
// are summed as integers, so WHEN the renormalisation happens cannot change a bit of the result; the root always
// renormalises.  No transcendental is evaluated until the root's log.  The reference's log-space result is reproduced to rounding (~1e-16 relative on the
// likelihood); its clamp log(P > 0 ? P : Double.MIN_VALUE) (:1430) is mirrored by clamping the matrices when they are
// staged (SIEVE_FLAG_LOG_PARTIALS), its linear-mode clamp of the constant-site product (:1041-1047) likewise.
//
// Layout in HBM: partials [slot][site][K][4] (one site = 128 contiguous bytes at K = 4), scales [slot][site];
// leaves [slot][site][4] (category independent; no per-leaf scale, see leaf.cuh).  Thread mapping: lane = 4 * site_in_warp + k, so a warp
// reads/writes 8 sites * 128 B = 1 KB contiguous with one 256-bit access per lane.
//
// Sites are independent, so one launch can walk ANY number of dependent ops: each thread only ever re-reads what it
// (or, for the scales, a lane of its own block before a barrier) wrote.  The same kernel serves the level-batched
// schedule (grid.y = ops of one level, one op per block) and the single-launch walk (grid.y = 1, loop over all ops).
#pragma once
#include "common.cuh"

struct SvPruneArgs {
    double* part;          // [slots][S][4][4]   (category stride is always 4; lanes k >= K idle)
    int* scale;            // [slots][S] binary exponent of the slot's values
    double* cpart;         // constant-site twins (NULL when not maintained)
    int* cscale;
    const double* leafX;   // [leaf slots][S][4]; the leaves' log scales are not stored per leaf (leaf.cuh) ...
    const double* lsum;    // ... their sum over all cells, per site of the shard (+ out_offset): added at the root
    const double* mats;    // [matrix slots][4][16], already clamped when SIEVE_FLAG_LOG_PARTIALS
    const double2* qfac;   // [matrix slots][4] = (expm1(-0.2 d) / 8, expm1(-0.4 d) / 16): the Q-specialised path
    int clamp_zero_distance;  // SIEVE_FLAG_LOG_PARTIALS: a zero distance gets the Double.MIN_VALUE off-diagonals
    const SvOp* ops;
    int op_begin, op_end;  // walk: ops [op_begin, op_end); per-level: op = op_begin + blockIdx.y
    int per_level;
    int pf_dist;           // ops ahead that the L2 prefetch annotations look (0: none); the first pf_dist ops are announced by a prologue
    int S, K;              // S = sites of this launch; the ops' rows were formed with the pools' site capacity
    long long out_offset;  // first site of this launch in site_logl / site_const (streamed evaluation)
    const int* site_list;  // LEAFK kernel only, or NULL: the S sites of this launch are site_list[0 .. S) of the shard (in-place
                           // re-evaluation of the patterns whose (t, v) moved, sieve_private_retraverse_changed)
    int linear_const_clamp;  // !SIEVE_FLAG_LOG_PARTIALS: clamp of the constant-site product (:1041-1047)
    int const_genotype;
    int root_genotype;
    double prop[4];        // category proportions at the root
    double* site_logl;     // [S]
    double* site_const;    // [S]
    unsigned long long part_rows, leaf_rows;  // rows of the two pools (SV_BOUND only)
    int n_matrix_slots;
};

__device__ __forceinline__ void sv_matvec(const double* __restrict__ M, const double (&x)[4], double (&y)[4]) {
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const double2 a = *reinterpret_cast<const double2*>(M + 4 * p);
        const double2 b = *reinterpret_cast<const double2*>(M + 4 * p + 2);
        y[p] = fma(b.y, x[3], fma(b.x, x[2], fma(a.y, x[1], a.x * x[0])));
    }
}

__device__ __forceinline__ double sv_sel(const double (&x)[4], int g) {
    return g == 0 ? x[0] : (g == 1 ? x[1] : (g == 2 ? x[2] : x[3]));
}

#ifndef SV_PRUNE_THREADS
#define SV_PRUNE_THREADS 128
#endif
#ifndef SV_PRUNE_NS
#define SV_PRUNE_NS 2  // sites per lane: the op's matrices are read from shared memory once and applied to NS sites
#endif
#define SV_PRUNE_SITES ((SV_PRUNE_THREADS / 4) * SV_PRUNE_NS)
#ifndef SV_PRUNE_MINB_CONST
#define SV_PRUNE_MINB_CONST 4
#endif
#ifndef SV_PRUNE_MINB
#define SV_PRUNE_MINB 5
#endif
#ifndef SV_PRUNE_NS_Q
#define SV_PRUNE_NS_Q 2       // variant-only Q walk: 2 sites per lane at 6 blocks/SM (78 regs) measured best
#endif
#ifndef SV_PRUNE_NS_Q_TWIN
#define SV_PRUNE_NS_Q_TWIN 1  // with the per-site constant twin the state doubles: 1 site per lane at 8 blocks/SM
#endif
#ifndef SV_PRUNE_MINB_CONST_Q
#define SV_PRUNE_MINB_CONST_Q 8
#endif
#ifndef SV_PRUNE_MINB_Q
#define SV_PRUNE_MINB_Q 6
#endif
// the MINB figures count resident groups of four warps per SM (the block size the kernels were tuned with); a build
// with another block size keeps the same number of resident warps
__host__ __device__ constexpr int sv_prune_min_blocks(int minb4) { return minb4 * 128 / SV_PRUNE_THREADS; }
__host__ __device__ constexpr int sv_prune_ns(bool with_const, bool qfast) {
    return qfast ? (with_const ? SV_PRUNE_NS_Q_TWIN : SV_PRUNE_NS_Q) : SV_PRUNE_NS;
}
__host__ __device__ constexpr int sv_prune_sites(bool with_const, bool qfast) {
    return (SV_PRUNE_THREADS / 4) * sv_prune_ns(with_const, qfast);
}

// y_j = M x_j for the NS sites of a lane; every matrix element is read from shared memory once
template <int NS>
__device__ __forceinline__ void sv_matvec_n(const double* __restrict__ M, const double (&x)[NS][4], double (&y)[NS][4]) {
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const double2 a = *reinterpret_cast<const double2*>(M + 4 * p);
        const double2 b = *reinterpret_cast<const double2*>(M + 4 * p + 2);
#pragma unroll
        for (int j = 0; j < NS; j++) y[j][p] = fma(b.y, x[j][3], fma(b.x, x[j][2], fma(a.y, x[j][1], a.x * x[j][0])));
    }
}

// ---- the Q-specialised contraction ---------------------------------------------------------------------------------
// SIEVE's only usable substitution model has a FIXED rate matrix (substitutionmodel/ScsFiniteMuExtendedModel.java:62-86,
// normalised by ScsSubstitutionModelBase.java:200-231):  Q = 0.3 [[-1,1,0,0],[1/6,-2/3,1/6,1/3],[0,1/3,-1,2/3],[0,1/3,1/3,-2/3]]
// with eigenvalues 0, -0.2, -0.4, -0.4 and is diagonalisable, so with g1 = expm1(-0.2 d), g2 = expm1(-0.4 d)
//     P(d) = I + g1 * Pi1 + g2 * Pi2,   Pi1 = u v^T / 8 (u = (3,1,-1,-1), v = (1,2,-1,-2)),   Pi2 of rank 2,
// and  y = P x  needs two scalars per (node, category) instead of a 4 x 4 matrix from shared memory:
//     t1 = g1/8  (x0 + 2 x1 - x2 - 2 x3),  ta = g2/16 (-3 x0 + 6 x1 - x2 - 2 x3),  tb = g2/16 (x0 - 2 x1 - 5 x2 + 6 x3)
//     y0 = x0 + 3 (t1 - ta),  y1 = x1 + (t1 + ta),  y3 = x3 + (tb - t1),  y2 = x2 + (tb - t1) + g2 (x2 - x3)
// (expm1 keeps the O(d) entries accurate for short branches).  q = (g1 / 8, g2 / 16).
template <int NS>
__device__ __forceinline__ void sv_matvec_q(const double2 q, const double (&x)[NS][4], double (&y)[NS][4], bool clamp0) {
#pragma unroll
    for (int j = 0; j < NS; j++) {
        const double A = fma(2.0, x[j][1], x[j][0]);   // x0 + 2 x1
        const double B = fma(2.0, x[j][3], x[j][2]);   // x2 + 2 x3
        const double D = fma(-2.0, x[j][1], x[j][0]);  // x0 - 2 x1
        const double t1 = q.x * (A - B);
        const double ta = q.y * fma(-3.0, D, -B);
        const double tb = q.y * (D + fma(6.0, x[j][3], -5.0 * x[j][2]));
        const double u = tb - t1;
        y[j][0] = fma(3.0, t1 - ta, x[j][0]);
        y[j][1] = x[j][1] + (t1 + ta);
        y[j][3] = x[j][3] + u;
        y[j][2] = fma(16.0 * q.y, x[j][2] - x[j][3], x[j][2] + u);
        if (clamp0 && q.x == 0.0) {
            // d == 0: P = I; the reference takes log(P > 0 ? P : Double.MIN_VALUE) (:1430) for the 12 zero entries
            const double sum = (x[j][0] + x[j][1]) + (x[j][2] + x[j][3]);
#pragma unroll
            for (int p = 0; p < 4; p++) y[j][p] = fma(SV_MIN_VALUE, sum - x[j][p], x[j][p]);
        }
    }
}
// the constant-site contribution of a LEAF child: only the genotype-0 column of P survives (:1432-1436)
template <int NS>
__device__ __forceinline__ void sv_col0_q(const double2 q, const double (&x)[NS][4], double (&y)[NS][4], bool clamp0) {
#pragma unroll
    for (int j = 0; j < NS; j++) {
        const double x0 = x[j][0];
        const double t1 = q.x * x0, ta = -3.0 * (q.y * x0), tb = q.y * x0;
        y[j][0] = fma(3.0, t1 - ta, x0);
        y[j][1] = t1 + ta;
        y[j][2] = tb - t1;
        y[j][3] = tb - t1;
        if (clamp0 && q.x == 0.0) {
            y[j][1] = y[j][2] = y[j][3] = SV_MIN_VALUE * x0;
        }
    }
}

// GENO0: root genotype == constant genotype == 0 (always so for ScsFiniteMuExtendedModel.java:37-40); the general
// case takes the same kernel with run-time selects.
// QFAST: every matrix of the launch was given as a branch distance (sieve_set_distances): no shared memory at all.
// MINB: resident blocks per SM asked of the compiler; 0 = the default of the variant (SIEVE_B200_PRUNE_MINB instantiates
// the variant-only Q walk at 7 and 8 for comparison).
// LEAFK: the leaf pool holds K x 4 values per (cell, site) (SIEVE_FLAG_PRIVATE_SEQ_COV, leaf_private.cuh) in the layout
// of an internal partial; a leaf child is then read like one, without a scale.
template <bool WITH_CONST, bool GENO0, bool QFAST, int MINB = 0, bool LEAFK = false>
__global__ void __launch_bounds__(SV_PRUNE_THREADS, sv_prune_min_blocks(MINB ? MINB
                                                                        : QFAST ? (WITH_CONST ? SV_PRUNE_MINB_CONST_Q : SV_PRUNE_MINB_Q)
                                                                                : (WITH_CONST ? SV_PRUNE_MINB_CONST : SV_PRUNE_MINB)))
    k_prune(SvPruneArgs a) {
    constexpr int NS = sv_prune_ns(WITH_CONST, QFAST);
    constexpr int SITES_PER_BLOCK = (SV_PRUNE_THREADS / 4) * NS;
    // warp-private copies of the op's two 4 x 4 x 4 matrices: no block-wide barrier anywhere in the walk, so the
    // warps of an SM drift apart and overlap each other's memory latency
    __shared__ __align__(32) double sm[(SV_PRUNE_THREADS / 32) * 2 * SV_KMAX * SV_MAT_STRIDE];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int k = min(lane & 3, a.K - 1);
    const int klane = lane & 3;
    // a warp owns 8 * NS consecutive sites; pass j of a lane is site base + 8 j, so that every pass of the warp is one
    // contiguous 1 KB access
    const int site0 = blockIdx.x * SITES_PER_BLOCK + warp * (8 * NS) + (lane >> 2);
    // Lanes beyond the last site (or beyond the last category) are clamped onto the last valid site (category): they
    // load and compute duplicates of a real lane's values, so the loads and the arithmetic need no predicates; only the
    // stores of out-of-range sites are suppressed.
    bool site_ok[NS];
    unsigned srow[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) {
        const int site = site0 + 8 * j;
        site_ok[j] = site < a.S;
        srow[j] = (unsigned)(site_ok[j] ? site : a.S - 1);
        if constexpr (LEAFK)
            if (a.site_list) {
                srow[j] = (unsigned)__ldg(a.site_list + srow[j]);
                SV_BOUND(srow[j] < a.part_rows);  // (the row checks below bound it properly)
            }
    }
    double* wm = sm + warp * (2 * SV_KMAX * SV_MAT_STRIDE);
    // staging role of this lane: matrix `which` (child 1 / 2), category kk, row q  ->  4 doubles = one 256-bit load
    const int st_which = lane >> 4, st_kk = (lane >> 2) & 3, st_q = lane & 3;
    double* st_dst = wm + st_which * (SV_KMAX * SV_MAT_STRIDE) + st_kk * SV_MAT_STRIDE + st_q * 4;
    const int st_off = st_kk * 16 + st_q * 4;
    const double* M1 = wm + k * SV_MAT_STRIDE;
    const double* M2 = wm + SV_KMAX * SV_MAT_STRIDE + k * SV_MAT_STRIDE;
    const int cg = GENO0 ? 0 : a.const_genotype;
    const int rg = GENO0 ? 0 : a.root_genotype;

    int i0 = a.op_begin, i1 = a.op_end;
    if (a.per_level) {
        i0 = a.op_begin + blockIdx.y;
        i1 = i0 + 1;
    }
    // The previous op's result stays in registers (x / cx with scales xs / cxs): the host orders the children so that
    // a consumer of it has it as child 1 (SV_OP_FWD1).
    double x[NS][4], cx[NS][4];
    int xs[NS], cxs[NS];  // binary exponents of x / cx
#pragma unroll
    for (int j = 0; j < NS; j++) {
        xs[j] = cxs[j] = 0;
#pragma unroll
        for (int g = 0; g < 4; g++) x[j][g] = cx[j][g] = 0.0;
    }

    // L2 prefetch of leaf operands a few ops ahead (sv_link_prefetch).  The lanes with category index j < NS ask for the
    // line of their site of pass j: the same address arithmetic as the real load, 8 addresses 32 bytes apart = two lines
    // per pass; one predicated instruction per operand and lane, no register, no scoreboard, no branch.
    const int pf_on = klane < NS ? 1 : 0;
    auto announce = [&](unsigned row, int cond) {
        unsigned sr = srow[0];
#pragma unroll
        for (int j = 1; j < NS; j++) sr = klane == j ? srow[j] : sr;
        if (cond && pf_on) sv_prefetch_l2(a.leafX + (size_t)(row + sr) * 4);
    };
#ifndef SV_PRUNE_NO_PF
    if (!a.per_level && a.pf_dist > 0) {
        // the first pf_dist ops of the list: nobody could announce their operands earlier
        const int jn = min(i1, i0 + a.pf_dist);
#pragma unroll 1
        for (int j = i0; j < jn; j++) {
            const int4 d0 = __ldg(reinterpret_cast<const int4*>(a.ops + j));
            const int fl = __ldg(reinterpret_cast<const int*>(a.ops + j) + 5);
            announce((unsigned)d0.y, fl & SV_OP_FAR1);
            announce((unsigned)d0.z, fl & SV_OP_FAR2);
        }
    }
#endif

    // the descriptor of the next op is fetched while this op's operands are in flight: it is off the critical path
    // (descriptor -> addresses -> operand loads) of every op but the first
    int4 n0 = make_int4(0, 0, 0, 0), n1 = make_int4(0, 0, 0, 0);
    if (i0 < i1) {
        const int4* opp = reinterpret_cast<const int4*>(a.ops + i0);
        n0 = __ldg(opp);
        n1 = __ldg(opp + 1);
    }
    for (int i = i0; i < i1; i++) {
        __syncwarp();  // scratch slots are reused (streamed mode): keep the warp's reads of an op ahead of the next op's writes
        const int4 o0 = n0, o1 = n1;
        if (i + 1 < i1) {
            const int4* opp = reinterpret_cast<const int4*>(a.ops + i + 1);
            n0 = __ldg(opp);
            n1 = __ldg(opp + 1);
        }
        const unsigned dst_base = (unsigned)o0.x, c1_base = (unsigned)o0.y, c2_base = (unsigned)o0.z;
        const int m1 = o0.w, m2 = o1.x, flags = o1.y;
        const bool unary = flags & SV_OP_UNARY;
        const bool leaf1 = flags & SV_OP_LEAF1, leaf2 = flags & SV_OP_LEAF2;
        SV_BOUND(m1 >= 0 && m1 < a.n_matrix_slots && (unary || (m2 >= 0 && m2 < a.n_matrix_slots)));
#ifndef SV_PRUNE_NO_PF
        // leaf operands of the op pf_dist ahead: into L2 now
        announce((unsigned)o1.z, flags & SV_OP_PF1);
        announce((unsigned)o1.w, flags & SV_OP_PF2);
#endif

        // ---- issue every global load of this op up front: matrices (one 256-bit load per lane) and children ----
        double m[4] = {0, 0, 0, 0};
        double2 q1 = make_double2(0.0, 0.0), q2 = make_double2(0.0, 0.0);
        if (QFAST) {
            q1 = __ldg(a.qfac + (size_t)m1 * SV_KMAX + k);
            if (!unary) q2 = __ldg(a.qfac + (size_t)m2 * SV_KMAX + k);
        } else if (!(st_which == 1 && unary))
            sv_ld256_nc(a.mats + (size_t)(st_which ? m2 : m1) * (SV_KMAX * 16) + st_off, m);
        double w[NS][4], cw[NS][4];   // child 2 and its constant-site twin
        int ws[NS], cws[NS];
#pragma unroll
        for (int j = 0; j < NS; j++) {
            ws[j] = cws[j] = 0;
#pragma unroll
            for (int g = 0; g < 4; g++) w[j][g] = cw[j][g] = 0.0;
        }
#pragma unroll
        for (int j = 0; j < NS; j++) {
            const size_t c1_row = c1_base + srow[j], c2_row = c2_base + srow[j];  // 32-bit sums, widened once
            SV_BOUND((flags & SV_OP_FWD1) || c1_row < (leaf1 ? a.leaf_rows : a.part_rows));
            SV_BOUND(unary || c2_row < (leaf2 ? a.leaf_rows : a.part_rows));
            if (!(flags & SV_OP_FWD1)) {
                if (leaf1) {
                    sv_ld256_nc(LEAFK ? a.leafX + c1_row * 16 + k * 4 : a.leafX + c1_row * 4, x[j]);
                    xs[j] = 0;
                } else {
                    sv_ld256(a.part + c1_row * 16 + k * 4, x[j]);
                    xs[j] = a.scale[c1_row];
                    if (WITH_CONST) {
                        sv_ld256(a.cpart + c1_row * 16 + k * 4, cx[j]);
                        cxs[j] = a.cscale[c1_row];
                    }
                }
            }
            if (!unary) {
                if (leaf2) {
                    sv_ld256_nc(LEAFK ? a.leafX + c2_row * 16 + k * 4 : a.leafX + c2_row * 4, w[j]);
                } else {
                    sv_ld256(a.part + c2_row * 16 + k * 4, w[j]);
                    ws[j] = a.scale[c2_row];
                    if (WITH_CONST) {
                        sv_ld256(a.cpart + c2_row * 16 + k * 4, cw[j]);
                        cws[j] = a.cscale[c2_row];
                    }
                }
            }
        }
        if (!QFAST) {
            __syncwarp();  // every lane is done with the previous op's matrices
            *reinterpret_cast<double2*>(st_dst) = make_double2(m[0], m[1]);
            *reinterpret_cast<double2*>(st_dst + 2) = make_double2(m[2], m[3]);
            __syncwarp();
        }
        const bool cz0 = flags & SV_OP_ZEROD;  // set by the host only in log mode and only when a distance is exactly 0

        // ================= variant partials =================
        double y[NS][4];
        if (QFAST) sv_matvec_q<NS>(q1, x, y, cz0); else sv_matvec_n<NS>(M1, x, y);
        if (!unary) {
            double z[NS][4];
            if (QFAST) sv_matvec_q<NS>(q2, w, z, cz0); else sv_matvec_n<NS>(M2, w, z);
#pragma unroll
            for (int j = 0; j < NS; j++)
#pragma unroll
                for (int p = 0; p < 4; p++) y[j][p] *= z[j][p];
        }
        // ================= constant-site partials =================
        double cy[NS][4];
        int cls[NS];
#pragma unroll
        for (int j = 0; j < NS; j++) cls[j] = 0;
        if (WITH_CONST) {
            if (leaf1 && QFAST) {
                sv_col0_q<NS>(q1, x, cy, cz0);
#pragma unroll
                for (int j = 0; j < NS; j++) cls[j] = xs[j];
            } else if (leaf1) {
                // leaf: only the constant genotype's term survives (:1432-1436); its scale is the leaf's
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const double mm = M1[4 * p + cg];
#pragma unroll
                    for (int j = 0; j < NS; j++) cy[j][p] = mm * (GENO0 ? x[j][0] : sv_sel(x[j], cg));
                }
#pragma unroll
                for (int j = 0; j < NS; j++) cls[j] = xs[j];
            } else {
                if (QFAST) sv_matvec_q<NS>(q1, cx, cy, cz0); else sv_matvec_n<NS>(M1, cx, cy);
#pragma unroll
                for (int j = 0; j < NS; j++) cls[j] = cxs[j];
            }
            if (!unary) {
                double cz[NS][4];
                if (leaf2 && QFAST) {
                    sv_col0_q<NS>(q2, w, cz, cz0);
#pragma unroll
                    for (int j = 0; j < NS; j++) cls[j] += ws[j];
                } else if (leaf2) {
#pragma unroll
                    for (int p = 0; p < 4; p++) {
                        const double mm = M2[4 * p + cg];
#pragma unroll
                        for (int j = 0; j < NS; j++) cz[j][p] = mm * (GENO0 ? w[j][0] : sv_sel(w[j], cg));
                    }
#pragma unroll
                    for (int j = 0; j < NS; j++) cls[j] += ws[j];
                } else {
                    if (QFAST) sv_matvec_q<NS>(q2, cw, cz, cz0); else sv_matvec_n<NS>(M2, cw, cz);
#pragma unroll
                    for (int j = 0; j < NS; j++) cls[j] += cws[j];
                }
#pragma unroll
                for (int j = 0; j < NS; j++)
#pragma unroll
                    for (int p = 0; p < 4; p++) cy[j][p] *= cz[j][p];
            }
            if (a.linear_const_clamp) {
#pragma unroll
                for (int j = 0; j < NS; j++)
#pragma unroll
                    for (int p = 0; p < 4; p++) cy[j][p] = cy[j][p] > 0.0 ? cy[j][p] : SV_MIN_VALUE;
            }
        }
        // ---- the result's exponent; where the host asks for it, renormalise each site's 4 x 4 values by a power of two
        //      (integer max over the high words) ----
        if (flags & SV_OP_RENORM) {
#pragma unroll
            for (int j = 0; j < NS; j++) {
                int hi = sv_maxhi4(y[j]);
                hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, 1));
                hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, 2));
                int sh;
                const double f = sv_pow2_rescale_hi_int(hi, &sh);
#pragma unroll
                for (int p = 0; p < 4; p++) x[j][p] = y[j][p] * f;
                xs[j] = (xs[j] + ws[j]) + sh;
                if (WITH_CONST) {
                    int chi = sv_maxhi4(cy[j]);
                    chi = max(chi, __shfl_xor_sync(0xffffffffu, chi, 1));
                    chi = max(chi, __shfl_xor_sync(0xffffffffu, chi, 2));
                    int csh;
                    const double cf = sv_pow2_rescale_hi_int(chi, &csh);
#pragma unroll
                    for (int p = 0; p < 4; p++) cx[j][p] = cy[j][p] * cf;
                    cxs[j] = cls[j] + csh;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < NS; j++) {
#pragma unroll
                for (int p = 0; p < 4; p++) x[j][p] = y[j][p];
                xs[j] = xs[j] + ws[j];
                if (WITH_CONST) {
#pragma unroll
                    for (int p = 0; p < 4; p++) cx[j][p] = cy[j][p];
                    cxs[j] = cls[j];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NS; j++) {
            // clamped duplicate lanes never store: with reused scratch slots (streamed mode) a duplicate in another warp
            // could otherwise overwrite a slot that the real lane of that site has not consumed yet
            if ((flags & SV_OP_NOSTORE) || !site_ok[j]) continue;
            const size_t dst_row = dst_base + srow[j];
            SV_BOUND(dst_row < a.part_rows);
            sv_st256(a.part + dst_row * 16 + k * 4, x[j]);
            // every category lane writes the (identical) scale: a lane only ever re-reads what it wrote itself
            a.scale[dst_row] = xs[j];
            if (WITH_CONST) {
                sv_st256(a.cpart + dst_row * 16 + k * 4, cx[j]);
                a.cscale[dst_row] = cxs[j];
            }
        }

        // ---- root: integrate over categories and take the log (:153-218, :255-269) ----
        if (flags & SV_OP_ROOT) {
#pragma unroll
            for (int j = 0; j < NS; j++) {
                double v = (klane < a.K ? a.prop[klane] : 0.0) * (GENO0 ? x[j][0] : sv_sel(x[j], rg));
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                // the scales of ALL leaves, which the walk never carried, come in here
                const long long os = a.out_offset + (long long)srow[j];  // srow = the site itself wherever site_ok
                const double lsum = site_ok[j] ? __ldg(a.lsum + os) : 0.0;
                if (site_ok[j] && klane == 0) a.site_logl[os] = log(v) + fma((double)xs[j], SV_LN2, lsum);
                if (WITH_CONST) {
                    double cv = (klane < a.K ? a.prop[klane] : 0.0) * (GENO0 ? cx[j][0] : sv_sel(cx[j], rg));
                    cv += __shfl_xor_sync(0xffffffffu, cv, 1);
                    cv += __shfl_xor_sync(0xffffffffu, cv, 2);
                    if (site_ok[j] && klane == 0) a.site_const[os] = log(cv) + fma((double)cxs[j], SV_LN2, lsum);
                }
            }
        }
    }
}

// ---- BeerLikelihoodCore.getNodePartials in the reference layout [K][S][G] ------------------------------------------
// scale: the node's own binary exponents, NULL for a leaf; lsum: the sum of the scales of the leaves below the node.
// One thread per site: the site's K x 4 values are renormalised here, so that the exported bits do not depend on where
// the walk happened to renormalise (which differs between the dense, sparse and streamed schedules).
// underflow_log (leaves in log mode, or NULL): the leaf's log-space genotype likelihoods [S][4] ([S][K][4] when
// underflow_per_cat), taken where the scaled linear value has underflowed to 0 (a genotype more than 2^-1000 below the best
// one of its (cell, site)): the reference holds a finite log there (RawReadCountsModelFiniteMuDM.java:121-260)
__global__ void k_export_partials(const double* __restrict__ x, const int* __restrict__ scale,
                                  const double* __restrict__ lsum, int S, int K, int is_leaf, int as_log,
                                  int clamp_min, double* __restrict__ out, const double* __restrict__ underflow_log = nullptr,
                                  int underflow_per_cat = 0) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double v[SV_KMAX][4];
    int hi = 0;
    for (int k = 0; k < K; k++)
        for (int g = 0; g < 4; g++) {
            v[k][g] = is_leaf ? x[(size_t)s * 4 + g] : x[((size_t)s * SV_KMAX + k) * 4 + g];
            hi = max(hi, __double2hiint(v[k][g]));
        }
    int e;
    const double f = sv_pow2_rescale_hi_int(hi, &e);
    const double ls = fma((double)((scale ? scale[s] : 0) + e), SV_LN2, lsum[s]);
    for (int k = 0; k < K; k++)
        for (int g = 0; g < 4; g++) {
            const double vn = v[k][g] * f;
            const double lin = vn * exp(ls);
            // clamp_min: linear-mode constant-site partials, clamped like ScsBeerLikelihoodCore.java:1041-1047
            double r = as_log ? log(vn) + ls : (clamp_min && !(lin > 0.0) ? SV_MIN_VALUE : lin);
            if (underflow_log && as_log && vn == 0.0)
                r = underflow_log[underflow_per_cat ? ((size_t)s * K + k) * 4 + g : (size_t)s * 4 + g];
            out[((size_t)k * S + s) * 4 + g] = r;
        }
}
