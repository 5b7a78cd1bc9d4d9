// prune.cuh -- Felsenstein pruning over dirty nodes, root integration and per-site log-likelihood.
//
// Replaces ScsBeerLikelihoodCore.calculateLogPartials / calculatePartials -> calculate(Log)PartialPartialPruning
// (likelihood/ScsBeerLikelihoodCore.java:1283-1501 / :857-1055), integratePartials (:153-218) and
// calculateLogLikelihoods (:255-269).
//
// Arithmetic: scaled linear space.  A node's partial for (site, category k) is 4 doubles x[g] and ONE natural-log
// scale per (node, site) shared by all categories and states: true value = x[g] * exp(lscale).  After every product the
// 16 values of a site are renormalised by an exact power of two (sv_pow2_rescale), so no transcendental is evaluated
// until the root's log.  The reference's log-space result is reproduced to rounding (~1e-16 relative on the
// likelihood); its clamp log(P > 0 ? P : Double.MIN_VALUE) (:1430) is mirrored by clamping the matrices when they are
// staged (SIEVE_FLAG_LOG_PARTIALS), its linear-mode clamp of the constant-site product (:1041-1047) likewise.
//
// Layout in HBM: partials [slot][site][K][4] (one site = 128 contiguous bytes at K = 4), scales [slot][site];
// leaves [slot][site][4] + [slot][site] (category independent).  Thread mapping: lane = 4 * site_in_warp + k, so a warp
// reads/writes 8 sites * 128 B = 1 KB contiguous with one 256-bit access per lane.
//
// Sites are independent, so one launch can walk ANY number of dependent ops: each thread only ever re-reads what it
// (or, for the scales, a lane of its own block before a barrier) wrote.  The same kernel serves the level-batched
// schedule (grid.y = ops of one level, one op per block) and the single-launch walk (grid.y = 1, loop over all ops).
#pragma once
#include "common.cuh"

struct SvPruneArgs {
    double* part;          // [slots][S][K][4]
    double* scale;         // [slots][S]
    double* cpart;         // constant-site twins (NULL when not maintained)
    double* cscale;
    const double* leafX;   // [leaf slots][S][4]
    const double* leafS;   // [leaf slots][S]
    const double* mats;    // [matrix slots][K][16]
    const SvOp* ops;
    int op_begin, op_end;  // walk: ops [op_begin, op_end); per-level: op = op_begin + blockIdx.y
    int per_level;
    int S, K;
    int clamp_matrices;    // SIEVE_FLAG_LOG_PARTIALS
    int const_genotype;
    int root_genotype;
    double prop[4];        // category proportions at the root
    double* site_logl;     // [S]
    double* site_const;    // [S]
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

// One child's contribution for one (site, category) lane: y = M x.  Either forwarded from registers (the child is the
// op this warp finished last), or loaded: a leaf (category independent, read-only path) or an internal node.
#define SV_PRUNE_THREADS 128
#define SV_PRUNE_SITES (SV_PRUNE_THREADS / 4)

template <bool WITH_CONST>
__global__ void __launch_bounds__(SV_PRUNE_THREADS, WITH_CONST ? 6 : 8) k_prune(SvPruneArgs a) {
    // warp-private copies of the op's two K x 4 x 4 matrices: no block-wide barrier anywhere in the walk, so the
    // warps of an SM drift apart and overlap each other's memory latency
    __shared__ __align__(32) double sm[(SV_PRUNE_THREADS / 32) * 2 * SV_KMAX * SV_MAT_STRIDE];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int k = lane & 3;
    const int site = blockIdx.x * SV_PRUNE_SITES + (tid >> 2);
    const int K = a.K;
    const size_t S = (size_t)a.S;
    const bool site_ok = site < a.S;
    const bool active = site_ok && (k < K);
    double* wm = sm + warp * (2 * SV_KMAX * SV_MAT_STRIDE);
    // staging role of this lane: matrix `which` (child 1 / 2), category kk, row q  ->  4 doubles = one 256-bit load
    const int st_which = lane >> 4, st_kk = (lane >> 2) & 3, st_q = lane & 3;
    double* st_dst = wm + st_which * (SV_KMAX * SV_MAT_STRIDE) + st_kk * SV_MAT_STRIDE + st_q * 4;
    const double* M1 = wm + k * SV_MAT_STRIDE;
    const double* M2 = wm + SV_KMAX * SV_MAT_STRIDE + k * SV_MAT_STRIDE;
    const size_t tk = ((size_t)(site_ok ? site : 0) * K + (k < K ? k : 0)) * 4;  // offset of this lane inside a node's block
    const size_t node_stride = S * K * 4;

    int i0 = a.op_begin, i1 = a.op_end;
    if (a.per_level) {
        i0 = a.op_begin + blockIdx.y;
        i1 = i0 + 1;
    }
    // what this lane produced last: forwarded in registers when the next op consumes it (in a post-order walk at least
    // one child of every op with an internal child is the op finished just before)
    double fy[4] = {0, 0, 0, 0}, fcy[4] = {0, 0, 0, 0}, fls = 0.0, fcls = 0.0;
    int fdst = -1;

    for (int i = i0; i < i1; i++) {
        const SvOp op = a.ops[i];
        const bool unary = op.flags & SV_OP_UNARY;
        const bool leaf1 = op.flags & SV_OP_LEAF1, leaf2 = op.flags & SV_OP_LEAF2;
        // ---- stage both matrices of the op: one 256-bit load per lane ----
        {
            double m[4] = {0, 0, 0, 0};
            if (st_kk < K && !(st_which == 1 && unary))
                sv_ld256_nc(a.mats + (size_t)(st_which ? op.m2 : op.m1) * K * 16 + st_kk * 16 + st_q * 4, m);
            if (a.clamp_matrices) {
#pragma unroll
                for (int e = 0; e < 4; e++) m[e] = m[e] > 0.0 ? m[e] : SV_MIN_VALUE;
            }
            __syncwarp();  // every lane is done with the previous op's matrices
            *reinterpret_cast<double2*>(st_dst) = make_double2(m[0], m[1]);
            *reinterpret_cast<double2*>(st_dst + 2) = make_double2(m[2], m[3]);
            __syncwarp();
        }
        const bool fwd1 = !leaf1 && op.c1 == fdst;
        const bool fwd2 = !unary && !leaf2 && op.c2 == fdst;

        double y[4], ls;
        double cy[4], cls = 0.0;
        double xc1 = 0.0, xc2 = 0.0;  // the leaf children's constant-genotype entries
        // ================= variant partials =================
        {
            double x[4] = {0, 0, 0, 0}, s1 = 0.0, s2 = 0.0;
            if (fwd1) {
#pragma unroll
                for (int g = 0; g < 4; g++) x[g] = fy[g];
                s1 = fls;
            } else if (active) {
                if (leaf1) {
                    sv_ld256_nc(a.leafX + ((size_t)op.c1 * S + site) * 4, x);
                    s1 = __ldg(a.leafS + (size_t)op.c1 * S + site);
                } else {
                    sv_ld256(a.part + (size_t)op.c1 * node_stride + tk, x);
                    s1 = a.scale[(size_t)op.c1 * S + site];
                }
            }
            if (WITH_CONST) xc1 = sv_sel(x, a.const_genotype);
            sv_matvec(M1, x, y);
            if (!unary) {
                double w[4] = {0, 0, 0, 0}, z[4];
                if (fwd2) {
#pragma unroll
                    for (int g = 0; g < 4; g++) w[g] = fy[g];
                    s2 = fls;
                } else if (active) {
                    if (leaf2) {
                        sv_ld256_nc(a.leafX + ((size_t)op.c2 * S + site) * 4, w);
                        s2 = __ldg(a.leafS + (size_t)op.c2 * S + site);
                    } else {
                        sv_ld256(a.part + (size_t)op.c2 * node_stride + tk, w);
                        s2 = a.scale[(size_t)op.c2 * S + site];
                    }
                }
                if (WITH_CONST) xc2 = sv_sel(w, a.const_genotype);
                sv_matvec(M2, w, z);
#pragma unroll
                for (int p = 0; p < 4; p++) y[p] *= z[p];
            }
            // renormalise the site's K x 4 values by a power of two
            double mx = sv_max4(y);
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            double sh;
            const double f = sv_pow2_rescale(mx, &sh);
#pragma unroll
            for (int p = 0; p < 4; p++) y[p] *= f;
            ls = (s1 + s2) + sh;
            if (WITH_CONST) {  // leaf children: scale of the constant twin == the leaf's scale
                if (leaf1) cls = s1;
                if (!unary && leaf2) cls += s2;
            }
        }
        // ================= constant-site partials =================
        if (WITH_CONST) {
            double cs = 0.0;
            if (leaf1) {
                // leaf: only the constant genotype's term survives (:1432-1436)
#pragma unroll
                for (int p = 0; p < 4; p++) cy[p] = M1[4 * p + a.const_genotype] * xc1;
            } else {
                double c[4] = {0, 0, 0, 0};
                if (fwd1) {
#pragma unroll
                    for (int g = 0; g < 4; g++) c[g] = fcy[g];
                    cs = fcls;
                } else if (active) {
                    sv_ld256(a.cpart + (size_t)op.c1 * node_stride + tk, c);
                    cs = a.cscale[(size_t)op.c1 * S + site];
                }
                sv_matvec(M1, c, cy);
                cls += cs;
            }
            if (!unary) {
                double cz[4];
                if (leaf2) {
#pragma unroll
                    for (int p = 0; p < 4; p++) cz[p] = M2[4 * p + a.const_genotype] * xc2;
                } else {
                    double c[4] = {0, 0, 0, 0};
                    cs = 0.0;
                    if (fwd2) {
#pragma unroll
                        for (int g = 0; g < 4; g++) c[g] = fcy[g];
                        cs = fcls;
                    } else if (active) {
                        sv_ld256(a.cpart + (size_t)op.c2 * node_stride + tk, c);
                        cs = a.cscale[(size_t)op.c2 * S + site];
                    }
                    sv_matvec(M2, c, cz);
                    cls += cs;
                }
#pragma unroll
                for (int p = 0; p < 4; p++) cy[p] *= cz[p];
            }
            if (!a.clamp_matrices) {
                // linear-mode clamp of the constant-site product (:1041-1047)
#pragma unroll
                for (int p = 0; p < 4; p++) cy[p] = cy[p] > 0.0 ? cy[p] : (active ? SV_MIN_VALUE : 0.0);
            }
            double cmx = sv_max4(cy);
            cmx = fmax(cmx, __shfl_xor_sync(0xffffffffu, cmx, 1));
            cmx = fmax(cmx, __shfl_xor_sync(0xffffffffu, cmx, 2));
            double csh;
            const double cf = sv_pow2_rescale(cmx, &csh);
#pragma unroll
            for (int p = 0; p < 4; p++) cy[p] *= cf;
            cls += csh;
        }

        if (active) {
            sv_st256(a.part + (size_t)op.dst * node_stride + tk, y);
            // every category lane writes the (identical) scale: a lane only ever re-reads what it wrote itself
            a.scale[(size_t)op.dst * S + site] = ls;
            if (WITH_CONST) {
                sv_st256(a.cpart + (size_t)op.dst * node_stride + tk, cy);
                a.cscale[(size_t)op.dst * S + site] = cls;
            }
        }
#pragma unroll
        for (int g = 0; g < 4; g++) {
            fy[g] = y[g];
            if (WITH_CONST) fcy[g] = cy[g];
        }
        fls = ls;
        fcls = cls;
        fdst = op.dst;

        // ---- root: integrate over categories and take the log (:153-218, :255-269) ----
        if (op.flags & SV_OP_ROOT) {
            double v = active ? a.prop[k] * sv_sel(y, a.root_genotype) : 0.0;
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (site_ok && k == 0) a.site_logl[site] = log(v) + ls;
            if (WITH_CONST) {
                double cv = active ? a.prop[k] * sv_sel(cy, a.root_genotype) : 0.0;
                cv += __shfl_xor_sync(0xffffffffu, cv, 1);
                cv += __shfl_xor_sync(0xffffffffu, cv, 2);
                if (site_ok && k == 0) a.site_const[site] = log(cv) + cls;
            }
        }
    }
}

// ---- BeerLikelihoodCore.getNodePartials in the reference layout [K][S][G] ------------------------------------------
__global__ void k_export_partials(const double* __restrict__ x, const double* __restrict__ scale, int S, int K,
                                  int is_leaf, int as_log, double* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over K * S * 4
    if (idx >= K * S * 4) return;
    const int g = idx & 3;
    const int s = (idx >> 2) % S;
    const int k = (idx >> 2) / S;
    const double v = is_leaf ? x[(size_t)s * 4 + g] : x[((size_t)s * K + k) * 4 + g];
    const double ls = scale[s];
    out[idx] = as_log ? log(v) + ls : v * exp(ls);
}
