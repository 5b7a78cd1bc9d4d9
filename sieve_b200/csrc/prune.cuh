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

template <bool WITH_CONST>
__global__ void __launch_bounds__(256) k_prune(SvPruneArgs a) {
    __shared__ __align__(16) double sm[2 * SV_KMAX * SV_MAT_STRIDE];
    const int tid = threadIdx.x;
    const int k = tid & 3;
    const int site = blockIdx.x * 64 + (tid >> 2);
    const int K = a.K;
    const size_t S = (size_t)a.S;
    const bool active = (site < a.S) && (k < K);
    const bool site_ok = site < a.S;

    int i0 = a.op_begin, i1 = a.op_end;
    if (a.per_level) {
        i0 = a.op_begin + blockIdx.y;
        i1 = i0 + 1;
    }
    for (int i = i0; i < i1; i++) {
        const SvOp op = a.ops[i];
        const bool unary = op.flags & SV_OP_UNARY;
        // ---- stage the (up to) two K x 4 x 4 transition matrices of this op in shared memory ----
        __syncthreads();  // previous op's readers are done; its global writes are visible block-wide
        if (tid < 128) {
            const int which = tid >> 6, idx = tid & 63, kk = idx >> 4, e = idx & 15;
            double v = 0.0;
            if (kk < K && !(which == 1 && unary)) {
                v = a.mats[(size_t)(which ? op.m2 : op.m1) * K * 16 + kk * 16 + e];
                if (a.clamp_matrices) v = v > 0.0 ? v : SV_MIN_VALUE;
            }
            sm[which * (SV_KMAX * SV_MAT_STRIDE) + kk * SV_MAT_STRIDE + e] = v;
        }
        __syncthreads();
        const double* M1 = sm + k * SV_MAT_STRIDE;
        const double* M2 = sm + SV_KMAX * SV_MAT_STRIDE + k * SV_MAT_STRIDE;

        double x1[4] = {0, 0, 0, 0}, x2[4] = {0, 0, 0, 0}, s1 = 0.0, s2 = 0.0;
        double y[4], z[4];
        double cy[4], cz[4], cs1 = 0.0, cs2 = 0.0;

        // ---- child 1 ----
        if (active) {
            if (op.flags & SV_OP_LEAF1) {
                sv_ld256_nc(a.leafX + ((size_t)op.c1 * S + site) * 4, x1);
                s1 = __ldg(a.leafS + (size_t)op.c1 * S + site);
            } else {
                sv_ld256(a.part + (((size_t)op.c1 * S + site) * K + k) * 4, x1);
                s1 = a.scale[(size_t)op.c1 * S + site];
            }
        }
        sv_matvec(M1, x1, y);
        if (WITH_CONST) {
            if (op.flags & SV_OP_LEAF1) {
                // leaf: only the constant genotype's term survives (:1432-1436)
                const double xc = sv_sel(x1, a.const_genotype);
#pragma unroll
                for (int p = 0; p < 4; p++) cy[p] = M1[4 * p + a.const_genotype] * xc;
                cs1 = s1;
            } else {
                double c1[4] = {0, 0, 0, 0};
                if (active) {
                    sv_ld256(a.cpart + (((size_t)op.c1 * S + site) * K + k) * 4, c1);
                    cs1 = a.cscale[(size_t)op.c1 * S + site];
                }
                sv_matvec(M1, c1, cy);
            }
        }
        // ---- child 2 ----
        if (!unary) {
            if (active) {
                if (op.flags & SV_OP_LEAF2) {
                    sv_ld256_nc(a.leafX + ((size_t)op.c2 * S + site) * 4, x2);
                    s2 = __ldg(a.leafS + (size_t)op.c2 * S + site);
                } else {
                    sv_ld256(a.part + (((size_t)op.c2 * S + site) * K + k) * 4, x2);
                    s2 = a.scale[(size_t)op.c2 * S + site];
                }
            }
            sv_matvec(M2, x2, z);
#pragma unroll
            for (int p = 0; p < 4; p++) y[p] *= z[p];
            if (WITH_CONST) {
                if (op.flags & SV_OP_LEAF2) {
                    const double xc = sv_sel(x2, a.const_genotype);
#pragma unroll
                    for (int p = 0; p < 4; p++) cz[p] = M2[4 * p + a.const_genotype] * xc;
                    cs2 = s2;
                } else {
                    double c2[4] = {0, 0, 0, 0};
                    if (active) {
                        sv_ld256(a.cpart + (((size_t)op.c2 * S + site) * K + k) * 4, c2);
                        cs2 = a.cscale[(size_t)op.c2 * S + site];
                    }
                    sv_matvec(M2, c2, cz);
                }
#pragma unroll
                for (int p = 0; p < 4; p++) cy[p] *= cz[p];
            }
        }
        if (WITH_CONST && !a.clamp_matrices) {
            // linear-mode clamp of the constant-site product (:1041-1047)
#pragma unroll
            for (int p = 0; p < 4; p++) cy[p] = cy[p] > 0.0 ? cy[p] : (active ? SV_MIN_VALUE : 0.0);
        }

        // ---- renormalise the site's K x 4 values by a power of two ----
        double mx = sv_max4(y);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        double sh;
        const double f = sv_pow2_rescale(mx, &sh);
#pragma unroll
        for (int p = 0; p < 4; p++) y[p] *= f;
        const double ls = (s1 + s2) + sh;
        double cls = 0.0;
        if (WITH_CONST) {
            double cmx = sv_max4(cy);
            cmx = fmax(cmx, __shfl_xor_sync(0xffffffffu, cmx, 1));
            cmx = fmax(cmx, __shfl_xor_sync(0xffffffffu, cmx, 2));
            double csh;
            const double cf = sv_pow2_rescale(cmx, &csh);
#pragma unroll
            for (int p = 0; p < 4; p++) cy[p] *= cf;
            cls = (cs1 + cs2) + csh;
        }

        if (active) {
            sv_st256(a.part + (((size_t)op.dst * S + site) * K + k) * 4, y);
            if (WITH_CONST) sv_st256(a.cpart + (((size_t)op.dst * S + site) * K + k) * 4, cy);
            if (k == 0) {
                a.scale[(size_t)op.dst * S + site] = ls;
                if (WITH_CONST) a.cscale[(size_t)op.dst * S + site] = cls;
            }
        }

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
