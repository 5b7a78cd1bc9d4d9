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
    double* part;          // [slots][S][4][4]   (category stride is always 4; lanes k >= K idle)
    double* scale;         // [slots][S]
    double* cpart;         // constant-site twins (NULL when not maintained)
    double* cscale;
    const double* leafX;   // [leaf slots][S][4]
    const double* leafS;   // [leaf slots][S]
    const double* mats;    // [matrix slots][4][16], already clamped when SIEVE_FLAG_LOG_PARTIALS
    const SvOp* ops;
    int op_begin, op_end;  // walk: ops [op_begin, op_end); per-level: op = op_begin + blockIdx.y
    int per_level;
    int S, K;
    int linear_const_clamp;  // !SIEVE_FLAG_LOG_PARTIALS: clamp of the constant-site product (:1041-1047)
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

#define SV_PRUNE_THREADS 128
#define SV_PRUNE_SITES (SV_PRUNE_THREADS / 4)
#ifndef SV_PRUNE_MINB_CONST
#define SV_PRUNE_MINB_CONST 5
#endif

// GENO0: root genotype == constant genotype == 0 (always so for ScsFiniteMuExtendedModel.java:37-40); the general
// case takes the same kernel with run-time selects.
template <bool WITH_CONST, bool GENO0>
__global__ void __launch_bounds__(SV_PRUNE_THREADS, WITH_CONST ? SV_PRUNE_MINB_CONST : 8) k_prune(SvPruneArgs a) {
    // warp-private copies of the op's two 4 x 4 x 4 matrices: no block-wide barrier anywhere in the walk, so the
    // warps of an SM drift apart and overlap each other's memory latency
    __shared__ __align__(32) double sm[(SV_PRUNE_THREADS / 32) * 2 * SV_KMAX * SV_MAT_STRIDE];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int k = lane & 3;
    const int site = blockIdx.x * SV_PRUNE_SITES + (tid >> 2);
    const bool site_ok = site < a.S;
    const bool active = site_ok && (k < a.K);
    const long long srow = site_ok ? site : 0;
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
    double x[4] = {0, 0, 0, 0}, cx[4] = {0, 0, 0, 0}, xs = 0.0, cxs = 0.0;

    for (int i = i0; i < i1; i++) {
        const int4* opp = reinterpret_cast<const int4*>(a.ops + i);
        const int4 o0 = __ldg(opp), o1 = __ldg(opp + 1), o2 = __ldg(opp + 2);
        const long long dst_row = ((long long)(unsigned)o0.x | ((long long)o0.y << 32)) + srow;
        const long long c1_row = ((long long)(unsigned)o0.z | ((long long)o0.w << 32)) + srow;
        const long long c2_row = ((long long)(unsigned)o1.x | ((long long)o1.y << 32)) + srow;
        const int m1 = o1.z, m2 = o1.w, flags = o2.x;
        const bool unary = flags & SV_OP_UNARY;
        const bool leaf1 = flags & SV_OP_LEAF1, leaf2 = flags & SV_OP_LEAF2;

        // ---- issue every global load of this op up front: matrices (one 256-bit load per lane) and children ----
        double m[4] = {0, 0, 0, 0};
        if (!(st_which == 1 && unary))
            sv_ld256_nc(a.mats + (size_t)(st_which ? m2 : m1) * (SV_KMAX * 16) + st_off, m);
        double w[4] = {0, 0, 0, 0}, ws = 0.0;      // child 2
        double cw[4] = {0, 0, 0, 0}, cws = 0.0;    // child 2, constant-site twin
        if (active) {
            if (!(flags & SV_OP_FWD1)) {
                if (leaf1) {
                    sv_ld256_nc(a.leafX + c1_row * 4, x);
                    xs = __ldg(a.leafS + c1_row);
                } else {
                    sv_ld256(a.part + c1_row * 16 + k * 4, x);
                    xs = a.scale[c1_row];
                    if (WITH_CONST) {
                        sv_ld256(a.cpart + c1_row * 16 + k * 4, cx);
                        cxs = a.cscale[c1_row];
                    }
                }
            }
            if (!unary) {
                if (leaf2) {
                    sv_ld256_nc(a.leafX + c2_row * 4, w);
                    ws = __ldg(a.leafS + c2_row);
                } else {
                    sv_ld256(a.part + c2_row * 16 + k * 4, w);
                    ws = a.scale[c2_row];
                    if (WITH_CONST) {
                        sv_ld256(a.cpart + c2_row * 16 + k * 4, cw);
                        cws = a.cscale[c2_row];
                    }
                }
            }
        }
        __syncwarp();  // every lane is done with the previous op's matrices
        *reinterpret_cast<double2*>(st_dst) = make_double2(m[0], m[1]);
        *reinterpret_cast<double2*>(st_dst + 2) = make_double2(m[2], m[3]);
        __syncwarp();

        // ================= variant partials =================
        double y[4], z[4];
        sv_matvec(M1, x, y);
        if (!unary) {
            sv_matvec(M2, w, z);
#pragma unroll
            for (int p = 0; p < 4; p++) y[p] *= z[p];
        }
        // ================= constant-site partials =================
        double cy[4];
        double cls = 0.0;
        if (WITH_CONST) {
            if (leaf1) {
                // leaf: only the constant genotype's term survives (:1432-1436); its scale is the leaf's
                const double xc = GENO0 ? x[0] : sv_sel(x, cg);
#pragma unroll
                for (int p = 0; p < 4; p++) cy[p] = M1[4 * p + cg] * xc;
                cls = xs;
            } else {
                sv_matvec(M1, cx, cy);
                cls = cxs;
            }
            if (!unary) {
                double cz[4];
                if (leaf2) {
                    const double wc = GENO0 ? w[0] : sv_sel(w, cg);
#pragma unroll
                    for (int p = 0; p < 4; p++) cz[p] = M2[4 * p + cg] * wc;
                    cls += ws;
                } else {
                    sv_matvec(M2, cw, cz);
                    cls += cws;
                }
#pragma unroll
                for (int p = 0; p < 4; p++) cy[p] *= cz[p];
            }
            if (a.linear_const_clamp) {
#pragma unroll
                for (int p = 0; p < 4; p++) cy[p] = cy[p] > 0.0 ? cy[p] : (active ? SV_MIN_VALUE : 0.0);
            }
        }
        // ---- renormalise the site's 4 x 4 values by a power of two (integer max over the high words) ----
        {
            int hi = sv_maxhi4(y);
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, 1));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, 2));
            double sh;
            const double f = sv_pow2_rescale_hi(hi, &sh);
#pragma unroll
            for (int p = 0; p < 4; p++) x[p] = y[p] * f;
            xs = (xs + ws) + sh;
        }
        if (WITH_CONST) {
            int hi = sv_maxhi4(cy);
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, 1));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, 2));
            double sh;
            const double f = sv_pow2_rescale_hi(hi, &sh);
#pragma unroll
            for (int p = 0; p < 4; p++) cx[p] = cy[p] * f;
            cxs = cls + sh;
        }

        if (active) {
            sv_st256(a.part + dst_row * 16 + k * 4, x);
            // every category lane writes the (identical) scale: a lane only ever re-reads what it wrote itself
            a.scale[dst_row] = xs;
            if (WITH_CONST) {
                sv_st256(a.cpart + dst_row * 16 + k * 4, cx);
                a.cscale[dst_row] = cxs;
            }
        }

        // ---- root: integrate over categories and take the log (:153-218, :255-269) ----
        if (flags & SV_OP_ROOT) {
            double v = active ? a.prop[k] * (GENO0 ? x[0] : sv_sel(x, rg)) : 0.0;
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (site_ok && k == 0) a.site_logl[site] = log(v) + xs;
            if (WITH_CONST) {
                double cv = active ? a.prop[k] * (GENO0 ? cx[0] : sv_sel(cx, rg)) : 0.0;
                cv += __shfl_xor_sync(0xffffffffu, cv, 1);
                cv += __shfl_xor_sync(0xffffffffu, cv, 2);
                if (site_ok && k == 0) a.site_const[site] = log(cv) + cxs;
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
    const double v = is_leaf ? x[(size_t)s * 4 + g] : x[((size_t)s * SV_KMAX + k) * 4 + g];
    const double ls = scale[s];
    out[idx] = as_log ? log(v) + ls : v * exp(ls);
}
