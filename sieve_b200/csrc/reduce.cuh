// reduce.cuh -- per-shard totals: sum of weighted site log-likelihoods and the constant-site aggregates that the
// acquisition-bias correction needs.
//
// Replaces ScsTreeLikelihood.calcLogP (likelihood/ScsTreeLikelihood.java:2621-2651) and the per-worker part of
// ScsAlignment.getAscBiasCorrection* (alignment/ScsAlignment.java:746-782).  The reference sums sites left to right;
// this is a fixed-shape tree (thread-strided -> warp shuffle -> block -> blocks in index order), deterministic run to
// run, and differs from the sequential sum only in the last bits (budget: 1e-9 relative on the total).
#pragma once
#include "common.cuh"

struct SvAcc {
    double sum;    // sum w * logL
    double m, se;  // running max and sum w * exp(c - m) of the constant-site logs
    double csum;   // sum w * c
    double lewis;  // sum w * log(1 - exp(c))
    double wsum;   // sum w
};

__device__ __forceinline__ void sv_lse_merge(double& m, double& se, double m2, double se2) {
    const double mm = fmax(m, m2);
    const double a = (m == mm) ? se : se * exp(m - mm);
    const double b = (m2 == mm) ? se2 : se2 * exp(m2 - mm);
    m = mm;
    se = a + b;
}

__device__ __forceinline__ void sv_acc_merge(SvAcc& a, const SvAcc& b) {
    a.sum += b.sum;
    sv_lse_merge(a.m, a.se, b.m, b.se);
    a.csum += b.csum;
    a.lewis += b.lewis;
    a.wsum += b.wsum;
}

__device__ __forceinline__ SvAcc sv_acc_shfl_down(const SvAcc& a, int d) {
    SvAcc b;
    b.sum = __shfl_down_sync(0xffffffffu, a.sum, d);
    b.m = __shfl_down_sync(0xffffffffu, a.m, d);
    b.se = __shfl_down_sync(0xffffffffu, a.se, d);
    b.csum = __shfl_down_sync(0xffffffffu, a.csum, d);
    b.lewis = __shfl_down_sync(0xffffffffu, a.lewis, d);
    b.wsum = __shfl_down_sync(0xffffffffu, a.wsum, d);
    return b;
}

__device__ __forceinline__ SvAcc sv_acc_zero() {
    SvAcc a;
    a.sum = 0.0;
    a.m = __longlong_as_double(0xfff0000000000000LL);
    a.se = 0.0;
    a.csum = 0.0;
    a.lewis = 0.0;
    a.wsum = 0.0;
    return a;
}

// out layout == sieve_shard_result_t {sum_log_l, const_lse, const_sum, lewis_sum, weight_sum}
// The constant-site log-likelihood of site s is  cst[s] + cst_shift:  with the algebraic constant-site path cst is
// Lambda (sum over cells of the leaves' constant-genotype log-likelihoods) and cst_shift = log C_tree (one number per
// tree, computed on the host); with the per-site twin cst is logConstRoot itself and cst_shift = 0.
__global__ void __launch_bounds__(256) k_reduce(const double* __restrict__ logl, const double* __restrict__ cst,
                                                double cst_shift, const int* __restrict__ w, int S, int with_const,
                                                SvAcc* __restrict__ partials, unsigned int* __restrict__ counter,
                                                double* __restrict__ out) {
    __shared__ SvAcc sm[8];
    __shared__ bool last;
    SvAcc acc = sv_acc_zero();
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
        const double wt = w ? (double)w[s] : 1.0;
        acc.sum += wt * logl[s];
        acc.wsum += wt;
        if (with_const) {
            const double c = cst[s] + cst_shift;
            sv_lse_merge(acc.m, acc.se, c, wt);
            acc.csum += wt * c;
            acc.lewis += wt * log(1.0 - exp(c));
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const SvAcc b = sv_acc_shfl_down(acc, d);
        sv_acc_merge(acc, b);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sm[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        SvAcc t = sm[0];
        for (int i = 1; i < (int)(blockDim.x >> 5); i++) sv_acc_merge(t, sm[i]);
        partials[blockIdx.x] = t;
        __threadfence();
        const unsigned int done = atomicAdd(counter, 1u);
        last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        // the last block to finish merges the per-block partials: fixed-shape tree over block indices (deterministic)
        __shared__ SvAcc fin[256];
        __threadfence();
        SvAcc t = sv_acc_zero();
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) sv_acc_merge(t, partials[i]);
        fin[threadIdx.x] = t;
        __syncthreads();
        for (int d = 128; d > 0; d >>= 1) {
            if ((int)threadIdx.x < d) sv_acc_merge(fin[threadIdx.x], fin[threadIdx.x + d]);
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            t = fin[0];
            out[0] = t.sum;
            out[1] = with_const ? log(t.se) + t.m : 0.0;
            out[2] = t.csum;
            out[3] = -t.lewis;
            out[4] = t.wsum;
            *counter = 0u;
        }
    }
}


// The leaf kernel leaves one partial sum per (cell group, site) of (a) the leaves' natural-log scales and (b) their
// constant-genotype log-likelihoods; this adds the groups in index order (deterministic), one thread per site:
//   lsum[s]   = sum over cells of the scale of leaf (cell, s): added to the site's log-likelihood once, at the root
//   lambda[s] = sum over cells of log leaf[cell][s][constant genotype]
// The constant-site likelihood factorises: every constant-site partial is (a site-independent 4-vector per node and
// category) x (the product over the leaves below of their constant-genotype likelihood), so the per-site part of the
// whole acquisition-bias computation is this one column sum, and it only changes when the leaves change.
__global__ void __launch_bounds__(256) k_colsum(const double* __restrict__ sumS, const double* __restrict__ sum0,
                                                int n_groups, int S, long long stride, double* __restrict__ lsum,
                                                double* __restrict__ lambda) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double a = 0.0, b = 0.0;
    for (int g = 0; g < n_groups; g++) {
        a += __ldg(sumS + (size_t)g * stride + s);
        if (sum0) b += __ldg(sum0 + (size_t)g * stride + s);
    }
    lsum[s] = a;
    if (sum0) lambda[s] = b;
}

// Constant-site partials of one node in the reference layout [K][S][G], rebuilt from the factorisation for
// BeerLikelihoodCore-style read-back: out = G[k][g] * exp(gscale) * prod over the node's leaves of their constant-
// genotype likelihood (lam = the column sum over exactly those leaves).
// lam_stride_k: 0 when the column sum is the same for every category, else the distance between the categories' sums
// (per-pattern coverage model, leaf_private.cuh).
__global__ void k_export_const(const double* __restrict__ lam, long long lam_stride_k, const double* __restrict__ G /*[4][4]*/,
                               double gscale, int S, int K, int as_log, double* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    for (int k = 0; k < K; k++)
        for (int g = 0; g < 4; g++) {
            const double l = lam[(size_t)k * lam_stride_k + s];
            const double v = log(G[k * 4 + g]) + gscale + l;
            // linear mode: the reference clamps a constant-site product that underflowed to Double.MIN_VALUE
            // (ScsBeerLikelihoodCore.java:1041-1047); the read-back shows the same
            const double e = exp(v);
            out[((size_t)k * S + s) * 4 + g] = as_log ? v : (e > 0.0 ? e : SV_MIN_VALUE);
        }
}
