// background.cuh -- stage-1 background likelihood (a "next" row).
//
// Replaces ScsBackgroundLikelihood.calculateLogPDirichletMultinomial (likelihood/ScsBackgroundLikelihood.java:169-268):
//   logP = M n (lgG(a4) - sum_{h<4} lgG(a_h)) - sum_c N_c lgG(a4 + c) + sum_{h<4} sum_m N_{h,m} lgG(a_h + m)
// over the five background histograms (variant1, variant2, variant3, normal, coverage) with the wild-type
// Dirichlet-multinomial parameters of NucReadCountsModelFiniteMuDM.getWildTypeNucReadCountModelParams (:99-112).
// A few hundred logGamma per evaluation: one block, one thread per histogram entry, fixed-order tree reduction.
#pragma once
#include "common.cuh"
#include "leaf.cuh"

__global__ void __launch_bounds__(256) k_background(const long long* __restrict__ values,
                                                    const long long* __restrict__ counts,
                                                    const int* __restrict__ off, long long n_background_sites, int n_taxa,
                                                    double eff_seq_err, double shape1, double* __restrict__ out) {
    __shared__ double sm[256];
    double a[5];
    a[0] = a[1] = a[2] = (eff_seq_err / 3) * shape1;
    a[3] = (1 - eff_seq_err) * shape1;
    a[4] = shape1;
    const int total = off[5];
    double acc = 0.0;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        int h = 0;
        while (h < 4 && i >= off[h + 1]) h++;
        const double t = (double)counts[i] * sv_log_gamma(a[h] + (double)values[i]);
        acc += h == 4 ? -t : t;
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) sm[threadIdx.x] += sm[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double part1 = (double)n_background_sites * n_taxa *
                             (sv_log_gamma(a[4]) - sv_log_gamma(a[0]) - sv_log_gamma(a[1]) - sv_log_gamma(a[2]) - sv_log_gamma(a[3]));
        out[0] = part1 + sm[0];
    }
}
