// sizefactors.cuh -- per-cell size factors (DESeq-style median of ratios) on the device, run once at start-up.
//
// Replaces SeqCovModelInterface.Base.computeSizeFactors (rawreadcountsmodel/seqcovmodel/SeqCovModelInterface.java:291-351)
// with MathFunctions.geometricMean (math/util/MathFunctions.java:73-92) and DiscreteStatistics.median/mean/variance
// (BEAST, external).  s_j = median over sites (expanded by pattern weight) of cov_js / geomean_j'(cov_j's), entries with
// zero coverage or a zero geometric mean skipped when zeroCovMode == 0.
// The sort is CUB's segmented radix sort (one-off initialisation, not on the per-step path).
#pragma once
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "common.cuh"

// one thread per site: geometric mean over cells of the (processed) coverage; exact integer moments for :322-323
__global__ void k_sf_geo_mean(const int4* __restrict__ counts, int n_cells, int S, int zero_cov_mode,
                              double* __restrict__ geo, unsigned long long* __restrict__ moments) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long m1 = 0, m2 = 0;
    if (s < S) {
        double sum = 0.0;
        int len = 0;
        for (int j = 0; j < n_cells; j++) {
            int c = __ldg(&counts[(size_t)j * S + s].w);
            if (zero_cov_mode == 1) c += 1;
            if (c > 0) {
                ++len;
                sum += log((double)c);
            }
            m1 += (unsigned long long)c;
            m2 += (unsigned long long)c * (unsigned long long)c;
        }
        geo[s] = len == 0 ? 0.0 : exp(sum / len);
    }
    // block totals -> two atomics per block (integer, so the result is exact and order independent)
    for (int d = 16; d > 0; d >>= 1) {
        m1 += __shfl_down_sync(0xffffffffu, m1, d);
        m2 += __shfl_down_sync(0xffffffffu, m2, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&moments[0], m1);
        atomicAdd(&moments[1], m2);
    }
}

// ratios of a batch of cells; skipped entries get +inf so that they sort to the end
__global__ void k_sf_ratios(const int4* __restrict__ counts, const double* __restrict__ geo,
                            const int* __restrict__ weights, int cell0, int n_batch, int S, int zero_cov_mode,
                            double* __restrict__ keys, int* __restrict__ vals) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (s >= S || b >= n_batch) return;
    int c = __ldg(&counts[(size_t)(cell0 + b) * S + s].w);
    if (zero_cov_mode == 1) c += 1;
    const double g = geo[s];
    const bool keep = zero_cov_mode == 1 || (c > 0 && g > 0.0);
    keys[(size_t)b * S + s] = keep ? (double)c / g : __longlong_as_double(0x7ff0000000000000LL);
    vals[(size_t)b * S + s] = keep ? (weights ? weights[s] : 1) : 0;
}

// one block per cell: weighted median of its sorted ratios
__global__ void __launch_bounds__(256) k_sf_median(const double* __restrict__ keys, const int* __restrict__ vals, int S,
                                                   int cell0, double* __restrict__ out) {
    __shared__ long long chunk_sum[256];
    __shared__ double picked[2];
    const int b = blockIdx.x;
    const double* k = keys + (size_t)b * S;
    const int* v = vals + (size_t)b * S;
    const int per = (S + blockDim.x - 1) / blockDim.x;
    const int lo = min(S, (int)threadIdx.x * per), hi = min(S, lo + per);
    long long mine = 0;
    for (int i = lo; i < hi; i++) mine += v[i];
    chunk_sum[threadIdx.x] = mine;
    if (threadIdx.x < 2) picked[threadIdx.x] = __longlong_as_double(0x7ff8000000000000LL);
    __syncthreads();
    long long before = 0, total = 0;
    for (int i = 0; i < (int)blockDim.x; i++) {
        if (i < (int)threadIdx.x) before += chunk_sum[i];
        total += chunk_sum[i];
    }
    if (total > 0) {
        // expanded (by weight) positions of the one or two middle elements, DiscreteStatistics.median
        const long long t1 = total / 2, t0 = (total % 2 == 1) ? t1 : t1 - 1;
        long long cum = before;
        for (int i = lo; i < hi; i++) {
            const long long nxt = cum + v[i];
            if (t0 >= cum && t0 < nxt) picked[0] = k[i];
            if (t1 >= cum && t1 < nxt) picked[1] = k[i];
            cum = nxt;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) out[cell0 + b] = total > 0 ? (picked[0] + picked[1]) / 2.0 : picked[0];
}

// Host driver; returns a cudaError_t.  stats_out[2] = {mean / 2, variance / 4} of the processed coverages.
static cudaError_t sv_compute_size_factors(const int4* d_counts, const int* d_weights, int n_cells, int S,
                                           int zero_cov_mode, double* d_sf_out, double* stats_out, cudaStream_t st,
                                           unsigned long long* launches) {
    cudaError_t err;
    double* d_geo = nullptr;
    unsigned long long* d_mom = nullptr;
    double *d_keys = nullptr, *d_keys2 = nullptr;
    int *d_vals = nullptr, *d_vals2 = nullptr, *d_off = nullptr;
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0;
    const int batch = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_cells, ((size_t)192 << 20) / ((size_t)S * 8)));
#define SF_CHECK(x)                 \
    do {                            \
        err = (x);                  \
        if (err != cudaSuccess) goto done; \
    } while (0)
    SF_CHECK(cudaMalloc(&d_geo, sizeof(double) * (size_t)S));
    SF_CHECK(cudaMalloc(&d_mom, 2 * sizeof(unsigned long long)));
    SF_CHECK(cudaMemsetAsync(d_mom, 0, 2 * sizeof(unsigned long long), st));
    k_sf_geo_mean<<<(S + 255) / 256, 256, 0, st>>>(d_counts, n_cells, S, zero_cov_mode, d_geo, d_mom);
    (*launches)++;
    SF_CHECK(cudaGetLastError());
    SF_CHECK(cudaMalloc(&d_keys, sizeof(double) * (size_t)batch * S));
    SF_CHECK(cudaMalloc(&d_keys2, sizeof(double) * (size_t)batch * S));
    SF_CHECK(cudaMalloc(&d_vals, sizeof(int) * (size_t)batch * S));
    SF_CHECK(cudaMalloc(&d_vals2, sizeof(int) * (size_t)batch * S));
    SF_CHECK(cudaMalloc(&d_off, sizeof(int) * (size_t)(batch + 1)));
    {
        std::vector<int> off(batch + 1);
        for (int i = 0; i <= batch; i++) off[i] = i * S;
        SF_CHECK(cudaMemcpyAsync(d_off, off.data(), sizeof(int) * (batch + 1), cudaMemcpyHostToDevice, st));
        SF_CHECK(cudaStreamSynchronize(st));
    }
    SF_CHECK(cub::DeviceSegmentedRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, batch * S,
                                                      batch, d_off, d_off + 1, 0, 64, st));
    SF_CHECK(cudaMalloc(&d_tmp, tmp_bytes));
    for (int c0 = 0; c0 < n_cells; c0 += batch) {
        const int nb = std::min(batch, n_cells - c0);
        dim3 g((S + 255) / 256, nb);
        k_sf_ratios<<<g, 256, 0, st>>>(d_counts, d_geo, d_weights, c0, nb, S, zero_cov_mode, d_keys, d_vals);
        (*launches)++;
        SF_CHECK(cudaGetLastError());
        SF_CHECK(cub::DeviceSegmentedRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, nb * S, nb,
                                                          d_off, d_off + 1, 0, 64, st));
        k_sf_median<<<nb, 256, 0, st>>>(d_keys2, d_vals2, S, c0, d_sf_out);
        (*launches)++;
        SF_CHECK(cudaGetLastError());
    }
    if (stats_out) {
        unsigned long long mom[2];
        SF_CHECK(cudaMemcpyAsync(mom, d_mom, sizeof(mom), cudaMemcpyDeviceToHost, st));
        SF_CHECK(cudaStreamSynchronize(st));
        const double N = (double)n_cells * (double)S;
        const double mean = (double)mom[0] / N;
        const double var = ((double)mom[1] - (double)mom[0] * mean) / (N - 1.0);
        stats_out[0] = mean / 2;
        stats_out[1] = var / 4;
    }
    SF_CHECK(cudaStreamSynchronize(st));
done:
#undef SF_CHECK
    cudaFree(d_geo);
    cudaFree(d_mom);
    cudaFree(d_keys);
    cudaFree(d_keys2);
    cudaFree(d_vals);
    cudaFree(d_vals2);
    cudaFree(d_off);
    cudaFree(d_tmp);
    return err;
}
