// sizefactors_global.cuh -- per-cell size factors over a SITE-SHARDED alignment, exact and pool-free.
//
// The reference computes the size factors once over the FULL alignment before it splits the sites over its workers
// (likelihood/ThreadedScsTreeLikelihood.java:119-131 -> rawreadcountsmodel/seqcovmodel/SeqCovModelInterface.java:291-351)
// and hands every worker a copy (ExploredSharedAllelicSeqCovModel.java:103-109).  With one GPU per site range no device
// ever sees the full alignment (10 000 cells x 1 000 000 sites are 160 GB of counts), so the median of a cell's
// coverage ratios over ALL sites is found by exact distributed selection instead of by sorting:
//
//   * geometric means are per site, i.e. local to the part that owns the site (k_sf_geo_mean, sizefactors.cuh);
//   * a ratio cov / geomean is a positive finite double, and for those the order of the values is the order of their
//     bit patterns: the value at (weight-expanded) position t of a cell is the smallest bit pattern u whose
//     "weight of the ratios <= u", summed over all parts, exceeds t -- found by bisection on u, 63 rounds;
//   * one round = one counting pass over every part's counts (k_sf_weight_le: reads the 16-byte count tuples, nothing
//     is materialised or sorted) + one integer all-reduce of 2 numbers per cell.  Integer sums: the result does not
//     depend on how the sites are split, and equals the single-handle result of sieve_compute_size_factors bit for bit.
//
// A "part" (sieve_sf_t) borrows a device-resident count tensor; it needs no likelihood core and no pools, so it can run
// before the cores are created (sparse cores size their pools from the memory that is free at that time).
// sieve_sf_solve is pure host logic over callbacks, so the protocol is testable without a device.
#pragma once
#include "sizefactors.cuh"

#define SV_SF_MAX_FINITE 0x7FEFFFFFFFFFFFFFull

struct sieve_sf {
    int device = 0;
    int n = 0, S = 0, zero_cov_mode = 0;
    const int4* d_counts = nullptr;  // borrowed
    int* d_weights = nullptr;        // owned copy, or NULL
    double* d_geo = nullptr;         // [S]
    unsigned long long* d_mom = nullptr;
    unsigned long long* d_piv = nullptr;  // [n][2]
    unsigned long long* d_cnt = nullptr;  // [n][2]
    unsigned long long mom[2] = {0, 0};
    cudaStream_t stream = nullptr;
    unsigned long long launches = 0;
};

// weight of the kept ratios of cell blockIdx.y that are <= each of the cell's n_piv pivots (bit patterns), over the
// sites of this part.  Integer atomics: the sum does not depend on the order.
template <int NPIV>
__global__ void __launch_bounds__(256) k_sf_weight_le(const int4* __restrict__ counts, const double* __restrict__ geo,
                                                      const int* __restrict__ weights, int S, int zero_cov_mode,
                                                      const unsigned long long* __restrict__ piv,
                                                      unsigned long long* __restrict__ out) {
    const int j = blockIdx.y;
    long long p[NPIV];
#pragma unroll
    for (int i = 0; i < NPIV; i++) p[i] = (long long)piv[(size_t)j * NPIV + i];
    unsigned long long acc[NPIV];
#pragma unroll
    for (int i = 0; i < NPIV; i++) acc[i] = 0;
    const int4* row = counts + (size_t)j * S;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
        int c = __ldg(&row[s].w);
        if (zero_cov_mode == 1) c += 1;
        const double g = __ldg(geo + s);
        const bool keep = zero_cov_mode == 1 || (c > 0 && g > 0.0);
        if (!keep) continue;
        const long long key = __double_as_longlong((double)c / g);  // the same expression as k_sf_ratios
        const unsigned long long w = weights ? (unsigned long long)__ldg(weights + s) : 1ull;
#pragma unroll
        for (int i = 0; i < NPIV; i++) acc[i] += key <= p[i] ? w : 0ull;
    }
#pragma unroll
    for (int i = 0; i < NPIV; i++) {
        unsigned long long v = acc[i];
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&out[(size_t)j * NPIV + i], v);
    }
}

static int sv_sf_open(int device, int n_cells, int n_sites, const void* d_counts, const int32_t* weights, int zero_cov_mode,
                      sieve_sf** out) {
    SV_REQUIRE(out && d_counts && n_cells >= 1 && n_sites >= 1, "sieve_sf_open: bad argument");
    SV_REQUIRE(zero_cov_mode == 0 || zero_cov_mode == 1, "zeroCovMode must be 0 or 1");
    int ndev = 0;
    SV_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) return sv_fail(SIEVE_ERR_CUDA, "sieve_sf_open: no CUDA device (there is no CPU fallback)");
    if (device < 0) SV_CUDA(cudaGetDevice(&device));
    SV_REQUIRE(device < ndev, "sieve_sf_open: device ordinal out of range");
    SV_CUDA(cudaSetDevice(device));
    sieve_sf* p = new sieve_sf();
    p->device = device;
    p->n = n_cells;
    p->S = n_sites;
    p->zero_cov_mode = zero_cov_mode;
    p->d_counts = (const int4*)d_counts;
    cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_geo, sizeof(double) * (size_t)n_sites);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_mom, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_piv, 2 * sizeof(unsigned long long) * (size_t)n_cells);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_cnt, 2 * sizeof(unsigned long long) * (size_t)n_cells);
    if (e == cudaSuccess && weights) {
        e = cudaMalloc(&p->d_weights, sizeof(int) * (size_t)n_sites);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(p->d_weights, weights, sizeof(int) * (size_t)n_sites, cudaMemcpyHostToDevice, p->stream);
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(p->d_mom, 0, 2 * sizeof(unsigned long long), p->stream);
    if (e == cudaSuccess) {
        // geometric mean of every site of this part over the cells + exact integer moments of the coverages
        k_sf_geo_mean<<<(n_sites + 255) / 256, 256, 0, p->stream>>>(p->d_counts, n_cells, n_sites, zero_cov_mode, p->d_geo,
                                                                   p->d_mom);
        p->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->mom, p->d_mom, sizeof(p->mom), cudaMemcpyDeviceToHost, p->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    if (e != cudaSuccess) {
        cudaFree(p->d_geo);
        cudaFree(p->d_mom);
        cudaFree(p->d_piv);
        cudaFree(p->d_cnt);
        cudaFree(p->d_weights);
        if (p->stream) cudaStreamDestroy(p->stream);
        delete p;
        cudaGetLastError();
        return sv_fail(SIEVE_ERR_CUDA, std::string("sieve_sf_open: ") + cudaGetErrorString(e));
    }
    *out = p;
    return SIEVE_OK;
}

static int sv_sf_weight_le(sieve_sf* p, int n_piv, const uint64_t* pivots, int64_t* out) {
    SV_REQUIRE(p && pivots && out && (n_piv == 1 || n_piv == 2), "sieve_sf_weight_le: bad argument");
    SV_CUDA(cudaSetDevice(p->device));
    const size_t bytes = sizeof(unsigned long long) * (size_t)p->n * n_piv;
    SV_CUDA(cudaMemcpyAsync(p->d_piv, pivots, bytes, cudaMemcpyHostToDevice, p->stream));
    SV_CUDA(cudaMemsetAsync(p->d_cnt, 0, bytes, p->stream));
    // enough blocks per cell to fill the device when there are few cells, few enough that a block has real work
    int bx = (p->S + 256 * 16 - 1) / (256 * 16);
    bx = std::max(1, std::min(bx, 64));
    dim3 grid((unsigned)bx, (unsigned)p->n);
    if (n_piv == 1)
        k_sf_weight_le<1><<<grid, 256, 0, p->stream>>>(p->d_counts, p->d_geo, p->d_weights, p->S, p->zero_cov_mode, p->d_piv, p->d_cnt);
    else
        k_sf_weight_le<2><<<grid, 256, 0, p->stream>>>(p->d_counts, p->d_geo, p->d_weights, p->S, p->zero_cov_mode, p->d_piv, p->d_cnt);
    p->launches++;
    SV_CUDA(cudaGetLastError());
    SV_CUDA(cudaMemcpyAsync(out, p->d_cnt, bytes, cudaMemcpyDeviceToHost, p->stream));
    SV_CUDA(cudaStreamSynchronize(p->stream));
    return SIEVE_OK;
}

// ---- the selection protocol: host logic over callbacks ------------------------------------------------------------
static int sv_sf_solve(const sieve_sf_ops_t* ops, double* out_sf, double* out_stats) {
    SV_REQUIRE(ops && ops->moments && ops->weight_le && out_sf && ops->n_cells >= 1, "sieve_sf_solve: bad argument");
    const int n = ops->n_cells;
    auto reduce = [&](int64_t* buf, int count) -> int {
        if (!ops->allreduce_sum_i64) return SIEVE_OK;
        const int r = ops->allreduce_sum_i64(ops->user, buf, count);
        return r == 0 ? SIEVE_OK : sv_fail(SIEVE_ERR_STATE, "sieve_sf_solve: the all-reduce callback failed");
    };
    auto local = [&](int r, const char* what) -> int {
        return r == 0 ? SIEVE_OK : (g_err.empty() || g_err.find(what) == std::string::npos
                                        ? sv_fail(SIEVE_ERR_STATE, std::string("sieve_sf_solve: ") + what + " failed")
                                        : r);
    };
    // 1. exact integer moments of the (processed) coverages over all parts: sum, sum of squares, number of entries
    int64_t mom[3] = {0, 0, 0};
    {
        uint64_t m[3] = {0, 0, 0};
        SV_TRY(local(ops->moments(ops->user, m), "moments"));
        for (int i = 0; i < 3; i++) mom[i] = (int64_t)m[i];
    }
    SV_TRY(reduce(mom, 3));
    if (out_stats) {
        // DiscreteStatistics.mean / variance of all coverages, halved / quartered (SeqCovModelInterface.java:322-325)
        const double N = (double)mom[2];
        const double mean = (double)(uint64_t)mom[0] / N;
        const double var = ((double)(uint64_t)mom[1] - (double)(uint64_t)mom[0] * mean) / (N - 1.0);
        out_stats[0] = mean / 2;
        out_stats[1] = var / 4;
    }
    // 2. per cell: the weight of the ratios that take part (zero coverage / zero geometric mean are skipped, :303-313)
    std::vector<uint64_t> piv((size_t)2 * n, SV_SF_MAX_FINITE);
    std::vector<int64_t> cnt((size_t)2 * n, 0);
    SV_TRY(local(ops->weight_le(ops->user, 2, piv.data(), cnt.data()), "weight_le"));
    SV_TRY(reduce(cnt.data(), 2 * n));
    // 3. weight-expanded positions of the one or two middle elements (DiscreteStatistics.median)
    std::vector<int64_t> target((size_t)2 * n, 0), total(n, 0);
    for (int j = 0; j < n; j++) {
        total[j] = cnt[(size_t)2 * j];
        const int64_t t1 = total[j] / 2, t0 = (total[j] % 2 == 1) ? t1 : t1 - 1;
        target[(size_t)2 * j] = t0;
        target[(size_t)2 * j + 1] = t1;
    }
    // 4. bisection on the bit pattern: the smallest u with weight(<= u) > target
    std::vector<uint64_t> lo((size_t)2 * n, 0), hi((size_t)2 * n, SV_SF_MAX_FINITE);
    for (int j = 0; j < n; j++)
        if (total[j] <= 0) hi[(size_t)2 * j] = hi[(size_t)2 * j + 1] = 0;  // nothing to select
    for (int round = 0; round < 64; round++) {
        bool any = false;
        for (size_t i = 0; i < (size_t)2 * n; i++) {
            if (lo[i] < hi[i]) {
                piv[i] = lo[i] + (hi[i] - lo[i]) / 2;
                any = true;
            } else
                piv[i] = lo[i];
        }
        if (!any) break;
        SV_TRY(local(ops->weight_le(ops->user, 2, piv.data(), cnt.data()), "weight_le"));
        SV_TRY(reduce(cnt.data(), 2 * n));
        for (size_t i = 0; i < (size_t)2 * n; i++) {
            if (lo[i] >= hi[i]) continue;
            if (cnt[i] > target[i])
                hi[i] = piv[i];
            else
                lo[i] = piv[i] + 1;
        }
    }
    for (int j = 0; j < n; j++) {
        if (total[j] <= 0) {
            out_sf[j] = NAN;
            continue;
        }
        double a, b;
        memcpy(&a, &lo[(size_t)2 * j], 8);
        memcpy(&b, &lo[(size_t)2 * j + 1], 8);
        out_sf[j] = (a + b) / 2.0;
    }
    return SIEVE_OK;
}

// parts of THIS process (one per GPU it drives) + an optional all-reduce across processes
struct SvSfParts {
    int n_parts;
    sieve_sf* const* parts;
    sieve_sf_allreduce_fn ar;
    void* ar_user;
    std::vector<int64_t> tmp;
};
static int sv_sf_parts_moments(void* user, uint64_t* out3) {
    SvSfParts* P = (SvSfParts*)user;
    out3[0] = out3[1] = out3[2] = 0;
    for (int i = 0; i < P->n_parts; i++) {
        out3[0] += P->parts[i]->mom[0];
        out3[1] += P->parts[i]->mom[1];
        out3[2] += (uint64_t)P->parts[i]->n * (uint64_t)P->parts[i]->S;
    }
    return 0;
}
static int sv_sf_parts_weight_le(void* user, int32_t n_piv, const uint64_t* pivots, int64_t* out) {
    SvSfParts* P = (SvSfParts*)user;
    const size_t len = (size_t)P->parts[0]->n * n_piv;
    P->tmp.resize(len);
    for (size_t k = 0; k < len; k++) out[k] = 0;
    for (int i = 0; i < P->n_parts; i++) {
        const int r = sv_sf_weight_le(P->parts[i], n_piv, pivots, P->tmp.data());
        if (r != SIEVE_OK) return r;
        for (size_t k = 0; k < len; k++) out[k] += P->tmp[k];
    }
    return 0;
}
static int sv_sf_parts_allreduce(void* user, int64_t* buf, int32_t count) {
    SvSfParts* P = (SvSfParts*)user;
    return P->ar ? P->ar(P->ar_user, buf, count) : 0;
}
