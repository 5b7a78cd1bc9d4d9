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

// The one exchange step of the multi-GPU path, fused into the reduction (comm.cuh sets it up): every rank owns a mailbox
// [2 parities][world][8 doubles] in its HBM that its peers have mapped through CUDA IPC.  The last block of k_reduce writes
// this rank's five totals straight into every peer's mailbox over NVLink (slot = this rank), fences, raises the slot's flag
// (the evaluation counter), then waits for the world's flags in its OWN mailbox and copies the rows into `gather` -- the
// all-gather of likelihood/ThreadedScsTreeLikelihood.java:366-379 without a second kernel or a library call.  Two
// parities: a rank can be at most one evaluation ahead of a peer (it needs the peer's flag of evaluation t to finish t),
// so the rows of t + 1 never overwrite rows of t that are still being read.  peers == NULL: single process, no exchange.
struct SvMail {
    double* const* peers;        // [world] device pointers to the mailboxes (own one at [rank])
    double* gather;              // [world][5] on this device
    unsigned int* error;         // set to 1 when a peer did not answer within the time limit
    unsigned long long step;     // evaluation counter, the same on every rank; starts at 1
    int rank, world;
};
#define SV_MAIL_ROW 8
#define SV_MAIL_TIMEOUT_NS 600000000000ull  // 10 minutes: a peer may legitimately be busy on its host for a while

__device__ __forceinline__ unsigned long long sv_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void sv_mail_exchange(const SvMail& m, const double* res /*shared, [5]*/) {
    const int r = threadIdx.x;
    if (r >= m.world) return;
    const size_t par = (size_t)(m.step & 1ull) * m.world;
    volatile double* dst = m.peers[r] + (par + m.rank) * SV_MAIL_ROW;
    for (int i = 0; i < 5; i++) dst[i] = res[i];
    __threadfence_system();
    reinterpret_cast<volatile unsigned long long*>(dst)[5] = m.step;
    volatile double* src = m.peers[m.rank] + (par + r) * SV_MAIL_ROW;
    const unsigned long long t0 = sv_globaltimer();
    bool ok = true;
    while (reinterpret_cast<volatile unsigned long long*>(src)[5] != m.step) {
        if (sv_globaltimer() - t0 > SV_MAIL_TIMEOUT_NS) {
            ok = false;
            break;
        }
        __nanosleep(100);
    }
    __threadfence_system();
    if (!ok) *m.error = 1u;
    for (int i = 0; i < 5; i++) m.gather[r * 5 + i] = ok ? src[i] : __longlong_as_double(0x7ff8000000000000LL);
}

// out layout == sieve_shard_result_t {sum_log_l, const_lse, const_sum, lewis_sum, weight_sum}
// The constant-site log-likelihood of site s is  cst[s] + cst_shift:  with the algebraic constant-site path cst is
// Lambda (sum over cells of the leaves' constant-genotype log-likelihoods) and cst_shift = log C_tree (one number per
// tree, computed on the host); with the per-site twin cst is logConstRoot itself and cst_shift = 0.
__global__ void __launch_bounds__(256) k_reduce(const double* __restrict__ logl, const double* __restrict__ cst,
                                                double cst_shift, const int* __restrict__ w, int S, int with_const,
                                                SvAcc* __restrict__ partials, unsigned int* __restrict__ counter,
                                                double* __restrict__ out, SvMail mail = SvMail{nullptr, nullptr, nullptr, 0, 0, 1}) {
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
        __shared__ double res[5];
        if (threadIdx.x == 0) {
            t = fin[0];
            res[0] = t.sum;
            res[1] = with_const ? log(t.se) + t.m : 0.0;
            res[2] = t.csum;
            res[3] = -t.lewis;
            res[4] = t.wsum;
            for (int i = 0; i < 5; i++) out[i] = res[i];
            *counter = 0u;
        }
        if (mail.peers) {
            __syncthreads();
            sv_mail_exchange(mail, res);
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
