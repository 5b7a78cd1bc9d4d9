// sieve_core.cu -- the C-ABI of include/sieve_b200.h: device pools, buffer-index bookkeeping and kernel launches.
//
// Host-side state mirrors BeerLikelihoodCore's double buffering (external BEAST class; SIEVE's additions in
// likelihood/ScsBeerLikelihoodCore.java:91-128, 4053-4065): per node a "current" and a "stored" buffer index for the
// partials and for the transition matrices; reject == swap of the index arrays, nothing is copied on the device.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <chrono>
#include <cstdio>
#include "../../include/sieve_b200.h"
#include "common.cuh"
#include "leaf.cuh"
#include "leaf_private.cuh"
#include "prune.cuh"
#include "walk.cuh"
#include "reduce.cuh"
#include "maxsum.cuh"
#include "sizefactors.cuh"
#include "background.cuh"
#include "comm.cuh"

#include "core_state.cuh"
#include "sizefactors_global.cuh"

extern "C" const char* sieve_last_error(void) { return g_err.c_str(); }
extern "C" int sieve_version(void) { return 1; }
extern "C" int sieve_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int sieve_create(const sieve_config_t* cfg, sieve_core_t** out) {
    SV_REQUIRE(cfg && out, "sieve_create: null argument");
    SV_REQUIRE(cfg->n_states == 4, "sieve_create: n_states must be 4 (ScsFiniteMuExtendedModel)");
    SV_REQUIRE(cfg->n_categories >= 1 && cfg->n_categories <= SV_KMAX, "sieve_create: n_categories must be in 1..4");
    SV_REQUIRE(cfg->n_leaves >= 1 && cfg->n_sites >= 1, "sieve_create: n_leaves and n_sites must be positive");
    int ndev = 0;
    SV_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) return sv_fail(SIEVE_ERR_CUDA, "sieve_create: no CUDA device (there is no CPU fallback)");
    int dev = cfg->device;
    if (dev < 0) SV_CUDA(cudaGetDevice(&dev));
    SV_REQUIRE(dev < ndev, "sieve_create: device ordinal out of range");
    SV_CUDA(cudaSetDevice(dev));

    sieve_core* h = new sieve_core();
    h->cfg = *cfg;
    h->device = dev;
    h->n = cfg->n_leaves;
    h->S = cfg->n_sites;
    h->K = cfg->n_categories;
    h->n_nodes = 2 * h->n;
    h->n_internal = h->n;
    h->log_mode = cfg->flags & SIEVE_FLAG_LOG_PARTIALS;
    h->with_const = cfg->flags & SIEVE_FLAG_CONST_PARTIALS;
    if (const char* e = getenv("SIEVE_B200_CONST_TWIN")) h->const_twin = atoi(e) != 0;
    if (const char* e = getenv("SIEVE_B200_PRUNE_MINB")) h->prune_minb = atoi(e);
    if (const char* e = getenv("SIEVE_B200_PF_DIST")) h->pf_dist = std::max(0, std::min(64, atoi(e)));
    if (const char* e = getenv("SIEVE_B200_LEAF_TAB")) h->leaf_tab = std::min(2, std::max(0, atoi(e)));
    if (const char* e = getenv("SIEVE_B200_LEAF_LIN")) h->leaf_lin_allowed = atoi(e) != 0;
    if (const char* e = getenv("SIEVE_B200_LEAF_CHECKED")) h->leaf_force_checked = atoi(e) != 0;
    const bool twin = h->with_const && h->const_twin;
    const bool algebraic = h->with_const && !h->const_twin;
    h->single_ado = cfg->flags & SIEVE_FLAG_SINGLE_ADO;
    h->ephemeral = cfg->flags & SIEVE_FLAG_EPHEMERAL;
    h->cur_part.assign(h->n_nodes, 0);
    h->stored_part.assign(h->n_nodes, 0);
    h->cur_mat.assign(h->n_nodes, 0);
    h->stored_mat.assign(h->n_nodes, 0);
    h->root = h->n_nodes - 1;
    if (const char* e = getenv("SIEVE_B200_GENERIC")) h->force_generic = atoi(e) != 0;

    auto fail = [&](int code) {
        sieve_destroy(h);
        return code;
    };
    cudaError_t ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) {
        g_err = std::string("cudaStreamCreate: ") + cudaGetErrorString(ce);
        return fail(SIEVE_ERR_CUDA);
    }
    const size_t n = h->n, S = h->S;
    size_t pool_sites = S, part_slots = 2 * n, leaf_bufs = 2;
    if (h->ephemeral) {
        // scratch sized so that the per-chunk leaf partials stay around 4 GB; slots: Sethi-Ullman bound of the walk
        // chunk: large enough to fill the GPU a few times over (>= 3 full waves of 32-site blocks), bounded by a
        // leaf-scratch budget of a quarter of the free memory (at most 48 GB)
        size_t free0 = 0, total0 = 0;
        cudaMemGetInfo(&free0, &total0);
        const size_t counts_bytes = (cfg->flags & SIEVE_FLAG_EXTERNAL_COUNTS) ? 0 : n * S * 16;
        const size_t avail = free0 > counts_bytes ? free0 - counts_bytes : 0;
        const size_t budget = std::min<size_t>((size_t)48 << 30, avail / 4);
        size_t cap = std::max<size_t>(4096, budget / (40 * n));
        cap = std::min<size_t>(cap, (size_t)148 * 8 * 32 * 4);
        if (const char* e = getenv("SIEVE_B200_CHUNK_SITES")) cap = std::max(32, atoi(e));
        cap = std::min(S, cap);
        cap = (cap + 31) / 32 * 32;
        h->chunk_cap = (int)cap;
        int lg = 1;
        while (((size_t)1 << lg) < n) lg++;
        h->n_slots = lg + 4;
        pool_sites = cap;
        part_slots = h->n_slots;
        leaf_bufs = 1;
    }
    h->private_cov = cfg->flags & SIEVE_FLAG_PRIVATE_SEQ_COV;
    if (h->private_cov && ((cfg->flags & (SIEVE_FLAG_EPHEMERAL | SIEVE_FLAG_SPARSE | SIEVE_FLAG_SINGLE_LEAF_BUFFER)) || twin)) {
        g_err = "sieve_create: SIEVE_FLAG_PRIVATE_SEQ_COV needs a dense resident core (no streamed / sparse / single-leaf-buffer mode, no per-site constant twin)";
        return fail(SIEVE_ERR_INVALID);
    }
    const size_t leaf_vals = h->private_cov ? SV_KMAX * 4 : 4;  // doubles per (cell, site) in the leaf pool
    h->sparse = (cfg->flags & SIEVE_FLAG_SPARSE) && !h->ephemeral;
    h->single_leaf = (cfg->flags & SIEVE_FLAG_SINGLE_LEAF_BUFFER) && !h->ephemeral;
    if (h->single_leaf) leaf_bufs = 1;
    if (h->sparse) {
        // everything but the internal partials is resident as usual; the pool takes what is left (minus head room)
        size_t free0 = 0, total0 = 0;
        cudaMemGetInfo(&free0, &total0);
        const size_t fixed = ((cfg->flags & SIEVE_FLAG_EXTERNAL_COUNTS) ? 0 : n * S * 16) + leaf_bufs * n * S * 32 + ((n + 31) / 32) * S * 16 + ((size_t)3 << 30);
        const size_t per_slot = S * (SV_KMAX * 32 + 4) * (twin ? 2 : 1);
        size_t slots = free0 > fixed ? (free0 - fixed) / per_slot : 0;
        if (const char* e = getenv("SIEVE_B200_POOL_SLOTS")) slots = (size_t)std::max(0, atoi(e));
        slots = std::min<size_t>(slots, 2 * n);
        if (slots < std::min<size_t>(96, 2 * n)) {
            g_err = "sieve_create: SIEVE_FLAG_SPARSE needs room for at least 96 partial slots";
            return fail(SIEVE_ERR_OOM);
        }
        part_slots = slots;
    }
    h->pool_slots = (int)part_slots;
    h->slot_of.assign((size_t)2 * h->n_nodes, -1);
    if (h->sparse) {
        for (int i = (int)part_slots - 1; i >= 0; i--) h->free_slots.push_back(i);
    } else if (!h->ephemeral) {
        for (int b = 0; b < 2; b++)
            for (int v = h->n; v < h->n_nodes; v++) h->slot_of[(size_t)b * h->n_nodes + v] = b * h->n_internal + (v - h->n);
    }
    h->eph_c1.assign(h->n_nodes, -2);
    h->eph_c2.assign(h->n_nodes, -2);
    h->eph_c1_stored = h->eph_c1;
    h->eph_c2_stored = h->eph_c2;
    h->valid.assign(h->n_nodes, 0);
    h->valid_stored.assign(h->n_nodes, 0);
    if (const char* e = getenv("SIEVE_B200_TRANSIENT_LEAVES")) h->transient_T = std::max(0, atoi(e));
    h->keep_T = std::max(1, h->transient_T);
    if ((leaf_bufs * n + 2) * pool_sites >= ((size_t)1 << 32) || part_slots * pool_sites + 2 * n + 2 >= ((size_t)1 << 32)) {
        g_err = "sieve_create: more than 2^32 (slot, site) rows per pool; use SIEVE_FLAG_EPHEMERAL (streamed evaluation)";
        return fail(SIEVE_ERR_OOM);
    }
    // does the residency fit?
    size_t need = ((cfg->flags & SIEVE_FLAG_EXTERNAL_COUNTS) ? 0 : n * S * 16) + leaf_bufs * n * pool_sites * 8 * leaf_vals + ((n + 31) / 32) * pool_sites * 16 +
                  part_slots * pool_sites * (SV_KMAX * 32 + 4) * (twin ? 2 : 1) + ((size_t)64 << 20);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (need > free_b) {
        g_err = "sieve_create: the pools need " + std::to_string(need >> 20) + " MiB, the device has " +
                std::to_string(free_b >> 20) + " MiB free" +
                (h->ephemeral ? "" : " (SIEVE_FLAG_EPHEMERAL streams the evaluation instead of keeping partials resident)");
        return fail(SIEVE_ERR_OOM);
    }
    int r;
#define CR(x)                              \
    do {                                   \
        r = (x);                           \
        if (r != SIEVE_OK) return fail(r); \
    } while (0)
    if (cfg->flags & SIEVE_FLAG_EXTERNAL_COUNTS)
        h->counts_owned = false;  // adopted by sieve_set_counts_device
    else
        CR(sv_alloc(h, &h->d_counts, n * S));
    CR(sv_alloc(h, &h->d_sf, n));
    CR(sv_alloc(h, &h->d_leafX, leaf_bufs * n * pool_sites * leaf_vals));
    h->pool_rows_leaf = leaf_bufs * n * pool_sites;
    h->pool_rows_part = part_slots * pool_sites;
    const size_t leaf_groups = (n + SV_LEAF_CELLS - 1) / SV_LEAF_CELLS;
    CR(sv_alloc(h, &h->d_sumS, leaf_groups * pool_sites));
    CR(sv_alloc(h, &h->d_lsum, 2 * S));
    CR(sv_alloc(h, &h->d_part, part_slots * pool_sites * SV_KMAX * 4));
    CR(sv_alloc(h, &h->d_scale, part_slots * pool_sites));
    if (twin) {
        CR(sv_alloc(h, &h->d_cpart, part_slots * pool_sites * SV_KMAX * 4));
        CR(sv_alloc(h, &h->d_cscale, part_slots * pool_sites));
    }
    if (h->private_cov) {
        CR(sv_alloc(h, &h->d_tv, (size_t)SV_KMAX * S));
        CR(sv_alloc(h, &h->d_tv_leaf, (size_t)SV_KMAX * S));
        if (algebraic) {
            CR(sv_alloc(h, &h->d_sum0K, leaf_groups * SV_KMAX * pool_sites));
            CR(sv_alloc(h, &h->d_lambdaK, (size_t)2 * SV_KMAX * S));
        }
        h->pf_dist = 0;  // the prefetch annotations assume 32-byte leaf rows
    }
    if (algebraic) {
        CR(sv_alloc(h, &h->d_sum0, leaf_groups * pool_sites));
        CR(sv_alloc(h, &h->d_lambda, 2 * S));
        h->G.assign((size_t)2 * h->n_nodes * SV_KMAX * 4, 0.0);
        h->Gs.assign((size_t)2 * h->n_nodes, 0.0);
    }
    if (algebraic) h->hP.assign((size_t)2 * h->n_nodes * SV_KMAX * 16, 0.0);
    CR(sv_alloc(h, &h->d_mats, (size_t)2 * h->n_nodes * SV_KMAX * 16));
    CR(sv_alloc(h, &h->d_qfac, (size_t)2 * h->n_nodes * SV_KMAX));
    h->mat_ver.assign((size_t)2 * h->n_nodes, 0);
    h->mat_mirror.assign((size_t)2 * h->n_nodes * SV_KMAX * 16, 0.0);
    h->dist_mirror.assign((size_t)2 * h->n_nodes * SV_KMAX, 0.0);
    h->mat_mirror_q.assign((size_t)2 * h->n_nodes, 0);
    h->g_ver.assign((size_t)2 * h->n_nodes, 0);
    h->buf_has.assign((size_t)2 * h->n_nodes, 0);
    h->buf_ver.assign((size_t)2 * h->n_nodes, 0);
    h->buf_fp.assign((size_t)2 * h->n_nodes, sieve_core::SvFp());
    h->ver_cur.assign(h->n_nodes, 0);
    h->ver_stored.assign(h->n_nodes, 0);
    h->fp_cur.assign(h->n_nodes, sieve_core::SvFp());
    h->fp_stored.assign(h->n_nodes, sieve_core::SvFp());
    if (const char* e = getenv("SIEVE_B200_ELIDE")) h->elide = atoi(e) != 0;
    h->nb_lo.assign((size_t)2 * h->n_nodes, 0);
    h->nb_hi.assign((size_t)2 * h->n_nodes, 1);
    h->nb_mlo.assign((size_t)2 * h->n_nodes, 0);
    h->slot_b.assign((size_t)2 * h->n_nodes, SV_BOUND_NONE);
    h->slot_bd.assign((size_t)2 * h->n_nodes, SV_BOUND_NONE);
    if (const char* e = getenv("SIEVE_B200_RENORM_ALWAYS")) h->renorm_always = atoi(e) != 0;
    if (const char* e = getenv("SIEVE_B200_WALK")) h->use_k_prune = atoi(e) == 0;  // 0: the general kernel for everything (A/B, tests)
    h->slot_is_q.assign((size_t)2 * h->n_nodes, 0);
    h->slot_zero.assign((size_t)2 * h->n_nodes, 0);
    CR(sv_alloc(h, &h->d_site_logl, S));
    CR(sv_alloc(h, &h->d_site_const, S));
    CR(sv_alloc(h, &h->d_partials, (size_t)SV_REDUCE_BLOCKS));
    CR(sv_alloc(h, &h->d_counter, 1));
    CR(sv_alloc(h, &h->d_result, 8));
#undef CR
    if (cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned int), h->stream) != cudaSuccess ||
        cudaMemsetAsync(h->d_mats, 0, sizeof(double) * 2 * h->n_nodes * SV_KMAX * 16, h->stream) != cudaSuccess ||
        cudaMemsetAsync(h->d_site_const, 0, sizeof(double) * S, h->stream) != cudaSuccess ||
        cudaMallocHost((void**)&h->h_result, 8 * sizeof(double)) != cudaSuccess) {
        g_err = "sieve_create: initialisation of device buffers failed";
        return fail(SIEVE_ERR_CUDA);
    }
    for (int i = 0; i < 6; i++) cudaEventCreate(&h->ev[i]);
    for (int i = 0; i < 2; i++) cudaEventCreateWithFlags(&h->ev_stage[i], cudaEventDisableTiming);
    cudaStreamSynchronize(h->stream);
    *out = h;
    return SIEVE_OK;
}

extern "C" int sieve_destroy(sieve_core_t* h) {
    if (!h) return SIEVE_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->counts_owned) cudaFree(h->d_counts);
    cudaFree(h->d_weights);
    cudaFree(h->d_sf);
    cudaFree(h->d_dmA);
    cudaFree(h->d_dmB);
    cudaFree(h->d_cellT);
    cudaFree(h->d_linA);
    cudaFree(h->d_linB);
    cudaFree(h->d_leafX);
    cudaFree(h->d_sumS);
    cudaFree(h->d_sum0);
    cudaFree(h->d_lsum);
    cudaFree(h->d_part);
    cudaFree(h->d_scale);
    cudaFree(h->d_cpart);
    cudaFree(h->d_lambda);
    cudaFree(h->d_tv);
    cudaFree(h->d_tv_leaf);
    cudaFree(h->d_lgx1);
    cudaFree(h->d_cellT2);
    cudaFree(h->d_lambdaK);
    cudaFree(h->d_sum0K);
    cudaFree(h->d_cscale);
    cudaFree(h->d_mats);
    cudaFree(h->d_qfac);
    cudaFree(h->d_site_logl);
    cudaFree(h->d_site_const);
    cudaFree(h->d_partials);
    cudaFree(h->d_counter);
    cudaFree(h->d_result);
    cudaFree(h->d_mlL);
    cudaFree(h->d_ml);
    cudaFree(h->d_mllogm);
    cudaFree(h->d_mlrawm);
    cudaFree(h->d_mlcellL);
    cudaFree(h->d_mlado);
    cudaFree(h->d_mlback);
    cudaFree(h->d_mlcoll);
    cudaFree(h->d_mladosel);
    cudaFree(h->d_mlamb);
    cudaFree(h->d_mlcats);
    cudaFree(h->d_mlops);
    for (size_t r = 0; r < h->peer_mail.size(); r++)
        if (h->peer_mail[r] && h->peer_mail[r] != (void*)h->d_mail) cudaIpcCloseMemHandle(h->peer_mail[r]);
    cudaFree(h->d_mail);
    cudaFree(h->d_peers);
    cudaFree(h->d_mail_err);
    if (h->comm && sv_nccl().ok) sv_nccl().CommDestroy(h->comm);
    cudaFree(h->d_gather);
    if (h->h_gather) cudaFreeHost(h->h_gather);
    cudaFree(h->d_stage_base);
    if (h->h_stage_base) cudaFreeHost(h->h_stage_base);
    for (int i = 0; i < 2; i++)
        if (h->ev_stage[i]) cudaEventDestroy(h->ev_stage[i]);
    if (h->h_result) cudaFreeHost(h->h_result);
    for (int i = 0; i < 6; i++)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h;
    return SIEVE_OK;
}

extern "C" int sieve_device_bytes(sieve_core_t* h, uint64_t* out) {
    SV_REQUIRE(h && out, "null argument");
    *out = h->bytes;
    return SIEVE_OK;
}

// ---- counts ---------------------------------------------------------------------------------------------------
// [site][cell][T] chunk -> [cell][site] int4, tracking the largest coverage
__global__ void k_transpose_counts(const int* __restrict__ in, int tuple_len, int s0, int n_chunk_sites, int n_cells,
                                   int S, int4* __restrict__ out, int* __restrict__ max_cov) {
    // a 32 x 32 (site x cell) tile per block so that both the read and the write are coalesced
    __shared__ int4 tile[32][33];
    const int cs = blockIdx.x * 32, cc = blockIdx.y * 32;
    int mx = 0;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int s = cs + r, c = cc + threadIdx.x;
        if (s < n_chunk_sites && c < n_cells) {
            const int* p = in + ((size_t)s * n_cells + c) * tuple_len;
            int4 v;
            v.x = p[0];
            v.y = p[1];
            v.z = p[2];
            v.w = tuple_len == 5 ? p[4] : p[3];
            tile[r][threadIdx.x] = v;
            mx = max(mx, v.w);
        }
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = cc + r, s = cs + threadIdx.x;
        if (s < n_chunk_sites && c < n_cells) out[(size_t)c * S + s0 + s] = tile[threadIdx.x][r];
    }
    for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_down_sync(0xffffffffu, mx, d));
    if (threadIdx.x == 0 && mx > 0) atomicMax(max_cov, mx);
}

__global__ void k_max_cov(const int4* __restrict__ c, size_t N, int* __restrict__ max_cov) {
    int mx = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x)
        mx = max(mx, c[i].w);
    for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_down_sync(0xffffffffu, mx, d));
    if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax(max_cov, mx);
}

// does every count of the shard index the look-up tables?  (0 <= alt1, alt2, alt3, cov < C; then the reference count
// cov - sum(alt), clamped at 0, does too.)  If so the leaf kernel runs without its out-of-table paths.
__global__ void k_counts_outside_tables(const int4* __restrict__ c, size_t N, int C, int* __restrict__ outside) {
    int bad = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x) {
        const int4 v = c[i];
        bad |= (unsigned)v.x >= (unsigned)C || (unsigned)v.y >= (unsigned)C || (unsigned)v.z >= (unsigned)C ||
               (unsigned)v.w >= (unsigned)C;
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(outside, 1);
}

static int sv_alloc_tables(sieve_core* h) {
    int cap = SV_TABLE_CAP_DEFAULT;
    if (const char* e = getenv("SIEVE_B200_TABLE_CAP")) cap = std::max(2, atoi(e));
    const int C = std::min(h->max_cov + 1, cap);
    if (C != h->C || !h->d_dmA) {
        cudaFree(h->d_dmA);
        cudaFree(h->d_dmB);
        cudaFree(h->d_cellT);
        cudaFree(h->d_linA);
        cudaFree(h->d_linB);
        h->d_dmA = h->d_dmB = h->d_cellT = h->d_linA = h->d_linB = nullptr;
        h->C = std::max(C, 1);
        SV_TRY(sv_alloc(h, &h->d_dmA, (size_t)h->C * 4));
        SV_TRY(sv_alloc(h, &h->d_dmB, (size_t)h->C * 2));
        SV_TRY(sv_alloc(h, &h->d_linA, (size_t)h->C * SV_LINA_STRIDE));
        SV_TRY(sv_alloc(h, &h->d_linB, (size_t)h->C * SV_LINB_STRIDE));
        SV_TRY(sv_alloc(h, &h->d_cellT, (size_t)h->n * h->C * 4));
        cudaFree(h->d_lgx1);
        cudaFree(h->d_cellT2);
        h->d_lgx1 = nullptr;
        h->d_cellT2 = nullptr;
        if (h->single_ado && !h->private_cov) SV_TRY(sv_alloc(h, &h->d_cellT2, (size_t)h->n * h->C));
        SV_TRY(sv_alloc(h, &h->d_lgx1, (size_t)h->C));
        k_build_lgx1<<<(h->C + 127) / 128, 128, 0, h->stream>>>(h->d_lgx1, h->C);
        sv_count(h);
    }
    h->dm_dirty = h->cov_dirty = true;
    // one pass over the counts (ingest time): are the tables complete for this shard?
    int* d_flag = nullptr;
    int flag = 1;
    SV_CUDA(cudaMalloc(&d_flag, sizeof(int)));
    cudaError_t e = cudaMemsetAsync(d_flag, 0, sizeof(int), h->stream);
    if (e == cudaSuccess) {
        k_counts_outside_tables<<<592, 256, 0, h->stream>>>(h->d_counts, (size_t)h->n * h->S, h->C, d_flag);
        sv_count(h);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_flag);
    SV_CUDA(e);
    h->tables_cover_counts = flag == 0 && !h->leaf_force_checked;
    return SIEVE_OK;
}

static int sv_finish_counts(sieve_core* h, int* d_max) {
    int mx = 0;
    SV_CUDA(cudaMemcpyAsync(&mx, d_max, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    SV_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(d_max);
    h->max_cov = mx;
    h->have_counts = true;
    h->leaves_ready = false;
    return sv_alloc_tables(h);
}

extern "C" int sieve_set_counts(sieve_core_t* h, const int32_t* counts, int32_t tuple_len) {
    SV_REQUIRE(h && counts, "sieve_set_counts: null argument");
    SV_REQUIRE(tuple_len == 4 || tuple_len == 5, "sieve_set_counts: tuple_len must be 4 or 5");
    if (!h->counts_owned) return sv_fail(SIEVE_ERR_STATE, "sieve_set_counts: SIEVE_FLAG_EXTERNAL_COUNTS handles take sieve_set_counts_device only");
    SV_CUDA(cudaSetDevice(h->device));
    const size_t n = h->n, S = h->S;
    // bounded staging: at most ~256 MiB of the host tensor on the device at a time
    size_t chunk_sites = std::max<size_t>(32, std::min<size_t>(S, ((size_t)256 << 20) / (n * tuple_len * 4)));
    chunk_sites = (chunk_sites + 31) / 32 * 32;
    int* d_in = nullptr;
    int* d_max = nullptr;
    cudaError_t e = cudaMalloc(&d_in, chunk_sites * n * tuple_len * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&d_max, sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_max, 0, sizeof(int), h->stream);
    for (size_t s0 = 0; e == cudaSuccess && s0 < S; s0 += chunk_sites) {
        const size_t ns = std::min(chunk_sites, S - s0);
        e = cudaMemcpyAsync(d_in, counts + s0 * n * tuple_len, ns * n * tuple_len * sizeof(int), cudaMemcpyHostToDevice,
                            h->stream);
        if (e != cudaSuccess) break;
        dim3 g((unsigned)((ns + 31) / 32), (unsigned)((n + 31) / 32)), b(32, 8);
        k_transpose_counts<<<g, b, 0, h->stream>>>(d_in, tuple_len, (int)s0, (int)ns, (int)n, (int)S, h->d_counts, d_max);
        sv_count(h);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    }
    cudaFree(d_in);
    if (e != cudaSuccess) {
        cudaFree(d_max);
        SV_CUDA(e);
    }
    return sv_finish_counts(h, d_max);
}

// [cell][site][4] uint16 -> int4, tracking the largest coverage
__global__ void k_widen_counts_u16(const ushort4* __restrict__ in, size_t N, int4* __restrict__ out, int* __restrict__ max_cov) {
    int mx = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x) {
        const ushort4 v = in[i];
        out[i] = make_int4(v.x, v.y, v.z, v.w);
        mx = max(mx, (int)v.w);
    }
    for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_down_sync(0xffffffffu, mx, d));
    if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax(max_cov, mx);
}

// The ingest format's tensor (sieve_b200/ingest.py): host memory, already in the core's own order [cell][site][4], as
// int32 (dtype_code 0) or uint16 (dtype_code 1).  Copied in bounded chunks so that a memory-mapped file can be streamed.
extern "C" int sieve_set_counts_cellmajor(sieve_core_t* h, const void* counts, int32_t dtype_code) {
    SV_REQUIRE(h && counts, "sieve_set_counts_cellmajor: null argument");
    SV_REQUIRE(dtype_code == 0 || dtype_code == 1, "sieve_set_counts_cellmajor: dtype_code must be 0 (int32) or 1 (uint16)");
    if (!h->counts_owned) return sv_fail(SIEVE_ERR_STATE, "sieve_set_counts_cellmajor: SIEVE_FLAG_EXTERNAL_COUNTS handles take sieve_set_counts_device only");
    SV_CUDA(cudaSetDevice(h->device));
    const size_t N = (size_t)h->n * h->S;
    int* d_max = nullptr;
    ushort4* d_in = nullptr;
    cudaError_t e = cudaMalloc(&d_max, sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_max, 0, sizeof(int), h->stream);
    const size_t chunk = (size_t)16 << 20;  // tuples per chunk
    if (dtype_code == 0) {
        for (size_t i0 = 0; e == cudaSuccess && i0 < N; i0 += chunk) {
            const size_t m = std::min(chunk, N - i0);
            e = cudaMemcpyAsync(h->d_counts + i0, (const int4*)counts + i0, m * sizeof(int4), cudaMemcpyHostToDevice, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        }
        if (e == cudaSuccess) {
            k_max_cov<<<592, 256, 0, h->stream>>>(h->d_counts, N, d_max);
            sv_count(h);
        }
    } else {
        if (e == cudaSuccess) e = cudaMalloc(&d_in, chunk * sizeof(ushort4));
        for (size_t i0 = 0; e == cudaSuccess && i0 < N; i0 += chunk) {
            const size_t m = std::min(chunk, N - i0);
            e = cudaMemcpyAsync(d_in, (const ushort4*)counts + i0, m * sizeof(ushort4), cudaMemcpyHostToDevice, h->stream);
            if (e != cudaSuccess) break;
            k_widen_counts_u16<<<592, 256, 0, h->stream>>>(d_in, m, h->d_counts + i0, d_max);
            sv_count(h);
            e = cudaStreamSynchronize(h->stream);
        }
        cudaFree(d_in);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        cudaFree(d_max);
        SV_CUDA(e);
    }
    return sv_finish_counts(h, d_max);
}

extern "C" int sieve_set_counts_device(sieve_core_t* h, const void* d_counts) {
    SV_REQUIRE(h && d_counts, "sieve_set_counts_device: null argument");
    SV_CUDA(cudaSetDevice(h->device));
    const size_t N = (size_t)h->n * h->S;
    if (!h->counts_owned)
        h->d_counts = (int4*)const_cast<void*>(d_counts);  // SIEVE_FLAG_EXTERNAL_COUNTS: adopted, not copied
    else
        SV_CUDA(cudaMemcpyAsync(h->d_counts, d_counts, N * sizeof(int4), cudaMemcpyDeviceToDevice, h->stream));
    int* d_max = nullptr;
    SV_CUDA(cudaMalloc(&d_max, sizeof(int)));
    SV_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), h->stream));
    k_max_cov<<<592, 256, 0, h->stream>>>(h->d_counts, N, d_max);
    sv_count(h);
    SV_CUDA(cudaGetLastError());
    return sv_finish_counts(h, d_max);
}

extern "C" int sieve_set_pattern_weights(sieve_core_t* h, const int32_t* weights) {
    SV_REQUIRE(h, "null handle");
    SV_CUDA(cudaSetDevice(h->device));
    if (!weights) {
        cudaFree(h->d_weights);
        h->d_weights = nullptr;
        return SIEVE_OK;
    }
    if (!h->d_weights) SV_TRY(sv_alloc(h, &h->d_weights, (size_t)h->S));
    SV_CUDA(cudaMemcpyAsync(h->d_weights, weights, sizeof(int) * h->S, cudaMemcpyHostToDevice, h->stream));
    SV_CUDA(cudaStreamSynchronize(h->stream));
    return SIEVE_OK;
}

extern "C" int sieve_compute_size_factors(sieve_core_t* h, int32_t zero_cov_mode, double* out_sf, double* out_stats) {
    SV_REQUIRE(h, "null handle");
    SV_REQUIRE(zero_cov_mode == 0 || zero_cov_mode == 1, "zeroCovMode must be 0 or 1");
    if (!h->have_counts) return sv_fail(SIEVE_ERR_STATE, "sieve_compute_size_factors: counts not set");
    SV_CUDA(cudaSetDevice(h->device));
    unsigned long long l = 0;
    SV_CUDA(sv_compute_size_factors(h->d_counts, h->d_weights, h->n, h->S, zero_cov_mode, h->d_sf, out_stats, h->stream, &l));
    h->launches += l;
    h->step_launches += (int)l;
    if (out_sf) {
        SV_CUDA(cudaMemcpyAsync(out_sf, h->d_sf, sizeof(double) * h->n, cudaMemcpyDeviceToHost, h->stream));
        SV_CUDA(cudaStreamSynchronize(h->stream));
    }
    h->have_sf = true;
    h->cov_dirty = true;
    return SIEVE_OK;
}

extern "C" int sieve_set_size_factors(sieve_core_t* h, const double* sf) {
    SV_REQUIRE(h && sf, "null argument");
    SV_CUDA(cudaSetDevice(h->device));
    SV_CUDA(cudaMemcpyAsync(h->d_sf, sf, sizeof(double) * h->n, cudaMemcpyHostToDevice, h->stream));
    SV_CUDA(cudaStreamSynchronize(h->stream));
    h->have_sf = true;
    h->cov_dirty = true;
    return SIEVE_OK;
}

// ---- size factors over a site-sharded alignment (sizefactors_global.cuh) ------------------------------------------
extern "C" int sieve_sf_open(int32_t device, int32_t n_cells, int32_t n_sites, const void* d_counts, const int32_t* weights,
                             int32_t zero_cov_mode, sieve_sf_t** out) {
    return sv_sf_open(device, n_cells, n_sites, d_counts, weights, zero_cov_mode, out);
}
extern "C" int sieve_sf_open_core(sieve_core_t* h, int32_t zero_cov_mode, sieve_sf_t** out) {
    SV_REQUIRE(h && out, "sieve_sf_open_core: null argument");
    if (!h->have_counts) return sv_fail(SIEVE_ERR_STATE, "sieve_sf_open_core: counts not set");
    std::vector<int32_t> w;
    if (h->d_weights) {
        w.resize(h->S);
        SV_CUDA(cudaSetDevice(h->device));
        SV_CUDA(cudaMemcpy(w.data(), h->d_weights, sizeof(int) * (size_t)h->S, cudaMemcpyDeviceToHost));
    }
    return sv_sf_open(h->device, h->n, h->S, h->d_counts, h->d_weights ? w.data() : nullptr, zero_cov_mode, out);
}
extern "C" int sieve_sf_close(sieve_sf_t* p) {
    if (!p) return SIEVE_OK;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    cudaFree(p->d_geo);
    cudaFree(p->d_mom);
    cudaFree(p->d_piv);
    cudaFree(p->d_cnt);
    cudaFree(p->d_weights);
    if (p->stream) cudaStreamDestroy(p->stream);
    cudaGetLastError();
    delete p;
    return SIEVE_OK;
}
extern "C" int sieve_sf_moments(sieve_sf_t* p, uint64_t* out3) {
    SV_REQUIRE(p && out3, "sieve_sf_moments: null argument");
    out3[0] = p->mom[0];
    out3[1] = p->mom[1];
    out3[2] = (uint64_t)p->n * (uint64_t)p->S;
    return SIEVE_OK;
}
extern "C" int sieve_sf_weight_le(sieve_sf_t* p, int32_t n_piv, const uint64_t* pivots, int64_t* out) {
    return sv_sf_weight_le(p, n_piv, pivots, out);
}
extern "C" int sieve_sf_solve(const sieve_sf_ops_t* ops, double* out_sf, double* out_stats) {
    return sv_sf_solve(ops, out_sf, out_stats);
}
extern "C" int sieve_sf_solve_parts(int32_t n_parts, sieve_sf_t* const* parts, sieve_sf_allreduce_fn ar, void* user,
                                    double* out_sf, double* out_stats) {
    SV_REQUIRE(n_parts >= 1 && parts && out_sf, "sieve_sf_solve_parts: bad argument");
    for (int i = 0; i < n_parts; i++) {
        SV_REQUIRE(parts[i], "sieve_sf_solve_parts: null part");
        SV_REQUIRE(parts[i]->n == parts[0]->n && parts[i]->zero_cov_mode == parts[0]->zero_cov_mode,
                   "sieve_sf_solve_parts: the parts disagree on the number of cells or on zeroCovMode");
    }
    SvSfParts P{n_parts, parts, ar, user, {}};
    sieve_sf_ops_t ops;
    ops.user = &P;
    ops.n_cells = parts[0]->n;
    ops.moments = sv_sf_parts_moments;
    ops.weight_le = sv_sf_parts_weight_le;
    ops.allreduce_sum_i64 = sv_sf_parts_allreduce;
    return sv_sf_solve(&ops, out_sf, out_stats);
}

// ---- PrivateAllelicSeqCovModel: (t, v) per (category, site) -----------------------------------------------------------
extern "C" int sieve_set_private_seq_cov(sieve_core_t* h, const double* t_ks, const double* v_ks) {
    SV_REQUIRE(h && t_ks && v_ks, "sieve_set_private_seq_cov: null argument");
    if (!h->private_cov) return sv_fail(SIEVE_ERR_STATE, "sieve_set_private_seq_cov: the core was not created with SIEVE_FLAG_PRIVATE_SEQ_COV");
    SV_CUDA(cudaSetDevice(h->device));
    const size_t S = h->S;
    std::vector<double2> tv((size_t)SV_KMAX * S, make_double2(1.0, 1.0));
    for (int k = 0; k < h->K; k++)
        for (size_t s = 0; s < S; s++) {
            const double t = t_ks[(size_t)k * S + s], v = v_ks[(size_t)k * S + s];
            if (!(t >= 0.0) || !(v >= 0.0))
                return sv_fail(SIEVE_ERR_INVALID, "sieve_set_private_seq_cov: allelicSeqCov and allelicSeqCovRawVar must not be negative");
            tv[(size_t)k * S + s] = make_double2(t, v);
        }
    SV_CUDA(cudaStreamSynchronize(h->stream));
    SV_CUDA(cudaMemcpy(h->d_tv, tv.data(), tv.size() * sizeof(double2), cudaMemcpyHostToDevice));
    h->have_tv = true;
    return SIEVE_OK;
}
extern "C" int sieve_get_private_seq_cov(sieve_core_t* h, double* t_ks, double* v_ks) {
    SV_REQUIRE(h && t_ks && v_ks, "sieve_get_private_seq_cov: null argument");
    if (!h->private_cov || !h->have_tv) return sv_fail(SIEVE_ERR_STATE, "sieve_get_private_seq_cov: no per-pattern coverage parameters set");
    SV_CUDA(cudaSetDevice(h->device));
    const size_t S = h->S;
    std::vector<double2> tv((size_t)SV_KMAX * S);
    SV_CUDA(cudaStreamSynchronize(h->stream));
    SV_CUDA(cudaMemcpy(tv.data(), h->d_tv, tv.size() * sizeof(double2), cudaMemcpyDeviceToHost));
    for (int k = 0; k < h->K; k++)
        for (size_t s = 0; s < S; s++) {
            t_ks[(size_t)k * S + s] = tv[(size_t)k * S + s].x;
            v_ks[(size_t)k * S + s] = tv[(size_t)k * S + s].y;
        }
    return SIEVE_OK;
}

// ---- leaves ---------------------------------------------------------------------------------------------------
extern "C" int sieve_set_leaf_params(sieve_core_t* h, const sieve_leaf_params_t* p) {
    SV_REQUIRE(h && p, "null argument");
    h->lp = *p;
    h->have_params = true;
    return SIEVE_OK;
}

// brings the look-up tables of the leaf kernel up to date with h->lp
// Do the linear-space tables hold every entry?  |LP(c, alpha)| grows monotonically with c, so the last row decides (and the
// first, for a tiny alpha); a non-positive alpha makes the entries NaN, which the packed rows cannot carry (the exponent
// field holds the binary exponent): the log-space kernel runs then and propagates the NaN as the reference does.
static bool sv_lin_tables_ok(const SvDmAlphas& al, int C) {
    const double alphas[6] = {al.fw1, al.omw1, al.hw2, al.fw2, al.w1, al.w2};
    for (double a : alphas) {
        if (!(a > 0.0)) return false;
        for (double c : {1.0, (double)std::max(C - 1, 1)}) {
            const double lp = std::log(c) + std::lgamma(a) + std::lgamma(c) - std::lgamma(a + c);
            if (!(std::fabs(lp) <= SV_LIN_MAX_LOG)) return false;
        }
    }
    return true;
}
// Does the default leaf kernel (k_leaf_lin) apply?  Every count inside the tables, single dropout, constant genotype 0, the
// tables in range and small enough for two blocks per SM to stage them in shared memory.
static bool sv_leaf_lin_ok(const sieve_core* h) {
    return h->leaf_lin_allowed && h->tables_cover_counts && h->single_ado && h->d_cellT2 && h->const_genotype == 0 && h->lin_range_ok &&
           h->leaf_tab == 1 && (size_t)h->C * (SV_LINA_STRIDE + SV_LINB_STRIDE) * sizeof(double) <= 100 * 1024;
}

static int sv_refresh_leaf_tables(sieve_core* h, SvDmAlphas* al_out, SvCovParams* cp_out) {
    const SvDmAlphas al = sv_dm_alphas(h->lp.eff_seq_err_rate, h->lp.shape_ctrl1, h->lp.shape_ctrl2);
    SvCovParams cp;
    cp.t = h->lp.allelic_seq_cov;
    cp.v = h->lp.allelic_seq_cov_raw_var;
    cp.theta = h->lp.ado_rate;
    cp.single_ado = h->single_ado ? 1 : 0;
    const int C = h->C;
    // tables are rebuilt only for the sub-model whose parameters moved (the reference's isDirtyCalculation checks,
    // rawreadcountsmodel/RawReadCountsModelInterface.java:476-490)
    if (h->lp.eff_seq_err_rate != h->tab_lp.eff_seq_err_rate || h->lp.shape_ctrl1 != h->tab_lp.shape_ctrl1 ||
        h->lp.shape_ctrl2 != h->tab_lp.shape_ctrl2)
        h->dm_dirty = true;
    if (h->lp.allelic_seq_cov != h->tab_lp.allelic_seq_cov ||
        h->lp.allelic_seq_cov_raw_var != h->tab_lp.allelic_seq_cov_raw_var || h->lp.ado_rate != h->tab_lp.ado_rate)
        h->cov_dirty = true;
    h->tab_lp = h->lp;
    if (h->dm_dirty) {
        k_build_dm_tables<<<(6 * C + 127) / 128, 128, 0, h->stream>>>(h->d_dmA, h->d_dmB, h->d_linA, h->d_linB, C, al);
        sv_count(h);
        h->dm_dirty = false;
        h->lin_range_ok = sv_lin_tables_ok(al, C);
    }
    if (h->cov_dirty && h->private_cov) h->cov_dirty = false;  // the coverage terms are evaluated directly (leaf_private.cuh)
    if (h->cov_dirty) {
        dim3 g((C + 127) / 128, h->n);
        k_build_cell_tables<<<g, 128, 0, h->stream>>>(h->d_cellT, h->d_sf, h->n, C, cp, h->d_lgx1, h->d_cellT2);
        sv_count(h);
        h->cov_dirty = false;
    }
    SV_CUDA(cudaGetLastError());
    *al_out = al;
    *cp_out = cp;
    return SIEVE_OK;
}

// One k_leaf + k_colsum pair over `n_entries` cells (d_cell_list NULL: all cells in index order) and the site range
// [site_begin, site_begin + n_sites): leaf likelihoods into leafX (NULL: none), the sums of the leaves' log scales into
// lsum_out[0 .. n_sites) and, when lambda_out is given, of their constant-genotype log-likelihoods into lambda_out.
// sumS / sum0: scratch for the per-cell-group partial sums, ceil(n_entries / SV_LEAF_CELLS) x sum_stride doubles each.
static int sv_launch_leaf_kernels(sieve_core* h, long long site_begin, int n_sites, long long out_stride, double* leafX,
                                  const int* d_cell_list, int n_entries, double* sumS, double* sum0, long long sum_stride,
                                  double* lsum_out, double* lambda_out) {
    SvLeafArgs a;
    SV_TRY(sv_refresh_leaf_tables(h, &a.al, &a.cp));
    const int C = h->C;
    a.counts = h->d_counts;
    a.dmA = h->d_dmA;
    a.dmB = h->d_dmB;
    a.cellT = h->d_cellT;
    a.cellT2 = h->d_cellT2;
    a.linA = h->d_linA;
    a.linB = h->d_linB;
    a.size_factors = h->d_sf;
    a.cell_list = d_cell_list;
    a.leafX = leafX;
    a.sumS = sumS;
    a.sum0 = lambda_out ? sum0 : nullptr;
    a.const_genotype = h->const_genotype;
    a.n_cells = n_entries;
    a.S = n_sites;
    a.counts_stride = h->S;
    a.site_begin = site_begin;
    a.out_stride = out_stride;
    a.sum_stride = sum_stride;
    a.C = C;
    const int groups = (n_entries + SV_LEAF_CELLS - 1) / SV_LEAF_CELLS;
    // SIEVE_B200_LEAF_TAB: 1 (default) Dirichlet-multinomial tables in shared memory, cell table through L1; 2 cell table
    // double-buffered in shared memory; 0 (forced when the tables do not fit) everything through L1
    int tab = h->leaf_tab;
    if ((size_t)C * 14 * sizeof(double) > 100 * 1024 && tab == 2) tab = 1;
    if ((size_t)C * 6 * sizeof(double) > 100 * 1024) tab = 0;
    const size_t smem = (size_t)C * (tab == 2 ? 14 : tab == 1 ? 6 : 0) * sizeof(double);
    dim3 grid((n_sites + SV_LEAF_SITES - 1) / SV_LEAF_SITES, groups);
    if (!h->leaf_attr_set) {
        SV_CUDA(cudaFuncSetAttribute(k_leaf<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        SV_CUDA(cudaFuncSetAttribute(k_leaf<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        SV_CUDA(cudaFuncSetAttribute(k_leaf<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        SV_CUDA(cudaFuncSetAttribute(k_leaf_lin, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        h->leaf_attr_set = true;
    }
    if (sv_leaf_lin_ok(h)) {
        dim3 gl((n_sites + SV_LEAFL_SITES - 1) / SV_LEAFL_SITES, groups);
        k_leaf_lin<<<gl, SV_LEAFL_THREADS, (size_t)C * (SV_LINA_STRIDE + SV_LINB_STRIDE) * sizeof(double), h->stream>>>(a);
    }
    else if (tab == 2)
        k_leaf<2><<<grid, SV_LEAF_THREADS, smem, h->stream>>>(a);
    else if (tab == 1 && h->tables_cover_counts)
        k_leaf<1, true><<<grid, SV_LEAF_THREADS, smem, h->stream>>>(a);
    else if (tab == 1)
        k_leaf<1><<<grid, SV_LEAF_THREADS, smem, h->stream>>>(a);
    else
        k_leaf<0><<<grid, SV_LEAF_THREADS, 0, h->stream>>>(a);
    sv_count(h);
    k_colsum<<<(n_sites + 255) / 256, 256, 0, h->stream>>>(sumS, a.sum0, groups, n_sites, sum_stride, lsum_out, lambda_out);
    sv_count(h);
    SV_CUDA(cudaGetLastError());
    return SIEVE_OK;
}

// the same for a core with the per-pattern coverage model: k_leaf_private + k_colsum_private; lambdaK_out is [KMAX][lambda_stride]
// d_site_list (or NULL: the whole shard): only these n_list sites are evaluated and written, everything else stays as it is
static int sv_launch_leaf_private(sieve_core* h, double* leafXK, const int* d_cell_list, int n_entries, double* sumS,
                                  double* sum0K, double* lsum_out, double* lambdaK_out, long long lambda_stride,
                                  const int* d_site_list = nullptr, int n_list = 0) {
    if (!h->have_tv) return sv_fail(SIEVE_ERR_STATE, "sieve_set_private_seq_cov must be called before the leaves are computed");
    SvLeafPrivArgs a;
    SvCovParams cp_unused;
    SV_TRY(sv_refresh_leaf_tables(h, &a.al, &cp_unused));
    a.counts = h->d_counts;
    a.dmA = h->d_dmA;
    a.dmB = h->d_dmB;
    a.size_factors = h->d_sf;
    a.cell_list = d_cell_list;
    a.site_list = d_site_list;
    a.n_list = n_list;
    a.tv = h->d_tv;
    a.leafXK = leafXK;
    a.sumS = sumS;
    a.sum0 = lambdaK_out ? sum0K : nullptr;
    a.const_genotype = h->const_genotype;
    a.n_cells = n_entries;
    a.S = h->S;
    a.C = h->C;
    a.K = h->K;
    a.counts_stride = h->S;
    a.site_begin = 0;
    a.out_stride = h->S;
    a.sum_stride = h->S;
    a.tv_stride = h->S;
    a.theta = h->lp.ado_rate;
    a.single_ado = h->single_ado ? 1 : 0;
    const int groups = (n_entries + SV_LEAF_CELLS - 1) / SV_LEAF_CELLS;
    const size_t smem = (size_t)h->C * 6 * sizeof(double) <= 96 * 1024 ? (size_t)h->C * 6 * sizeof(double) : 0;
    static bool attr_set = false;
    if (!attr_set) {
        SV_CUDA(cudaFuncSetAttribute(k_leaf_private, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_set = true;
    }
    const int n_eval = d_site_list ? n_list : h->S;
    dim3 grid((n_eval + SV_LEAFP_THREADS - 1) / SV_LEAFP_THREADS, groups);
    k_leaf_private<<<grid, SV_LEAFP_THREADS, smem, h->stream>>>(a);
    sv_count(h);
    k_colsum_private<<<(n_eval + 255) / 256, 256, 0, h->stream>>>(sumS, a.sum0, groups, h->S, h->K, h->S, lsum_out, lambdaK_out,
                                                                 lambda_stride, d_site_list, n_list);
    sv_count(h);
    SV_CUDA(cudaGetLastError());
    if (leafXK && !d_cell_list) {
        // the leaf buffer now holds the current pairs (for all sites: an unlisted site's pair did not move)
        SV_CUDA(cudaMemcpyAsync(h->d_tv_leaf, h->d_tv, sizeof(double2) * SV_KMAX * h->S, cudaMemcpyDeviceToDevice, h->stream));
        h->tv_leaf_ok = true;
    }
    return SIEVE_OK;
}

// site_begin / n_sites: the range of the shard to evaluate; out_stride: site capacity of the leaf pool it goes to
static int sv_launch_leaves(sieve_core* h, long long site_begin, int n_sites, long long out_stride) {
    const bool algebraic = h->with_const && !h->const_twin;
    if (h->private_cov)
        return sv_launch_leaf_private(h, h->d_leafX + (size_t)h->cur_leaf * h->n * h->S * SV_KMAX * 4, nullptr, h->n, h->d_sumS,
                                      h->d_sum0K, h->d_lsum + (size_t)h->cur_leaf * h->S,
                                      algebraic ? h->d_lambdaK + (size_t)h->cur_leaf * SV_KMAX * h->S : nullptr, h->S);
    // the per-site sums: streamed mode indexes them by the shard's site, resident mode per leaf buffer
    const long long sum_at = h->ephemeral ? site_begin : (long long)h->cur_leaf * h->S;
    return sv_launch_leaf_kernels(h, site_begin, n_sites, out_stride,
                                  h->d_leafX + (h->ephemeral ? 0 : (size_t)h->cur_leaf * h->n * h->S * 4), nullptr, h->n,
                                  h->d_sumS, h->d_sum0, out_stride, h->d_lsum + sum_at,
                                  algebraic ? h->d_lambda + sum_at : nullptr);
}

// the sums over the leaves below `node` (the node itself when it is a leaf) of the leaves' log scales and, when lam is
// given, of their constant-genotype log-likelihoods: what a read-back adds to the node's own scale.  d_ls / d_lam: [S].
static int sv_subtree_leaf_sums(sieve_core* h, int node, double* d_ls, double* d_lam) {
    std::vector<int> leaves, st;
    st.push_back(node);
    while (!st.empty()) {
        const int v = st.back();
        st.pop_back();
        if (v < h->n) {
            leaves.push_back(v);
            continue;
        }
        if (h->eph_c1[v] == -2) return sv_fail(SIEVE_ERR_STATE, "node was never computed");
        st.push_back(h->eph_c1[v]);
        if (h->eph_c2[v] >= 0) st.push_back(h->eph_c2[v]);
    }
    std::sort(leaves.begin(), leaves.end());
    const size_t S = h->S;
    const size_t groups = (leaves.size() + SV_LEAF_CELLS - 1) / SV_LEAF_CELLS;
    int* d_list = nullptr;
    double* d_sums = nullptr;
    cudaError_t e = cudaMalloc(&d_list, sizeof(int) * leaves.size());
    if (e == cudaSuccess) e = cudaMalloc(&d_sums, sizeof(double) * (1 + (h->private_cov ? SV_KMAX : 1)) * groups * S);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_list, leaves.data(), sizeof(int) * leaves.size(), cudaMemcpyHostToDevice, h->stream);
    int r = SIEVE_OK;
    if (e == cudaSuccess) {
        // the sums must belong to the leaf buffer that is read back: evaluate them with the parameters THAT buffer was
        // computed with, not with parameters set since (sieve_set_leaf_params without sieve_update_leaves)
        const sieve_leaf_params_t pending = h->lp;
        h->lp = h->leaf_lp[h->cur_leaf];
        if (h->private_cov)  // d_lam is [KMAX][S] here
            r = sv_launch_leaf_private(h, nullptr, d_list, (int)leaves.size(), d_sums, d_sums + groups * S, d_ls, d_lam, (long long)S);
        else
            r = sv_launch_leaf_kernels(h, 0, (int)S, (long long)S, nullptr, d_list, (int)leaves.size(), d_sums,
                                       d_sums + groups * S, (long long)S, d_ls, d_lam);
        h->lp = pending;
    }
    if (e == cudaSuccess && r == SIEVE_OK) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_list);
    cudaFree(d_sums);
    if (r != SIEVE_OK) return r;
    SV_CUDA(e);
    return SIEVE_OK;
}

extern "C" int sieve_update_leaves(sieve_core_t* h, int32_t for_update) {
    SV_REQUIRE(h, "null handle");
    if (!h->have_counts || !h->have_sf || !h->have_params)
        return sv_fail(SIEVE_ERR_STATE, "sieve_update_leaves: counts, size factors and leaf parameters must be set first");
    SV_CUDA(cudaSetDevice(h->device));
    if (h->ephemeral) {
        // streamed mode: the leaves are recomputed chunk by chunk inside every evaluation; nothing to do now
        h->leaves_ready = true;
        h->eph_leaves_dirty = true;
        h->leaf_lp[0] = h->leaf_lp[1] = h->lp;
        return SIEVE_OK;
    }
    if (for_update && !h->single_leaf) h->cur_leaf = 1 - h->cur_leaf;
    h->ml_ready = false;
    h->leaf_ver[h->cur_leaf] = ++h->ver_counter;
    if (!for_update) h->leaf_ver[1 - h->cur_leaf] = h->leaf_ver[h->cur_leaf];
    if (h->single_leaf) h->leaf_changed_since_store = true;
    if (h->timing) {
        SV_CUDA(cudaEventRecord(h->ev[0], h->stream));
    }
    SV_TRY(sv_launch_leaves(h, 0, h->S, h->S));
    h->leaf_lp[h->cur_leaf] = h->lp;
    if (!for_update) h->leaf_lp[1 - h->cur_leaf] = h->lp;
    if (h->timing) {
        SV_CUDA(cudaEventRecord(h->ev[1], h->stream));
        h->leaf_timed = true;
    }
    if (!for_update && !h->single_leaf) {
        // initialisation: both buffers hold the leaves (storeLeafPartials, ScsBeerLikelihoodCore.java:437-445)
        const size_t nx = (size_t)h->n * h->S * (h->private_cov ? SV_KMAX * 4 : 4);
        if (h->d_lambdaK)
            SV_CUDA(cudaMemcpyAsync(h->d_lambdaK + (size_t)(1 - h->cur_leaf) * SV_KMAX * h->S,
                                    h->d_lambdaK + (size_t)h->cur_leaf * SV_KMAX * h->S,
                                    (size_t)SV_KMAX * h->S * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        SV_CUDA(cudaMemcpyAsync(h->d_leafX + (size_t)(1 - h->cur_leaf) * nx, h->d_leafX + (size_t)h->cur_leaf * nx,
                                nx * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        SV_CUDA(cudaMemcpyAsync(h->d_lsum + (size_t)(1 - h->cur_leaf) * h->S, h->d_lsum + (size_t)h->cur_leaf * h->S,
                                (size_t)h->S * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        if (h->d_lambda)
            SV_CUDA(cudaMemcpyAsync(h->d_lambda + (size_t)(1 - h->cur_leaf) * h->S, h->d_lambda + (size_t)h->cur_leaf * h->S,
                                    (size_t)h->S * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
    h->leaves_ready = true;
    return SIEVE_OK;
}

// ---- matrices -------------------------------------------------------------------------------------------------
// staging [count][K][16] -> pool [slot][4][16]; the Double.MIN_VALUE clamp of the log-space reference
// (ScsBeerLikelihoodCore.java:1430: log(P > 0 ? P : MIN_VALUE)) is applied here, once per uploaded matrix
__global__ void k_scatter_matrices(const double* __restrict__ stage, const int* __restrict__ slots, int count, int K,
                                   int clamp, double* __restrict__ mats) {
    const int i = blockIdx.x;
    const int e = threadIdx.x;
    if (i < count && e < K * 16) {
        double v = stage[(size_t)i * K * 16 + e];
        if (clamp) v = v > 0.0 ? v : SV_MIN_VALUE;
        mats[(size_t)slots[i] * (SV_KMAX * 16) + e] = v;
    }
}

// Branch distances -> (g1/8, g2/16) for the Q-specialised walk, and the full matrices P = I + g1 Pi1 + g2 Pi2 for the
// generic kernel and the getters.  One thread per (entry, category).
__global__ void k_scatter_distances(const double* __restrict__ stage, const int* __restrict__ slots, int count, int K,
                                    int clamp, double2* __restrict__ qfac, double* __restrict__ mats) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count * K) return;
    const int i = idx / K, k = idx % K;
    const double d = stage[(size_t)i * K + k];
    const double g1 = expm1(-0.2 * d), g2 = expm1(-0.4 * d);
    qfac[(size_t)slots[i] * SV_KMAX + k] = make_double2(g1 / 8.0, g2 / 16.0);
    const double P1[16] = {3, 6, -3, -6, 1, 2, -1, -2, -1, -2, 1, 2, -1, -2, 1, 2};          // 8 Pi1
    const double P2[16] = {9, -18, 3, 6, -3, 6, -1, -2, 1, -2, 11, -10, 1, -2, -5, 6};       // 16 Pi2
    double* out = mats + (size_t)slots[i] * (SV_KMAX * 16) + k * 16;
    for (int e = 0; e < 16; e++) {
        double v = ((e % 5 == 0) ? 1.0 : 0.0) + (g1 / 8.0) * P1[e] + (g2 / 16.0) * P2[e];
        if (d == 0.0) v = (e % 5 == 0) ? 1.0 : 0.0;
        if (clamp) v = v > 0.0 ? v : SV_MIN_VALUE;
        out[e] = v;
    }
}

static bool sv_node_ok(const sieve_core* h, int node) { return node >= 0 && node < h->n_nodes; }

extern "C" int sieve_set_node_matrix_for_update(sieve_core_t* h, int32_t node) {
    SV_REQUIRE(h && sv_node_ok(h, node), "sieve_set_node_matrix_for_update: bad node");
    h->cur_mat[node] = 1 - h->cur_mat[node];
    return SIEVE_OK;
}

static PendingMatrix* sv_pending_matrix(sieve_core* h, int slot) {
    // (an earlier DISTANCE entry for the slot only has room for K doubles: it is superseded by a new matrix entry)
    for (auto it = h->pending_mats.rbegin(); it != h->pending_mats.rend(); ++it)
        if (it->slot == slot) {
            if (it->is_distance) break;
            return &*it;
        }
    PendingMatrix pm;
    pm.slot = slot;
    pm.is_distance = 0;
    pm.off = h->pending_vals.size();
    h->pending_vals.resize(pm.off + (size_t)SV_KMAX * 16, 0.0);
    h->pending_mats.push_back(pm);
    return &h->pending_mats.back();
}

extern "C" int sieve_set_node_matrix(sieve_core_t* h, int32_t node, int32_t category, const double* p16) {
    SV_REQUIRE(h && p16 && sv_node_ok(h, node), "sieve_set_node_matrix: bad argument");
    SV_REQUIRE(category >= 0 && category < h->K, "sieve_set_node_matrix: bad category");
    // the K categories of one node arrive one by one (ScsTreeLikelihood.java:567-578); they share a pending entry
    PendingMatrix* pm = nullptr;
    const int slot = h->cur_mat[node] * h->n_nodes + node;
    if (!h->pending_mats.empty() && h->pending_mats.back().slot == slot && !h->pending_mats.back().is_distance)
        pm = &h->pending_mats.back();
    else
        pm = sv_pending_matrix(h, slot);
    pm->is_distance = 0;
    memcpy(h->pending_vals.data() + pm->off + category * 16, p16, 16 * sizeof(double));
    return SIEVE_OK;
}

extern "C" int sieve_set_matrices(sieve_core_t* h, int32_t count, const int32_t* nodes, const double* P, int32_t for_update) {
    SV_REQUIRE(h && count >= 0 && (count == 0 || (nodes && P)), "sieve_set_matrices: bad argument");
    const int K = h->K;
    for (int i = 0; i < count; i++) SV_REQUIRE(sv_node_ok(h, nodes[i]), "sieve_set_matrices: bad node");
    h->pending_mats.reserve(h->pending_mats.size() + count);
    for (int i = 0; i < count; i++) {
        if (for_update) h->cur_mat[nodes[i]] = 1 - h->cur_mat[nodes[i]];
        PendingMatrix pm;
        pm.slot = h->cur_mat[nodes[i]] * h->n_nodes + nodes[i];
        pm.is_distance = 0;
        pm.off = h->pending_vals.size();
        h->pending_vals.insert(h->pending_vals.end(), P + (size_t)i * K * 16, P + (size_t)(i + 1) * K * 16);
        h->pending_mats.push_back(pm);
    }
    return SIEVE_OK;
}

extern "C" int sieve_set_distances(sieve_core_t* h, int32_t count, const int32_t* nodes, const double* distances,
                                   int32_t for_update) {
    SV_REQUIRE(h && count >= 0 && (count == 0 || (nodes && distances)), "sieve_set_distances: bad argument");
    const int K = h->K;
    // every input is checked before any buffer index or queue is touched: a rejected call leaves the handle as it was
    for (int i = 0; i < count; i++) {
        SV_REQUIRE(sv_node_ok(h, nodes[i]), "sieve_set_distances: bad node");
        for (int k = 0; k < K; k++)
            if (!(distances[(size_t)i * K + k] >= 0.0)) return sv_fail(SIEVE_ERR_INVALID, "sieve_set_distances: negative distance");
    }
    h->pending_mats.reserve(h->pending_mats.size() + count);
    for (int i = 0; i < count; i++) {
        if (for_update) h->cur_mat[nodes[i]] = 1 - h->cur_mat[nodes[i]];
        PendingMatrix pm;
        pm.slot = h->cur_mat[nodes[i]] * h->n_nodes + nodes[i];
        pm.is_distance = 1;
        pm.off = h->pending_vals.size();
        h->pending_vals.insert(h->pending_vals.end(), distances + (size_t)i * K, distances + (size_t)(i + 1) * K);
        h->pending_mats.push_back(pm);
    }
    return SIEVE_OK;
}

// ---- ops ------------------------------------------------------------------------------------------------------
extern "C" int sieve_set_node_partials_for_update(sieve_core_t* h, int32_t node) {
    SV_REQUIRE(h && sv_node_ok(h, node), "sieve_set_node_partials_for_update: bad node");
    if (node < h->n) return sv_fail(SIEVE_ERR_INVALID, "leaves flip together: use sieve_update_leaves(for_update = 1)");
    if (h->ephemeral) return SIEVE_OK;  // no resident partials to flip
    h->cur_part[node] = 1 - h->cur_part[node];
    return SIEVE_OK;
}

#include "core_planner.cuh"

extern "C" int sieve_calculate_partials(sieve_core_t* h, int32_t child1, int32_t is_leaf1, int32_t child2, int32_t is_leaf2,
                                        int32_t parent, int32_t const_genotype) {
    SV_REQUIRE(h, "null handle");
    SV_REQUIRE(const_genotype >= 0 && const_genotype < 4, "bad const genotype");
    SV_REQUIRE((child1 < h->n) == (is_leaf1 != 0), "sieve_calculate_partials: is_leaf1 contradicts the node numbering");
    SV_REQUIRE(child2 < 0 || ((child2 < h->n) == (is_leaf2 != 0)),
               "sieve_calculate_partials: is_leaf2 contradicts the node numbering");
    h->const_genotype = const_genotype;
    SV_TRY(sv_record_children(h, child1, child2, parent));
    if (h->ephemeral) return SIEVE_OK;
    if (h->pending_per_level) return sv_fail(SIEVE_ERR_STATE, "sieve_calculate_partials: a per-level update is pending");
    h->dirty_list.push_back(parent);
    return SIEVE_OK;
}

extern "C" int sieve_update_nodes(sieve_core_t* h, int32_t n_ops, const int32_t* ops, const int32_t* level_offsets,
                                  int32_t n_levels, uint32_t mode) {
    SV_REQUIRE(h && n_ops >= 0 && (n_ops == 0 || ops), "sieve_update_nodes: bad argument");
    if (h->ephemeral) {
        // streamed mode: every evaluation walks the whole tree; an update only edits the parent -> children map
        for (int i = 0; i < n_ops; i++) SV_TRY(sv_record_children(h, ops[3 * i + 1], ops[3 * i + 2], ops[3 * i]));
        return SIEVE_OK;
    }
    const bool per_level = (mode & SIEVE_UPDATE_PER_LEVEL) && level_offsets && n_levels > 0;
    if (per_level && h->sparse) return sv_fail(SIEVE_ERR_STATE, "sieve_update_nodes: the per-level schedule needs a dense core");
    if (per_level) {
        SV_REQUIRE(h->pending_ops.empty() && !h->pending_per_level, "sieve_update_nodes: per-level update with other ops pending");
        SV_REQUIRE(level_offsets[0] == 0 && level_offsets[n_levels] == n_ops, "sieve_update_nodes: bad level offsets");
    }
    if (mode & SIEVE_UPDATE_FLIP)
        for (int i = 0; i < n_ops; i++) {
            const int p = ops[3 * i];
            SV_REQUIRE(p >= h->n && p < h->n_nodes, "sieve_update_nodes: parent must be an internal node");
            h->cur_part[p] = 1 - h->cur_part[p];
        }
    for (int i = 0; i < n_ops; i++) SV_TRY(sv_record_children(h, ops[3 * i + 1], ops[3 * i + 2], ops[3 * i]));
    if (!per_level) {
        // the planner orders the walk (sv_plan_resident_ops); the list only names the dirty nodes
        if (h->pending_per_level) return sv_fail(SIEVE_ERR_STATE, "sieve_update_nodes: a per-level update is pending");
        for (int i = 0; i < n_ops; i++) h->dirty_list.push_back(ops[3 * i]);
        return SIEVE_OK;
    }
    // level-batched schedule: the caller's order and grouping are kept, every result is stored.  Only the (parent, child1,
    // child2) triples are queued here: the ops themselves are made by sv_flush AFTER its matrix pass, because that pass
    // may still retarget a node's matrix buffer (change tracking) and decides the zero-distance flags the ops carry.
    SV_REQUIRE(h->dirty_list.empty(), "sieve_update_nodes: per-level update with other ops pending");
    h->pending_level_triples.assign(ops, ops + (size_t)3 * n_ops);
    for (int i = 0; i < n_ops; i++) {
        const int p = ops[3 * i];
        SV_REQUIRE(p >= h->n && p < h->n_nodes, "sieve_update_nodes: parent must be an internal node");
        h->g_nodes.push_back(p);
        h->valid[p] = 1;
        h->ver_cur[p] = ++h->ver_counter;  // computed outside the planner: a new value of unknown provenance
        h->fp_cur[p] = sieve_core::SvFp();
        h->prov_touched.push_back(p);
        {
            const size_t bi = (size_t)h->cur_part[p] * h->n_nodes + p;
            h->buf_has[bi] = 1;
            h->buf_ver[bi] = h->ver_cur[p];
            h->buf_fp[bi] = sieve_core::SvFp();
        }
    }
    h->pending_levels.assign(level_offsets, level_offsets + n_levels + 1);
    h->pending_per_level = true;
    return SIEVE_OK;
}

extern "C" int sieve_set_root(sieve_core_t* h, int32_t root, const double* proportions, int32_t root_genotype) {
    SV_REQUIRE(h && proportions, "null argument");
    SV_REQUIRE(root >= h->n && root < h->n_nodes, "sieve_set_root: the root must be an internal node");
    SV_REQUIRE(root_genotype >= 0 && root_genotype < 4, "sieve_set_root: bad root genotype");
    h->root = root;
    h->root_genotype = root_genotype;
    for (int k = 0; k < 4; k++) h->prop[k] = k < h->K ? proportions[k] : 0.0;
    return SIEVE_OK;
}

// Resident blocks per SM for the variant-only Q walk.  Every block walks the whole op list, so a launch takes as long as
// its waves of blocks: full waves cost their share, the last partial wave (a fraction phi of the slots) costs about
// 0.3 + 0.65 phi of a full one (its blocks run faster on a less crowded SM, measured at 6 blocks/SM over 1.0 .. 3.0
// waves).  The kernel compiles to 78 registers (6 blocks/SM); capped at 72 / 64 registers it spills 8 / 48 bytes and
// fits 7 / 8 blocks, which at whole waves is 3 / 6 % slower per site -- but what decides is how the launch's blocks
// divide into waves: 100 000 sites are 1.76 / 1.51 / 1.32 waves (6 wins: 2.50 / 2.66 / 2.86 ms), 125 000 sites 2.20 /
// 1.89 / 1.65 (7 wins: 3.29 / 3.19 / 3.50 ms).  Same PTX, same arithmetic: the choice cannot change a bit of the results.
// profiles/README.md has the measurements.
static int sv_pick_prune_blocks(unsigned warps) {
    // the model was fitted in units of four warps (64 sites): resident units per SM = 6, 7 or 8
    const unsigned blocks = (warps + 3) / 4;
    static const int m[3] = {6, 7, 8};
    static const double per_wave[3] = {6.0, 7.0 * 1.026, 8.0 * 1.058};
    int best = 6;
    double best_t = 0.0;
    if (blocks <= 148u * 6u) return best;  // everything is resident at once anyway
    for (int i = 0; i < 3; i++) {
        const double w = (double)blocks / (148.0 * m[i]);
        const double full = floor(w), phi = w - full;
        const double t = per_wave[i] * (full + (phi > 0.0 ? 0.3 + 0.65 * phi : 0.0));
        if (i == 0 || t < best_t * 0.99) {  // a higher block count has to win by more than 1 %
            best = m[i];
            best_t = t;
        }
    }
    return best;
}

extern "C" int sieve_walk_blocks_per_sm(int64_t n_sites) {
    if (n_sites <= 0) return 6;
    const int spb = sv_prune_sites(false, true);
    return sv_pick_prune_blocks((unsigned)((n_sites + spb - 1) / spb) * (SV_PRUNE_THREADS / 32));
}

template <bool QFAST>
static void sv_launch_prune_t(sieve_core* h, dim3 grid, const SvPruneArgs& a) {
    const bool g0 = a.const_genotype == 0 && a.root_genotype == 0;
    if (h->private_cov) {  // per-category leaves (sieve_create rejects the twin; sv_launch_staged checks the genotypes)
        k_prune<false, true, QFAST, 0, true><<<grid, SV_PRUNE_THREADS, 0, h->stream>>>(a);
        return;
    }
    if (h->with_const && h->const_twin) {
        if (g0)
            k_prune<true, true, QFAST><<<grid, SV_PRUNE_THREADS, 0, h->stream>>>(a);
        else
            k_prune<true, false, false><<<grid, SV_PRUNE_THREADS, 0, h->stream>>>(a);
    } else {
        const int minb = QFAST && g0 && grid.y == 1 ? (h->prune_minb ? h->prune_minb : sv_pick_prune_blocks(grid.x * (SV_PRUNE_THREADS / 32))) : 0;
        if (QFAST && g0 && grid.y == 1 && !a.per_level && !h->use_k_prune) {
            // the default: the lean single-launch walk (walk.cuh); same bits as k_prune<false, true, true>
            if (minb == 7)
                k_walk<7><<<grid, SV_PRUNE_THREADS, 0, h->stream>>>(a);
            else if (minb == 8)
                k_walk<8><<<grid, SV_PRUNE_THREADS, 0, h->stream>>>(a);
            else
                k_walk<0><<<grid, SV_PRUNE_THREADS, 0, h->stream>>>(a);
        } else if (minb == 7)
            k_prune<false, true, QFAST, 7><<<grid, SV_PRUNE_THREADS, 0, h->stream>>>(a);
        else if (minb == 8)
            k_prune<false, true, QFAST, 8><<<grid, SV_PRUNE_THREADS, 0, h->stream>>>(a);
        else if (g0)
            k_prune<false, true, QFAST><<<grid, SV_PRUNE_THREADS, 0, h->stream>>>(a);
        else
            k_prune<false, false, false><<<grid, SV_PRUNE_THREADS, 0, h->stream>>>(a);
    }
}

// launches the scatters + the pruning for what is staged on the device
static int sv_launch_staged(sieve_core* h) {
    const int K = h->K;
    const size_t nm = h->last_nm, nd = h->last_nd, no = h->last_no;
    if (nm) {
        k_scatter_matrices<<<(unsigned)nm, 64, 0, h->stream>>>((const double*)(h->d_stage + h->last_off_mv),
                                                              (const int*)(h->d_stage + h->last_off_ms), (int)nm, K,
                                                              h->log_mode ? 1 : 0, h->d_mats);
        sv_count(h);
    }
    if (nd) {
        k_scatter_distances<<<(unsigned)((nd * K + 127) / 128), 128, 0, h->stream>>>(
            (const double*)(h->d_stage + h->last_off_dv), (const int*)(h->d_stage + h->last_off_ds), (int)nd, K,
            h->log_mode ? 1 : 0, h->d_qfac, h->d_mats);
        sv_count(h);
    }
    if (no) {
        SvPruneArgs a;
        a.part = h->d_part;
        a.scale = h->d_scale;
        a.cpart = h->d_cpart;
        a.cscale = h->d_cscale;
        a.leafX = h->d_leafX;
        a.part_rows = h->pool_rows_part;
        a.leaf_rows = h->pool_rows_leaf;
        a.n_matrix_slots = 2 * h->n_nodes;
        a.lsum = h->d_lsum + (h->ephemeral ? 0 : (size_t)h->cur_leaf * h->S);
        a.mats = h->d_mats;
        a.qfac = h->d_qfac;
        a.clamp_zero_distance = h->log_mode ? 1 : 0;
        a.ops = (const SvOp*)(h->d_stage);
        a.pf_dist = h->last_per_level ? 0 : h->pf_dist;
        a.S = h->S;
        a.out_offset = 0;
        a.K = K;
        a.linear_const_clamp = h->log_mode ? 0 : 1;
        a.const_genotype = h->const_genotype;
        a.root_genotype = h->root_genotype;
        for (int k = 0; k < 4; k++) a.prop[k] = h->prop[k];
        a.site_logl = h->d_site_logl;
        a.site_const = h->d_site_const;
        a.site_list = h->partial_list;
        if (h->partial_list) a.S = h->partial_n;
        if (h->private_cov && (a.const_genotype != 0 || a.root_genotype != 0))
            return sv_fail(SIEVE_ERR_STATE, "SIEVE_FLAG_PRIVATE_SEQ_COV cores support root / constant genotype 0 only (ScsFiniteMuExtendedModel)");
        // the Q-specialised walk needs every matrix of the launch to have been given as a distance (and genotype 0)
        const bool qfast = h->last_qfast && a.const_genotype == 0 && a.root_genotype == 0 && !h->force_generic;
        const int spb = sv_prune_sites(h->with_const && h->const_twin, qfast);
        const unsigned gx = (unsigned)(((h->partial_list ? (size_t)h->partial_n : h->S) + spb - 1) / spb);
        auto launch = [&](dim3 grid) {
            if (qfast)
                sv_launch_prune_t<true>(h, grid, a);
            else
                sv_launch_prune_t<false>(h, grid, a);
            sv_count(h);
        };
        if (h->ephemeral) {
            // streamed evaluation: per site chunk, leaves from the counts, then the whole walk on scratch slots
            a.op_begin = 0;
            a.op_end = (int)no;
            a.per_level = 0;
            for (long long s0 = 0; s0 < h->S; s0 += h->chunk_cap) {
                const int ns = (int)std::min<long long>(h->chunk_cap, h->S - s0);
                SV_TRY(sv_launch_leaves(h, s0, ns, h->chunk_cap));
                a.S = ns;
                a.out_offset = s0;
                launch(dim3((unsigned)((ns + spb - 1) / spb)));
            }
        } else if (h->last_per_level) {
            for (size_t l = 0; l + 1 < h->last_levels.size(); l++) {
                const int b = h->last_levels[l], e = h->last_levels[l + 1];
                if (e <= b) continue;
                a.op_end = e;
                a.per_level = 1;
                for (int b0 = b; b0 < e; b0 += 65535) {  // grid.y is limited to 65535
                    a.op_begin = b0;
                    launch(dim3(gx, (unsigned)std::min(65535, e - b0)));
                }
            }
        } else {
            a.op_begin = 0;
            a.op_end = (int)no;
            a.per_level = 0;
            launch(dim3(gx));
        }
    }
    SV_CUDA(cudaGetLastError());
    return SIEVE_OK;
}

// uploads pending matrices / distances + ops with ONE host->device copy and launches scatter + pruning
// SIEVE_B200_HOST_PROF=1: wall time of the host phases of a flush on stderr (diagnostic)
struct SvHostProf {
    bool on;
    std::chrono::steady_clock::time_point t0;
    double us[6] = {0, 0, 0, 0, 0, 0};
    SvHostProf() : on(getenv("SIEVE_B200_HOST_PROF") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void lap(int i) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        us[i] += std::chrono::duration<double, std::micro>(t1 - t0).count();
        t0 = t1;
    }
    void report(size_t no, size_t nm) {
        if (on && no > 8)
            fprintf(stderr, "sv_flush host us: matrices %.0f plan %.0f link %.0f stage+launch %.0f const-factors %.0f (ops %zu, matrices %zu)\n",
                    us[0], us[1], us[2], us[3], us[4], no, nm);
    }
};

static int sv_flush(sieve_core* h) {
    const int K = h->K;
    SvHostProf prof;
    // slot attributes first: the op builders look at them.  Only the last entry of a slot counts.  With change tracking on,
    // an entry that repeats what its slot already holds is dropped, and one that repeats what the node's OTHER matrix
    // buffer holds (the usual case when a caller refreshes the matrices of unchanged branches after flipping them) is
    // dropped too and the node is pointed at that buffer again: nothing is uploaded, scattered or mirrored for it.
    if (!h->pending_mats.empty()) {
        const int N = h->n_nodes;
        std::vector<int>& last_idx = h->scratch_idx;
        last_idx.assign((size_t)2 * N, -1);
        for (size_t i = 0; i < h->pending_mats.size(); i++) last_idx[h->pending_mats[i].slot] = (int)i;
        for (size_t i = 0; i < h->pending_mats.size(); i++) {
            PendingMatrix& pm = h->pending_mats[i];
            const double* pv = h->pending_vals.data() + pm.off;
            pm.skip = 0;
            if (last_idx[pm.slot] != (int)i) {
                pm.skip = 1;  // superseded by a later entry for the same slot
                continue;
            }
            const size_t len = (pm.is_distance ? (size_t)K : (size_t)K * 16) * sizeof(double);
            const int other = pm.slot < N ? pm.slot + N : pm.slot - N;
            const int node = pm.slot < N ? pm.slot : pm.slot - N;
            // distances are mirrored compactly with the node's two buffers side by side (one cache line per node)
            auto mirror = [&](int sl) -> double* {
                if (pm.is_distance) return h->dist_mirror.data() + ((size_t)(sl < N ? sl : sl - N) * 2 + (sl < N ? 0 : 1)) * SV_KMAX;
                return h->mat_mirror.data() + (size_t)sl * SV_KMAX * 16;
            };
            auto same = [&](int sl) {
                return h->mat_ver[sl] != 0 && h->mat_mirror_q[sl] == (char)pm.is_distance && memcmp(mirror(sl), pv, len) == 0;
            };
            pm.reuse = same(pm.slot) ? 1 : (same(other) ? 2 : 0);
            if (h->elide && pm.reuse == 1) {
                pm.skip = 1;
                continue;
            }
            if (h->elide && pm.reuse == 2 && last_idx[other] < 0 && h->cur_mat[node] * N + node == pm.slot) {
                h->cur_mat[node] = 1 - h->cur_mat[node];
                pm.skip = 1;
                continue;
            }
            h->slot_is_q[pm.slot] = pm.is_distance ? 1 : 0;
            char z = 0;
            if (pm.is_distance)
                for (int k = 0; k < K; k++) z |= pv[k] == 0.0;
            h->slot_zero[pm.slot] = z;
            if (pm.reuse == 2) h->mat_ver[pm.slot] = h->mat_ver[other];
            if (pm.reuse == 0) h->mat_ver[pm.slot] = ++h->ver_counter;
            h->mat_mirror_q[pm.slot] = (char)pm.is_distance;
            memcpy(mirror(pm.slot), pv, len);
            if (pm.reuse == 2) {
                h->slot_b[pm.slot] = h->slot_b[other];
                h->slot_bd[pm.slot] = h->slot_bd[other];
            } else
                sv_slot_bounds(h, pm);
        }
    }
    prof.lap(0);
    if (h->ephemeral)
        SV_TRY(sv_build_streamed_ops(h));
    else if (h->pending_per_level) {
        // the level-batched schedule's ops, made now that every node points at the matrix buffer it will be read from
        const std::vector<int>& t = h->pending_level_triples;
        h->pending_ops.reserve(t.size() / 3);
        for (size_t i = 0; i + 2 < t.size(); i += 3) {
            SvOp op;
            SV_TRY(sv_make_op(h, t[i + 1], t[i + 2], t[i], &op));
            sv_bound_op(h, t[i], t[i + 1], t[i + 2], &op);
            h->pending_ops.push_back(op);
        }
        h->pending_level_triples.clear();
    } else if (!h->dirty_list.empty() || !h->demand_list.empty())
        SV_TRY(sv_plan_resident_ops(h));
    prof.lap(1);
    const size_t no = h->pending_ops.size();
    size_t nm = 0, nd = 0;
    for (const PendingMatrix& pm : h->pending_mats)
        if (!pm.skip) (pm.is_distance ? nd : nm)++;
    if (nm == 0 && nd == 0 && no == 0) {
        h->pending_mats.clear();
        h->pending_vals.clear();
        h->g_nodes.clear();
        return SIEVE_OK;
    }
    if (no > 0 && !h->leaves_ready) return sv_fail(SIEVE_ERR_STATE, "ops queued before the leaves were computed");
    // staging layout: [ops][matrix values][distance values][matrix slots][distance slots]
    const size_t off_mv = no * sizeof(SvOp);
    const size_t off_dv = off_mv + nm * K * 16 * sizeof(double);
    const size_t off_ms = off_dv + nd * K * sizeof(double);
    const size_t off_ds = off_ms + nm * sizeof(int);
    const size_t total = off_ds + nd * sizeof(int);
    SV_TRY(sv_ensure_stage(h, total));  // the other half: the previous step's kernels may still be reading theirs
    sv_link_forwarding(h->pending_ops, h->pending_per_level);
    if (no) {
        std::vector<SvOp>& ops = h->pending_ops;
        sv_link_prefetch(ops, h->pending_per_level ? 0 : h->pf_dist);
        memcpy(h->h_stage, ops.data(), no * sizeof(SvOp));
    }
    prof.lap(2);
    size_t im = 0, id = 0;
    for (const PendingMatrix& pm : h->pending_mats) {
        if (pm.skip) continue;
        if (pm.is_distance) {
            memcpy(h->h_stage + off_dv + id * K * sizeof(double), h->pending_vals.data() + pm.off, K * sizeof(double));
            ((int*)(h->h_stage + off_ds))[id++] = pm.slot;
        } else {
            memcpy(h->h_stage + off_mv + im * K * 16 * sizeof(double), h->pending_vals.data() + pm.off, K * 16 * sizeof(double));
            ((int*)(h->h_stage + off_ms))[im++] = pm.slot;
        }
    }
    bool qfast = true;
    for (const SvOp& op : h->pending_ops) {
        if (!h->slot_is_q[op.m1] || (!(op.flags & SV_OP_UNARY) && !h->slot_is_q[op.m2])) {
            qfast = false;
            break;
        }
    }
    SV_CUDA(cudaMemcpyAsync(h->d_stage, h->h_stage, total, cudaMemcpyHostToDevice, h->stream));
    h->last_no = no;
    h->last_nm = nm;
    h->last_nd = nd;
    h->last_off_mv = off_mv;
    h->last_off_ms = off_ms;
    h->last_off_dv = off_dv;
    h->last_off_ds = off_ds;
    h->last_qfast = qfast;
    h->last_stage_bytes = total;
    h->last_levels = h->pending_levels;
    h->last_per_level = h->pending_per_level;
    SV_TRY(sv_launch_staged(h));
    if (!h->partial_list)
        for (const SvOp& op : h->pending_ops)
            if (op.flags & SV_OP_ROOT) h->site_out_stale = false;  // every site's log-likelihood was just written
    SV_CUDA(cudaEventRecord(h->ev_stage[h->stage_sel], h->stream));
    h->ev_stage_set[h->stage_sel] = true;
    prof.lap(3);
    // the site-independent constant-site factor is host arithmetic; done here it overlaps the kernels just launched
    // (only the reduce kernel, launched later, needs its result log C_tree)
    if (!h->hP.empty())
        for (const PendingMatrix& pm : h->pending_mats) {
            if (pm.skip || pm.reuse == 1) continue;
            if (pm.reuse == 2) {
                const int other = pm.slot < h->n_nodes ? pm.slot + h->n_nodes : pm.slot - h->n_nodes;
                memcpy(h->hP.data() + (size_t)pm.slot * SV_KMAX * 16, h->hP.data() + (size_t)other * SV_KMAX * 16,
                       sizeof(double) * SV_KMAX * 16);
            } else
                sv_host_matrix(h, pm);
        }
    if (!h->G.empty() && !h->g_nodes.empty()) sv_update_G(h);
    h->g_nodes.clear();
    prof.lap(4);
    prof.report(no, nm + nd);
    if (!h->pending_ops.empty() || !h->pending_mats.empty()) h->ml_ready = false;
    h->pending_mats.clear();
    h->pending_vals.clear();
    h->pending_ops.clear();
    h->pending_levels.clear();
    h->pending_per_level = false;
    return SIEVE_OK;
}

// logConstRoot per site of a private-model core into d_site_const (leaf_private.cuh, k_const_private)
static void sv_launch_const_private(sieve_core* h) {
    SvConstPrivW w;
    for (int k = 0; k < SV_KMAX; k++) w.logw[k] = h->logw[k];
    k_const_private<<<(h->S + 255) / 256, 256, 0, h->stream>>>(h->d_lambdaK + (size_t)h->cur_leaf * SV_KMAX * h->S, h->S, h->S, h->K,
                                                              w, h->d_site_const);
    sv_count(h);
}

// the totals of the shard and -- bound to a communicator -- their exchange with the other ranks: fused into the reduction over
// peer memory (reduce.cuh, SvMail) or, when the mailboxes could not be mapped, as an ncclAllGather behind it
static int sv_launch_reduce(sieve_core* h) {
    const int blocks = std::min(SV_REDUCE_BLOCKS, (h->S + 255) / 256);
    SvMail mail{nullptr, nullptr, nullptr, 0, 0, 1};
    if (h->comm && h->p2p) {
        mail.peers = h->d_peers;
        mail.gather = h->d_gather;
        mail.error = h->d_mail_err;
        mail.step = ++h->mail_step;
        mail.rank = h->comm_rank;
        mail.world = h->comm_world;
    }
    if (h->private_cov && h->with_const) {
        sv_launch_const_private(h);
        k_reduce<<<blocks, 256, 0, h->stream>>>(h->d_site_logl, h->d_site_const, 0.0, h->d_weights, h->S, 1, h->d_partials,
                                                h->d_counter, h->d_result, mail);
    } else {
        const bool algebraic = h->with_const && !h->const_twin;
        const double* cst = algebraic ? h->d_lambda + (h->ephemeral ? 0 : (size_t)h->cur_leaf * h->S) : h->d_site_const;
        k_reduce<<<blocks, 256, 0, h->stream>>>(h->d_site_logl, cst, algebraic ? h->log_c : 0.0, h->d_weights, h->S,
                                                h->with_const ? 1 : 0, h->d_partials, h->d_counter, h->d_result, mail);
    }
    sv_count(h);
    if (h->comm && !h->p2p) {
        ncclResult_t nr = sv_nccl().AllGather(h->d_result, h->d_gather, 5, ncclDouble, h->comm, h->stream);
        if (nr != ncclSuccess) return sv_fail(SIEVE_ERR_CUDA, std::string("ncclAllGather: ") + sv_nccl().GetErrorString(nr));
    }
    return SIEVE_OK;
}

static int sv_evaluate(sieve_core* h, sieve_shard_result_t* out) {
    if (h->timing) SV_CUDA(cudaEventRecord(h->ev[2], h->stream));
    SV_TRY(sv_flush(h));
    if (h->timing) SV_CUDA(cudaEventRecord(h->ev[3], h->stream));
    const int blocks = std::min(SV_REDUCE_BLOCKS, (h->S + 255) / 256);
    SV_TRY(sv_launch_reduce(h));  // + the one exchange of the multi-GPU path: every rank's 5 doubles to every rank
    SV_CUDA(cudaGetLastError());
    if (h->timing) SV_CUDA(cudaEventRecord(h->ev[4], h->stream));
    SV_CUDA(cudaMemcpyAsync(h->h_result, h->d_result, 5 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    unsigned int mail_err = 0;
    if (h->comm) {
        SV_CUDA(cudaMemcpyAsync(h->h_gather, h->d_gather, (size_t)h->comm_world * 5 * sizeof(double), cudaMemcpyDeviceToHost,
                                h->stream));
        if (h->p2p) SV_CUDA(cudaMemcpyAsync(&mail_err, h->d_mail_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    }
    SV_CUDA(cudaStreamSynchronize(h->stream));
    if (mail_err) return sv_fail(SIEVE_ERR_CUDA, "the peer-memory exchange timed out: a rank did not reach this evaluation");
    if (out) {
        out->sum_log_l = h->h_result[0];
        out->const_lse = h->h_result[1];
        out->const_sum = h->h_result[2];
        out->lewis_sum = h->h_result[3];
        out->weight_sum = h->h_result[4];
    }
    if (h->timing) {
        float leaf = 0, prune = 0, red = 0;
        if (h->leaf_timed) cudaEventElapsedTime(&leaf, h->ev[0], h->ev[1]);
        cudaEventElapsedTime(&prune, h->ev[2], h->ev[3]);
        cudaEventElapsedTime(&red, h->ev[3], h->ev[4]);
        h->last_ms[0] = leaf;
        h->last_ms[1] = prune;
        h->last_ms[2] = red;
        h->last_ms[3] = leaf + prune + red;
        h->leaf_timed = false;
    }
    return SIEVE_OK;
}

extern "C" int sieve_evaluate(sieve_core_t* h, sieve_shard_result_t* out) {
    SV_REQUIRE(h, "null handle");
    SV_CUDA(cudaSetDevice(h->device));
    int r = sv_evaluate(h, out);
    h->last_launches = h->step_launches;
    h->step_launches = 0;
    return r;
}

extern "C" int sieve_step(sieve_core_t* h, int32_t update_leaves, int32_t n_matrices, const int32_t* matrix_nodes,
                          const double* P, int32_t n_ops, const int32_t* ops, sieve_shard_result_t* out) {
    SV_REQUIRE(h, "null handle");
    SV_CUDA(cudaSetDevice(h->device));
    if (update_leaves) SV_TRY(sieve_update_leaves(h, 1));
    SV_TRY(sieve_set_matrices(h, n_matrices, matrix_nodes, P, 1));
    SV_TRY(sieve_update_nodes(h, n_ops, ops, nullptr, 0, SIEVE_UPDATE_FLIP));
    int r = sv_evaluate(h, out);
    h->last_launches = h->step_launches;
    h->step_launches = 0;
    return r;
}

extern "C" int sieve_step_distances(sieve_core_t* h, int32_t update_leaves, int32_t n_nodes, const int32_t* nodes,
                                    const double* distances, int32_t n_ops, const int32_t* ops,
                                    sieve_shard_result_t* out) {
    SV_REQUIRE(h, "null handle");
    SV_CUDA(cudaSetDevice(h->device));
    if (update_leaves) SV_TRY(sieve_update_leaves(h, 1));
    SV_TRY(sieve_set_distances(h, n_nodes, nodes, distances, 1));
    SV_TRY(sieve_update_nodes(h, n_ops, ops, nullptr, 0, SIEVE_UPDATE_FLIP));
    int r = sv_evaluate(h, out);
    h->last_launches = h->step_launches;
    h->step_launches = 0;
    return r;
}

extern "C" int sieve_last_stage_bytes(sieve_core_t* h, uint64_t* out) {
    SV_REQUIRE(h && out, "null argument");
    *out = h->last_stage_bytes;
    return SIEVE_OK;
}

extern "C" int sieve_replay(sieve_core_t* h, int32_t update_leaves, int32_t n_times) {
    SV_REQUIRE(h && n_times >= 0, "sieve_replay: bad argument");
    SV_CUDA(cudaSetDevice(h->device));
    if (!h->pending_ops.empty() || !h->pending_mats.empty() || h->pending_per_level || !h->dirty_list.empty())
        return sv_fail(SIEVE_ERR_STATE, "sieve_replay: there is un-launched work queued");
    if (h->last_no == 0) return sv_fail(SIEVE_ERR_STATE, "sieve_replay: nothing was staged yet");
    const int blocks = std::min(SV_REDUCE_BLOCKS, (h->S + 255) / 256);
    for (int i = 0; i < n_times; i++) {
        if (update_leaves) {
            h->dm_dirty = h->cov_dirty = true;  // a parameter proposal rebuilds the tables
            if (!h->ephemeral) SV_TRY(sv_launch_leaves(h, 0, h->S, h->S));
        }
        SV_TRY(sv_launch_staged(h));
        SV_TRY(sv_launch_reduce(h));  // with the per-evaluation exchange of a multi-process run, as in sv_evaluate
    }
    SV_CUDA(cudaGetLastError());
    return SIEVE_OK;
}

extern "C" int sieve_get_site_log_likelihoods(sieve_core_t* h, double* out_l, double* out_c) {
    SV_REQUIRE(h && out_l, "null argument");
    SV_CUDA(cudaSetDevice(h->device));
    SV_TRY(sv_flush(h));
    SV_CUDA(cudaMemcpyAsync(out_l, h->d_site_logl, sizeof(double) * h->S, cudaMemcpyDeviceToHost, h->stream));
    const bool algebraic = h->with_const && !h->const_twin && !h->private_cov;
    if (out_c && h->private_cov && h->with_const) sv_launch_const_private(h);
    if (out_c) {
        const double* src = algebraic ? h->d_lambda + (h->ephemeral ? 0 : (size_t)h->cur_leaf * h->S) : h->d_site_const;
        SV_CUDA(cudaMemcpyAsync(out_c, src, sizeof(double) * h->S, cudaMemcpyDeviceToHost, h->stream));
    }
    SV_CUDA(cudaStreamSynchronize(h->stream));
    if (out_c && algebraic)  // logConstRoot[s] = log C_tree + Lambda[s]
        for (int s = 0; s < h->S; s++) out_c[s] += h->log_c;
    return SIEVE_OK;
}

extern "C" int sieve_get_result_device_ptr(sieve_core_t* h, void** d_result) {
    SV_REQUIRE(h && d_result, "null argument");
    *d_result = h->d_result;
    return SIEVE_OK;
}

// The log-space genotype likelihoods of ONE leaf straight from the log-space mixture (k_leaf_ml / k_leaf_ml_private: the
// reference's own operation order, no common linear scale), into *d_L ([S][4], private model [S][K][4]; the caller frees
// it).  sieve_get_node_partials takes them where the scaled linear leaf value has underflowed to 0.
static int sv_leaf_log_exact(sieve_core* h, int node, double** d_L) {
    const size_t S = h->S, K = h->K;
    const int C = h->C;
    const size_t per = h->private_cov ? K : 1;
    double* d_cell = nullptr;
    unsigned short* d_ado = nullptr;
    *d_L = nullptr;
    cudaError_t e = cudaMalloc(d_L, sizeof(double) * S * per * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_ado, sizeof(unsigned short) * S * per);
    if (e == cudaSuccess && !h->private_cov) e = cudaMalloc(&d_cell, sizeof(double) * (size_t)C * 4);
    int rr = SIEVE_OK;
    if (e == cudaSuccess) {
        SvDmAlphas al;
        SvCovParams cp;
        // the tables of the parameters the CURRENT leaf buffer was computed with (as sv_subtree_leaf_sums)
        const sieve_leaf_params_t pending = h->lp;
        h->lp = h->leaf_lp[h->cur_leaf];
        rr = sv_refresh_leaf_tables(h, &al, &cp);
        h->lp = pending;
        if (rr == SIEVE_OK && h->private_cov) {
            if (!h->have_tv)
                rr = sv_fail(SIEVE_ERR_STATE, "sieve_get_node_partials: sieve_set_private_seq_cov has not been called");
            else {
                SvMlLeafPrivArgs pa;
                pa.counts = h->d_counts + (size_t)node * S;
                pa.dmA = h->d_dmA;
                pa.dmB = h->d_dmB;
                pa.size_factors = h->d_sf + node;
                pa.tv = h->d_tv;
                pa.leafL = *d_L;
                pa.ado = d_ado;
                pa.n_cells = 1;
                pa.S = (int)S;
                pa.C = C;
                pa.K = (int)K;
                pa.al = al;
                pa.theta = cp.theta;
                pa.single_ado = cp.single_ado;
                k_leaf_ml_private<<<dim3((unsigned)((S + 127) / 128), 1), 128, 0, h->stream>>>(pa);
                sv_count(h);
            }
        } else if (rr == SIEVE_OK) {
            k_build_cell_log_tables<<<dim3((unsigned)((C + 127) / 128), 1), 128, 0, h->stream>>>(d_cell, h->d_sf + node, 1, C, cp);
            sv_count(h);
            SvMlLeafArgs la;
            la.counts = h->d_counts + (size_t)node * S;
            la.dmA = h->d_dmA;
            la.dmB = h->d_dmB;
            la.cellL = d_cell;
            la.size_factors = h->d_sf + node;
            la.leafL = *d_L;
            la.ado = d_ado;
            la.n_cells = 1;
            la.S = (int)S;
            la.C = C;
            la.counts_stride = (long long)S;
            la.site_begin = 0;
            la.al = al;
            la.cp = cp;
            k_leaf_ml<<<dim3((unsigned)((S + 127) / 128), 1), 128, 0, h->stream>>>(la);
            sv_count(h);
        }
        if (rr == SIEVE_OK) e = cudaGetLastError();
        if (rr == SIEVE_OK && e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    }
    cudaFree(d_cell);
    cudaFree(d_ado);
    if (rr != SIEVE_OK || e != cudaSuccess) {
        cudaFree(*d_L);
        *d_L = nullptr;
    }
    if (rr != SIEVE_OK) return rr;
    SV_CUDA(e);
    return SIEVE_OK;
}

extern "C" int sieve_get_node_partials(sieve_core_t* h, int32_t node, int32_t which, double* out) {
    SV_REQUIRE(h && out && sv_node_ok(h, node), "sieve_get_node_partials: bad argument");
    SV_REQUIRE(which == 0 || which == 1, "sieve_get_node_partials: which must be 0 or 1");
    if (h->ephemeral) return sv_fail(SIEVE_ERR_STATE, "sieve_get_node_partials: a streamed (SIEVE_FLAG_EPHEMERAL) core keeps no partials");
    SV_CUDA(cudaSetDevice(h->device));
    SV_TRY(sv_flush(h));
    if (node >= h->n && !h->valid[node]) {
        // a transient node: its partial was consumed from registers and never written; recompute it into its buffer
        h->demand_list.push_back(node);
        SV_TRY(sv_flush(h));
    }
    const size_t S = h->S, K = h->K, n = h->n;
    const double* x;
    const int* sc;
    int is_leaf = node < h->n;
    // the leaves' scales live only as sums (leaf.cuh): re-derive the sum over the leaves below this node
    double* d_sub = nullptr;  // [1 + (private model: K, else 1)][S]: scales, constant-genotype log-likelihoods (per category)
    SV_CUDA(cudaMalloc(&d_sub, sizeof(double) * (1 + (h->private_cov ? SV_KMAX : 1)) * S));
    double* d_exact = nullptr;  // a leaf in log mode: its log-space likelihoods, for the entries the linear scale lost
    struct SubFree {
        double* p;
        double** q;
        ~SubFree() {
            cudaFree(p);
            cudaFree(*q);
        }
    } sub_free{d_sub, &d_exact};
    if (is_leaf) {
        SV_REQUIRE(which == 0, "leaves have no constant-site partials");
        if (!h->leaves_ready) return sv_fail(SIEVE_ERR_STATE, "leaves not computed yet");
        SV_TRY(sv_subtree_leaf_sums(h, node, d_sub, nullptr));
        x = h->d_leafX + ((size_t)h->cur_leaf * n + node) * S * (h->private_cov ? SV_KMAX * 4 : 4);
        sc = nullptr;
        if (h->private_cov) is_leaf = 0;  // K x 4 values per site, the layout of an internal partial (no exponent: sc == NULL)
        if (h->log_mode) SV_TRY(sv_leaf_log_exact(h, node, &d_exact));
    } else {
        const int sl = h->slot_of[(size_t)h->cur_part[node] * h->n_nodes + node];
        if (sl < 0) return sv_fail(SIEVE_ERR_STATE, "sieve_get_node_partials: the node could not be materialised");
        const size_t slot = (size_t)sl;
        if (which == 1 && !h->with_const) return sv_fail(SIEVE_ERR_STATE, "constant-site partials are not maintained");
        const bool from_factors = which == 1 && !h->const_twin;
        SV_TRY(sv_subtree_leaf_sums(h, node, d_sub, from_factors ? d_sub + S : nullptr));
        if (from_factors) {
            // rebuilt from the factorisation: G_node x product over the node's leaves of their constant-genotype likelihood
            double *d_G = nullptr, *d_o = nullptr;
            const size_t NN = K * S * 4;
            cudaError_t e = cudaMalloc(&d_G, sizeof(double) * 16);
            if (e == cudaSuccess) e = cudaMalloc(&d_o, sizeof(double) * NN);
            const double* g = h->G.data() + ((size_t)h->cur_part[node] * h->n_nodes + node) * SV_KMAX * 4;
            if (e == cudaSuccess) e = cudaMemcpyAsync(d_G, g, sizeof(double) * 16, cudaMemcpyHostToDevice, h->stream);
            if (e == cudaSuccess) {
                k_export_const<<<(unsigned)((S + 255) / 256), 256, 0, h->stream>>>(
                    d_sub + S, h->private_cov ? (long long)S : 0LL, d_G, h->Gs[(size_t)h->cur_part[node] * h->n_nodes + node],
                    (int)S, (int)K, h->log_mode ? 1 : 0, d_o);
                sv_count(h);
                e = cudaGetLastError();
            }
            if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_o, NN * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            cudaFree(d_G);
            cudaFree(d_o);
            SV_CUDA(e);
            return SIEVE_OK;
        }
        x = (which ? h->d_cpart : h->d_part) + slot * S * SV_KMAX * 4;
        sc = (which ? h->d_cscale : h->d_scale) + slot * S;
    }
    double* d_out = nullptr;
    const size_t N = K * S * 4;
    SV_CUDA(cudaMalloc(&d_out, N * sizeof(double)));
    k_export_partials<<<(unsigned)((S + 255) / 256), 256, 0, h->stream>>>(x, sc, d_sub, (int)S, (int)K, is_leaf,
                                                                          h->log_mode ? 1 : 0, which == 1 ? 1 : 0, d_out, d_exact,
                                                                          h->private_cov ? 1 : 0);
    sv_count(h);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, N * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_out);
    SV_CUDA(e);
    return SIEVE_OK;
}

#include "core_maxsum.cuh"

// ---- store / restore ------------------------------------------------------------------------------------------
extern "C" int sieve_store(sieve_core_t* h) {
    SV_REQUIRE(h, "null handle");
    h->stored_part = h->cur_part;
    h->stored_mat = h->cur_mat;
    h->stored_leaf = h->cur_leaf;
    h->stored_lp = h->leaf_lp[h->cur_leaf];  // the parameters the current leaves were computed with
    h->leaf_ver_stored = h->leaf_ver[h->cur_leaf];
    h->leaf_changed_since_store = false;
    h->eph_c1_stored = h->eph_c1;
    h->eph_c2_stored = h->eph_c2;
    h->valid_stored = h->valid;
    h->log_c_stored = h->log_c;
    memcpy(h->logw_stored, h->logw, sizeof(h->logw));
    // value versions / provenance: the current and the stored arrays differ only where a plan has written since the last
    // store (restore swaps the arrays, which keeps that true), so only those entries are copied
    for (int v : h->prov_touched) {
        h->ver_stored[v] = h->ver_cur[v];
        h->fp_stored[v] = h->fp_cur[v];
    }
    h->prov_touched.clear();
    if (h->sparse)  // what only the previous stored state referred to gives its slot back
        for (int v = h->n; v < h->n_nodes; v++) sv_slot_release(h, 1 - h->cur_part[v], v);
    return SIEVE_OK;
}
// single leaf buffer: a rejected leaf-parameter proposal has overwritten the leaves -- put the stored state's back
static int sv_leaves_back_to_stored(sieve_core* h) {
    if (!h->single_leaf || !h->leaf_changed_since_store) return SIEVE_OK;
    SV_CUDA(cudaSetDevice(h->device));
    h->lp = h->stored_lp;  // the state's parameters are the stored ones again
    SV_TRY(sv_launch_leaves(h, 0, h->S, h->S));
    h->leaf_lp[h->cur_leaf] = h->stored_lp;
    h->leaf_ver[h->cur_leaf] = h->leaf_ver_stored;  // same parameters, same kernel: the same bits as before the proposal
    h->leaf_changed_since_store = false;
    return SIEVE_OK;
}

extern "C" int sieve_unstore(sieve_core_t* h) {
    SV_REQUIRE(h, "null handle");
    SV_TRY(sv_leaves_back_to_stored(h));
    h->cur_part = h->stored_part;
    h->cur_mat = h->stored_mat;
    h->cur_leaf = h->stored_leaf;
    h->eph_c1 = h->eph_c1_stored;
    h->eph_c2 = h->eph_c2_stored;
    h->valid = h->valid_stored;
    h->log_c = h->log_c_stored;
    memcpy(h->logw, h->logw_stored, sizeof(h->logw));
    for (int v : h->prov_touched) {
        h->ver_cur[v] = h->ver_stored[v];
        h->fp_cur[v] = h->fp_stored[v];
    }
    h->prov_touched.clear();
    h->topo_cache_ok = false;
    h->ml_ready = false;
    return SIEVE_OK;
}
extern "C" int sieve_restore(sieve_core_t* h) {
    SV_REQUIRE(h, "null handle");
    SV_TRY(sv_leaves_back_to_stored(h));
    h->cur_part.swap(h->stored_part);
    h->cur_mat.swap(h->stored_mat);
    std::swap(h->cur_leaf, h->stored_leaf);
    h->eph_c1.swap(h->eph_c1_stored);
    h->eph_c2.swap(h->eph_c2_stored);
    h->valid.swap(h->valid_stored);
    std::swap(h->log_c, h->log_c_stored);
    for (int k = 0; k < SV_KMAX; k++) std::swap(h->logw[k], h->logw_stored[k]);
    h->ver_cur.swap(h->ver_stored);
    h->fp_cur.swap(h->fp_stored);
    if (h->eph_c1 != h->eph_c1_stored || h->eph_c2 != h->eph_c2_stored) h->topo_cache_ok = false;
    h->ml_ready = false;
    h->site_out_stale = true;
    // after a rejected parameter proposal the tables hold the rejected values; they are compared with the values of
    // the next sieve_set_leaf_params at the next leaf update, so nothing has to be rebuilt here
    return SIEVE_OK;
}

// ---- combining shards (host arithmetic on a handful of doubles) --------------------------------------------------
extern "C" double sieve_combine_shards(const sieve_shard_result_t* sh, int32_t n_shards, int32_t asc_mode, double M) {
    if (!sh || n_shards <= 0) return NAN;
    double logP = 0.0;
    for (int i = 0; i < n_shards; i++) logP += sh[i].sum_log_l;  // ThreadedScsTreeLikelihood.java:371-373
    double loci = 0.0;
    for (int i = 0; i < n_shards; i++) loci += sh[i].weight_sum;
    switch (asc_mode) {
        case SIEVE_ASC_FELSENSTEIN_MEAN: {
            // one shard: ScsAlignment.java:768-772; several: :811-815 -- logSumExp over the per-shard log-sums
            double mx = sh[0].const_lse;
            for (int i = 1; i < n_shards; i++) mx = sh[i].const_lse > mx ? sh[i].const_lse : mx;
            double s = 0.0;
            for (int i = 0; i < n_shards; i++) s += exp(sh[i].const_lse - mx);
            logP += (log(s) + mx - log(loci)) * M;
            break;
        }
        case SIEVE_ASC_FELSENSTEIN_GEOMETRIC: {
            double s = 0.0;
            for (int i = 0; i < n_shards; i++) s += sh[i].const_sum;
            // :773-776 divides by the number of loci for one worker, :816 by the number of workers for several
            logP += n_shards == 1 ? s * M / loci : s * M / (double)n_shards;
            break;
        }
        case SIEVE_ASC_LEWIS:
            for (int i = 0; i < n_shards; i++) logP += sh[i].lewis_sum;  // :746-752, applied per worker
            break;
        default:
            break;
    }
    return logP;
}

extern "C" void sieve_pattern_points(int32_t n_sites, int32_t n_shards, int32_t* points) {
    points[0] = 0;
    for (int i = 0; i < n_shards; i++) points[i + 1] = points[i] + n_sites / n_shards + (i < n_sites % n_shards ? 1 : 0);
}

// ---- multi-process shards -----------------------------------------------------------------------------------------------
extern "C" int sieve_comm_unique_id(uint8_t* id_out128) {
    SV_REQUIRE(id_out128, "null argument");
    if (!sv_nccl().ok) return sv_fail(SIEVE_ERR_CUDA, "libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    ncclResult_t r = sv_nccl().GetUniqueId(&id);
    if (r != ncclSuccess) return sv_fail(SIEVE_ERR_CUDA, std::string("ncclGetUniqueId: ") + sv_nccl().GetErrorString(r));
    memcpy(id_out128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return SIEVE_OK;
}

// Mailboxes for the exchange fused into k_reduce (reduce.cuh, SvMail): every rank allocates one, the CUDA IPC handles go
// round with one ncclAllGather, every rank maps its peers' mailboxes (NVLink peer access) and a second all-gather makes
// sure that EVERY rank succeeded -- otherwise all of them keep the ncclAllGather per evaluation (ranks inside one process
// cannot open each other's handles, a device pair without peer access cannot map).  SIEVE_B200_P2P=0 forces that path.
static int sv_mail_setup(sieve_core* h) {
    const int world = h->comm_world, rank = h->comm_rank;
    h->p2p = false;
    const char* env = getenv("SIEVE_B200_P2P");
    int want = !(env && atoi(env) == 0);
    const size_t mail_bytes = (size_t)2 * world * SV_MAIL_ROW * sizeof(double);
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (want && (cudaMalloc(&h->d_mail, mail_bytes) != cudaSuccess || cudaMemset(h->d_mail, 0, mail_bytes) != cudaSuccess ||
                 cudaMalloc(&h->d_peers, sizeof(double*) * world) != cudaSuccess ||
                 cudaMalloc(&h->d_mail_err, sizeof(unsigned int)) != cudaSuccess ||
                 cudaMemset(h->d_mail_err, 0, sizeof(unsigned int)) != cudaSuccess))
        want = 0;
    if (want && world > 1 && cudaIpcGetMemHandle(&mine, h->d_mail) != cudaSuccess) want = 0;
    cudaGetLastError();
    // round 1: (want, handle) of every rank
    const size_t rec = 8 + sizeof(cudaIpcMemHandle_t);
    std::vector<unsigned char> all(rec * world, 0);
    unsigned char* d_all = nullptr;
    SV_CUDA(cudaMalloc(&d_all, rec * world));
    {
        std::vector<unsigned char> me(rec, 0);
        me[0] = (unsigned char)want;
        memcpy(me.data() + 8, &mine, sizeof(mine));
        SV_CUDA(cudaMemcpyAsync(d_all + rec * rank, me.data(), rec, cudaMemcpyHostToDevice, h->stream));
        ncclResult_t nr = sv_nccl().AllGather(d_all + rec * rank, d_all, rec, ncclChar, h->comm, h->stream);
        if (nr != ncclSuccess) {
            cudaFree(d_all);
            return sv_fail(SIEVE_ERR_CUDA, std::string("ncclAllGather: ") + sv_nccl().GetErrorString(nr));
        }
        SV_CUDA(cudaMemcpyAsync(all.data(), d_all, rec * world, cudaMemcpyDeviceToHost, h->stream));
        SV_CUDA(cudaStreamSynchronize(h->stream));
    }
    int ok = want;
    for (int r = 0; r < world; r++) ok = ok && all[rec * r] != 0;
    h->peer_mail.assign(world, nullptr);
    if (ok) {
        for (int r = 0; r < world && ok; r++) {
            if (r == rank) {
                h->peer_mail[r] = h->d_mail;
                continue;
            }
            cudaIpcMemHandle_t hd;
            memcpy(&hd, all.data() + rec * r + 8, sizeof(hd));
            if (cudaIpcOpenMemHandle(&h->peer_mail[r], hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                h->peer_mail[r] = nullptr;
                ok = 0;
            }
        }
        cudaGetLastError();
    }
    // round 2: did everybody map everybody?
    {
        unsigned char flag = (unsigned char)ok;
        SV_CUDA(cudaMemcpyAsync(d_all + rank, &flag, 1, cudaMemcpyHostToDevice, h->stream));
        ncclResult_t nr = sv_nccl().AllGather(d_all + rank, d_all, 1, ncclChar, h->comm, h->stream);
        if (nr != ncclSuccess) {
            cudaFree(d_all);
            return sv_fail(SIEVE_ERR_CUDA, std::string("ncclAllGather: ") + sv_nccl().GetErrorString(nr));
        }
        SV_CUDA(cudaMemcpyAsync(all.data(), d_all, world, cudaMemcpyDeviceToHost, h->stream));
        SV_CUDA(cudaStreamSynchronize(h->stream));
        for (int r = 0; r < world; r++) ok = ok && all[r] != 0;
    }
    cudaFree(d_all);
    if (ok) {
        SV_CUDA(cudaMemcpy(h->d_peers, h->peer_mail.data(), sizeof(double*) * world, cudaMemcpyHostToDevice));
        h->p2p = true;
        h->mail_step = 0;
    }
    if (getenv("SIEVE_B200_COMM_DEBUG"))
        fprintf(stderr, "sieve_comm_init: rank %d of %d, exchange %s\n", rank, world,
                h->p2p ? "fused into k_reduce over peer memory (CUDA IPC mailboxes)" : "ncclAllGather per evaluation");
    return SIEVE_OK;
}

extern "C" int sieve_comm_exchange_mode(sieve_core_t* h) { return !h || !h->comm ? 0 : (h->p2p ? 2 : 1); }

extern "C" int sieve_comm_init(sieve_core_t* h, int32_t rank, int32_t world, const uint8_t* id128) {
    SV_REQUIRE(h && id128 && world >= 1 && rank >= 0 && rank < world, "sieve_comm_init: bad argument");
    if (!sv_nccl().ok) return sv_fail(SIEVE_ERR_CUDA, "libnccl.so.2 could not be loaded");
    SV_CUDA(cudaSetDevice(h->device));
    ncclUniqueId id;
    memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    ncclResult_t r = sv_nccl().CommInitRank(&h->comm, world, id, rank);
    if (r != ncclSuccess) {
        h->comm = nullptr;
        return sv_fail(SIEVE_ERR_CUDA, std::string("ncclCommInitRank: ") + sv_nccl().GetErrorString(r));
    }
    h->comm_rank = rank;
    h->comm_world = world;
    SV_TRY(sv_alloc(h, &h->d_gather, (size_t)world * 5));
    SV_CUDA(cudaMallocHost((void**)&h->h_gather, (size_t)world * 5 * sizeof(double)));
    return sv_mail_setup(h);
}

extern "C" int sieve_comm_size(sieve_core_t* h) { return h && h->comm ? h->comm_world : 1; }

extern "C" int sieve_last_global_results(sieve_core_t* h, sieve_shard_result_t* out_world) {
    SV_REQUIRE(h && out_world, "null argument");
    if (!h->comm) {
        out_world[0].sum_log_l = h->h_result[0];
        out_world[0].const_lse = h->h_result[1];
        out_world[0].const_sum = h->h_result[2];
        out_world[0].lewis_sum = h->h_result[3];
        out_world[0].weight_sum = h->h_result[4];
        return SIEVE_OK;
    }
    memcpy(out_world, h->h_gather, (size_t)h->comm_world * sizeof(sieve_shard_result_t));
    return SIEVE_OK;
}

// ---- stage-1 background likelihood (a "next" row of the scope table) ----------------------------------------------
extern "C" int sieve_background_log_likelihood(int32_t device, const int64_t* values, const int64_t* counts,
                                               const int32_t* offsets6, int64_t n_background_sites, int32_t n_taxa,
                                               double eff_seq_err_rate, double shape_ctrl1, double* out_log_p) {
    SV_REQUIRE(values && counts && offsets6 && out_log_p, "sieve_background_log_likelihood: null argument");
    SV_REQUIRE(offsets6[0] == 0, "sieve_background_log_likelihood: offsets must start at 0");
    for (int i = 0; i < 5; i++) SV_REQUIRE(offsets6[i + 1] >= offsets6[i], "sieve_background_log_likelihood: offsets must ascend");
    int ndev = 0;
    SV_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) return sv_fail(SIEVE_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
    if (device >= 0) SV_CUDA(cudaSetDevice(device));
    const int total = offsets6[5];
    long long *d_v = nullptr, *d_c = nullptr;
    int* d_o = nullptr;
    double* d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_v, sizeof(long long) * std::max(total, 1));
    if (e == cudaSuccess) e = cudaMalloc(&d_c, sizeof(long long) * std::max(total, 1));
    if (e == cudaSuccess) e = cudaMalloc(&d_o, sizeof(int) * 6);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(d_v, values, sizeof(long long) * total, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_c, counts, sizeof(long long) * total, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_o, offsets6, sizeof(int) * 6, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_background<<<1, 256>>>(d_v, d_c, d_o, n_background_sites, n_taxa, eff_seq_err_rate, shape_ctrl1, d_out);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out_log_p, d_out, sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_v);
    cudaFree(d_c);
    cudaFree(d_o);
    cudaFree(d_out);
    SV_CUDA(e);
    return SIEVE_OK;
}

// ---- instrumentation ------------------------------------------------------------------------------------------
extern "C" int sieve_sparse_info(sieve_core_t* h, int32_t* pool_slots, int32_t* keep_leaves, int32_t* free_slots) {
    SV_REQUIRE(h, "null handle");
    if (pool_slots) *pool_slots = h->pool_slots;
    if (keep_leaves) *keep_leaves = h->sparse ? h->keep_T : 0;
    if (free_slots) *free_slots = h->sparse ? (int32_t)h->free_slots.size() : 0;
    return SIEVE_OK;
}

extern "C" int sieve_set_elision(sieve_core_t* h, int32_t on) {
    SV_REQUIRE(h, "null handle");
    h->elide = on != 0;
    return SIEVE_OK;
}
extern "C" int sieve_elided_count(sieve_core_t* h, uint64_t* out) {
    SV_REQUIRE(h && out, "null argument");
    *out = h->elided_total;
    return SIEVE_OK;
}

extern "C" int sieve_enable_timing(sieve_core_t* h, int32_t on) {
    SV_REQUIRE(h, "null handle");
    h->timing = on != 0;
    return SIEVE_OK;
}
extern "C" int sieve_last_timing(sieve_core_t* h, float* out_ms4, int32_t* out_launches) {
    SV_REQUIRE(h, "null handle");
    if (out_ms4) memcpy(out_ms4, h->last_ms, sizeof(h->last_ms));
    if (out_launches) *out_launches = h->last_launches;
    return SIEVE_OK;
}
extern "C" int sieve_launch_count(sieve_core_t* h, uint64_t* out) {
    SV_REQUIRE(h && out, "null argument");
    *out = h->launches;
    return SIEVE_OK;
}
extern "C" int sieve_get_stream(sieve_core_t* h, void** out) {
    SV_REQUIRE(h && out, "null argument");
    *out = (void*)h->stream;
    return SIEVE_OK;
}
