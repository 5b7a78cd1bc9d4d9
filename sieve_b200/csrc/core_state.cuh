// core_state.cuh -- the handle behind sieve_core_t and the small host helpers every part of the core uses.
// Included by sieve_core.cu only (one translation unit); split out for readability.
#pragma once
#include <unordered_map>
static thread_local std::string g_err;

static int sv_fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define SV_CUDA(call)                                                                                      \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess)                                                                            \
            return sv_fail(SIEVE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));           \
    } while (0)
#define SV_REQUIRE(cond, msg) \
    do {                      \
        if (!(cond)) return sv_fail(SIEVE_ERR_INVALID, msg); \
    } while (0)

struct PendingMatrix {
    int slot;
    int reuse = 0;    // host mirror of the matrix: 0 = compute, 1 = the slot already holds these values, 2 = copy the twin slot's
    int skip = 0;     // not uploaded: superseded, or (change tracking) nothing new
    int is_distance;  // the values are K branch distances (length * branch rate * category rate) instead of K matrices
    size_t off;       // its values in sieve_core::pending_vals (K doubles or K * 16): the entries stay small, a whole-tree
                      // update of 10 000 cells queues 20 000 of them per proposal
};

struct sieve_core {
    sieve_config_t cfg;
    int n = 0, S = 0, K = 0, n_nodes = 0, n_internal = 0;
    bool log_mode = false, with_const = false, single_ado = true, ephemeral = false;
    int device = 0;
    cudaStream_t stream = nullptr;
    uint64_t bytes = 0;
    uint64_t launches = 0;
    int step_launches = 0;  // since the last evaluate
    bool leaf_attr_set = false;
    bool tables_cover_counts = false;  // every count of the shard is inside the look-up tables (checked at ingest)
    bool leaf_force_checked = false;   // SIEVE_B200_LEAF_CHECKED=1: keep the out-of-table paths in k_leaf regardless
    int leaf_tab = 1;    // SIEVE_B200_LEAF_TAB: where k_leaf reads its tables from (sv_launch_leaf_kernels)
    int prune_minb = 0;  // SIEVE_B200_PRUNE_MINB: 6, 7 or 8 resident blocks per SM for the variant-only Q walk (0: chosen per launch)

    // device pools
    int4* d_counts = nullptr;  // [cell][site]
    bool counts_owned = true;
    int* d_weights = nullptr;
    double* d_sf = nullptr;
    double *d_dmA = nullptr, *d_dmB = nullptr, *d_cellT = nullptr;
    double *d_linA = nullptr, *d_linB = nullptr;  // linear-space twins of dmA / dmB (leaf.cuh, k_leaf_lin)
    bool lin_range_ok = false;   // the current tables fit the 13-bit exponents of the linear tables
    bool leaf_lin_allowed = true;  // SIEVE_B200_LEAF_LIN=0: always the log-space leaf kernels (A/B, tests)
    int C = 0;  // table entries per table
    int max_cov = -1;
    double* d_leafX = nullptr;                                                           // [2][n][S][4]
    // the leaves' log scales are kept only as their sum over the cells (leaf.cuh): per-cell-group partial sums (scratch)
    // and, per leaf buffer, the per-site total that the root adds
    double *d_sumS = nullptr, *d_sum0 = nullptr;  // [cell groups][site]
    double* d_lsum = nullptr;                     // [leaf buffers][site]
    double *d_part = nullptr, *d_cpart = nullptr;  // [slot][S][K][4]
    int *d_scale = nullptr, *d_cscale = nullptr;   // [slot][S] binary exponents
    double* d_mats = nullptr;                                                            // [2][n_nodes][4][16]
    double2* d_qfac = nullptr;                                                           // [2][n_nodes][4]
    std::vector<char> slot_is_q;  // per matrix slot: was it given as a distance?
    std::vector<char> slot_zero;  // ... and is one of its K distances exactly 0?
    double *d_site_logl = nullptr, *d_site_const = nullptr;
    // algebraic constant-site path (default when SIEVE_FLAG_CONST_PARTIALS): see k_colsum in reduce.cuh
    bool const_twin = false;     // SIEVE_B200_CONST_TWIN=1: prune the constant-site partials per site like the reference
    double* d_lambda = nullptr;  // [leaf buffers][site] sum over cells of the leaves' constant-genotype log-likelihood
    std::vector<double> hP;      // [2][node][K][16] host mirror of the (clamped) transition matrices
    std::vector<double> G, Gs;   // [2][node][K][4] site-independent factor of the constant-site partials + log-scale
    double log_c = 0.0, log_c_stored = 0.0;  // log sum_k prop_k G_root[k][root genotype]
    std::vector<int> g_nodes;    // nodes whose G has to be refreshed, children first
    SvAcc* d_partials = nullptr;
    unsigned int* d_counter = nullptr;
    double* d_result = nullptr;
    // staging
    // staging of a step's ops and matrices: two halves used in turn, so that a flush can fill one while the kernels of the
    // previous step still read the other (no stream synchronisation between the leaf kernels and the walk's launch)
    char* h_stage = nullptr;  // pinned; the half the CURRENT step was staged in
    char* d_stage = nullptr;
    char *h_stage_base = nullptr, *d_stage_base = nullptr;
    int stage_sel = 0;
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};  // recorded after the last launch that reads the half
    bool ev_stage_set[2] = {false, false};
    size_t stage_cap = 0;     // bytes per half
    double* h_result = nullptr;  // pinned, 8 doubles

    // host bookkeeping
    std::vector<int> cur_part, stored_part, cur_mat, stored_mat;
    int cur_leaf = 0, stored_leaf = 0;
    std::vector<SvOp> pending_ops;
    std::vector<int> pending_levels;  // offsets into pending_ops, empty == one group
    bool pending_per_level = false;
    std::vector<int> pending_level_triples;  // (parent, child1, child2) of a queued per-level update; ops are made at flush
    std::vector<PendingMatrix> pending_mats;
    std::vector<double> pending_vals;
    int root = -1, root_genotype = 0, const_genotype = 0;
    double prop[4] = {0.25, 0.25, 0.25, 0.25};
    sieve_leaf_params_t lp{};
    bool have_counts = false, have_sf = false, have_params = false, leaves_ready = false;
    bool dm_dirty = true, cov_dirty = true;
    sieve_leaf_params_t tab_lp{};  // parameters the current tables were built with
    sieve_leaf_params_t leaf_lp[2]{};  // parameters each leaf buffer was computed with (read-backs, store)

    // what the last flush staged on the device (for sieve_replay)
    size_t last_no = 0, last_nm = 0, last_nd = 0, last_off_mv = 0, last_off_ms = 0, last_off_dv = 0, last_off_ds = 0;
    size_t last_stage_bytes = 0;
    bool last_qfast = false;
    bool force_generic = false;
    unsigned long long pool_rows_part = 0, pool_rows_leaf = 0;  // rows (slot x site) of the partial / leaf pools (bounds-checked build)
    int pf_dist = 3;  // SIEVE_B200_PF_DIST: how many ops ahead the walk asks the TMA unit to bring operands into L2 (0: off)
    // streamed (SIEVE_FLAG_EPHEMERAL) evaluation: site chunks, a handful of scratch slots, leaves recomputed per chunk
    int chunk_cap = 0;   // sites per chunk == site capacity of the scratch pools
    int n_slots = 0;     // scratch slots for internal partials
    std::vector<int> eph_c1, eph_c2, eph_c1_stored, eph_c2_stored;  // children of every internal node (-2 = unset)
    bool eph_leaves_dirty = true;
    // resident mode: the core owns the walk order.  Callers name the dirty nodes (and their children); the planner
    // orders them so that results are consumed from registers, and leaves small subtrees' results unsaved
    // ("transient": at most transient_T leaves, consumed by the very next op); they are recomputed when needed again.
    int transient_T = 8;
    std::vector<char> valid, valid_stored;  // does the node's current buffer hold its partial?
    std::vector<int> dirty_list;            // internal nodes queued for recomputation
    std::vector<int> demand_list;           // nodes whose partial must be materialised (getNodePartials)  // SIEVE_B200_GENERIC=1: never take the Q-specialised walk (A/B measurements)
    std::vector<int> last_levels;
    bool last_per_level = false;

    // multi-process shards: NCCL communicator bound to this handle's stream
    ncclComm_t comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    double* d_gather = nullptr;   // [world][5]
    double* h_gather = nullptr;   // pinned
    // peer-memory exchange fused into k_reduce (reduce.cuh, SvMail); p2p == false: ncclAllGather per evaluation instead
    bool p2p = false;
    double* d_mail = nullptr;             // own mailbox [2][world][8]
    std::vector<void*> peer_mail;         // IPC mappings of the peers' mailboxes (own slot: d_mail)
    double** d_peers = nullptr;           // the same pointers on the device
    unsigned int* d_mail_err = nullptr;
    unsigned long long mail_step = 0;

    // Change tracking (resident planner): every matrix slot, leaf buffer and node value carries a version number, and a
    // node remembers the versions of the inputs its value was computed from.  A node that the caller marks dirty but
    // whose inputs are the ones its value in the stored state was computed from is NOT recomputed: the kernels are
    // deterministic, so the result would be the same bits (SIEVE forces whole-tree recomputation for every branch-rate
    // proposal under a relaxed clock with bias correction, ScsTreeLikelihood.java:2814-2821 -- here only the path from
    // the changed branch to the root is walked).  sieve_set_elision(h, 0) switches the short-cut off.
    struct SvFp {
        int c1 = -9, c2 = -9;
        uint64_t p1 = 0, p2 = 0, m1 = 0, m2 = 0;
        bool operator==(const SvFp& o) const {
            return c1 == o.c1 && c2 == o.c2 && p1 == o.p1 && p2 == o.p2 && m1 == o.m1 && m2 == o.m2;
        }
    };
    // Magnitude bounds (sv_bound_op): for the value each (buffer, node) holds, every entry lies in [2^lo, 2^hi) and the
    // largest is at least 2^mlo; per matrix slot floor(log2) of its smallest entry / smallest diagonal entry.  They decide
    // where the walk renormalises (SV_OP_RENORM); SIEVE_B200_RENORM_ALWAYS=1 renormalises after every op (A/B, tests).
    std::vector<int> nb_lo, nb_hi, nb_mlo;  // [2 * n_nodes]
    std::vector<int> slot_b, slot_bd;       // [2 * n_nodes]
    bool renorm_always = false;
    bool use_k_prune = false;  // SIEVE_B200_WALK=0: never take the lean walk kernel (walk.cuh)
    bool elide = true;
    uint64_t ver_counter = 0, elided_total = 0;
    std::vector<uint64_t> mat_ver;      // [2 * n_nodes] per matrix slot
    std::vector<double> mat_mirror;     // [2 * n_nodes][SV_KMAX * 16] the matrices a slot was last given ...
    std::vector<double> dist_mirror;    // [n_nodes][2][SV_KMAX] ... or its K distances
    std::vector<char> mat_mirror_q;     // ... and in which form
    uint64_t leaf_ver[2] = {0, 0};
    // Sparse residency (SIEVE_FLAG_SPARSE): the partial pool has fewer slots than 2 * n_internal.  Nodes with more than
    // keep_T leaves keep their partials (a slot per buffer that holds something the current or the stored state refers
    // to); smaller subtrees are recomputed from the resident leaves when an update needs them, through temporary slots
    // that are recycled inside the walk.  Dense cores use the same code with a fixed identity mapping.
    bool sparse = false, sparse_recheck = true;
    // SIEVE_FLAG_SINGLE_LEAF_BUFFER: one leaf buffer instead of two.  A leaf-parameter proposal overwrites the leaves; if it is
    // rejected, sieve_restore recomputes them from the stored parameters (same kernel, same bits) instead of flipping back.
    bool single_leaf = false, leaf_changed_since_store = false;
    sieve_leaf_params_t stored_lp{};
    uint64_t leaf_ver_stored = 0;
    int pool_slots = 0;
    int keep_T = 0;
    std::vector<int> slot_of;      // [2 * n_nodes] physical slot of (buffer, node), -1 = none
    std::vector<int> free_slots;
    std::vector<int> held;         // (buffer * n_nodes + node) of small nodes materialised on demand: released by the next plan
    std::vector<int> scratch_idx;       // per-flush scratch
    std::vector<int> topo_parent, topo_leaves;  // planner: parents and subtree leaf counts of the current children map
    bool topo_cache_ok = false;
    std::vector<uint64_t> g_ver;        // [2 * n_nodes] version of the node value whose constant-site factor G sits in that slot
    std::vector<char> buf_has;          // [2 * n_nodes] the partial buffer holds a value ...
    std::vector<uint64_t> buf_ver;      // ... of this version ...
    std::vector<SvFp> buf_fp;           // ... computed from these inputs
    std::vector<int> prov_touched;  // nodes whose ver_cur / fp_cur were written since the last store
    std::vector<uint64_t> ver_cur, ver_stored;  // [n_nodes] version of the node's value in the current / stored state
    std::vector<SvFp> fp_cur, fp_stored;        // ... and what it was computed from

    // PrivateAllelicSeqCovModel (SIEVE_FLAG_PRIVATE_SEQ_COV, leaf_private.cuh): (t, v) per (category, site); the leaf pool
    // then holds K x 4 values per (cell, site) and the constant-site column sums are kept per category
    bool private_cov = false, have_tv = false;
    double* d_lgx1 = nullptr;      // [C] logGamma(x + 1) of the table rows (k_build_lgx1)
    double2* d_cellT2 = nullptr;   // [cell][C] compact rows of the cell table for k_leaf_lin (single dropout)
    double2* d_tv = nullptr;       // [KMAX][S]
    double2* d_tv_leaf = nullptr;  // [KMAX][S] the pairs the CURRENT leaf buffer was computed with (sieve_private_retraverse_changed)
    bool tv_leaf_ok = false;
    bool site_out_stale = true;   // d_site_logl does not belong to the current state (nothing evaluated yet, or sieve_restore
                                  // since: like the reference's patternLogLikelihoods, the per-site outputs are not restored)
    const int* partial_list = nullptr;  // while sieve_private_retraverse_changed runs: the flush covers these sites only, in place
    int partial_n = 0;
    double* d_lambdaK = nullptr;   // [leaf buffers][KMAX][S]
    double* d_sum0K = nullptr;     // scratch [cell groups][KMAX][S]
    double logw[SV_KMAX] = {0, 0, 0, 0}, logw_stored[SV_KMAX] = {0, 0, 0, 0};  // log(prop_k G_root[k][rg]) + scale of G_root

    // max-sum pass for variant calling (maxsum.cuh); its pools are allocated by the first sieve_ml_trace
    bool ml_alloc = false, ml_ready = false, ml_cell_tables_ok = false;
    int ml_cap = 0;               // site capacity of the max-sum pools: S on a dense core, a chunk on a sparse one
    long long ml_site_begin = 0;  // the site range the pools currently hold
    int ml_n_sites = 0;
    double *d_mlL = nullptr, *d_ml = nullptr, *d_mllogm = nullptr, *d_mlcellL = nullptr, *d_mlrawm = nullptr;
    unsigned short* d_mlado = nullptr;
    unsigned* d_mlback = nullptr;
    unsigned char *d_mlcoll = nullptr, *d_mladosel = nullptr, *d_mlamb = nullptr, *d_mlcats = nullptr;
    SvMlOp* d_mlops = nullptr;
    int ml_n_ops = 0;

    // timing
    bool timing = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float last_ms[4] = {0, 0, 0, 0};
    int last_launches = 0;
    bool leaf_timed = false;
};

static inline void sv_count(sieve_core* h) {
    h->launches++;
    h->step_launches++;
}

#define SV_BOUND_NONE (-1000000)  // "no lower bound" of the magnitude bounds (core_planner.cuh, sv_bound_op)
static const int SV_TABLE_CAP_DEFAULT = 4096;
static const int SV_REDUCE_BLOCKS = 296;


template <typename T>
static int sv_alloc(sieve_core* h, T** p, size_t count) {
    const size_t b = count * sizeof(T);
    cudaError_t e = cudaMalloc((void**)p, b ? b : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return sv_fail(e == cudaErrorMemoryAllocation ? SIEVE_ERR_OOM : SIEVE_ERR_CUDA,
                       std::string("cudaMalloc of ") + std::to_string(b) + " bytes: " + cudaGetErrorString(e));
    }
    h->bytes += b;
    // SIEVE_B200_POISON=1 (tests; compute-sanitizer's initcheck is not available on this pool): every pool starts as 0xFF
    // bytes -- NaN as a double, -1 as an exponent or a count -- so a read of memory that no kernel or copy initialised
    // shows up in the results instead of passing by luck on zeroed pages
    static const bool poison = getenv("SIEVE_B200_POISON") && atoi(getenv("SIEVE_B200_POISON")) != 0;
    if (poison && b) cudaMemset(*p, 0xFF, b);
    return SIEVE_OK;
}
#define SV_TRY(x)                    \
    do {                             \
        int r__ = (x);               \
        if (r__ != SIEVE_OK) return r__; \
    } while (0)

// makes the OTHER half current, large enough for `bytes`, and waits until nothing in flight reads it any more
static int sv_ensure_stage(sieve_core* h, size_t bytes) {
    if (bytes > h->stage_cap) {
        size_t cap = std::max<size_t>(bytes, std::max<size_t>(h->stage_cap * 2, (size_t)1 << 16));
        cap = (cap + 255) & ~(size_t)255;
        SV_CUDA(cudaStreamSynchronize(h->stream));
        if (h->h_stage_base) cudaFreeHost(h->h_stage_base);
        if (h->d_stage_base) {
            cudaFree(h->d_stage_base);
            h->bytes -= 2 * h->stage_cap;
        }
        h->h_stage_base = h->d_stage_base = h->h_stage = h->d_stage = nullptr;
        h->stage_cap = 0;
        h->ev_stage_set[0] = h->ev_stage_set[1] = false;
        SV_CUDA(cudaMallocHost((void**)&h->h_stage_base, 2 * cap));
        SV_TRY(sv_alloc(h, &h->d_stage_base, 2 * cap));
        h->stage_cap = cap;
    }
    h->stage_sel ^= 1;
    h->h_stage = h->h_stage_base + (size_t)h->stage_sel * h->stage_cap;
    h->d_stage = h->d_stage_base + (size_t)h->stage_sel * h->stage_cap;
    if (h->ev_stage_set[h->stage_sel]) SV_CUDA(cudaEventSynchronize(h->ev_stage[h->stage_sel]));  // two flushes ago: long done
    return SIEVE_OK;
}
