// core_planner.cuh -- host side of the pruning walk: slots of the partial pools, op construction, the site-independent
// constant-site factor, the resident planner (walk order, forwarding, unsaved subtrees, change tracking, sparse
// residency) and the streamed walk builder.  Included by sieve_core.cu only; no kernels here.
#pragma once
// row of (buffer, internal node) in the partial pools; a node without a slot (its value only ever lives in registers)
// gets a unique row that is never dereferenced -- the forwarding link compares rows
static inline unsigned sv_part_row(const sieve_core* h, int buf, int node) {
    const int sl = h->slot_of[(size_t)buf * h->n_nodes + node];
    if (sl < 0) return 0xFFFFFFFFu - (unsigned)node;
    return (unsigned)((long long)sl * h->S);
}

static void sv_slot_release(sieve_core* h, int buf, int node) {
    const size_t bi = (size_t)buf * h->n_nodes + node;
    if (!h->sparse || h->slot_of[bi] < 0) return;
    h->free_slots.push_back(h->slot_of[bi]);
    h->slot_of[bi] = -1;
    h->buf_has[bi] = 0;
    if (h->cur_part[node] == buf) h->valid[node] = 0;
    if (h->stored_part[node] == buf) h->valid_stored[node] = 0;
}

// a free slot; when none is left, buffers that neither the current nor the stored state refers to are reclaimed first
static int sv_slot_alloc(sieve_core* h) {
    if (h->free_slots.empty()) {
        const int n = h->n, N = h->n_nodes;
        for (int v = n; v < N; v++)
            for (int b = 0; b < 2; b++) {
                if (h->slot_of[(size_t)b * N + v] < 0) continue;
                const bool used = (b == h->cur_part[v] && h->valid[v]) || (b == h->stored_part[v] && h->valid_stored[v]);
                if (!used) sv_slot_release(h, b, v);
            }
    }
    if (h->free_slots.empty()) return -1;
    const int sl = h->free_slots.back();
    h->free_slots.pop_back();
    return sl;
}

static int sv_make_op(sieve_core* h, int c1, int c2, int parent, SvOp* op) {
    const int n = h->n;
    const long long S = h->S;
    SV_REQUIRE(parent >= n && parent < h->n_nodes, "op: parent must be an internal node");
    SV_REQUIRE(c1 >= 0 && c1 < h->n_nodes && c1 != h->root, "op: bad child1");
    SV_REQUIRE(c2 >= -1 && c2 < h->n_nodes && c2 != h->root, "op: bad child2");
    op->flags = 0;
    op->pf_row1 = op->pf_row2 = 0;
    op->dst_row = sv_part_row(h, h->cur_part[parent], parent);
    auto child = [&](int c, int leaf_flag, unsigned* row, int* m) {
        if (c < n) {
            op->flags |= leaf_flag;
            *row = (unsigned)((long long)(h->cur_leaf * n + c) * S);
        } else
            *row = sv_part_row(h, h->cur_part[c], c);
        *m = h->cur_mat[c] * h->n_nodes + c;
        if (h->log_mode && h->slot_zero[*m]) op->flags |= SV_OP_ZEROD;
    };
    child(c1, SV_OP_LEAF1, &op->c1_row, &op->m1);
    if (c2 < 0) {
        op->flags |= SV_OP_UNARY;
        op->c2_row = 0;
        op->m2 = 0;
    } else
        child(c2, SV_OP_LEAF2, &op->c2_row, &op->m2);
    if (parent == h->root) op->flags |= SV_OP_ROOT;
    return SIEVE_OK;
}

// ---- where the walk renormalises -------------------------------------------------------------------------------------
// Scaling by a power of two is exact and the exponents are summed as integers, so a partial may stay un-normalised for as
// long as none of its 16 values per site can leave the normal range of a double.  The host bounds the magnitudes without
// looking at the data:
//   * a leaf partial has its largest value in [1, 2) (k_leaf normalises it), the others anywhere below;
//   * y = P x with a stochastic matrix: every y[p] <= max x (rows sum to 1), every y[p] >= min P * max x, and
//     max y >= min diag(P) * max x;
//   * the product of the two children's contributions multiplies the bounds.
// [2^lo, 2^hi) contains every entry of the node's value, 2^mlo is a lower bound of the largest one.  When the bounds
// come near the limits of the exponent range the op renormalises (SV_OP_RENORM), and so do: the root (its log is taken:
// the result must not depend on the schedule), a node below a zero-length branch in log mode (the Double.MIN_VALUE
// entries of that matrix must meet normalised values, as the reference's arithmetic has it), every op of a core that
// keeps the per-site constant-site twin (its values have no such bound).
static inline int sv_badd(int a, int b) { return std::max(SV_BOUND_NONE, a + b); }
static inline int sv_floor_log2(double x) {
    if (!(x > 0.0) || !std::isfinite(x)) return SV_BOUND_NONE;
    int e;
    frexp(x, &e);
    return e - 1;
}

// floor(log2) of the smallest entry / smallest diagonal entry of the K matrices just given to a slot (one less: slack for
// the rounding of the device's sums)
static void sv_slot_bounds(sieve_core* h, const PendingMatrix& pm) {
    const int K = h->K;
    const double* pv = h->pending_vals.data() + pm.off;
    static const double P1[16] = {3, 6, -3, -6, 1, 2, -1, -2, -1, -2, 1, 2, -1, -2, 1, 2};     // 8 Pi1
    static const double P2[16] = {9, -18, 3, 6, -3, 6, -1, -2, 1, -2, 11, -10, 1, -2, -5, 6};  // 16 Pi2
    double mn = INFINITY, mnd = INFINITY;
    for (int k = 0; k < K; k++) {
        // (one pair of expm1 per category: this runs for every new matrix of a step, 2 000 of them in a whole-tree move)
        const double d = pm.is_distance ? pv[k] : 0.0;
        const double g1 = pm.is_distance ? expm1(-0.2 * d) : 0.0;  // expm1(-0.4 d) = g1 (g1 + 2)
        const double q1 = g1 / 8.0, q2 = g1 * (g1 + 2.0) / 16.0;
        for (int e = 0; e < 16; e++) {
            double v;
            if (pm.is_distance) {
                v = ((e % 5 == 0) ? 1.0 : 0.0) + q1 * P1[e] + q2 * P2[e];
                if (d == 0.0) v = (e % 5 == 0) ? 1.0 : 0.0;
            } else
                v = pv[k * 16 + e];
            if (h->log_mode) v = v > 0.0 ? v : SV_MIN_VALUE;
            if (!(v > 0.0)) v = 0.0;
            mn = std::min(mn, v);
            if (e % 5 == 0) mnd = std::min(mnd, v);
        }
    }
    // (the closed form above differs from the device's by rounding only; two binades of slack cover it)
    const int b = sv_floor_log2(mn), bd = sv_floor_log2(mnd);
    h->slot_b[pm.slot] = b == SV_BOUND_NONE ? b : b - 2;
    h->slot_bd[pm.slot] = bd == SV_BOUND_NONE ? bd : bd - 2;
}

// bounds of the op's result from those of its operands; sets SV_OP_RENORM where needed and records the bounds of what
// the node's current buffer will hold
static void sv_bound_op(sieve_core* h, int parent, int c1, int c2, SvOp* op) {
    const int n = h->n, N = h->n_nodes;
    struct Bd {
        int lo, hi, mlo;
    };
    auto side = [&](int c) {
        Bd x;
        if (c < n) {
            x.lo = SV_BOUND_NONE;  // a genotype may be arbitrarily unlikely
            x.hi = 1;
            x.mlo = 0;
        } else {
            const size_t bi = (size_t)h->cur_part[c] * N + c;
            x.lo = h->nb_lo[bi];
            x.hi = h->nb_hi[bi];
            x.mlo = h->nb_mlo[bi];
        }
        const size_t m = (size_t)h->cur_mat[c] * N + c;
        const int b = h->slot_b[m], bd = h->slot_bd[m];
        Bd y;
        y.lo = std::max(x.lo, sv_badd(b, x.mlo));
        y.hi = x.hi + 1;  // rows of P sum to 1 up to rounding
        y.mlo = std::max(sv_badd(bd, x.mlo), y.lo);
        return y;
    };
    Bd r = side(c1);
    if (c2 >= 0) {
        const Bd z = side(c2);
        Bd t;
        t.lo = sv_badd(r.lo, z.lo);
        t.hi = r.hi + z.hi;
        t.mlo = std::max(sv_badd(r.mlo, z.lo), sv_badd(z.mlo, r.lo));
        r = t;
    }
    const size_t own_m = (size_t)h->cur_mat[parent] * N + parent;
    const bool below_zero_branch = parent != h->root && h->log_mode && h->slot_zero[own_m];
    const bool renorm = h->renorm_always || parent == h->root || (h->with_const && h->const_twin) || below_zero_branch ||
                        r.lo < -900 || r.hi > 900;
    if (renorm) {
        op->flags |= SV_OP_RENORM;
        r.lo = sv_badd(r.lo, -r.hi);  // the exponent taken out is at most hi
        r.hi = 1;
        r.mlo = 0;
    }
    const size_t bi = (size_t)h->cur_part[parent] * N + parent;
    h->nb_lo[bi] = r.lo;
    h->nb_hi[bi] = r.hi;
    h->nb_mlo[bi] = r.mlo;
}

// ---- algebraic constant-site path: the site-independent factor ----------------------------------------------------
// constPartial[v][k][s][p] = G_v[k][p] * prod_{leaves l below v} leaf_l[s][cg]   (induction over
// ScsBeerLikelihoodCore.java:1432-1440 / :1467-1490: a leaf child contributes P[p][cg] * leaf[cg], an internal child
// sum_c P[p][c] * const[c], and the leaf factor is common to the four parent states).  G is 16 doubles per node, kept on
// the host in the same double buffers as the partials, refreshed for the nodes an update recomputes.
static void sv_host_matrix(sieve_core* h, const PendingMatrix& pm) {
    const int K = h->K;
    const double* pv = h->pending_vals.data() + pm.off;
    double* dst = h->hP.data() + (size_t)pm.slot * SV_KMAX * 16;
    static const double P1[16] = {3, 6, -3, -6, 1, 2, -1, -2, -1, -2, 1, 2, -1, -2, 1, 2};     // 8 Pi1
    static const double P2[16] = {9, -18, 3, 6, -3, 6, -1, -2, 1, -2, 11, -10, 1, -2, -5, 6};  // 16 Pi2
    for (int k = 0; k < K; k++) {
        double d = 0.0, q1 = 0.0, q2 = 0.0;
        if (pm.is_distance) {
            d = pv[k];
            const double g1 = expm1(-0.2 * d);  // expm1(-0.4 d) = g1 (g1 + 2)
            q1 = g1 / 8.0;
            q2 = g1 * (g1 + 2.0) / 16.0;
        }
        for (int e = 0; e < 16; e++) {
            double v;
            if (pm.is_distance) {
                v = ((e % 5 == 0) ? 1.0 : 0.0) + q1 * P1[e] + q2 * P2[e];
                if (d == 0.0) v = (e % 5 == 0) ? 1.0 : 0.0;
            } else
                v = pv[k * 16 + e];
            if (h->log_mode) v = v > 0.0 ? v : SV_MIN_VALUE;
            dst[k * 16 + e] = v;
        }
    }
}

static void sv_update_G(sieve_core* h) {
    const int n = h->n, N = h->n_nodes, K = h->K, cg = h->const_genotype;
    for (int v : h->g_nodes) {
        const int a = h->eph_c1[v], b = h->eph_c2[v];
        double* g = h->G.data() + ((size_t)h->cur_part[v] * N + v) * SV_KMAX * 4;
        double gs = 0.0, mx = 0.0;
        auto side = [&](int c, int k, double* y) {
            const double* P = h->hP.data() + ((size_t)h->cur_mat[c] * N + c) * SV_KMAX * 16 + k * 16;
            if (c < n) {
                for (int p = 0; p < 4; p++) y[p] = P[4 * p + cg];
            } else {
                const double* gc = h->G.data() + ((size_t)h->cur_part[c] * N + c) * SV_KMAX * 4 + k * 4;
                for (int p = 0; p < 4; p++) y[p] = P[4 * p] * gc[0] + P[4 * p + 1] * gc[1] + P[4 * p + 2] * gc[2] + P[4 * p + 3] * gc[3];
            }
        };
        for (int k = 0; k < K; k++) {
            double y[4], z[4];
            side(a, k, y);
            if (b >= 0) {
                side(b, k, z);
                for (int p = 0; p < 4; p++) y[p] *= z[p];
            }
            for (int p = 0; p < 4; p++) {
                g[k * 4 + p] = y[p];
                mx = std::max(mx, y[p]);
            }
        }
        if (a >= n) gs += h->Gs[(size_t)h->cur_part[a] * N + a];
        if (b >= n) gs += h->Gs[(size_t)h->cur_part[b] * N + b];
        if (mx > 0.0 && std::isfinite(mx)) {
            for (int i = 0; i < K * 4; i++) g[i] /= mx;
            gs += std::log(mx);
        }
        h->Gs[(size_t)h->cur_part[v] * N + v] = gs;
        h->g_ver[(size_t)h->cur_part[v] * N + v] = h->ver_cur[v];
    }
    h->g_nodes.clear();
    const int root = h->root;
    const double* gr = h->G.data() + ((size_t)h->cur_part[root] * N + root) * SV_KMAX * 4;
    double sum = 0.0;
    for (int k = 0; k < K; k++) sum += h->prop[k] * gr[k * 4 + h->root_genotype];
    h->log_c = std::log(sum) + h->Gs[(size_t)h->cur_part[root] * N + root];
    for (int k = 0; k < SV_KMAX; k++)  // per category, for the per-pattern coverage model (k_const_private)
        h->logw[k] = k < K ? std::log(h->prop[k] * gr[k * 4 + h->root_genotype]) + h->Gs[(size_t)h->cur_part[root] * N + root]
                           : -INFINITY;
}

// Resident-mode planner.  Input: the dirty nodes (dirty_list), nodes to materialise (demand_list), the parent -> children
// map.  Output: pending_ops, children before parents, ordered so that as many results as possible are consumed by the
// very next op (register forwarding), with the results of small subtrees (<= transient_T leaves) that are consumed at
// once left unsaved (SV_OP_NOSTORE).  A clean node whose partial is not in memory is recomputed from its subtree on the
// way (at most transient_T - 1 extra ops reading leaves).  Clean recomputed nodes that do get stored go to their
// current buffer and become valid in both the current and the stored state: a clean subtree is the same in both.
static int sv_plan_resident_ops(sieve_core* h) {
    const int n = h->n, N = h->n_nodes, root = h->root;
    std::vector<char> dirty(N, 0), demanded(N, 0);
    for (int v : h->dirty_list) dirty[v] = 1;
    for (int v : h->demand_list) demanded[v] = 1;
    // parents and subtree leaf counts where the children are known
    // (cached between plans: they only change when an op names new children for a node)
    std::vector<int>&parent = h->topo_parent, &leaves = h->topo_leaves;
    if (!h->topo_cache_ok) {
    parent.assign(N, -1);
    leaves.assign(N, 0);
    for (int v = n; v < N; v++) {
        if (h->eph_c1[v] >= 0) parent[h->eph_c1[v]] = v;
        if (h->eph_c2[v] >= 0) parent[h->eph_c2[v]] = v;
    }
    {
        // bottom-up leaf counts via an explicit post-order from every known node without a known parent
        std::vector<char> done(N, 0);
        std::vector<std::pair<int, int>> st;
        for (int r = n; r < N; r++) {
            if (parent[r] >= 0 || h->eph_c1[r] == -2) continue;
            st.push_back({r, 0});
            while (!st.empty()) {
                auto [v, state] = st.back();
                st.pop_back();
                if (v < n) continue;
                const int a = h->eph_c1[v], b = h->eph_c2[v];
                if (a == -2) {  // never described: treat as large
                    leaves[v] = 1 << 28;
                    continue;
                }
                if (state == 0) {
                    st.push_back({v, 1});
                    if (b >= n) st.push_back({b, 0});
                    if (a >= n) st.push_back({a, 0});
                } else {
                    const long long la = a < n ? 1 : leaves[a], lb = b < 0 ? 0 : (b < n ? 1 : leaves[b]);
                    leaves[v] = (int)std::min<long long>(la + lb, 1 << 28);
                }
            }
        }
    }
    h->topo_cache_ok = true;
    h->sparse_recheck = true;  // leaf counts moved: a node may have dropped below the keep threshold
    }
    if (h->sparse) {
        // small nodes that were materialised for a read-back go back to "not in memory"
        for (int bi : h->held) {
            const int b = bi / N, v = bi % N;
            if (v != root && leaves[v] <= h->keep_T) sv_slot_release(h, b, v);
        }
        h->held.clear();
        // the nodes that keep their partials must fit twice (a whole-tree proposal flips every one of them) plus the
        // temporaries of the walk; when they do not, the threshold goes up and the nodes below it give their slots back
        auto kept_count = [&](int T) {
            int c = 0;
            for (int v = n; v < N; v++) c += (v == root || leaves[v] > T) ? 1 : 0;
            return c;
        };
        int T = std::max(h->keep_T, 1);
        // (a pool with a slot for every (buffer, node) keeps everything, like a dense core)
        while (h->pool_slots < 2 * h->n_internal && 2 * (long long)kept_count(T) + 64 > h->pool_slots && T < (1 << 28))
            T += std::max(1, T / 4);
        if (T != h->keep_T || h->sparse_recheck) {
            h->keep_T = T;
            for (int v = n; v < N; v++)
                if (v != root && leaves[v] <= T) {
                    sv_slot_release(h, 0, v);
                    sv_slot_release(h, 1, v);
                }
            h->sparse_recheck = false;
        }
    }
    // ---- change tracking: which of the dirty nodes really get a new value? (children before parents) ----
    if (h->prov_touched.size() > (size_t)8 * N) {  // a caller that never stores: keep the list a set
        std::sort(h->prov_touched.begin(), h->prov_touched.end());
        h->prov_touched.erase(std::unique(h->prov_touched.begin(), h->prov_touched.end()), h->prov_touched.end());
    }
    {
        std::vector<char> done(N, 0);
        std::vector<std::pair<int, int>> st;
        auto pver = [&](int c) -> uint64_t { return c < 0 ? 0 : (c < n ? h->leaf_ver[h->cur_leaf] : h->ver_cur[c]); };
        auto mver = [&](int c) -> uint64_t { return c < 0 ? 0 : h->mat_ver[(size_t)h->cur_mat[c] * N + c]; };
        for (int top : h->dirty_list) {
            if (done[top]) continue;
            st.push_back({top, 0});
            while (!st.empty()) {
                auto [v, state] = st.back();
                st.pop_back();
                if (v < n || !dirty[v] || done[v]) continue;
                const int a = h->eph_c1[v], b = h->eph_c2[v];
                if (a == -2) {
                    done[v] = 1;
                    continue;  // reported by the walk below
                }
                if (state == 0) {
                    st.push_back({v, 1});
                    if (b >= n) st.push_back({b, 0});
                    if (a >= n) st.push_back({a, 0});
                    continue;
                }
                done[v] = 1;
                sieve_core::SvFp f;
                f.c1 = a;
                f.c2 = b;
                f.p1 = pver(a);
                f.p2 = pver(b);
                f.m1 = mver(a);
                f.m2 = mver(b);
                const bool known = f.p1 != 0 && f.m1 != 0 && (b < 0 || (f.p2 != 0 && f.m2 != 0));
                if (h->elide && known && v != root) {
                    // 1. one of the node's two buffers holds the value of exactly these inputs
                    int hit = -1;
                    for (int t = 0; t < 2 && hit < 0; t++) {
                        const int bsel = t == 0 ? h->cur_part[v] : 1 - h->cur_part[v];
                        const size_t bi = (size_t)bsel * N + v;
                        if (h->buf_has[bi] && h->buf_fp[bi] == f) hit = bsel;
                    }
                    if (hit >= 0) {
                        h->cur_part[v] = hit;
                        h->valid[v] = 1;
                        h->ver_cur[v] = h->buf_ver[(size_t)hit * N + v];
                        h->fp_cur[v] = f;
                        h->prov_touched.push_back(v);
                        dirty[v] = 0;
                        h->elided_total++;
                        if (!h->G.empty() && h->g_ver[(size_t)h->cur_part[v] * N + v] != h->ver_cur[v]) h->g_nodes.push_back(v);
                        continue;
                    }
                    // 2. it is the stored state's value, which was consumed from registers and never written: the node
                    //    is clean but not in memory (recomputed from its small subtree if a dirty parent needs it)
                    if (h->ver_stored[v] != 0 && h->fp_stored[v] == f) {
                        h->cur_part[v] = h->stored_part[v];
                        h->valid[v] = 0;
                        h->ver_cur[v] = h->ver_stored[v];
                        h->fp_cur[v] = f;
                        h->prov_touched.push_back(v);
                        dirty[v] = 0;
                        h->elided_total++;
                        if (!h->G.empty() && h->g_ver[(size_t)h->cur_part[v] * N + v] != h->ver_cur[v]) h->g_nodes.push_back(v);
                        continue;
                    }
                }
                h->ver_cur[v] = ++h->ver_counter;
                h->fp_cur[v] = known ? f : sieve_core::SvFp();
                h->prov_touched.push_back(v);
                h->g_nodes.push_back(v);  // children first: this pass is a post-order over the dirty nodes
            }
        }
    }
    auto needs = [&](int c) { return c >= n && (dirty[c] || !h->valid[c]); };
    std::vector<int> emitted;
    struct Frame {
        int v, stage, first, last;
    };
    std::vector<Frame> st;
    std::vector<char> visited(N, 0);
    auto run_from = [&](int top) -> int {
        st.push_back({top, 0, -1, -1});
        while (!st.empty()) {
            Frame& f = st.back();
            const int v = f.v;
            if (f.stage == 0) {
                if (visited[v]) {
                    st.pop_back();
                    continue;
                }
                const int a = h->eph_c1[v], b = h->eph_c2[v];
                if (a == -2)
                    return sv_fail(SIEVE_ERR_STATE, "node " + std::to_string(v) + " is needed but was never computed (no children given)");
                const bool na = needs(a), nb = b >= 0 && needs(b);
                f.first = f.last = -1;
                if (na && nb) {
                    const int small = h->sparse ? h->keep_T : h->transient_T;
                    const bool ta = leaves[a] <= small, tb = leaves[b] <= small;
                    int last;
                    if (ta != tb)
                        last = ta ? a : b;  // the transient one is consumed from registers and never written
                    else
                        last = leaves[a] >= leaves[b] ? a : b;
                    f.last = last;
                    f.first = last == a ? b : a;
                } else if (na)
                    f.last = a;
                else if (nb)
                    f.last = b;
                f.stage = 1;
                if (f.first >= 0) {
                    const int nxt = f.first;
                    st.push_back({nxt, 0, -1, -1});
                }
            } else if (f.stage == 1) {
                f.stage = 2;
                if (f.last >= 0) {
                    const int nxt = f.last;
                    st.push_back({nxt, 0, -1, -1});
                }
            } else {
                visited[v] = 1;
                emitted.push_back(v);  // the op itself is made below, when the destination slots are known
                st.pop_back();
            }
        }
        return SIEVE_OK;
    };
    // tops: dirty / demanded nodes without a dirty parent; ancestors of a dirty node are expected to be dirty too
    for (int v = N - 1; v >= n; v--) {
        if (!(dirty[v] || demanded[v]) || visited[v]) continue;
        if (parent[v] >= 0 && dirty[parent[v]]) continue;  // reached through its parent
        if (!dirty[v] && h->valid[v]) continue;             // demanded but already in memory
        SV_TRY(run_from(v));
    }
    // a dirty node whose parent was dirty but never reached it (inconsistent input) would be skipped: check
    for (int v : h->dirty_list)
        if (dirty[v] && !visited[v]) SV_TRY(run_from(v));
    // store decisions and ops, in walk order
    std::vector<char> is_temp(h->sparse ? N : 0, 0);
    for (size_t i = 0; i < emitted.size(); i++) {
        const int v = emitted[i];
        bool consumed_next = false;
        if (i + 1 < emitted.size()) {
            const int p = emitted[i + 1];
            consumed_next = h->eph_c1[p] == v || h->eph_c2[p] == v;
        }
        bool store;
        const size_t bi = (size_t)h->cur_part[v] * N + v;
        if (!h->sparse) {
            const bool transient = h->transient_T > 0 && v != root && leaves[v] <= h->transient_T;
            store = !(transient && consumed_next) || demanded[v];
        } else {
            const bool kept = v == root || leaves[v] > h->keep_T;
            store = kept || demanded[v] || !consumed_next;
            if (store && h->slot_of[bi] < 0) {
                const int sl = sv_slot_alloc(h);
                if (sl < 0)
                    return sv_fail(SIEVE_ERR_OOM, "the sparse partial pool is exhausted (" + std::to_string(h->pool_slots) +
                                                      " slots): raise SIEVE_B200_POOL_SLOTS or free device memory");
                h->slot_of[bi] = sl;
                if (!kept) {
                    if (demanded[v])
                        h->held.push_back((int)bi);  // read back after this flush, released by the next plan
                    else
                        is_temp[v] = 1;              // an intermediate of a small subtree: recycled once its parent has consumed it
                }
            }
        }
        SvOp op;
        SV_TRY(sv_make_op(h, h->eph_c1[v], h->eph_c2[v], v, &op));
        sv_bound_op(h, v, h->eph_c1[v], h->eph_c2[v], &op);
        if (!store) op.flags |= SV_OP_NOSTORE;
        h->pending_ops.push_back(op);
        // a flush over a site list (sieve_private_retraverse_changed) writes the listed sites only: a node whose buffer was
        // not in memory before may be written for its consumer, but it does not become a resident value
        const bool partial_invalid = h->partial_list && !h->valid[v];
        h->valid[v] = store && !partial_invalid ? 1 : 0;
        if (partial_invalid) h->buf_has[bi] = 0;
        if (store && !partial_invalid) {
            h->buf_has[bi] = 1;
            h->buf_ver[bi] = h->ver_cur[v];
            h->buf_fp[bi] = h->fp_cur[v];
        }
        // a clean subtree is identical in the stored state; its buffer is shared when the indices agree
        if (store && !dirty[v] && h->cur_part[v] == h->stored_part[v]) h->valid_stored[v] = 1;
        if (h->sparse) {
            // this op was the consumer of its children's temporary slots
            for (int c : {h->eph_c1[v], h->eph_c2[v]})
                if (c >= n && is_temp[c]) {
                    is_temp[c] = 0;
                    sv_slot_release(h, h->cur_part[c], c);
                }
        }
    }
    if (h->sparse)
        for (int v : emitted)
            if (is_temp[v]) h->held.push_back(h->cur_part[v] * N + v);  // no consumer in this plan (inconsistent dirty set)
    // (the constant-site factor follows the really dirty nodes only: queued by the change-tracking pass above)
    h->dirty_list.clear();
    h->demand_list.clear();
    return SIEVE_OK;
}

// Register forwarding for the single-launch walk: when an op consumes the destination of the op right before it, that
// child is made child 1 (the two children's contributions commute) and flagged.
static void sv_link_forwarding(std::vector<SvOp>& ops, bool per_level) {
    for (size_t i = 0; i < ops.size(); i++) {
        SvOp& op = ops[i];
        op.flags &= ~SV_OP_FWD1;
        if (per_level || i == 0) continue;
        const unsigned prev = ops[i - 1].dst_row;
        const bool c1_int = !(op.flags & SV_OP_LEAF1);
        const bool c2_int = !(op.flags & SV_OP_UNARY) && !(op.flags & SV_OP_LEAF2);
        if (c2_int && op.c2_row == prev && !(c1_int && op.c1_row == prev)) {
            std::swap(op.c1_row, op.c2_row);
            std::swap(op.m1, op.m2);
            const int l1 = op.flags & SV_OP_LEAF1;
            op.flags &= ~(SV_OP_LEAF1 | SV_OP_LEAF2);
            if (l1) op.flags |= SV_OP_LEAF2;
        }
        if (!(op.flags & SV_OP_LEAF1) && op.c1_row == prev) op.flags |= SV_OP_FWD1;
    }
}

// L2 prefetch annotations for the single-launch walk (common.cuh, SV_OP_PF*): op i announces the leaf operands of op
// i + D.  Leaf partials are the walk's far reads (each is read exactly once per evaluation, so it always comes from DRAM:
// 3.2 of the 3.8 GB the walk reads at 1 000 cells x 100 000 sites); stored partials of internal nodes are mostly still in
// L2 when they are consumed and are not announced.
static void sv_link_prefetch(std::vector<SvOp>& ops, int D) {
    const int n = (int)ops.size();
    for (int j = 0; j < n; j++) {
        SvOp& op = ops[j];
        op.flags &= ~(SV_OP_PF1 | SV_OP_PF2 | SV_OP_FAR1 | SV_OP_FAR2);
        op.pf_row1 = op.pf_row2 = 0;
        if (D <= 0) continue;
        if (op.flags & SV_OP_LEAF1) op.flags |= SV_OP_FAR1;
        if (!(op.flags & SV_OP_UNARY) && (op.flags & SV_OP_LEAF2)) op.flags |= SV_OP_FAR2;
    }
    if (D <= 0) return;
    for (int i = 0; i + D < n; i++) {
        const SvOp& t = ops[i + D];
        SvOp& op = ops[i];
        if (t.flags & SV_OP_FAR1) {
            op.flags |= SV_OP_PF1;
            op.pf_row1 = t.c1_row;
        }
        if (t.flags & SV_OP_FAR2) {
            op.flags |= SV_OP_PF2;
            op.pf_row2 = t.c2_row;
        }
    }
}

// ---- streamed (ephemeral) evaluation: the host keeps parent -> children and regenerates the walk every time --------
static int sv_record_children(sieve_core* h, int c1, int c2, int parent) {
    const int n = h->n;
    SV_REQUIRE(parent >= n && parent < h->n_nodes, "op: parent must be an internal node");
    SV_REQUIRE(c1 >= 0 && c1 < h->n_nodes && c1 != h->root, "op: bad child1");
    SV_REQUIRE(c2 >= -1 && c2 < h->n_nodes && c2 != h->root, "op: bad child2");
    if (h->eph_c1[parent] != c1 || h->eph_c2[parent] != c2) h->topo_cache_ok = false;
    h->eph_c1[parent] = c1;
    h->eph_c2[parent] = c2;
    return SIEVE_OK;
}

// Post-order over the whole tree with the child that needs more scratch first (Sethi-Ullman), scratch slots assigned
// from a free list: a binary tree of n leaves needs at most ~log2(n) + 2 live partials.
static int sv_build_streamed_ops(sieve_core* h) {
    const int n = h->n, N = h->n_nodes, root = h->root;
    const long long cap = h->chunk_cap;
    std::vector<int> order;  // children before parents
    order.reserve(n);
    std::vector<std::pair<int, int>> st;
    st.push_back({root, 0});
    while (!st.empty()) {
        auto [v, state] = st.back();
        st.pop_back();
        if (v < n) continue;
        if (h->eph_c1[v] == -2) return sv_fail(SIEVE_ERR_STATE, "streamed evaluation: node " + std::to_string(v) + " has no children set yet");
        if (state == 0) {
            st.push_back({v, 1});
            if (h->eph_c2[v] >= 0) st.push_back({h->eph_c2[v], 0});
            st.push_back({h->eph_c1[v], 0});
        } else
            order.push_back(v);
    }
    std::vector<int> need(N, 0);
    for (int v : order) {
        int a = h->eph_c1[v], b = h->eph_c2[v];
        int na = a >= n ? need[a] : 0, nb = (b >= n) ? need[b] : 0;
        if (na < nb) std::swap(na, nb);
        const int internal = (a >= n) + (b >= n);
        need[v] = internal == 0 ? 1 : (internal == 1 ? std::max(na, 2) : std::max(std::max(na, nb + 1), 3));
    }
    if (need[root] > h->n_slots)
        return sv_fail(SIEVE_ERR_STATE, "streamed evaluation: the tree needs " + std::to_string(need[root]) + " scratch slots, " +
                                            std::to_string(h->n_slots) + " allocated");
    std::vector<int> free_slots;
    for (int s = h->n_slots - 1; s >= 0; s--) free_slots.push_back(s);
    std::vector<int> slot_of(N, -1);
    h->pending_ops.clear();
    h->pending_levels.clear();
    h->pending_per_level = false;
    st.push_back({root, 0});
    while (!st.empty()) {
        auto [v, state] = st.back();
        st.pop_back();
        if (v < n) continue;
        int a = h->eph_c1[v], b = h->eph_c2[v];
        if (state == 0) {
            st.push_back({v, 1});
            // heavier child first == pushed last
            const int na = a >= n ? need[a] : 0, nb = b >= n ? need[b] : 0;
            if (na >= nb) {
                if (b >= n) st.push_back({b, 0});
                if (a >= n) st.push_back({a, 0});
            } else {
                if (a >= n) st.push_back({a, 0});
                if (b >= n) st.push_back({b, 0});
            }
        } else {
            SvOp op;
            op.flags = 0;
            op.pf_row1 = op.pf_row2 = 0;
            const int dst = free_slots.back();
            free_slots.pop_back();
            slot_of[v] = dst;
            op.dst_row = (unsigned)((long long)dst * cap);
            auto child = [&](int c, int leaf_flag, unsigned* row, int* m) {
                if (c < n) {
                    op.flags |= leaf_flag;
                    *row = (unsigned)((long long)c * cap);
                } else
                    *row = (unsigned)((long long)slot_of[c] * cap);
                *m = h->cur_mat[c] * N + c;
                if (h->log_mode && h->slot_zero[*m]) op.flags |= SV_OP_ZEROD;
            };
            child(a, SV_OP_LEAF1, &op.c1_row, &op.m1);
            if (b < 0) {
                op.flags |= SV_OP_UNARY;
                op.c2_row = 0;
                op.m2 = 0;
            } else
                child(b, SV_OP_LEAF2, &op.c2_row, &op.m2);
            if (v == root) op.flags |= SV_OP_ROOT;
            sv_bound_op(h, v, a, b, &op);
            h->pending_ops.push_back(op);
            h->g_nodes.push_back(v);
            if (a >= n) free_slots.push_back(slot_of[a]);
            if (b >= n) free_slots.push_back(slot_of[b]);
        }
    }
    return SIEVE_OK;
}
