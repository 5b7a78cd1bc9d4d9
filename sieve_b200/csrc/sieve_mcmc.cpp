// sieve_mcmc.cpp -- native MCMC replay driver (include/sieve_b200_mcmc.h).  Uses only the public C-ABI of
// sieve_b200.h; compiled into the same shared library for convenience.
//
// What it mirrors of the reference (paths relative to /root/reference/src/beast/):
//   operator schedule and weights   examples/templates/smc_stage_1.xml:139-160, smc_stage_2.xml:85-95, rmc_stage_2.xml:123-176
//   trunk-aware tree moves          evolution/operators/{ScsUniform,ScsExchange,ScsSubtreeSlide,ScsWilsonBalding,
//                                   ScsTreeScaleOperator}.java (BEAST's moves with the unary root kept above the MRCA)
//   dirty logic                     evolution/likelihood/ScsTreeLikelihood.java:2786-2824 (requiresRecalculation),
//                                   :545-628 (traverse: matrix refresh when the branch time changed, parent recompute)
//   store / restore                 :2827-2880
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../include/sieve_b200_mcmc.h"

namespace {

struct Tree {
    int n = 0;
    std::vector<int> left, right, parent;
    std::vector<double> height;
    int root() const { return 2 * n - 1; }
    int node_count() const { return 2 * n; }
    int mrca() const { return left[root()]; }
    bool is_leaf(int i) const { return i < n; }
    void set_parents() {
        parent.assign(node_count(), -1);
        for (int p = n; p < node_count(); p++) {
            if (left[p] >= 0) parent[left[p]] = p;
            if (right[p] >= 0) parent[right[p]] = p;
        }
    }
    int sibling(int i) const {
        const int p = parent[i];
        return left[p] == i ? right[p] : left[p];
    }
    void replace_child(int p, int old_c, int new_c) {
        if (left[p] == old_c)
            left[p] = new_c;
        else
            right[p] = new_c;
        parent[new_c] = p;
    }
};

// ---- gamma quantiles (for the BEAST1-style discrete gamma of sitemodel/ScsSiteModel.java:16) ----
double reg_lower_gamma(double a, double x) {
    if (x <= 0) return 0.0;
    const double gln = std::lgamma(a);
    if (x < a + 1.0) {  // series
        double ap = a, sum = 1.0 / a, del = sum;
        for (int i = 0; i < 1000; i++) {
            ap += 1.0;
            del *= x / ap;
            sum += del;
            if (std::fabs(del) < std::fabs(sum) * 1e-17) break;
        }
        return sum * std::exp(-x + a * std::log(x) - gln);
    }
    // continued fraction (modified Lentz)
    double b = x + 1.0 - a, c = 1e300, d = 1.0 / b, h = d;
    for (int i = 1; i < 1000; i++) {
        const double an = -i * (i - a);
        b += 2.0;
        d = an * d + b;
        if (std::fabs(d) < 1e-300) d = 1e-300;
        c = b + an / c;
        if (std::fabs(c) < 1e-300) c = 1e-300;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (std::fabs(del - 1.0) < 1e-16) break;
    }
    return 1.0 - std::exp(-x + a * std::log(x) - gln) * h;
}

double gamma_quantile(double p, double shape, double scale) {
    // bisection on the regularised incomplete gamma: robust for the shape range the model allows (1e-3 .. 1e3)
    double lo = 0.0, hi = std::max(1.0, shape) * 4.0;
    while (reg_lower_gamma(shape, hi) < p) hi *= 2.0;
    for (int i = 0; i < 200; i++) {
        const double mid = 0.5 * (lo + hi);
        if (reg_lower_gamma(shape, mid) < p)
            lo = mid;
        else
            hi = mid;
        if (hi - lo <= 1e-16 * hi) break;
    }
    return 0.5 * (lo + hi) * scale;
}

void gamma_rates(double shape, int K, double* out) {
    double mean = 0.0;
    for (int i = 0; i < K; i++) {
        out[i] = gamma_quantile((2.0 * i + 1.0) / (2.0 * K), shape, 1.0 / shape);
        mean += out[i];
    }
    mean /= K;
    for (int i = 0; i < K; i++) out[i] /= mean;
}

struct OpWeight {
    int cls;
    double w;
    int sub;  // which parameter for the parameter classes
};

}  // namespace

struct sieve_mcmc {
    sieve_mcmc_config_t cfg;
    sieve_core_t* core = nullptr;
    Tree tree, stored_tree;
    std::vector<double> rate, stored_rate;          // per-branch rate multipliers (relaxed clock), 1 for strict
    std::vector<double> branch_time, stored_branch_time;  // m_branchLengths of ScsTreeLikelihood.java:205
    sieve_leaf_params_t leaf, stored_leaf;
    double gamma_shape = 1.0, stored_gamma_shape = 1.0;
    double cat_rates[4] = {1, 1, 1, 1};
    double logp = 0.0;
    std::mt19937_64 rng;
    std::vector<OpWeight> ops;
    double total_weight = 0.0;
    sieve_mcmc_stats_t st{};
    // scratch
    std::vector<char> dirty;      // operator-flagged nodes (Node.makeDirty)
    std::vector<char> need;       // internal nodes to recompute
    std::vector<int> mat_nodes;
    std::vector<double> dist;
    std::vector<int> op_list;
    std::vector<int> stack, order;
    std::vector<int> cand, rate_touched;  // evaluate(): nodes whose branch time may have moved
    bool scan_all = true;                 // ... or all of them (whole-tree moves, first evaluation)
    std::string err;
    std::vector<sieve_shard_result_t> world_results;
    sieve_combine_fn combine = nullptr;
    void* combine_user = nullptr;

    double unif() { return std::uniform_real_distribution<double>(0.0, 1.0)(rng); }
    int randint(int n) { return (int)std::uniform_int_distribution<int>(0, n - 1)(rng); }
};

namespace {

double branch_len(const Tree& t, int i) { return t.parent[i] < 0 ? 0.0 : t.height[t.parent[i]] - t.height[i]; }

// post-order list of the internal nodes flagged in m->need, children before parents, following only flagged nodes
void collect_ops(sieve_mcmc* m) {
    const Tree& t = m->tree;
    m->op_list.clear();
    if (!m->need[t.root()]) return;
    // iterative DFS from the root through flagged internal nodes
    std::vector<int>& out = m->order;
    std::vector<int>& work = m->stack;
    out.clear();
    work.clear();
    work.push_back(t.root());
    while (!work.empty()) {
        const int v = work.back();
        work.pop_back();
        out.push_back(v);
        const int l = t.left[v], r = t.right[v];
        if (l >= t.n && m->need[l]) work.push_back(l);
        if (r >= t.n && m->need[r]) work.push_back(r);
    }
    // `out` is a pre-order (parent before children); reversed it is a valid children-first order
    for (auto it = out.rbegin(); it != out.rend(); ++it) {
        m->op_list.push_back(*it);
        m->op_list.push_back(t.left[*it]);
        m->op_list.push_back(t.right[*it]);
    }
}

// evaluates the current state on the device; full: everything dirty (hasDirt != CLEAN)
int evaluate(sieve_mcmc* m, bool full, bool update_leaves, double* out_logp, int* n_ops, int* n_mats) {
    Tree& t = m->tree;
    const int N = t.node_count(), K = m->cfg.n_categories;
    std::fill(m->need.begin(), m->need.end(), 0);
    m->mat_nodes.clear();
    m->dist.clear();
    auto visit = [&](int i) {
        if (i == t.root()) return;
        const double bt = branch_len(t, i) * m->cfg.clock_rate * m->rate[i];
        // traverse, :563: refresh when the node is dirty or its branch time differs from the stored one
        if (full || m->dirty[i] || bt != m->branch_time[i]) {
            m->branch_time[i] = bt;
            m->mat_nodes.push_back(i);
            for (int k = 0; k < K; k++) m->dist.push_back(bt * m->cat_rates[k]);
            // its parent (and all ancestors) must be recomputed, :621-628
            for (int p = t.parent[i]; p >= 0 && !m->need[p]; p = t.parent[p]) m->need[p] = 1;
        }
    };
    if (full || m->scan_all) {
        for (int i = 0; i < N; i++) visit(i);
    } else {
        // The reference's traverse visits every node; a branch time can only have moved where an operator touched the
        // node (height or parent link: every move flags those) or its parent's height -- so the flagged nodes and their
        // children are the only candidates.  Same outcome, O(flagged) instead of O(nodes) per state (node order kept).
        m->cand = m->rate_touched;  // a rate proposal changes a branch time without flagging the node
        for (int i = 0; i < N; i++)
            if (m->dirty[i]) {
                m->cand.push_back(i);
                if (!t.is_leaf(i)) {
                    m->cand.push_back(t.left[i]);
                    if (t.right[i] >= 0) m->cand.push_back(t.right[i]);
                }
            }
        std::sort(m->cand.begin(), m->cand.end());
        m->cand.erase(std::unique(m->cand.begin(), m->cand.end()), m->cand.end());
        for (int i : m->cand) visit(i);
    }
    if (full || update_leaves)
        for (int i = t.n; i < N; i++) m->need[i] = 1;
    collect_ops(m);
    *n_ops = (int)m->op_list.size() / 3;
    *n_mats = (int)m->mat_nodes.size();
    if (*n_ops == 0 && !update_leaves) {
        *out_logp = m->logp;
        return SIEVE_OK;
    }
    if (update_leaves) {
        int rc = sieve_set_leaf_params(m->core, &m->leaf);
        if (rc != SIEVE_OK) return rc;
    }
    sieve_shard_result_t r;
    int rc = sieve_step_distances(m->core, update_leaves ? 1 : 0, *n_mats, m->mat_nodes.data(), m->dist.data(), *n_ops,
                                  m->op_list.data(), &r);
    if (rc != SIEVE_OK) return rc;
    if (m->combine)
        *out_logp = m->combine(&r, m->combine_user);
    else {
        // all ranks' results when a communicator is bound to the core (NCCL all-gather inside the step), else just ours
        const int world = sieve_comm_size(m->core);
        m->world_results.resize(world);
        rc = sieve_last_global_results(m->core, m->world_results.data());
        if (rc != SIEVE_OK) return rc;
        *out_logp = sieve_combine_shards(m->world_results.data(), world, m->cfg.asc_mode, m->cfg.n_background_sites);
    }
    return SIEVE_OK;
}

void store(sieve_mcmc* m) {
    sieve_store(m->core);
    m->stored_tree = m->tree;
    m->stored_rate = m->rate;
    m->stored_branch_time = m->branch_time;
    m->stored_leaf = m->leaf;
    m->stored_gamma_shape = m->gamma_shape;
}

void restore(sieve_mcmc* m) {
    sieve_restore(m->core);
    m->tree.left.swap(m->stored_tree.left);
    m->tree.right.swap(m->stored_tree.right);
    m->tree.parent.swap(m->stored_tree.parent);
    m->tree.height.swap(m->stored_tree.height);
    m->rate.swap(m->stored_rate);
    m->branch_time.swap(m->stored_branch_time);
    const bool leaf_changed = memcmp(&m->leaf, &m->stored_leaf, sizeof(m->leaf)) != 0;
    m->leaf = m->stored_leaf;
    if (leaf_changed) sieve_set_leaf_params(m->core, &m->leaf);
    if (m->gamma_shape != m->stored_gamma_shape) {
        m->gamma_shape = m->stored_gamma_shape;
        gamma_rates(m->gamma_shape, m->cfg.n_categories, m->cat_rates);
    }
}

// ---- tree moves: return false when the proposal is invalid (counted as rejected), *log_hastings otherwise ----
double scale_draw(sieve_mcmc* m, double sf) { return sf + m->unif() * (1.0 / sf - sf); }

bool move_tree_scale(sieve_mcmc* m, double* lh) {
    Tree& t = m->tree;
    const double s = scale_draw(m, 0.95);
    for (int i = t.n; i < t.node_count(); i++) t.height[i] *= s;  // leaves stay at height 0
    m->scan_all = true;  // every branch time moves and no node is flagged: evaluate() looks at all of them
    *lh = (t.n - 2) * std::log(s);
    return true;
}

bool move_root_scale(sieve_mcmc* m, double* lh) {
    Tree& t = m->tree;
    const double s = scale_draw(m, 0.95);
    const double nh = t.height[t.root()] * s;
    if (nh <= t.height[t.mrca()]) return false;
    t.height[t.root()] = nh;
    m->dirty[t.root()] = 1;
    *lh = -std::log(s);
    return true;
}

bool move_uniform(sieve_mcmc* m, double* lh) {
    Tree& t = m->tree;
    if (t.n < 2) return false;
    const int node = t.n + m->randint(t.n - 1);  // an internal node below the root
    const double lo = std::max(t.height[t.left[node]], t.height[t.right[node]]);
    const double hi = t.height[t.parent[node]];
    t.height[node] = lo + m->unif() * (hi - lo);
    m->dirty[node] = 1;
    *lh = 0.0;
    return true;
}

// BEAST's SubtreeSlide on the binary part; the parent of the chosen node slides, possibly changing the topology
bool move_subtree_slide(sieve_mcmc* m, double* lh) {
    Tree& t = m->tree;
    if (t.n < 3) return false;
    int i;
    do {
        i = m->randint(t.node_count() - 1);
    } while (i == t.mrca() || i == t.root());
    const int p = t.parent[i];
    const int cip = t.sibling(i);
    const int pip = t.parent[p];
    const double size = 0.1 * t.height[t.mrca()];
    const double delta = std::normal_distribution<double>(0.0, size)(m->rng);
    const double old_h = t.height[p];
    const double new_h = old_h + delta;
    *lh = 0.0;
    if (delta > 0) {
        if (pip >= 0 && t.height[pip] < new_h) {
            // slide up past ancestors: find the edge (new_child, new_parent) that spans new_h
            int new_parent = pip, new_child = p;
            while (t.height[new_parent] < new_h) {
                new_child = new_parent;
                new_parent = t.parent[new_parent];
                if (new_parent < 0) break;
            }
            if (new_parent < 0) return false;  // above the root
            // detach p (with i), connect cip to pip
            t.replace_child(pip, p, cip);
            // insert p between new_child and new_parent
            t.replace_child(new_parent, new_child, p);
            if (t.left[p] == i)
                t.right[p] = new_child;
            else
                t.left[p] = new_child;
            t.parent[new_child] = p;
            t.height[p] = new_h;
            m->dirty[p] = m->dirty[cip] = m->dirty[new_child] = m->dirty[pip] = m->dirty[new_parent] = 1;
            // number of edges the reverse move could choose from
            int cnt = 0;
            std::vector<int>& st = m->stack;
            st.clear();
            st.push_back(new_child);
            while (!st.empty()) {
                const int v = st.back();
                st.pop_back();
                if (t.height[v] <= old_h) {
                    cnt++;
                    continue;
                }
                if (!t.is_leaf(v)) {
                    st.push_back(t.left[v]);
                    if (t.right[v] >= 0) st.push_back(t.right[v]);
                }
            }
            *lh = std::log(1.0 / std::max(cnt, 1));
        } else {
            if (pip >= 0 && new_h >= t.height[pip]) return false;
            t.height[p] = new_h;
            m->dirty[p] = 1;
        }
    } else {
        if (t.height[i] > new_h) return false;
        if (t.height[cip] > new_h) {
            // slide down into the sibling's subtree: edges spanning new_h
            std::vector<int> cand;
            std::vector<int>& st = m->stack;
            st.clear();
            st.push_back(cip);
            while (!st.empty()) {
                const int v = st.back();
                st.pop_back();
                if (t.height[v] <= new_h) {
                    cand.push_back(v);
                    continue;
                }
                if (!t.is_leaf(v)) {
                    st.push_back(t.left[v]);
                    st.push_back(t.right[v]);
                }
            }
            if (cand.empty()) return false;
            const int new_child = cand[m->randint((int)cand.size())];
            const int new_parent = t.parent[new_child];
            // detach p, connect cip to pip
            t.replace_child(pip, p, cip);
            t.replace_child(new_parent, new_child, p);
            if (t.left[p] == i)
                t.right[p] = new_child;
            else
                t.left[p] = new_child;
            t.parent[new_child] = p;
            t.height[p] = new_h;
            m->dirty[p] = m->dirty[cip] = m->dirty[new_child] = m->dirty[pip] = m->dirty[new_parent] = 1;
            *lh = std::log((double)cand.size());
        } else {
            t.height[p] = new_h;
            m->dirty[p] = 1;
        }
    }
    return true;
}

void exchange_nodes(sieve_mcmc* m, int i, int j) {
    Tree& t = m->tree;
    const int ip = t.parent[i], jp = t.parent[j];
    // swap the subtrees rooted at i and j
    if (t.left[ip] == i)
        t.left[ip] = j;
    else
        t.right[ip] = j;
    if (t.left[jp] == j)
        t.left[jp] = i;
    else
        t.right[jp] = i;
    t.parent[i] = jp;
    t.parent[j] = ip;
    m->dirty[i] = m->dirty[j] = m->dirty[ip] = m->dirty[jp] = 1;
}

bool move_narrow(sieve_mcmc* m, double* lh) {
    Tree& t = m->tree;
    if (t.n < 3) return false;
    // grandparent with at least one internal child, below the root (BEAST Exchange.narrow)
    int gp, tries = 0;
    do {
        gp = t.n + m->randint(t.n - 1);
        if (++tries > 10 * t.n) return false;
    } while (t.is_leaf(t.left[gp]) && t.is_leaf(t.right[gp]));
    int p = t.left[gp], uncle = t.right[gp];
    if (t.height[p] < t.height[uncle]) std::swap(p, uncle);
    if (t.is_leaf(p)) return false;
    const int i = m->unif() < 0.5 ? t.left[p] : t.right[p];
    exchange_nodes(m, i, uncle);
    *lh = 0.0;
    return true;
}

bool move_wide(sieve_mcmc* m, double* lh) {
    Tree& t = m->tree;
    if (t.n < 4) return false;
    int i, j;
    do {
        i = m->randint(t.node_count() - 1);
    } while (i == t.mrca() || i == t.root());
    do {
        j = m->randint(t.node_count() - 1);
    } while (j == i || j == t.mrca() || j == t.root());
    const int ip = t.parent[i], jp = t.parent[j];
    if (ip == jp || i == jp || j == ip) return false;
    if (t.height[j] >= t.height[ip] || t.height[i] >= t.height[jp]) return false;
    exchange_nodes(m, i, j);
    *lh = 0.0;
    return true;
}

bool move_wilson_balding(sieve_mcmc* m, double* lh) {
    Tree& t = m->tree;
    if (t.n < 4) return false;
    int i;
    do {
        i = m->randint(t.node_count() - 1);
    } while (i == t.mrca() || i == t.root());
    const int ip = t.parent[i];
    int j, jp, tries = 0;
    do {
        j = m->randint(t.node_count() - 1);
        jp = j == t.root() ? -1 : t.parent[j];
        if (++tries > 100) return false;
    } while (j == t.root() || j == t.mrca() || jp < 0 || t.height[jp] <= t.height[i] || j == i || jp == t.root());
    if (jp == ip || j == ip || jp == i) return false;
    const int cip = t.sibling(i);
    const int pip = t.parent[ip];
    if (pip < 0 || pip == t.root()) return false;  // keep the MRCA / trunk structure out of this simplified move
    const double new_min = std::max(t.height[i], t.height[j]);
    const double new_range = t.height[jp] - new_min;
    const double new_age = new_min + m->unif() * new_range;
    const double old_min = std::max(t.height[i], t.height[cip]);
    const double old_range = t.height[pip] - old_min;
    // prune ip (with i) and regraft it on the edge (j, jp)
    t.replace_child(pip, ip, cip);
    t.replace_child(jp, j, ip);
    if (t.left[ip] == i)
        t.right[ip] = j;
    else
        t.left[ip] = j;
    t.parent[j] = ip;
    t.height[ip] = new_age;
    m->dirty[ip] = m->dirty[cip] = m->dirty[j] = m->dirty[pip] = m->dirty[jp] = 1;
    *lh = std::log(new_range / std::fabs(old_range));
    return true;
}

bool move_branch_rate(sieve_mcmc* m, double* lh) {
    Tree& t = m->tree;
    int i;
    do {
        i = m->randint(t.node_count() - 1);
    } while (i == t.root());
    const double s = scale_draw(m, 0.75);
    m->rate[i] *= s;
    m->rate_touched.push_back(i);
    *lh = -std::log(s);
    return true;
}

bool move_param(sieve_mcmc* m, int cls, int sub, double* lh) {
    const double s = scale_draw(m, 0.9);
    *lh = -std::log(s);
    double* p = nullptr;
    double lo = 0.0, hi = 1e300;
    if (cls == SIEVE_OP_COVERAGE) {
        p = sub == 0 ? &m->leaf.allelic_seq_cov : &m->leaf.allelic_seq_cov_raw_var;
        hi = 1000.0;  // bounds of smc_stage_1.xml:31-32
    } else if (cls == SIEVE_OP_NUC) {
        p = sub == 0 ? &m->leaf.eff_seq_err_rate : (sub == 1 ? &m->leaf.shape_ctrl1 : &m->leaf.shape_ctrl2);
        hi = sub == 0 ? 1.0 : 1e4;
    } else if (cls == SIEVE_OP_ADO) {
        p = &m->leaf.ado_rate;
        hi = 1.0;
    } else if (cls == SIEVE_OP_GAMMA_SHAPE) {
        p = &m->gamma_shape;
        lo = 1e-3;
        hi = 1e3;
    }
    const double v = *p * s;
    if (!(v > lo) || !(v < hi)) return false;
    *p = v;
    if (cls == SIEVE_OP_GAMMA_SHAPE) gamma_rates(m->gamma_shape, m->cfg.n_categories, m->cat_rates);
    return true;
}

void build_mix(sieve_mcmc* m) {
    auto add = [&](int cls, double w, int sub = 0) {
        m->ops.push_back({cls, w, sub});
        m->total_weight += w;
    };
    // tree moves, weights of smc_stage_1.xml:139-147 (identical in stage 2, smc_stage_2.xml:85-95)
    add(SIEVE_OP_TREE_SCALE, 3);
    add(SIEVE_OP_ROOT_SCALE, 3);
    add(SIEVE_OP_UNIFORM, 30);
    add(SIEVE_OP_SUBTREE_SLIDE, 15);
    add(SIEVE_OP_NARROW, 15);
    add(SIEVE_OP_WIDE, 10);
    add(SIEVE_OP_WILSON_BALDING, 5);
    if (m->cfg.mix == SIEVE_MIX_STAGE1) {
        add(SIEVE_OP_PRIOR_ONLY, 5);        // ePopSize 3 + growthRate 2, :148-149
        add(SIEVE_OP_COVERAGE, 1, 0);       // allelicSeqCov, :152
        add(SIEVE_OP_COVERAGE, 1, 1);       // allelicSeqCovRawVar, :153
        add(SIEVE_OP_NUC, 2, 0);            // effSeqErrRate
        add(SIEVE_OP_NUC, 2, 1);            // shapeCtrl1
        add(SIEVE_OP_NUC, 2, 2);            // shapeCtrl2
        add(SIEVE_OP_ADO, 3);               // adoRate
        add(SIEVE_OP_GAMMA_SHAPE, 2);       // gammaShape
    } else if (m->cfg.mix == SIEVE_MIX_STAGE2) {
        add(SIEVE_OP_PRIOR_ONLY, 5);
    } else {
        add(SIEVE_OP_PRIOR_ONLY, 5);
        add(SIEVE_OP_BRANCH_RATE, 38);      // ORC rate operators, rmc_stage_2.xml:123-176 (20 + 15 + 3)
    }
}

}  // namespace

extern "C" int sieve_gamma_category_rates(double shape, int32_t k, double* rates_out) {
    if (!(shape > 0) || k < 1 || k > 16 || !rates_out) return SIEVE_ERR_INVALID;
    gamma_rates(shape, k, rates_out);
    return SIEVE_OK;
}

extern "C" int sieve_mcmc_create(const sieve_mcmc_config_t* cfg, sieve_core_t* core, const int32_t* left,
                                 const int32_t* right, const double* height, const double* branch_rates,
                                 sieve_mcmc_t** out) {
    if (!cfg || !core || !left || !right || !height || !out) return SIEVE_ERR_INVALID;
    if (cfg->n_leaves < 2 || cfg->n_categories < 1 || cfg->n_categories > 4) return SIEVE_ERR_INVALID;
    sieve_mcmc* m = new sieve_mcmc();
    m->cfg = *cfg;
    m->core = core;
    const int n = cfg->n_leaves, N = 2 * n;
    m->tree.n = n;
    m->tree.left.assign(left, left + N);
    m->tree.right.assign(right, right + N);
    m->tree.height.assign(height, height + N);
    m->tree.set_parents();
    m->rate.assign(N, 1.0);
    if (branch_rates) m->rate.assign(branch_rates, branch_rates + N);
    m->branch_time.assign(N, -1.0);
    m->leaf = cfg->leaf;
    m->gamma_shape = cfg->gamma_shape;
    gamma_rates(m->gamma_shape, cfg->n_categories, m->cat_rates);
    m->rng.seed(cfg->seed);
    m->dirty.assign(N, 0);
    m->need.assign(N, 0);
    build_mix(m);
    double prop[4] = {0, 0, 0, 0};
    for (int k = 0; k < cfg->n_categories; k++) prop[k] = 1.0 / cfg->n_categories;
    int rc = sieve_set_root(core, m->tree.root(), prop, 0);
    if (rc == SIEVE_OK) rc = sieve_set_leaf_params(core, &m->leaf);
    if (rc == SIEVE_OK) rc = sieve_update_leaves(core, 0);
    int no = 0, nm = 0;
    if (rc == SIEVE_OK) rc = evaluate(m, true, false, &m->logp, &no, &nm);
    if (rc != SIEVE_OK) {
        delete m;
        return rc;
    }
    m->stored_tree = m->tree;
    m->stored_rate = m->rate;
    m->stored_branch_time = m->branch_time;
    m->stored_leaf = m->leaf;
    m->stored_gamma_shape = m->gamma_shape;
    *out = m;
    return SIEVE_OK;
}

extern "C" int sieve_mcmc_set_combine(sieve_mcmc_t* m, sieve_combine_fn combine, void* user) {
    if (!m) return SIEVE_ERR_INVALID;
    m->combine = combine;
    m->combine_user = user;
    // the cached log-likelihood must be the global one from now on
    int no = 0, nm = 0;
    sieve_store(m->core);
    return evaluate(m, true, false, &m->logp, &no, &nm);
}

extern "C" int sieve_mcmc_destroy(sieve_mcmc_t* m) {
    delete m;
    return SIEVE_OK;
}

extern "C" int sieve_mcmc_run(sieve_mcmc_t* m, int64_t n_states, int32_t reset, sieve_mcmc_stats_t* out) {
    if (!m || n_states < 0) return SIEVE_ERR_INVALID;
    if (reset) m->st = sieve_mcmc_stats_t{};
    const auto t_begin = std::chrono::steady_clock::now();
    for (int64_t s = 0; s < n_states; s++) {
        const auto t0 = std::chrono::steady_clock::now();
        // draw the operator
        double u = m->unif() * m->total_weight;
        size_t oi = 0;
        for (; oi + 1 < m->ops.size(); oi++) {
            if (u < m->ops[oi].w) break;
            u -= m->ops[oi].w;
        }
        const OpWeight& ow = m->ops[oi];
        m->st.proposed[ow.cls]++;
        m->st.states++;
        bool accepted = false;
        if (ow.cls == SIEVE_OP_PRIOR_ONLY) {
            // a prior-only parameter: the likelihood is not dirty, BEAST takes the cached value
            accepted = m->unif() < 0.3;
        } else {
            store(m);
            std::fill(m->dirty.begin(), m->dirty.end(), 0);
            m->rate_touched.clear();
            m->scan_all = false;
            double lh = 0.0;
            bool valid = false, full = false, leaves = false;
            switch (ow.cls) {
                case SIEVE_OP_TREE_SCALE: valid = move_tree_scale(m, &lh); break;
                case SIEVE_OP_ROOT_SCALE: valid = move_root_scale(m, &lh); break;
                case SIEVE_OP_UNIFORM: valid = move_uniform(m, &lh); break;
                case SIEVE_OP_SUBTREE_SLIDE: valid = move_subtree_slide(m, &lh); break;
                case SIEVE_OP_NARROW: valid = move_narrow(m, &lh); break;
                case SIEVE_OP_WIDE: valid = move_wide(m, &lh); break;
                case SIEVE_OP_WILSON_BALDING: valid = move_wilson_balding(m, &lh); break;
                case SIEVE_OP_BRANCH_RATE:
                    valid = move_branch_rate(m, &lh);
                    // isForcingTreeDirtyRMC: with Felsenstein's correction a rate change dirties the whole tree
                    full = m->cfg.asc_mode == SIEVE_ASC_FELSENSTEIN_MEAN || m->cfg.asc_mode == SIEVE_ASC_FELSENSTEIN_GEOMETRIC;
                    break;
                case SIEVE_OP_GAMMA_SHAPE:
                    valid = move_param(m, ow.cls, ow.sub, &lh);
                    full = true;  // siteModel dirty -> IS_DIRTY, :2809-2812
                    break;
                default:  // coverage / nucleotide / dropout parameters: rawReadCountsModel dirty, :2798-2802
                    valid = move_param(m, ow.cls, ow.sub, &lh);
                    full = true;
                    leaves = true;
                    break;
            }
            if (valid) {
                double lp = 0.0;
                int no = 0, nm = 0;
                const int rc = evaluate(m, full, leaves, &lp, &no, &nm);
                if (rc != SIEVE_OK) return rc;
                m->st.likelihood_evaluations += (no > 0 || leaves) ? 1 : 0;
                m->st.ops_total += no;
                m->st.matrices_total += nm;
                m->st.leaf_updates += leaves ? 1 : 0;
                const double log_alpha = lp - m->logp + lh;
                if (std::isfinite(lp) && (log_alpha >= 0.0 || std::log(m->unif()) < log_alpha)) {
                    accepted = true;
                    m->logp = lp;
                }
            }
            if (!accepted) restore(m);
        }
        if (accepted) {
            m->st.accepted++;
            m->st.accepted_by_op[ow.cls]++;
        }
        m->st.seconds_by_op[ow.cls] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    m->st.seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    m->st.log_p = m->logp;
    if (out) *out = m->st;
    return SIEVE_OK;
}

extern "C" int sieve_mcmc_get_state(sieve_mcmc_t* m, int32_t* left, int32_t* right, double* height, double* branch_rates,
                                    sieve_leaf_params_t* leaf, double* gamma_shape, double* log_p) {
    if (!m) return SIEVE_ERR_INVALID;
    const int N = m->tree.node_count();
    if (left) memcpy(left, m->tree.left.data(), sizeof(int32_t) * N);
    if (right) memcpy(right, m->tree.right.data(), sizeof(int32_t) * N);
    if (height) memcpy(height, m->tree.height.data(), sizeof(double) * N);
    if (branch_rates) memcpy(branch_rates, m->rate.data(), sizeof(double) * N);
    if (leaf) *leaf = m->leaf;
    if (gamma_shape) *gamma_shape = m->gamma_shape;
    if (log_p) *log_p = m->logp;
    return SIEVE_OK;
}

extern "C" int sieve_mcmc_full_recompute(sieve_mcmc_t* m, double* log_p) {
    if (!m || !log_p) return SIEVE_ERR_INVALID;
    // like BEAST's robust check: everything filthy, leaves included, into the flipped buffers; then adopt it
    sieve_store(m->core);
    int no = 0, nm = 0;
    const int rc = evaluate(m, true, true, log_p, &no, &nm);
    return rc;
}

// ---- the reference's own traverse sequence, call by call ----------------------------------------------------------------
// What ScsTreeLikelihood.traverse issues for a whole-tree update when the core behind the LikelihoodCore seam is swapped
// and NOTHING else in the plugin changes (likelihood/ScsTreeLikelihood.java:563-580 matrices, :583-601 leaves, :621-628 +
// :745-752 / :854-861 partials, :888-920 root; calcLogP :2621-2651): per non-root node setNodeMatrixForUpdate and K x
// setNodeMatrix (16 doubles each), per internal node setNodePartialsForUpdate and calculateLogPartials, then the
// evaluation and getPatternLogLikelihoods.  Every call goes through the public C-ABI, natively, as the JNI shim forwards
// it -- this is the number for the no-extra-edit integration (the batched sieve_step_distances needs one more edit).
extern "C" int sieve_traverse_full(sieve_core_t* h, int32_t n_nodes, int32_t n_categories, int32_t root, int32_t n_ops,
                                   const int32_t* ops_post_order, const double* P, int32_t update_leaves,
                                   double* site_log_l, sieve_shard_result_t* out) {
    if (!h || !ops_post_order || !P || n_ops < 0) return SIEVE_ERR_INVALID;
    int r;
    const int n_leaves = n_nodes / 2;
    for (int node = 0; node < n_nodes; node++) {
        if (node == root) continue;
        if ((r = sieve_set_node_matrix_for_update(h, node)) != SIEVE_OK) return r;
        for (int k = 0; k < n_categories; k++)
            if ((r = sieve_set_node_matrix(h, node, k, P + ((size_t)node * n_categories + k) * 16)) != SIEVE_OK) return r;
    }
    if (update_leaves && (r = sieve_update_leaves(h, 1)) != SIEVE_OK) return r;
    for (int i = 0; i < n_ops; i++) {
        const int parent = ops_post_order[3 * i], c1 = ops_post_order[3 * i + 1], c2 = ops_post_order[3 * i + 2];
        if ((r = sieve_set_node_partials_for_update(h, parent)) != SIEVE_OK) return r;
        if ((r = sieve_calculate_partials(h, c1, c1 < n_leaves, c2, c2 >= 0 && c2 < n_leaves, parent, 0)) != SIEVE_OK) return r;
    }
    if ((r = sieve_evaluate(h, out)) != SIEVE_OK) return r;
    if (site_log_l && (r = sieve_get_site_log_likelihoods(h, site_log_l, nullptr)) != SIEVE_OK) return r;
    return SIEVE_OK;
}
