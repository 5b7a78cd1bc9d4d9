"""Device side of the per-pattern coverage model (`PrivateAllelicSeqCovModel`) against
oracle/private_allelic.py: category-dependent leaf partials, site log-likelihoods and logConstRoot, the max-sum trace
with category-dependent leaves, and the accept-time re-estimation of (allelicSeqCov, allelicSeqCovRawVar).

Tolerances as for the shared model: per-site log-likelihood 1e-12 relative, leaf / node log partials 1e-12 relative
(+1e-12 absolute), re-estimated (t, v) 1e-12 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from sieve_b200 import substmodel  # noqa: E402
from sieve_b200.data import synthetic  # noqa: E402

PARAMS = (0.3, 50.0, 50.0, 0.008, 150.0, 2.0)


@pytest.fixture(scope="module")
def core_mod():
    from sieve_b200 import _lib, core
    assert _lib.lib().sieve_device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return core


def _pairs(K, S, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(15.0, 90.0, (K, S)), rng.uniform(5.0, 120.0, (K, S))


def _setup(core_mod, oracle, n, S, seed, single=True, K=4, asc=True, pairs_seed=None):
    d = synthetic(n, S, seed=seed)
    counts, tree = d["counts"], d["tree"]
    c5 = oracle.counts5(counts)
    sf, _ = oracle.size_factors(c5)
    rates, props = substmodel.gamma_category_rates(0.8, K)
    mats = substmodel.node_matrices(tree, rates)[:, :K]
    t, v = _pairs(K, S, seed if pairs_seed is None else pairs_seed)
    c = core_mod.ScsCudaLikelihoodCore(n, S, K, 4, True, asc, single, private_seq_cov=True)
    c.set_counts(counts)
    c.set_size_factors(sf)
    c.set_leaf_params(*PARAMS)
    c.set_private_seq_cov(t, v)
    c.update_leaves(for_update=False)
    c.set_root(tree.root, props, 0)
    nodes = [i for i in range(tree.node_count) if i != tree.root]
    c.set_distances(nodes, (tree.branch_lengths()[nodes][:, None] * rates[None, :K]), for_update=False)
    c.update_nodes(tree.ops(), flip=False)
    return c, tree, c5, sf, mats, props, t, v


def _close(a, b, rtol, atol=0.0):
    a, b = np.asarray(a), np.asarray(b)
    assert np.all(np.abs(a - b) <= rtol * np.abs(b) + atol), float(np.max(np.abs(a - b) / (np.abs(b) + atol + 1e-300)))


@pytest.mark.parametrize("single,K,asc", [(True, 4, True), (False, 4, True), (True, 2, False)])
def test_private_leaves_sites_and_constants_match_the_oracle(core_mod, oracle, single, K, asc):
    from oracle import private_allelic as pa
    n, S = 9, 70
    c, tree, c5, sf, mats, props, t, v = _setup(core_mod, oracle, n, S, 91, single=single, K=K, asc=asc)
    r = c.evaluate()
    P = oracle.leaf_params(*PARAMS, single_ado=single)
    sl, lc, partials = pa.full_eval_private(oracle, c5, sf, P, t, v, tree.ops(), mats, props, tree.root)
    got_l, got_c = c.site_log_likelihoods()
    _close(got_l, sl, 1e-12)
    if asc:
        _close(got_c, lc, 1e-12)
    assert abs(r.sum_log_l - sl.sum()) <= 1e-9 * abs(sl.sum())
    for node in (0, n // 2, n - 1, n + 1, tree.root - 1):
        _close(c.get_node_partials(node), partials[node], 1e-12, 1e-12)
    tt, vv = c.get_private_seq_cov()
    assert np.array_equal(tt, t) and np.array_equal(vv, v)
    c.close()


def test_equal_pairs_reduce_to_the_shared_model(core_mod, oracle):
    """With every (category, pattern) pair at the shared (t, v) the private model IS the shared model
    (PrivateAllelicSeqCovModel.java:159-161): same site log-likelihoods as a shared-model core (different kernels, so to
    rounding, not bit for bit), through a leaf-parameter proposal, a reject and a local move."""
    from sieve_b200.treelikelihood import LeafModelParams, ScsTreeLikelihood
    n, S = 30, 500
    d = synthetic(n, S, seed=17)
    counts, tree = d["counts"], d["tree"]
    sf, _ = oracle.size_factors(oracle.counts5(counts))
    out = {}
    for private in (False, True):
        tl = ScsTreeLikelihood(counts, tree.copy(), size_factors=sf, asc_mode="felsenstein_mean", n_background_sites=5e3,
                               private_seq_cov=private)
        seq = [tl.calculate_logp(), tl.pattern_log_likelihoods()]
        tl.store()
        tl.set_leaf_params(LeafModelParams(0.22, 50.0, 50.0, 0.012, 110.0, 2.4))
        seq += [tl.calculate_logp(), tl.pattern_log_likelihoods()]
        tl.restore()
        tl.store()
        moved = tree.copy()
        vtx = n + n // 3
        lo = max(moved.height[moved.left[vtx]], moved.height[moved.right[vtx]])
        moved.height[vtx] = lo + 0.3 * (moved.height[moved.parent[vtx]] - lo)
        tl.set_tree(moved, [vtx])
        seq += [tl.calculate_logp(), tl.pattern_log_likelihoods()]
        tl.restore()
        seq += [tl.calculate_logp(), tl.pattern_log_likelihoods()]
        out[private] = seq
        tl.close()
    for a, b in zip(out[True], out[False]):
        if isinstance(a, tuple):
            _close(a[0], b[0], 2e-13)
            _close(a[1], b[1], 2e-13)
        else:
            assert abs(a - b) <= 1e-11 * abs(b)
    assert out[True][0] == out[True][6]  # restore brings the first state back bit for bit


@pytest.mark.parametrize("single", [True, False])
def test_max_sum_trace_with_category_dependent_leaves(core_mod, oracle, single):
    from oracle import ml_genotypes as mlo
    from oracle import private_allelic as pa
    n, S = 7, 40
    c, tree, c5, sf, mats, props, t, v = _setup(core_mod, oracle, n, S, 23, single=single)
    c.evaluate()
    c.ml_trace()
    P = oracle.leaf_params(*PARAMS, single_ado=single)
    tr = pa.ml_trace_private(oracle, mlo, c5, sf, P, t, v, tree.ops(), mats)
    for leaf in range(n):
        p, b, _ = c.ml_node(leaf)
        _close(p, tr["leaf"][leaf], 1e-12, 1e-12)
        assert np.array_equal(b, tr["ado"][leaf])
    for node in range(n, 2 * n):
        p, b, _ = c.ml_node(node)
        _close(p, tr["ml"][node], 1e-12)
        assert (b != tr["back"][node]).mean() <= 2e-3
    c.close()


@pytest.mark.parametrize("single,zero_cov_mode", [(True, 0), (False, 0), (True, 1)])
def test_accept_time_reestimation_matches_the_oracle(core_mod, oracle, single, zero_cov_mode):
    """sieve_private_update_from_ml == ScsTreeLikelihood.postProcess steps 1-2 restated (oracle.private_allelic.post_process)
    on every pair with a unique ML combination; tied pairs are flagged and resolved by sieve_b200.private_cov exactly as
    the oracle's enumeration does; the recomputed state then equals the oracle's evaluation under the new pairs."""
    from oracle import ml_genotypes as mlo
    from oracle import private_allelic as pa
    from sieve_b200 import private_cov
    n, S = 8, 48
    c, tree, c5, sf, mats, props, t, v = _setup(core_mod, oracle, n, S, 57, single=single)
    c.evaluate()
    P = oracle.leaf_params(*PARAMS, single_ado=single)
    t_ref, v_ref, ncomb = pa.post_process(oracle, mlo, c5, sf, P, t, v, tree.ops(), mats, zero_cov_mode)
    changed, tied = c.private_update_from_ml(zero_cov_mode)
    assert np.array_equal(tied != 0, ncomb > 1)
    t_new, v_new = c.get_private_seq_cov()
    uniq = ncomb == 1
    assert uniq.mean() > 0.9
    _close(t_new[uniq], t_ref[uniq], 1e-12)
    _close(v_new[uniq], v_ref[uniq], 1e-12)
    assert changed == int(((t_new != t) | (v_new != v)).sum()) and changed > 0
    # groups: both allele numbers must occur among the ML states, or the test says little
    assert np.array_equal(t_new[~uniq], t[~uniq])
    counts4 = np.concatenate([c5[:, :, :3], c5[:, :, 4:5]], axis=2)
    moved = private_cov.resolve_tied_pairs(c, tree, counts4, sf, tied, zero_cov_mode, single)
    t_fin, v_fin = c.get_private_seq_cov()
    _close(t_fin, t_ref, 1e-12)
    _close(v_fin, v_ref, 1e-12)
    assert moved <= int((~uniq).sum())
    # the state under the new pairs (postProcess step 3)
    c.update_leaves(for_update=True)
    c.update_nodes(tree.ops(), flip=True)
    c.evaluate()
    sl, lc, _ = pa.full_eval_private(oracle, c5, sf, P, t_fin, v_fin, tree.ops(), mats, props, tree.root)
    got_l, got_c = c.site_log_likelihoods()
    _close(got_l, sl, 1e-12)
    _close(got_c, lc, 1e-12)
    # idempotent: a second pass over the same ML states moves nothing but pairs whose ML states changed with (t, v)
    c.close()


def test_accept_mirror_and_idempotence(core_mod, oracle):
    """ScsTreeLikelihood.accept() drives the whole post-process; repeating it converges (a pair whose ML numbers of
    sequenced alleles did not change reproduces its (t, v) exactly and is not counted)."""
    from sieve_b200.treelikelihood import ScsTreeLikelihood
    n, S = 16, 300
    d = synthetic(n, S, seed=5)
    counts, tree = d["counts"], d["tree"]
    sf, _ = oracle.size_factors(oracle.counts5(counts))
    tl = ScsTreeLikelihood(counts, tree.copy(), size_factors=sf, asc_mode="felsenstein_mean", n_background_sites=5e3,
                           private_seq_cov=True)
    lp0 = tl.calculate_logp()
    moved = [tl.accept() for _ in range(6)]
    assert moved[0] > 0.5 * 4 * S           # nearly every pair leaves the shared start values
    assert moved[-1] < 0.05 * moved[0]      # and the iteration settles
    lp1 = tl.logP
    assert np.isfinite(lp1) and lp1 != lp0
    t, v = tl.core.get_private_seq_cov()
    assert np.all(t > 0) and np.all(v > 0)
    tl.close()


def test_retraverse_of_changed_patterns_equals_a_whole_shard_pass(core_mod, oracle):
    """ScsTreeLikelihood.postProcess step 3 (:2493-2498): sieve_private_retraverse_changed re-evaluates, in place, only the
    patterns whose (t, v) moved.  Chain A uses it, chain B the whole-shard pass (leaf flip + every node): logP, both per-site
    vectors and node read-backs must agree bit for bit through accepted and rejected tree moves in between, and the later
    accepts must touch fewer sites than the shard has.  (An accept right after a REJECTED move is not a sequence BEAST produces;
    the per-site outputs then belong to the rejected proposal -- as the reference's patternLogLikelihoods would -- and the core
    re-evaluates every site.)"""
    from sieve_b200.treelikelihood import ScsTreeLikelihood
    n, S = 19, 333          # ragged: not a multiple of the kernels' site blocks
    d = synthetic(n, S, seed=11)
    counts, tree = d["counts"], d["tree"]
    sf, _ = oracle.size_factors(oracle.counts5(counts))
    mk = lambda: ScsTreeLikelihood(counts, tree.copy(), size_factors=sf, asc_mode="felsenstein_mean", n_background_sites=5e3,
                                   private_seq_cov=True)
    A, B = mk(), mk()
    assert A.retraverse_changed_only
    B.retraverse_changed_only = False
    assert A.calculate_logp() == B.calculate_logp()
    rng = np.random.default_rng(3)
    touched = []

    def same_state():
        assert A.logP == B.logP
        for x, y in zip(A.pattern_log_likelihoods(), B.pattern_log_likelihoods()):
            np.testing.assert_array_equal(x, y)
        ta, tb = A.core.get_private_seq_cov(), B.core.get_private_seq_cov()
        np.testing.assert_array_equal(ta[0], tb[0])
        np.testing.assert_array_equal(ta[1], tb[1])

    for it in range(6):
        ma, mb = A.accept(), B.accept()
        assert ma == mb
        if ma:
            touched.append(A.last_retraversed_sites)
            assert 0 < touched[-1] <= S
        same_state()
        # node read-backs: a leaf, a small (transient) node, a large one, the root -- the variant and the constant-site partials
        for node in (0, n - 1, tree.n, tree.n + 3, tree.root - 1, tree.root):
            np.testing.assert_array_equal(A.core.get_node_partials(node), B.core.get_node_partials(node))
            if node >= tree.n:
                np.testing.assert_array_equal(A.core.get_node_partials(node, const=True), B.core.get_node_partials(node, const=True))
        # a tree move in between: accepted on even rounds, rejected on odd ones
        t2 = A.tree.copy()
        node = int(rng.integers(tree.n, tree.node_count - 1))
        lo = max(t2.height[t2.left[node]], t2.height[t2.right[node]])
        t2.height[node] = lo + (t2.height[t2.parent[node]] - lo) * rng.random()
        for tl in (A, B):
            tl.store()
            tl.set_tree(t2.copy(), [node])
            tl.calculate_logp()
        same_state()
        if it % 2 == 1:
            for tl in (A, B):
                tl.restore()
                tl.calculate_logp()
            same_state()
    assert len(touched) >= 3 and min(touched[1:]) < S      # later accepts move a part of the patterns only
    assert touched[2] == S                                  # ... except right after a rejected move (stale per-site outputs)
    # and the state after one more accept (the loop ended on a rejected move: per-site outputs are the proposal's until then)
    # is the oracle's evaluation under the final pairs
    assert A.accept() == B.accept()
    same_state()
    from oracle import private_allelic as pa
    t_fin, v_fin = A.core.get_private_seq_cov()
    P = oracle.leaf_params(*A.params.as_tuple(), single_ado=True)
    mats = substmodel.node_matrices(A.tree, A.rates)
    sl, lc, _ = pa.full_eval_private(oracle, oracle.counts5(counts), sf, P, t_fin, v_fin, A.tree.ops(), mats, A.proportions, A.tree.root)
    got_l, got_c = A.pattern_log_likelihoods()
    _close(got_l, sl, 1e-12)
    _close(got_c, lc, 1e-12)
    # misuse: leaf parameters set without a leaf update -- the listed sites would be evaluated under other parameters
    pr = list(A.params.as_tuple())
    pr[3] *= 1.01
    A.core.set_leaf_params(*pr)
    with pytest.raises(RuntimeError):
        A.core.private_retraverse_changed(A.tree.ops())
    A.close()
    B.close()


def test_private_core_rejects_unsupported_modes(core_mod):
    for kw in (dict(sparse=True), dict(ephemeral=True), dict(single_leaf_buffer=True)):
        with pytest.raises(RuntimeError):
            core_mod.ScsCudaLikelihoodCore(6, 64, 4, 4, True, True, True, private_seq_cov=True, **kw)
    c = core_mod.ScsCudaLikelihoodCore(6, 64, 4, 4, True, True, True, private_seq_cov=True)
    with pytest.raises(RuntimeError):
        c.get_private_seq_cov()
    with pytest.raises(RuntimeError):
        c.set_private_seq_cov(np.full((4, 64), -1.0), np.ones((4, 64)))
    c.close()
    s = core_mod.ScsCudaLikelihoodCore(6, 64, 4, 4, True, True, True)
    with pytest.raises(RuntimeError):
        s.set_private_seq_cov(np.ones((4, 64)), np.ones((4, 64)))
    s.close()
