"""Pins the CPU oracle by derivations that are independent of its own code paths:
(i) brute-force enumeration of internal genotypes, (ii) log-mode == linear-mode, (iii) mpmath 40-digit densities,
(iv) thread-split invariance incl. both acquisition-bias code paths.  The reference itself holds no numeric vectors."""
import itertools

import mpmath
import numpy as np
import pytest

from sieve_b200 import substmodel
from sieve_b200.tree import caterpillar, random_coalescent, fixture_9_leaves as fixture9

mpmath.mp.dps = 40


def _mp_lgamma(x):
    return mpmath.loggamma(mpmath.mpf(x))


def test_log_gamma_matches_mpmath(oracle):
    xs = np.concatenate([np.linspace(0.01, 3, 50), np.linspace(3, 2000, 200), [1e-3, 0.4, 149.6, 150.0, 1e4, 1e5]])
    for x in xs:
        ref = float(_mp_lgamma(x))
        got = oracle.log_gamma(x)
        assert abs(got - ref) <= 5e-15 * max(1.0, abs(ref)) + 2e-15, (x, got, ref)
    assert np.isnan(oracle.log_gamma(0.0)) and np.isnan(oracle.log_gamma(-1.0))


def _mp_dm(counts5, alpha5):
    # DirichletMultinomial pmf written from its textbook form:
    # cov * B(a0, cov) / prod_{c_i>0} c_i * B(a_i, c_i)
    def lp(c, a):
        if c <= 0:
            return mpmath.mpf(0)
        return mpmath.log(c) + mpmath.loggamma(a) + mpmath.loggamma(c) - mpmath.loggamma(mpmath.mpf(a) + c)
    return lp(counts5[4], alpha5[4]) - sum(lp(counts5[i], alpha5[i]) for i in range(4))


def test_dm_density_matches_mpmath_and_normalises(oracle):
    eps, w1, w2 = 0.008, 150.0, 2.0
    f = eps / 3
    alphas = [[f * w1, f * w1, f * w1, (1 - eps) * w1, w1], [(0.5 - f) * w2, f * w2, f * w2, (0.5 - f) * w2, w2]]
    rng = np.random.default_rng(1)
    for a in alphas:
        for _ in range(40):
            cov = int(rng.integers(1, 300))
            alt = rng.multinomial(cov, [0.2, 0.05, 0.05, 0.7])
            c5 = [alt[0], alt[1], alt[2], alt[3], cov]
            ref = float(_mp_dm(c5, a))
            got = oracle.dm_log_density(c5, a)
            assert abs(got - ref) <= 2e-12 * max(1.0, abs(ref)), (c5, got, ref)
    # a pmf: sums to one over all compositions of a small total
    a = alphas[1]
    tot = 0.0
    cov = 6
    for c in itertools.product(range(cov + 1), repeat=3):
        if sum(c) <= cov:
            tot += np.exp(oracle.dm_log_density([c[0], c[1], c[2], cov - sum(c), cov], a))
    assert abs(tot - 1.0) < 1e-12
    # the reference's quirk: cov == 0 -> 1.0 (math/distributions/DirichletMultinomial.java:29-30)
    assert oracle.dm_log_density([0, 0, 0, 0, 0], a) == 1.0


def test_nb_density_matches_mpmath_and_normalises(oracle):
    rng = np.random.default_rng(2)
    for _ in range(60):
        r = float(rng.uniform(0.5, 200))
        p = float(rng.uniform(0.01, 0.99))
        x = int(rng.integers(0, 500))
        ref = float(mpmath.loggamma(r + x) - mpmath.loggamma(x + 1) - mpmath.loggamma(r)
                    + r * mpmath.log(p) + x * mpmath.log(1 - mpmath.mpf(p)))
        got = oracle.nb_log_density(x, r, p)
        assert abs(got - ref) <= 5e-12 * max(1.0, abs(ref)), (x, r, p, got, ref)
    tot = sum(np.exp(oracle.nb_log_density(x, 50.0, 0.33)) for x in range(2000))
    assert abs(tot - 1.0) < 1e-10


def test_log_sum_exp(oracle):
    a = np.array([-1000.0, -1001.0, -1002.5])
    ref = float(mpmath.log(sum(mpmath.exp(mpmath.mpf(x)) for x in a)))
    assert abs(oracle.log_sum_exp(a) - ref) < 1e-12
    assert np.isnan(oracle.log_sum_exp([-np.inf, -np.inf]))   # MathFunctions.java:105-111


def test_size_factors_median_of_ratios(oracle):
    rng = np.random.default_rng(3)
    S, n = 37, 6
    cov = rng.integers(0, 60, size=(S, n))
    cov[rng.random((S, n)) < 0.1] = 0
    c4 = np.zeros((S, n, 4), dtype=np.int32)
    c4[..., 3] = cov
    sf, st = oracle.size_factors(oracle.counts5(c4))
    # independent numpy restatement of DESeq's median-of-ratios with zeros skipped
    ref = np.empty(n)
    geo = np.array([np.exp(np.log(row[row > 0]).mean()) if (row > 0).any() else 0.0 for row in cov])
    for j in range(n):
        ok = (cov[:, j] > 0) & (geo > 0)
        ref[j] = np.median(cov[ok, j] / geo[ok])
    np.testing.assert_allclose(sf, ref, rtol=1e-14)
    assert abs(st[0] - cov.mean() / 2) < 1e-12 and abs(st[1] - cov.var(ddof=1) / 4) < 1e-9


def _random_leaf_logs(rng, n, S):
    return rng.uniform(-30, -1, size=(n, S, 4))


def _brute_force(tree, mats, leaf_lin, props, const=False):
    """Sum over all genotype assignments of the internal nodes (root fixed to 0/0). mats [node][K][4][4]."""
    n, S, K = tree.n, leaf_lin.shape[1], mats.shape[1]
    internal = [i for i in range(n, tree.node_count) if i != tree.root]
    out = np.zeros(S)
    for k in range(K):
        acc = np.zeros(S)
        for assign in itertools.product(range(4), repeat=len(internal)):
            g = {tree.root: 0}
            g.update(dict(zip(internal, assign)))
            w = 1.0
            for node in internal:
                w *= mats[node, k][g[int(tree.parent[node])], g[node]]
            if w == 0.0:
                continue
            term = np.full(S, w)
            for leaf in range(n):
                pg = g[int(tree.parent[leaf])]
                if const:
                    term = term * (mats[leaf, k][pg, 0] * leaf_lin[leaf, :, 0])
                else:
                    term = term * (mats[leaf, k][pg, :] @ leaf_lin[leaf].T)
            acc += term
        out += props[k] * acc
    return out


def _run_core(oracle, tree, mats, leaf, use_log, props):
    n, S, K = tree.n, leaf.shape[1], mats.shape[1]
    core = oracle.Core(tree.node_count, n, S, K, 4, use_log)
    for i in range(n):
        core.set_node_partials(i, leaf[i] if use_log else np.exp(leaf[i]))
    for node in range(tree.node_count):
        if node != tree.root:
            for k in range(K):
                core.set_node_matrix(node, k, mats[node, k])
    for p, c1, c2 in tree.ops():
        core.calculate_partials(c1, c1 < n, c2, 0 <= c2 < n, p, 0)
    rp, cr = core.integrate_partials(tree.root, props, 0)
    return core.calculate_log_likelihoods(rp, cr) + (core,)


@pytest.mark.parametrize("n", [2, 3, 5, 6])
def test_pruning_matches_brute_force_enumeration(oracle, n):
    rng = np.random.default_rng(10 + n)
    tree = random_coalescent(n, rng)
    tree.height *= 0.4 / tree.height[tree.root]
    rates, props = substmodel.gamma_category_rates(0.7, 4)
    mats = substmodel.node_matrices(tree, rates)
    S = 7
    leaf = _random_leaf_logs(rng, n, S)
    ll, lc, _ = _run_core(oracle, tree, mats, leaf, True, props)
    m4 = mats.reshape(tree.node_count, 4, 4, 4)
    bf = np.log(_brute_force(tree, m4, np.exp(leaf), props))
    bfc = np.log(_brute_force(tree, m4, np.exp(leaf), props, const=True))
    np.testing.assert_allclose(ll, bf, rtol=1e-12)
    np.testing.assert_allclose(lc, bfc, rtol=1e-12)


def test_log_mode_equals_linear_mode(oracle):
    rng = np.random.default_rng(5)
    tree = fixture9()
    rates, props = substmodel.gamma_category_rates(1.0, 4)
    mats = substmodel.node_matrices(tree, rates)
    leaf = _random_leaf_logs(rng, tree.n, 33)
    ll_log, lc_log, _ = _run_core(oracle, tree, mats, leaf, True, props)
    ll_lin, lc_lin, _ = _run_core(oracle, tree, mats, leaf, False, props)
    np.testing.assert_allclose(ll_log, ll_lin, rtol=1e-13)
    np.testing.assert_allclose(lc_log, lc_lin, rtol=1e-13)


def test_leaf_log_mode_equals_linear_mode_and_mixture_by_hand(oracle):
    rng = np.random.default_rng(6)
    S, n = 50, 4
    c4 = np.zeros((S, n, 4), dtype=np.int32)
    cov = rng.integers(0, 120, size=(S, n))
    alt = (cov * rng.random((S, n)) * (rng.random((S, n)) < 0.5)).astype(int)
    c4[..., 0] = alt
    c4[..., 1] = ((cov - alt) * 0.05 * rng.random((S, n))).astype(int)
    c4[..., 3] = cov
    c5 = oracle.counts5(c4)
    sf, _ = oracle.size_factors(c5)
    for single in (True, False):
        pl = oracle.leaf_params(single_ado=single, use_log=True)
        pn = oracle.leaf_params(single_ado=single, use_log=False)
        for j in range(n):
            a = oracle.leaf_likelihood(c5, j, sf[j], pl)
            b = oracle.leaf_likelihood(c5, j, sf[j], pn)
            np.testing.assert_allclose(a, np.log(b), rtol=1e-11, atol=1e-11)
    # single-ADO mixture re-derived from the densities
    P = oracle.leaf_params()
    th, t, v, eps, w1, w2 = 0.3, 50.0, 50.0, 0.008, 150.0, 2.0
    f = eps / 3
    j = 1
    got = oracle.leaf_likelihood(c5, j, sf[j], P)
    for s in range(S):
        r5 = c5[s, j]
        N0 = oracle.dm_log_density(r5, [f * w1, f * w1, f * w1, (1 - eps) * w1, w1])
        N1 = oracle.dm_log_density(r5, [(1 - eps) * w1, f * w1, f * w1, f * w1, w1])
        N2 = oracle.dm_log_density(r5, [(.5 - f) * w2, (.5 - f) * w2, f * w2, f * w2, w2])
        N3 = oracle.dm_log_density(r5, [(.5 - f) * w2, f * w2, f * w2, (.5 - f) * w2, w2])
        C = []
        for k in (1, 2):
            mean = (k + 1e-6) * t * sf[j]
            var = mean + sf[j] ** 2 * (k + 1e-6) ** 2 * v
            C.append(oracle.nb_log_density(r5[4], mean ** 2 / (var - mean), mean / var))
        lse = np.logaddexp
        ref = [lse(np.log(1 - th) + N0 + C[1], np.log(th) + N0 + C[0]),
               lse(np.log(1 - th) + N3 + C[1], np.log(th / 2) + lse(N0, N1) + C[0]),
               lse(np.log(1 - th) + N1 + C[1], np.log(th) + N1 + C[0]),
               lse(np.log(1 - th) + N2 + C[1], np.log(th) + N1 + C[0])]
        np.testing.assert_allclose(got[s], ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("asc", ["none", "felsenstein_mean", "lewis"])
def test_thread_split_invariance(oracle, asc):
    from sieve_b200.data import synthetic
    d = synthetic(7, 41, seed=11)
    tree = d["tree"]
    c5 = oracle.counts5(d["counts"])
    sf, _ = oracle.size_factors(c5)
    rates, props = substmodel.gamma_category_rates(1.0, 4)
    mats = substmodel.node_matrices(tree, rates)
    P = oracle.leaf_params()
    M = 9727.0
    one, sl1, sc1 = oracle.full_eval(c5, sf, P, tree.ops(), mats, props, tree.root, asc, M, 1)
    for T in (2, 3, 5):
        many, slT, scT = oracle.full_eval(c5, sf, P, tree.ops(), mats, props, tree.root, asc, M, T)
        np.testing.assert_allclose(slT, sl1, rtol=0, atol=0)
        assert abs(many - one) <= 1e-11 * abs(one)


def test_pattern_points(oracle):
    # likelihood/ThreadedScsTreeLikelihood.java:227-238
    assert list(oracle.pattern_points(10, 3)) == [0, 4, 7, 10]
    assert list(oracle.pattern_points(273, 8))[-1] == 273


def test_store_restore_is_an_index_swap(oracle):
    rng = np.random.default_rng(7)
    tree = caterpillar(4)
    rates, props = substmodel.gamma_category_rates(1.0, 4)
    mats = substmodel.node_matrices(tree, rates)
    leaf = _random_leaf_logs(rng, 4, 5)
    ll0, _, core = _run_core(oracle, tree, mats, leaf, True, props)
    core.store()
    # propose: new matrix on node 1, recompute its ancestors into the flipped buffers
    core.set_node_matrix_for_update(1)
    for k in range(4):
        core.set_node_matrix(1, k, substmodel.transition_probabilities(0.9 * rates[k]).reshape(16))
    for p, c1, c2 in tree.ops(tree.dirty_closure([1])):
        core.set_node_partials_for_update(p)
        core.calculate_partials(c1, c1 < 4, c2, 0 <= c2 < 4, p, 0)
    ll1, _ = core.calculate_log_likelihoods(*core.integrate_partials(tree.root, props, 0))
    assert np.abs(ll1 - ll0).max() > 1e-6
    core.restore()
    ll2, _ = core.calculate_log_likelihoods(*core.integrate_partials(tree.root, props, 0))
    np.testing.assert_array_equal(ll2, ll0)
