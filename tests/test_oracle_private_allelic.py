"""Pins oracle/private_allelic.py (the per-pattern coverage model) by independent
derivations -- the reference holds no vector for this class and no JVM is available: sample-moment identities, a
vectorised re-implementation written from the formulas of the method's javadoc rather than from its loops, and scipy's
negative binomial.  CPU only; there is no CUDA counterpart of this row yet."""
import math

import numpy as np
import pytest
from scipy import stats

from oracle import private_allelic as pa
from oracle import sieve_oracle as so


def test_nb_term_is_a_negative_binomial_with_the_models_moments():
    rng = np.random.default_rng(3)
    for _ in range(50):
        al = int(rng.integers(0, 3))
        s, t, v = float(np.exp(0.3 * rng.standard_normal())), float(rng.uniform(5, 80)), float(rng.uniform(1, 120))
        cov = int(rng.integers(0, 400))
        mean, var = pa.seq_cov_mean_var(al, s, t, v)
        assert var > mean > 0
        got = pa.seq_cov_log_likelihood_per_allele(cov, al, s, t, v, so.nb_log_density)
        r, p = mean * mean / (var - mean), mean / var
        want = stats.nbinom.logpmf(cov, r, p)
        assert abs(got - want) <= 1e-10 * max(1.0, abs(want))
        # the moments of that distribution are the model's
        assert abs(stats.nbinom.mean(r, p) - mean) <= 1e-9 * mean
        assert abs(stats.nbinom.var(r, p) - var) <= 1e-9 * var


def test_deterministic_estimate_is_the_method_of_moments():
    rng = np.random.default_rng(5)
    n = 40
    sf = np.exp(0.3 * rng.standard_normal(n))
    cov = rng.poisson(90 * sf).astype(np.int64)
    t, v = pa.deterministic_allelic_seq_cov_and_var(cov, sf, 2)
    x = cov / sf
    q = x.mean()
    z = x.var(ddof=1) - q * (1 / sf).mean()
    assert abs(t - q / (2 + 1e-6)) <= 1e-13 * t
    if z > 0:
        assert abs(v - z / (2 + 1e-6) ** 2) <= 1e-10 * abs(v)
    # a sample without over-dispersion keeps the previous variance
    t2, v2 = pa.deterministic_allelic_seq_cov_and_var(np.full(n, 7), np.ones(n), 2, prev_v=3.25)
    assert v2 == 3.25 and abs(t2 - 7 / (2 + 1e-6)) < 1e-13
    # a single cell has no variance estimate at all
    assert pa.deterministic_allelic_seq_cov_and_var([5], [1.0], 1, prev_v=1.5) == (5 / (1 + 1e-6), 1.5)


def _vectorised(cov, sf, alleles, combos, weights, prev_v):
    """The estimator written from its definition: per combination, group the cells by their number of sequenced
    alleles; t = mean over the non-empty groups of (group mean of cov / s) / alleles; v = mean over the groups with a
    positive moment estimate of (sample variance - group mean * mean(1 / s)) / alleles^2; then weighted means over the
    combinations (for v only over those that have such a group)."""
    x = np.asarray(cov) / np.asarray(sf)
    ts, vs, wt, wv = [], [], [], []
    for comb, w in zip(combos, weights):
        comb = np.asarray(comb)
        tq, vq = [], []
        for al in alleles:
            m = comb == al
            if m.sum() == 0:
                continue
            q = x[m].mean()
            tq.append(q / (al + 1e-6))
            if m.sum() > 1:
                z = x[m].var(ddof=1) - q * (1 / np.asarray(sf)[m]).mean()
                if z > 0:
                    vq.append(z / (al + 1e-6) ** 2)
        ts.append(sum(tq) / len(tq))
        wt.append(w)
        if vq:
            vs.append(sum(vq) / len(vq))
            wv.append(w)
    t = float(np.dot(ts, wt) / sum(wt))
    v = float(np.dot(vs, wv) / sum(wv)) if vs and np.dot(vs, wv) > 0 else prev_v
    return t, v


@pytest.mark.parametrize("alleles", [[1, 2], [0, 1, 2]])
def test_update_matches_the_definition(alleles):
    rng = np.random.default_rng(11 + len(alleles))
    n = 30
    sf = np.exp(0.3 * rng.standard_normal(n))
    for trial in range(40):
        combos = [rng.choice(alleles, size=n) for _ in range(int(rng.integers(1, 5)))]
        weights = [int(w) for w in rng.integers(1, 6, size=len(combos))]
        base = np.where(combos[0] == 2, 100.0, np.where(combos[0] == 1, 50.0, 0.2))
        cov = rng.negative_binomial(8, 8 / (8 + base * sf)).astype(np.int64)
        got = pa.update_allelic_seq_cov_and_raw_var(cov, sf, alleles, combos, weights, prev_v=17.0)
        want = _vectorised(cov, sf, alleles, combos, weights, prev_v=17.0)
        assert abs(got[0] - want[0]) <= 1e-12 * abs(want[0])
        assert abs(got[1] - want[1]) <= 1e-10 * abs(want[1])


def test_update_quirks():
    sf = np.array([1.0, 1.0, 1.0, 1.0])
    # one combination, two groups: alleles 2 -> cells {0, 1} (cov 10, 30), alleles 1 -> cells {2, 3} (cov 4, 4: z <= 0)
    t, v = pa.update_allelic_seq_cov_and_raw_var([10, 30, 4, 4], sf, [1, 2], [np.array([2, 2, 1, 1])], [3], prev_v=9.0)
    q2, q1 = 20.0, 4.0
    assert abs(t - 0.5 * (q1 / (1 + 1e-6) + q2 / (2 + 1e-6))) < 1e-13          # divided by the two non-empty groups
    z2 = ((10 - 20) ** 2 + (30 - 20) ** 2) / 1 - 20 * 1.0
    assert abs(v - z2 / (2 + 1e-6) ** 2) < 1e-12                               # divided by the ONE group with z > 0
    # no group with z > 0 anywhere: the variance keeps its previous value
    t, v = pa.update_allelic_seq_cov_and_raw_var([4, 4, 4, 4], sf, [1, 2], [np.array([2, 2, 1, 1])], [1], prev_v=9.0)
    assert v == 9.0
    # weights: the same combination twice with weights (1, 3) equals once with any weight
    a = pa.update_allelic_seq_cov_and_raw_var([10, 30, 4, 9], sf, [1, 2], [np.array([2, 2, 1, 1])] * 2, [1, 3], 0.0)
    b = pa.update_allelic_seq_cov_and_raw_var([10, 30, 4, 9], sf, [1, 2], [np.array([2, 2, 1, 1])], [7], 0.0)
    assert a[0] == pytest.approx(b[0], rel=1e-15) and a[1] == pytest.approx(b[1], rel=1e-15)


def test_leaf_terms_depend_on_category_and_pattern():
    rng = np.random.default_rng(2)
    cov = rng.integers(0, 150, size=(3, 5))
    sf = np.exp(0.2 * rng.standard_normal(5))
    t = np.array([[40.0, 50.0, 60.0], [45.0, 55.0, 65.0]])
    v = np.array([[30.0, 60.0, 90.0], [35.0, 65.0, 95.0]])
    out = pa.leaf_seq_cov_log_likelihoods(cov, sf, [1, 2], t, v, so.nb_log_density)
    assert out.shape == (2, 3, 5, 2) and np.isfinite(out).all()
    k, p, j, a = 1, 2, 3, 1
    mean, var = pa.seq_cov_mean_var(2, sf[j], t[k, p], v[k, p])
    assert abs(out[k, p, j, a] - stats.nbinom.logpmf(cov[p, j], mean * mean / (var - mean), mean / var)) < 1e-9
    assert not np.allclose(out[0], out[1])
    assert pa.processed_seq_cov([0, 3], 1).tolist() == [1, 4] and pa.processed_seq_cov([0, 3], 0).tolist() == [0, 3]
    assert math.isclose(pa.NUMERICAL_STABILIZER, 1e-6)


def test_private_leaf_partials_reduce_to_the_shared_model():
    """With every (matrix, pattern) pair at the shared values the private model's leaf partials are K copies of the
    shared model's; a pair that differs changes exactly its own (matrix, pattern) rows."""
    from sieve_b200.data import synthetic
    d = synthetic(6, 40, seed=4)
    c5 = so.counts5(d["counts"])
    sf, _ = so.size_factors(c5)
    P = so.leaf_params()
    K, S = 3, 40
    t = np.full((K, S), P.allelic_cov)
    v = np.full((K, S), P.allelic_raw_var)
    shared = so.leaf_likelihood(c5, 2, sf[2], P)
    got = pa.leaf_likelihood_private(so, c5, 2, sf[2], P, t, v)
    for k in range(K):
        np.testing.assert_array_equal(got[k], shared)
    t[1, 7] *= 1.3
    v[1, 7] *= 0.6
    got2 = pa.leaf_likelihood_private(so, c5, 2, sf[2], P, t, v)
    changed = np.argwhere((got2 != got).any(axis=2))
    assert changed.tolist() in ([[1, 7]], [])      # (an empty site has no coverage term to change)
    if d["counts"][7, 2, 3] > 0:
        assert changed.tolist() == [[1, 7]]
