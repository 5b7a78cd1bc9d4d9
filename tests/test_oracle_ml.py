"""Pins the max-sum / ML-genotype oracle (oracle/ml_genotypes.py).

* `update_max_likelihood_genotypes` against the golden cases of the reference's own unit test
  (test/beast/evolution/likelihood/ScsTreeLikelihoodTest.java:153-292, tree :40-143) -- the only vectors the
  reference holds for this path;
* the max-sum partial and the enumerated combinations against a brute force over all genotype assignments."""
import itertools

import numpy as np
import pytest

from oracle import ml_genotypes as mlo
from sieve_b200 import substmodel
from sieve_b200.data import synthetic

# ScsTreeLikelihoodTest.makeTree(): node -> children (left, right); 17 is the unary root, target node = 9
REF_TREE = {17: [16], 16: [15, 8], 15: [9, 6], 9: [10, 12], 10: [1, 13], 13: [3, 7], 12: [5, 14], 14: [4, 11],
            11: [2, 0]}

G1 = [1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 1, 1, 1, 1, 1, 0, 0, 0]
G2 = [0, 1, 0, 1, 0, 1, 1, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0]
E3 = [0, 1, 0, 1, 0, 1, 1, 1, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0]
E4 = [1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0]
E5 = [1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 1, 1, 1, 1, 1, 1, 0, 0]
E6 = [0, 1, 0, 1, 0, 1, 1, 1, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0]
E7 = [0, 1, 0, 1, 0, 1, 1, 1, 0, 0, 1, 0, 0, 1, 0, 1, 0, 0]
E8 = [1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 1, 1, 1, 1, 1, 0, 0]


def _run(start, genotype):
    return mlo.update_max_likelihood_genotypes([list(g) for g in start], REF_TREE, 9, genotype, 0, 18)


def test_update_ml_genotypes_reference_case_1():
    out = _run([G1, G2], 0)  # ScsTreeLikelihoodTest.java:153-178
    assert len(out) == 4
    for e in (G1, G2, E3, E4):
        assert e in out


def test_update_ml_genotypes_reference_case_2():
    out = _run([G1, G2, E5, E6], 0)  # :180-216
    assert len(out) == 8
    for e in (G1, G2, E3, E4, E5, E6, E7, E8):
        assert e in out


def test_update_ml_genotypes_reference_case_3():
    out = _run([G1, E6], 0)  # :218-250
    assert len(out) == 8
    for e in (G1, G2, E3, E4, E5, E6, E7, E8):
        assert e in out


def test_update_ml_genotypes_reference_case_4():
    g3 = [1, 1, 1, 1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0]
    g4 = [1, 1, 1, 0, 1, 1, 1, 2, 0, 1, 1, 1, 1, 0, 1, 0, 0, 0]
    out = _run([G1, G2, g3, g4], [0, 1])  # :252-292, genotypes {0, 1}
    assert len(out) == 6
    for e in (G1, G2, E3, E4, g3, g4):
        assert e in out


def _small_case(n, S, seed, single_ado=True):
    d = synthetic(n, S, seed=seed)
    tree, counts4 = d["tree"], d["counts"]
    from oracle import sieve_oracle as so
    c5 = so.counts5(counts4)
    sf, _ = so.size_factors(c5)
    params = so.leaf_params(single_ado=single_ado)
    rates, _ = substmodel.gamma_category_rates(1.0, 4)
    mats = substmodel.node_matrices(tree, rates)
    ops = tree.ops()
    return tree, c5, sf, params, mats, ops


@pytest.mark.parametrize("single_ado", [True, False])
def test_max_sum_equals_brute_force(single_ado):
    n, S = 4, 6
    tree, c5, sf, params, mats, ops = _small_case(n, S, 7, single_ado)
    tr = mlo.ml_trace(c5, sf, params, ops, mats, 0)
    internal = [int(p) for p, _, _ in ops]
    root = internal[-1]
    logm = np.log(np.where(mats > 0, mats, 4.9e-324)).reshape(2 * n, 4, 4, 4)
    for k in range(4):
        for s in range(S):
            best, arg = -np.inf, []
            # the root genotype is 0; every other node (leaves too) is free
            free = [x for x in range(2 * n) if x != root]
            for assign in itertools.product(range(4), repeat=len(free)):
                g = dict(zip(free, assign))
                g[root] = 0
                v = 0.0
                for p, c1, c2 in ops:
                    for ch in (int(c1), int(c2)):
                        if ch >= 0:
                            v += logm[ch, k, g[int(p)], g[ch]]
                for leaf in range(n):
                    v += tr["leaf"][leaf, s, g[leaf]]
                if v > best + 1e-9 * abs(best if np.isfinite(best) else 1.0):
                    best, arg = v, [g]
                elif abs(v - best) <= 1e-9 * abs(best):
                    arg.append(g)
            got = tr["ml"][root][k, s, 0]
            assert abs(got - best) <= 1e-12 * abs(best)
            back_at = {node: tr["back"][node][k, s] for node in tr["back"]}
            combos = mlo.enumerate_ml_genotypes(ops, n, back_at, tr["ado"][:, s, :], 0)
            geno_sets = sorted({tuple(c[n + x] for x in range(2 * n)) for c in combos})
            assert geno_sets == sorted(tuple(g[x] for x in range(2 * n)) for g in arg)


def test_leaf_components_and_ado_masks():
    from oracle import sieve_oracle as so
    tree, c5, sf, params, mats, ops = _small_case(5, 40, 3, single_ado=False)
    for j in range(5):
        plain = so.leaf_likelihood(c5, j, sf[j], params)
        ml, ado = mlo.leaf_likelihood_ml(c5, j, sf[j], params)
        assert np.array_equal(plain, ml)  # the mixed likelihood itself is the same number
        assert ((ado >= 1) & (ado <= 7)).all()
    # strongly covered hom-ref site: no dropout (component 0) wins for 0/0
    c = np.zeros((1, 1, 5), dtype=np.int32)
    c[0, 0] = [0, 0, 0, 100, 100]
    _, ado = mlo.leaf_likelihood_ml(c, 0, 1.0, so.leaf_params(single_ado=True))
    assert ado[0, 0] == 1


def test_enumeration_is_consistent_with_update_function():
    """The reference re-expands its list with updateMaxLikelihoodGenotypes before using it; on a list that already is
    the full product the function must be the identity (as a set)."""
    n, S = 5, 8
    tree, c5, sf, params, mats, ops = _small_case(n, S, 11)
    tr = mlo.ml_trace(c5, sf, params, ops, mats, 0)
    children = {int(p): [int(c) for c in (c1, c2) if c >= 0] for p, c1, c2 in ops}
    for s in range(S):
        back_at = {node: tr["back"][node][0, s] for node in tr["back"]}
        # force ties so that the list has several entries
        for node in back_at:
            back_at[node] = np.array([b | 0x33 if b >> 4 else b | 0x03 for b in back_at[node]], dtype=np.uint8)
        combos = [list(c) for c in mlo.enumerate_ml_genotypes(ops, n, back_at, tr["ado"][:, s, :], 0)]
        assert len(combos) > 1
        before = sorted(map(tuple, combos))
        for node in children:
            if len(children[node]) > 1:
                gs = sorted({c[n + node] for c in combos})
                mlo.update_max_likelihood_genotypes(combos, children, node, gs, n, 2 * n)
        assert sorted(map(tuple, combos)) == before


def test_host_enumerator_of_the_product_matches_the_oracle():
    """sieve_b200/variants.py (product code, host half of the tie-break) against the oracle's enumeration, on
    back-pointers with forced ties -- no device involved."""
    from sieve_b200 import variants
    n, S = 6, 10
    tree, c5, sf, params, mats, ops = _small_case(n, S, 13)
    tr = mlo.ml_trace(c5, sf, params, ops, mats, 0)
    children = {int(p): (int(c1), int(c2)) for p, c1, c2 in ops}
    for s in range(S):
        back = np.zeros((2 * n, 4), dtype=np.uint8)
        for node in range(n, 2 * n):
            b = tr["back"][node][1, s]
            back[node] = [x | (0x21 if x >> 4 else 0x01) for x in b]
        back[:n] = tr["ado"][:, s, :] | 1
        back_at = {node: back[node] for node in range(n, 2 * n)}
        want = mlo.enumerate_ml_genotypes(ops, n, back_at, back[:n], 0)
        got = variants.enumerate_combinations(children, tree.root, n, back, 0)
        assert got == want and len(got) >= 1


def test_linear_space_twin_of_the_max_sum_pruning():
    """ml_lin_pruning (ScsBeerLikelihoodCore.java:1759-2335) against the log-space twin where the two must agree (no
    underflow, no zero entry of P), and on what is specific to it: a LEAF child under a zero-length branch contributes
    log(0 * leaf) = -inf for the off-diagonal genotypes (no Double.MIN_VALUE clamp in :1948-1949), an INTERNAL child keeps
    the clamp (:2260-2263)."""
    import numpy as np
    from oracle import ml_genotypes as mlo
    rng = np.random.default_rng(5)
    K, S = 2, 9
    leafA = rng.uniform(1e-30, 1e-3, (K, S, 4))
    leafB = rng.uniform(1e-30, 1e-3, (K, S, 4))
    ml_int = rng.uniform(-300.0, -5.0, (K, S, 4))
    P = rng.dirichlet(np.ones(4), size=(K, 4)).reshape(K, 16)
    Q = rng.dirichlet(np.ones(4), size=(K, 4)).reshape(K, 16)
    # leaf-leaf, leaf-partial, partial-partial against the log twin fed log leaves
    for (a, la), (b, lb) in (((leafA, True), (leafB, True)), ((leafA, True), (ml_int, False)), ((ml_int, False), (ml_int[::-1], False))):
        got, gb = mlo.ml_lin_pruning(a, la, P, b, lb, Q)
        want, wb = mlo.ml_log_pruning(np.log(a) if la else a, P, np.log(b) if lb else b, Q)
        np.testing.assert_allclose(got, want, rtol=1e-13)
        assert (gb != wb).mean() < 0.02   # rounding-level near-ties only
    # zero-length branch above a leaf: identity matrix, unclamped
    I = np.tile(np.eye(4).reshape(16), (K, 1))
    got, gb = mlo.ml_lin_pruning(leafA, True, I, None, False, None)
    np.testing.assert_allclose(got, np.log(leafA), rtol=1e-15)
    assert np.array_equal(gb, np.broadcast_to(np.array([1, 2, 4, 8], dtype=np.uint8), gb.shape))
    # ... above an internal node: the clamp keeps every genotype a (hopeless) candidate, the diagonal wins
    got, gb = mlo.ml_lin_pruning(ml_int, False, I, None, False, None)
    np.testing.assert_allclose(got, ml_int, rtol=1e-15)
    # a leaf whose likelihoods all underflowed: every candidate is -inf, every genotype is listed (:1964-1978 compare -inf == -inf)
    got, gb = mlo.ml_lin_pruning(np.zeros((K, S, 4)), True, P, None, False, None)
    assert np.all(np.isneginf(got)) and np.all(gb == 15)
