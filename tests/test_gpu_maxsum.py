"""Parity of the device's max-sum pass (variant calling) with the oracle of oracle/ml_genotypes.py.

Floating point: ML partials within 1e-12 relative.  Arg-max lists (bit masks) must be identical wherever the oracle's
margin between the best and the runner-up is not a rounding-level near-tie: the reference decides ties by exact
equality of doubles, so a last-bit difference in a leaf likelihood may legitimately flip a tie between two candidates
that agree to 1e-13; such entries are counted and must stay a tiny fraction."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from sieve_b200 import substmodel, variants  # noqa: E402
from sieve_b200.data import synthetic  # noqa: E402

PARAMS = (0.3, 50.0, 50.0, 0.008, 150.0, 2.0)


@pytest.fixture(scope="module")
def core_mod():
    from sieve_b200 import _lib, core
    assert _lib.lib().sieve_device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return core


def _case(core_mod, oracle, n, S, seed, single=True, K=4, asc=True, use_log=True):
    from oracle import ml_genotypes as mlo
    d = synthetic(n, S, seed=seed)
    counts, tree = d["counts"], d["tree"]
    c5 = oracle.counts5(counts)
    sf, _ = oracle.size_factors(c5)
    rates, props = substmodel.gamma_category_rates(0.8, K)
    mats = substmodel.node_matrices(tree, rates)
    c = core_mod.ScsCudaLikelihoodCore(n, S, K, 4, use_log, asc, single)
    c.set_counts(counts)
    c.set_size_factors(sf)
    c.set_leaf_params(*PARAMS)
    c.update_leaves(for_update=False)
    c.set_root(tree.root, props, 0)
    nodes = [i for i in range(tree.node_count) if i != tree.root]
    c.set_matrices(nodes, mats[nodes][:, :K], for_update=False)
    c.update_nodes(tree.ops(), flip=False)
    c.evaluate()
    P = oracle.leaf_params(*PARAMS, single_ado=single, use_log=use_log)
    tr = mlo.ml_trace(c5, sf, P, tree.ops(), mats[:, :K], 0)
    return c, tree, tr


def _margin_ok(values, mask_a, mask_b, rtol=1e-11):
    """two arg-max sets over `values` differ only by candidates that are within rtol of the maximum"""
    diff = int(mask_a) ^ int(mask_b)
    mx = np.max(values)
    return all(abs(values[i] - mx) <= rtol * max(1.0, abs(mx)) for i in range(len(values)) if (diff >> i) & 1)


@pytest.mark.parametrize("single", [True, False])
def test_leaf_log_partials_and_dropout_argmax(core_mod, oracle, single):
    c, tree, tr = _case(core_mod, oracle, 12, 400, 21, single=single)
    c.ml_trace()
    flips = 0
    for leaf in range(tree.n):
        p, b, _ = c.ml_node(leaf)
        for k in range(c.K):
            np.testing.assert_allclose(p[k], tr["leaf"][leaf], rtol=1e-12, atol=1e-12)
            flips += int((b[k] != tr["ado"][leaf]).sum())
    assert flips == 0  # the dropout components of a leaf differ by whole units of log-likelihood, not by rounding
    c.close()


@pytest.mark.parametrize("single,K", [(True, 4), (False, 4), (True, 2)])
def test_max_sum_partials_and_back_pointers(core_mod, oracle, single, K):
    n, S = 17, 300
    c, tree, tr = _case(core_mod, oracle, n, S, 33, single=single, K=K)
    c.ml_trace()
    total = near = 0
    for node in range(n, 2 * n):
        p, b, _ = c.ml_node(node)
        ref, rb = tr["ml"][node], tr["back"][node]
        assert np.all(np.abs(p - ref) <= 1e-12 * np.abs(ref)), node
        total += b.size
        bad = np.argwhere(b != rb)
        for k, s, g in bad:
            near += 1
            # re-derive the candidates of this entry from the oracle's inputs and check that the flip is a near-tie
            left, right = int(tree.left[node]), int(tree.right[node])
            for side, child in ((0, left), (1, right)):
                if child < 0:
                    continue
                vals = tr["leaf"][child][s] if child < n else tr["ml"][child][k, s]
                m = np.asarray(substmodel.node_matrices(tree, substmodel.gamma_category_rates(0.8, K)[0])[child, k])
                cand = np.log(np.where(m > 0, m, 4.9e-324)).reshape(4, 4)[g] + vals
                assert _margin_ok(cand, (b[k, s, g] >> (4 * side)) & 15, (rb[k, s, g] >> (4 * side)) & 15)
    assert near <= 1e-3 * total, (near, total)
    # ML category per site (getPatternCategories)
    cats = c.ml_categories()
    mism = np.nonzero(cats != tr["categories"])[0]
    for s in mism:
        v = tr["partials"][tr["root"]][:, s, 0]
        assert _margin_ok(v, cats[s], tr["categories"][s])
    assert len(mism) <= 2
    c.close()


def test_collections_and_calls_match_enumeration(core_mod, oracle):
    from oracle import ml_genotypes as mlo
    n, S = 9, 250
    c, tree, tr = _case(core_mod, oracle, n, S, 44)
    calls = variants.call_variants(c, tree)
    ops = tree.ops()
    checked = 0
    for s in range(S):
        want = mlo.call_pattern(ops, n, tr, s, 0)
        if s in calls.tied:
            got = sorted(calls.tied[s])
            assert got == sorted(want)
        else:
            assert len(want) == 1, (s, want)
            k, comb = want[0]
            assert calls.category[s] == k
            assert list(calls.genotypes[:, s]) == list(comb[n:])
            assert list(calls.ado[:, s]) == list(comb[:n])
            checked += 1
    assert checked > 0.9 * S
    # the collection of a node = the set of genotypes it takes over all combinations of that (category, site)
    node = n + 3
    _, _, coll = c.ml_node(node)
    for s in range(0, S, 17):
        for k in range(4):
            back_at = {v: tr["back"][v][k, s] for v in tr["back"]}
            combos = mlo.enumerate_ml_genotypes(ops, n, back_at, tr["ado"][:, s, :], 0)
            want = 0
            for comb in combos:
                want |= 1 << comb[n + node]
            assert coll[k, s] == want
    c.close()


def test_trace_is_invalidated_by_changes_and_needs_resident_core(core_mod, oracle):
    c, tree, tr = _case(core_mod, oracle, 6, 64, 5)
    with pytest.raises(RuntimeError):
        c.ml_categories()  # no trace yet
    c.ml_trace()
    c.ml_categories()
    c.set_leaf_params(0.2, 50.0, 50.0, 0.008, 150.0, 2.0)
    c.update_leaves(for_update=True)
    with pytest.raises(RuntimeError):
        c.ml_call()
    c.update_nodes(tree.ops(), flip=True)
    c.evaluate()
    c.ml_trace()
    g1, a1, k1, m1 = c.ml_call()
    c.ml_trace()  # idempotent
    g2, a2, k2, m2 = c.ml_call()
    assert np.array_equal(g1, g2) and np.array_equal(a1, a2) and np.array_equal(k1, k2) and np.array_equal(m1, m2)
    c.close()
    e = core_mod.ScsCudaLikelihoodCore(6, 64, 4, 4, True, False, True, ephemeral=True)
    with pytest.raises(RuntimeError):
        e.ml_trace()
    e.close()


def test_called_genotypes_recover_simulated_truth(core_mod, oracle):
    """Sanity on the model level: with 50x coverage the ML leaf genotypes are mostly the simulated ones."""
    n, S = 20, 2000
    d = synthetic(n, S, seed=77)
    if "genotypes" not in d:
        pytest.skip("generator does not expose the simulated genotypes")
    counts, tree = d["counts"], d["tree"]
    sf, _ = oracle.size_factors(oracle.counts5(counts))
    rates, props = substmodel.gamma_category_rates(1.0, 4)
    mats = substmodel.node_matrices(tree, rates)
    c = core_mod.ScsCudaLikelihoodCore(n, S, 4, 4, True, False, True)
    c.set_counts(counts)
    c.set_size_factors(sf)
    c.set_leaf_params(*PARAMS)
    c.update_leaves(for_update=False)
    c.set_root(tree.root, props, 0)
    nodes = [i for i in range(tree.node_count) if i != tree.root]
    c.set_matrices(nodes, mats[nodes], for_update=False)
    c.update_nodes(tree.ops(), flip=False)
    c.evaluate()
    calls = variants.call_variants(c, tree)
    truth = np.asarray(d["genotypes"])[:n]  # [leaf][S]
    covered = counts[:, :, 3].T > 10
    agree = (calls.genotypes[:n] == truth)[covered].mean()
    assert agree > 0.9, agree
    c.close()


def test_sparse_core_calls_variants_in_site_chunks(core_mod, oracle, monkeypatch):
    """A sparse core (the residency of the 10 000-cell shapes) has no room for the max-sum pools of a whole shard: it traces
    chunk by chunk inside ml_call and keeps the calls.  Same calls, categories and tie flags as the dense core, with a ragged
    last chunk; the tied sites are resolved on a small dense core over just those sites (ScsTreeLikelihood.call_variants)."""
    from sieve_b200.treelikelihood import ScsTreeLikelihood
    n, S = 48, 700
    d = synthetic(n, S, seed=12)
    counts, tree = d["counts"], d["tree"]
    sf, _ = oracle.size_factors(oracle.counts5(counts))
    out = {}
    for mode in ("dense", "sparse"):
        if mode == "sparse":
            monkeypatch.setenv("SIEVE_B200_POOL_SLOTS", "96")
            monkeypatch.setenv("SIEVE_B200_ML_CHUNK", "256")
        tl = ScsTreeLikelihood(counts, tree.copy(), size_factors=sf, asc_mode="felsenstein_mean", n_background_sites=2e3,
                               sparse=(mode == "sparse"))
        tl.calculate_logp()
        calls = tl.call_variants()
        cats = tl.core.ml_categories()
        if mode == "sparse":
            assert tl.core.sparse_info()["keep_leaves"] > 0
            with pytest.raises(RuntimeError):
                tl.core.ml_site(0)
            with pytest.raises(RuntimeError):
                tl.core.ml_node(n + 1)
        # a local move, then again: the chunked trace follows the current state
        moved = tree.copy()
        v = n + n // 2
        lo = max(moved.height[moved.left[v]], moved.height[moved.right[v]])
        moved.height[v] = lo + 0.2 * (moved.height[moved.parent[v]] - lo)
        tl.set_tree(moved, [v])
        tl.calculate_logp()
        calls2 = tl.call_variants()
        out[mode] = (calls, cats, calls2)
        tl.close()
    for i in (0, 2):
        a, b = out["dense"][i], out["sparse"][i]
        np.testing.assert_array_equal(a.genotypes, b.genotypes)
        np.testing.assert_array_equal(a.ado, b.ado)
        np.testing.assert_array_equal(a.category, b.category)
        assert a.tied == b.tied
    np.testing.assert_array_equal(out["dense"][1], out["sparse"][1])
    assert (out["dense"][0].genotypes[:n] > 0).any()


@pytest.mark.parametrize("single", [True, False])
def test_linear_space_twin_of_the_max_sum_pass(core_mod, oracle, single):
    """A core with linear partials (useLogPartials = false) runs the linear-space twin of the max-sum pruning
    (ScsBeerLikelihoodCore.java:1759-2335): leaf children enter as log(P * leaf) with the unclamped P.  Against
    oracle.ml_genotypes.ml_lin_pruning, incl. a zero-length branch above a leaf, where the two twins differ."""
    from oracle import ml_genotypes as mlo
    n, S, K = 11, 260, 4
    d = synthetic(n, S, seed=29)
    counts, tree = d["counts"], d["tree"]
    # a zero-length branch above leaf 3: P = I, log(0 * leaf) = -inf for the off-diagonal candidates
    tree.height[tree.parent[3]] = tree.height[3]
    c5 = oracle.counts5(counts)
    sf, _ = oracle.size_factors(c5)
    rates, props = substmodel.gamma_category_rates(0.8, K)
    mats = substmodel.node_matrices(tree, rates)
    assert np.allclose(mats[3, 0].reshape(4, 4), np.eye(4))
    c = core_mod.ScsCudaLikelihoodCore(n, S, K, 4, False, True, single)
    c.set_counts(counts)
    c.set_size_factors(sf)
    c.set_leaf_params(*PARAMS)
    c.update_leaves(for_update=False)
    c.set_root(tree.root, props, 0)
    nodes = [i for i in range(tree.node_count) if i != tree.root]
    c.set_matrices(nodes, mats[nodes], for_update=False)
    c.update_nodes(tree.ops(), flip=False)
    c.evaluate()
    c.ml_trace()
    tr = mlo.ml_trace(c5, sf, oracle.leaf_params(*PARAMS, single_ado=single, use_log=False), tree.ops(), mats, 0)
    total = flips = 0
    for node in range(n, 2 * n):
        p, b, _ = c.ml_node(node)
        ref, rb = tr["ml"][node], tr["back"][node]
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(p), fin), node
        assert np.all(np.abs(p[fin] - ref[fin]) <= 1e-12 * np.abs(ref[fin])), node
        total += b.size
        flips += int((b != rb).sum())
    assert flips <= 2e-3 * total, (flips, total)
    # the parent of leaf 3: for every parent genotype p the only candidate of that child is p itself
    par = int(tree.parent[3])
    _, b, _ = c.ml_node(par)
    side = 0 if int(tree.left[par]) == 3 else 1
    want = np.array([1, 2, 4, 8], dtype=np.uint8)
    assert np.array_equal((b >> (4 * side)) & 15, np.broadcast_to(want, b.shape))
    # calls agree with the oracle's enumeration on the linear twin's back-pointers
    calls = variants.call_variants(c, tree)
    for s in range(0, S, 7):
        want_c = mlo.call_pattern(tree.ops(), n, tr, s, 0)
        if s not in calls.tied and len(want_c) == 1:
            k, comb = want_c[0]
            assert calls.category[s] == k and list(calls.genotypes[:, s]) == list(comb[n:])
    c.close()
