"""Parity of the CUDA path (through the C-ABI) with the CPU oracle on the same seeded inputs, and with the committed
golden fixtures.  Tolerances (BASELINE.json north_star): per-site log-likelihood 1e-12 relative, total 1e-9 relative."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from sieve_b200 import substmodel  # noqa: E402
from sieve_b200.data import synthetic  # noqa: E402
from sieve_b200.tree import caterpillar, fixture_9_leaves, random_coalescent  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SITE_RTOL = 1e-12
TOTAL_RTOL = 1e-9


@pytest.fixture(scope="module")
def core_mod():
    from sieve_b200 import _lib, core
    _lib.lib()
    assert _lib.lib().sieve_device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return core


def _setup(core_mod, counts, tree, sf, params, mats, props, asc=False, use_log=True, single=True, weights=None, K=4):
    S, n = counts.shape[:2]
    c = core_mod.ScsCudaLikelihoodCore(n, S, K, 4, use_log, asc, single)
    c.set_counts(counts)
    c.set_pattern_weights(weights)
    c.set_size_factors(sf)
    c.set_leaf_params(*params)
    c.update_leaves(for_update=False)
    c.set_root(tree.root, props, 0)
    nodes = [i for i in range(tree.node_count) if i != tree.root]
    c.set_matrices(nodes, mats[nodes][:, :K], for_update=False)
    return c


def _oracle_eval(oracle, counts, tree, sf, params, mats, props, asc, M=0.0, weights=None, single=True, use_log=True,
                 T=1):
    P = oracle.leaf_params(*params, single_ado=single, use_log=use_log)
    return oracle.full_eval(oracle.counts5(counts), sf, P, tree.ops(), mats, props, tree.root, asc, M, T, weights=weights)


PARAMS = (0.3, 50.0, 50.0, 0.008, 150.0, 2.0)


def _close(a, b, rtol, atol=0.0):
    a, b = np.asarray(a), np.asarray(b)
    assert np.all(np.abs(a - b) <= rtol * np.abs(b) + atol), (np.abs(a - b) / (np.abs(b) + atol)).max()


@pytest.mark.parametrize("tag", ["cells10", "cells40"])
def test_golden_example_data(core_mod, tag):
    z = np.load(os.path.join(GOLD, tag + ".npz"))
    g = json.load(open(os.path.join(GOLD, "golden_" + tag + ".json")))
    patterns, weights = z["patterns"], z["weights"]
    n, S, M = g["n"], g["S"], g["M"]
    tree = caterpillar(n, 0.0, 1.0)
    tree.height *= g["tree_scale"]
    mats = np.array(g["matrices"])
    props = np.array(g["proportions"])
    c = _setup(core_mod, patterns, tree, np.array(g["size_factors"]), PARAMS, mats, props, asc=True, weights=weights)
    c.update_nodes(tree.ops(), flip=False)
    r = c.evaluate()
    sl, sc = c.site_log_likelihoods()
    _close(sl, g["site_logl"], SITE_RTOL)
    _close(sc, g["site_log_const"], SITE_RTOL)
    for asc in ("none", "felsenstein_mean", "felsenstein_geometric", "lewis"):
        got = core_mod.combine_shards([r], asc, M)
        assert abs(got - g["logP_" + asc]) <= TOTAL_RTOL * abs(g["logP_" + asc]), (asc, got, g["logP_" + asc])
    # leaf partials of cell 0, reference layout [K][S][G] with the K copies identical (ScsBeerLikelihoodCore.java:418-423)
    lp = c.get_node_partials(0)
    ref = np.array(g["leaf0_log_partials"])
    for k in range(4):
        np.testing.assert_allclose(lp[k], ref, rtol=1e-12, atol=1e-12)
    # device size factors == oracle's (SeqCovModelInterface.java:291-351)
    sf, st = c.compute_size_factors(0)
    np.testing.assert_allclose(sf, g["size_factors"], rtol=1e-13)
    np.testing.assert_allclose(st, g["cov_stats"], rtol=1e-10)
    c.close()


@pytest.mark.parametrize("asc,use_log,single,K", [("none", True, True, 4), ("felsenstein_mean", True, True, 4),
                                                 ("lewis", True, False, 4), ("felsenstein_mean", False, True, 4),
                                                 ("felsenstein_geometric", True, True, 2), ("none", True, True, 1)])
def test_full_evaluation_matches_oracle(core_mod, oracle, asc, use_log, single, K):
    d = synthetic(23, 700, seed=5)
    counts, tree = d["counts"], d["tree"]
    sf, _ = oracle.size_factors(oracle.counts5(counts))
    rates, props = substmodel.gamma_category_rates(0.8, K)
    mats = substmodel.node_matrices(tree, rates)
    M = 5000.0
    c = _setup(core_mod, counts, tree, sf, PARAMS, mats, props, asc=(asc != "none"), use_log=use_log, single=single, K=K)
    c.update_nodes(tree.ops(), flip=False)
    r = c.evaluate()
    sl, sc = c.site_log_likelihoods()
    # the oracle in the SAME mode: log-space pruning (ScsBeerLikelihoodCore.java:1384-1501) or its linear twin (:958-1055)
    logp, osl, osc = _oracle_eval(oracle, counts, tree, sf, PARAMS, mats, props, asc, M, single=single, use_log=use_log)
    _close(sl, osl, SITE_RTOL)
    if asc != "none":
        _close(sc, osc, SITE_RTOL)
    got = core_mod.combine_shards([r], asc, M)
    assert abs(got - logp) <= TOTAL_RTOL * abs(logp)
    # every internal node's partials, reference layout: logs in log mode, plain likelihoods in linear mode
    if True:
        P = oracle.leaf_params(*PARAMS, single_ado=single, use_log=use_log)
        oc = oracle.Core(tree.node_count, tree.n, 700, K, 4, use_log)
        c5 = oracle.counts5(counts)
        for i in range(tree.n):
            oc.set_node_partials(i, oracle.leaf_likelihood(c5, i, sf[i], P))
        for node in range(tree.node_count):
            if node != tree.root:
                for k in range(K):
                    oc.set_node_matrix(node, k, mats[node, k])
        for p, c1, c2 in tree.ops():
            oc.calculate_partials(c1, c1 < tree.n, c2, 0 <= c2 < tree.n, p, 0)
        atol = 1e-11 if use_log else 1e-318     # linear mode: subnormal values carry only a few bits
        for node in (tree.n, tree.n + 7, tree.root - 1, tree.root):
            np.testing.assert_allclose(c.get_node_partials(node), oc.get_node_partials(node), rtol=2e-12, atol=atol)
            if asc != "none":
                np.testing.assert_allclose(c.get_node_partials(node, const=True), oc.get_const_partials(node),
                                           rtol=2e-12, atol=atol)
    c.close()


def test_linear_partials_specific_behaviour(core_mod, oracle):
    """useLogPartials = false (ScsBeerLikelihoodCore.java:857-1055), the cases where it differs from the log-space core:

    * a zero-length branch: the linear twin multiplies by the plain matrix, there is no Double.MIN_VALUE clamp of P
      (:1430 belongs to the log-space core only), so P = I passes the child's partial through unchanged;
    * a constant-site product that is exactly 0: the leaf likelihood of the constant genotype underflows in the
      reference's unscaled linear arithmetic, `cst1 * cst2 > 0 ? ... : Double.MIN_VALUE` (:1041-1047) fires and the site's
      logConstRoot ends at log(Double.MIN_VALUE)-ish or -inf.  The device arithmetic is always scaled and keeps the true
      value (documented deviation): the site contributes exp(.) = 0 to the Felsenstein sum either way,
      so the totals agree; only the read-back of that site's logConstRoot differs.
    """
    d = synthetic(19, 300, seed=77, max_cov=200)
    counts, tree = d["counts"].copy(), d["tree"]
    tree.height[4] = tree.height[tree.parent[4]]             # zero-length branch above leaf 4
    # site 11: the cell with the largest size factor has 1 500 reads, all of them the first alternative nucleotide: its
    # genotype 0/0 log-likelihood is about -790 (exp underflows to 0), the other genotypes about -290 (representable)
    sf0, _ = oracle.size_factors(oracle.counts5(counts))
    counts[11, int(np.argmax(sf0))] = (1500, 0, 0, 1500)
    sf, _ = oracle.size_factors(oracle.counts5(counts))
    rates, props = substmodel.gamma_category_rates(1.0, 4)
    mats = substmodel.node_matrices(tree, rates)
    assert np.array_equal(mats[4, 0].reshape(4, 4), np.eye(4))
    M = 3000.0
    c = _setup(core_mod, counts, tree, sf, PARAMS, mats, props, asc=True, use_log=False)
    c.update_nodes(tree.ops(), flip=False)
    r = c.evaluate()
    sl, sc = c.site_log_likelihoods()
    logp, osl, osc = _oracle_eval(oracle, counts, tree, sf, PARAMS, mats, props, "felsenstein_mean", M, use_log=False)
    assert np.isfinite(osl).all()
    _close(sl, osl, SITE_RTOL)
    # wherever the reference's unscaled constant-site product is representable the two agree ...
    ok = np.isfinite(osc) & (osc > -700.0)
    assert not ok[11] and ok.sum() >= 200
    _close(sc[ok], osc[ok], SITE_RTOL)
    # ... and where the reference's linear core lost it to underflow (site 11 by construction, and any site with a real
    # mutation carried by several well-covered cells) the device keeps the true, very negative value
    assert np.isfinite(sc[~ok]).all() and (sc[~ok] < -700.0).all()
    assert sc[11] < -745.0
    got = core_mod.combine_shards([r], "felsenstein_mean", M)
    assert abs(got - logp) <= TOTAL_RTOL * abs(logp)
    # parent of the zero-length branch: linear partials against the oracle's linear core
    P = oracle.leaf_params(*PARAMS, use_log=False)
    oc = oracle.Core(tree.node_count, tree.n, 300, 4, 4, False)
    c5 = oracle.counts5(counts)
    for i in range(tree.n):
        oc.set_node_partials(i, oracle.leaf_likelihood(c5, i, sf[i], P))
    for node in range(tree.node_count):
        if node != tree.root:
            for k in range(4):
                oc.set_node_matrix(node, k, mats[node, k])
    for p, c1, c2 in tree.ops():
        oc.calculate_partials(c1, c1 < tree.n, c2, 0 <= c2 < tree.n, p, 0)
    for node in (int(tree.parent[4]), tree.root):
        np.testing.assert_allclose(c.get_node_partials(node), oc.get_node_partials(node), rtol=2e-12, atol=1e-318)
    # and the same zero-length branch through the distance fast path (no clamp either in linear mode)
    c2_ = _setup(core_mod, counts, tree, sf, PARAMS, mats, props, asc=True, use_log=False)
    nodes = np.array([i for i in range(tree.node_count) if i != tree.root], dtype=np.int32)
    c2_.set_distances(nodes, substmodel.node_distances(tree, rates)[nodes], for_update=False)
    c2_.update_nodes(tree.ops(), flip=False)
    c2_.evaluate()
    sl2, _ = c2_.site_log_likelihoods()
    _close(sl2, osl, SITE_RTOL)
    c.close()
    c2_.close()


def test_leaf_scale_sums_over_cell_groups_and_site_blocks(core_mod, oracle, monkeypatch):
    """The leaves' log scales exist only as their per-site sum over the cells (leaf.cuh): 70 cells = three groups of 32
    with a ragged last one, 2 300 sites = three blocks of 1 024 with a ragged last one.  Site log-likelihoods against the
    oracle, read-backs of leaves / internal nodes (their scale sum is re-derived over the leaf list below the node), and
    bit-identity of the streamed evaluation, whose site chunks cut the blocks differently."""
    from sieve_b200.treelikelihood import ScsTreeLikelihood
    d = synthetic(70, 2300, seed=77)
    counts, tree = d["counts"], d["tree"]
    c5 = oracle.counts5(counts)
    sf, _ = oracle.size_factors(c5)
    rates, props = substmodel.gamma_category_rates(0.9, 4)
    mats = substmodel.node_matrices(tree, rates)
    M = 8000.0
    c = _setup(core_mod, counts, tree, sf, PARAMS, mats, props, asc=True)
    c.update_nodes(tree.ops(), flip=False)
    r = c.evaluate()
    sl, sc = c.site_log_likelihoods()
    logp, osl, osc = _oracle_eval(oracle, counts, tree, sf, PARAMS, mats, props, "felsenstein_mean", M)
    _close(sl, osl, SITE_RTOL)
    _close(sc, osc, SITE_RTOL)
    got = core_mod.combine_shards([r], "felsenstein_mean", M)
    assert abs(got - logp) <= TOTAL_RTOL * abs(logp)
    P = oracle.leaf_params(*PARAMS)
    for j in (0, 31, 32, 63, 64, 69):       # first / last cell of every group
        np.testing.assert_allclose(c.get_node_partials(j)[0], oracle.leaf_likelihood(c5, j, sf[j], P), rtol=2e-13, atol=2e-12)
    oc = oracle.Core(tree.node_count, tree.n, 2300, 4, 4, True)
    for i in range(tree.n):
        oc.set_node_partials(i, oracle.leaf_likelihood(c5, i, sf[i], P))
    for node in range(tree.node_count):
        if node != tree.root:
            for k in range(4):
                oc.set_node_matrix(node, k, mats[node, k])
    for p, c1, c2 in tree.ops():
        oc.calculate_partials(c1, c1 < tree.n, c2, 0 <= c2 < tree.n, p, 0)
    for node in (tree.n, tree.n + 33, tree.root - 1, tree.root):
        np.testing.assert_allclose(c.get_node_partials(node), oc.get_node_partials(node), rtol=2e-12, atol=1e-11)
        np.testing.assert_allclose(c.get_node_partials(node, const=True), oc.get_const_partials(node), rtol=2e-12, atol=1e-11)
    c.close()
    monkeypatch.setenv("SIEVE_B200_CHUNK_SITES", "736")     # chunks of 736 sites: 4 chunks, none aligned to a leaf block
    res = ScsTreeLikelihood(counts, tree.copy(), asc_mode="felsenstein_mean", n_background_sites=M, size_factors=sf)
    eph = ScsTreeLikelihood(counts, tree.copy(), asc_mode="felsenstein_mean", n_background_sites=M, ephemeral=True, size_factors=sf)
    assert res.calculate_logp() == eph.calculate_logp()
    np.testing.assert_array_equal(eph.pattern_log_likelihoods()[0], res.pattern_log_likelihoods()[0])
    np.testing.assert_array_equal(eph.pattern_log_likelihoods()[1], res.pattern_log_likelihoods()[1])
    res.close()
    eph.close()


def test_leaf_kernel_without_bounds_checks_is_bit_identical(core_mod, monkeypatch):
    """When the ingest scan finds every count inside the look-up tables the log-space leaf kernel runs without its
    out-of-table paths (k_leaf<1, true>); SIEVE_B200_LEAF_CHECKED=1 keeps them.  Same arithmetic, so the same bits -- also
    when the tables are capped below the largest count, where the unchecked kernel must not be chosen at all.
    (SIEVE_B200_LEAF_LIN=0: the log-space kernels on both sides.)  The default kernel, k_leaf_lin, forms the same
    likelihoods as products of tabulated mantissa / exponent pairs: equal to rounding, leaf by leaf and site by site."""
    from sieve_b200.treelikelihood import ScsTreeLikelihood
    d = synthetic(37, 1300, seed=19)
    counts, tree = d["counts"], d["tree"]
    out = {}
    monkeypatch.setenv("SIEVE_B200_LEAF_LIN", "0")
    for cap in (None, "24"):
        for checked in ("0", "1"):
            monkeypatch.setenv("SIEVE_B200_LEAF_CHECKED", checked)
            if cap:
                monkeypatch.setenv("SIEVE_B200_TABLE_CAP", cap)
            else:
                monkeypatch.delenv("SIEVE_B200_TABLE_CAP", raising=False)
            tl = ScsTreeLikelihood(counts, tree.copy(), asc_mode="felsenstein_mean", n_background_sites=1000.0)
            logp = tl.calculate_logp()
            out[(cap, checked)] = (logp, tl.pattern_log_likelihoods()[0].copy(), tl.pattern_log_likelihoods()[1].copy())
            tl.close()
    ref = out[(None, "1")]
    for key in ((None, "0"), ("24", "0"), ("24", "1")):
        got = out[key]
        if key[0] is None:
            assert got[0] == ref[0]
            np.testing.assert_array_equal(got[1], ref[1])
            np.testing.assert_array_equal(got[2], ref[2])
        else:   # direct evaluation beyond the capped tables: same formulas, equal to rounding
            _close(got[1], ref[1], SITE_RTOL)
            _close(got[2], ref[2], SITE_RTOL)
    np.testing.assert_array_equal(out[("24", "0")][1], out[("24", "1")][1])
    # the default kernel against the log-space one
    monkeypatch.setenv("SIEVE_B200_LEAF_CHECKED", "0")
    monkeypatch.delenv("SIEVE_B200_TABLE_CAP", raising=False)
    tl0 = ScsTreeLikelihood(counts, tree.copy(), asc_mode="felsenstein_mean", n_background_sites=1000.0)
    tl0.calculate_logp()
    monkeypatch.delenv("SIEVE_B200_LEAF_LIN", raising=False)
    tl = ScsTreeLikelihood(counts, tree.copy(), asc_mode="felsenstein_mean", n_background_sites=1000.0)
    logp = tl.calculate_logp()
    site, const = tl.pattern_log_likelihoods()
    for leaf in range(0, tree.n, 4):
        _close(tl.core.get_node_partials(leaf), tl0.core.get_node_partials(leaf), 1e-12, 1e-12)
    tl.close()
    tl0.close()
    _close(site, ref[1], 1e-13)
    _close(const, ref[2], 1e-13)
    assert abs(logp - ref[0]) <= 1e-12 * abs(ref[0])


def test_leaf_kernel_all_cells_and_edge_counts(core_mod, oracle, monkeypatch):
    rng = np.random.default_rng(8)
    S, n = 300, 5
    counts = np.zeros((S, n, 4), dtype=np.int32)
    cov = rng.integers(0, 90, size=(S, n))
    cov[:10] = 0                                   # empty sites: the 1.0 quirk of DirichletMultinomial.java:29-30
    cov[10:20, 0] = rng.integers(3000, 9000, 10)   # beyond the lookup tables (direct evaluation path)
    alt = (cov * rng.random((S, n)) * (rng.random((S, n)) < 0.4)).astype(np.int64)
    counts[..., 0] = alt
    counts[..., 1] = ((cov - alt) * 0.1 * rng.random((S, n))).astype(np.int64)
    counts[..., 2] = ((cov - alt - counts[..., 1]) * 0.05 * rng.random((S, n))).astype(np.int64)
    counts[..., 3] = cov
    c5 = oracle.counts5(counts)
    sf, _ = oracle.size_factors(c5)
    tree = caterpillar(n)
    for single in (True, False):
        for params in (PARAMS, (0.05, 20.0, 80.0, 0.03, 60.0, 5.0)):
            c = core_mod.ScsCudaLikelihoodCore(n, S, 4, 4, True, False, single)
            c.set_counts(counts)
            c.set_size_factors(sf)
            c.set_leaf_params(*params)
            c.update_leaves(for_update=False)
            P = oracle.leaf_params(*params, single_ado=single)
            for j in range(n):
                ref = oracle.leaf_likelihood(c5, j, sf[j], P)
                got = c.get_node_partials(j)[0]
                # the device keeps the 4 genotype likelihoods of a (cell, site) in linear space under one common scale:
                # a genotype more than e^-700 below the best one underflows to 0 there (it cannot influence a site
                # likelihood by more than 1e-300 relative; only at coverage in the thousands).  The read-back takes those
                # entries from the log-space mixture instead, so the reference's finite logs come back everywhere.
                lost = ref < ref.max(axis=1, keepdims=True) - 700.0
                assert lost.sum() < 0.02 * ref.size
                assert np.all(np.isfinite(got))
                np.testing.assert_allclose(got[~lost], ref[~lost], rtol=2e-13, atol=2e-12)
                np.testing.assert_allclose(got[lost], ref[lost], rtol=2e-13, atol=2e-12)
            c.close()


def test_small_table_cap_uses_direct_path(core_mod, oracle, monkeypatch):
    monkeypatch.setenv("SIEVE_B200_TABLE_CAP", "16")
    d = synthetic(6, 200, seed=9)
    counts = d["counts"]
    c5 = oracle.counts5(counts)
    sf, _ = oracle.size_factors(c5)
    c = core_mod.ScsCudaLikelihoodCore(6, 200, 4, 4, True, False, True)
    c.set_counts(counts)
    c.set_size_factors(sf)
    c.set_leaf_params(*PARAMS)
    c.update_leaves(for_update=False)
    P = oracle.leaf_params(*PARAMS)
    for j in range(6):
        np.testing.assert_allclose(c.get_node_partials(j)[0], oracle.leaf_likelihood(c5, j, sf[j], P), rtol=2e-13, atol=2e-12)
    c.close()


def test_nan_propagates_for_zero_error_rate(core_mod):
    # logBeta(0, c) is NaN in the reference (SURVEY.md section 8a' item 2): effSeqErrRate == 0 with an alt read
    counts = np.zeros((4, 2, 4), dtype=np.int32)
    counts[..., 0] = 3
    counts[..., 3] = 20
    c = core_mod.ScsCudaLikelihoodCore(2, 4, 4, 4, True, False, True)
    c.set_counts(counts)
    c.set_size_factors(np.ones(2))
    c.set_leaf_params(0.3, 50.0, 50.0, 0.0, 150.0, 2.0)
    c.update_leaves(for_update=False)
    assert np.isnan(c.get_node_partials(0)).all()
    c.close()


def test_per_level_launches_equal_single_walk(core_mod, oracle):
    d = synthetic(40, 333, seed=12)
    counts, tree = d["counts"], d["tree"]
    sf, _ = oracle.size_factors(oracle.counts5(counts))
    rates, props = substmodel.gamma_category_rates(1.0, 4)
    mats = substmodel.node_matrices(tree, rates)
    out = []
    for per_level in (False, True):
        c = _setup(core_mod, counts, tree, sf, PARAMS, mats, props, asc=True)
        if per_level:
            ops, off = tree.ops_by_level()
            c.update_nodes(ops, off, flip=False, per_level=True)
        else:
            c.update_nodes(tree.ops(), flip=False)
        r = c.evaluate()
        out.append((r.sum_log_l, r.const_lse) + c.site_log_likelihoods())
        c.close()
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1]
    np.testing.assert_array_equal(out[0][2], out[1][2])
    np.testing.assert_array_equal(out[0][3], out[1][3])


@pytest.mark.parametrize("use_distances", [True, False])
def test_per_level_steps_with_unchanged_matrices_and_change_tracking(core_mod, oracle, use_distances):
    """Regression (round-1 advisor finding): consecutive per-level steps that flip the matrix buffers and hand in the SAME
    values.  Change tracking points the node back at the buffer that already holds them -- the per-level ops must be made
    after that decision, or they read the slot that never received the data.  Bit-for-bit against the single-launch walk,
    with one branch of exactly zero length so that the zero-distance flag of the ops is exercised too."""
    d = synthetic(37, 411, seed=21)
    counts, tree = d["counts"], d["tree"]
    leaf = 5
    tree.height[leaf] = tree.height[tree.parent[leaf]]   # a zero-length branch (a cell sampled at its parent's height)
    assert tree.branch_lengths()[leaf] == 0.0
    sf, _ = oracle.size_factors(oracle.counts5(counts))
    rates, props = substmodel.gamma_category_rates(1.0, 4)
    mats = substmodel.node_matrices(tree, rates)
    nodes = np.array([i for i in range(tree.node_count) if i != tree.root], dtype=np.int32)
    D = substmodel.node_distances(tree, rates)[nodes]
    D2 = D.copy()
    D2[3] *= 1.25                                          # third step: one branch really changes
    mats2 = mats.copy()
    for k in range(4):
        mats2[nodes[3], k] = substmodel.transition_probabilities(D2[3, k]).reshape(16)
    lvl_ops, lvl_off = tree.ops_by_level()
    out = {}
    for per_level in (False, True):
        c = _setup(core_mod, counts, tree, sf, PARAMS, mats, props, asc=True)
        c.set_elision(True)
        res = []
        for step in range(4):
            Dn, Mn = (D2, mats2) if step == 2 else (D, mats)
            if use_distances:
                c.set_distances(nodes, Dn, for_update=True)
            else:
                c.set_matrices(nodes, Mn[nodes], for_update=True)
            if per_level:
                c.update_nodes(lvl_ops, lvl_off, flip=True, per_level=True)
            else:
                c.update_nodes(tree.ops(), flip=True)
            r = c.evaluate()
            res.append((r.sum_log_l, r.const_lse) + c.site_log_likelihoods())
        out[per_level] = res
        c.close()
    for step in range(4):
        a, b = out[False][step], out[True][step]
        assert np.isfinite(a[0]) and a[0] == b[0] and a[1] == b[1], (step, a[0], b[0])
        np.testing.assert_array_equal(a[2], b[2])
        np.testing.assert_array_equal(a[3], b[3])
    # steps 0, 1 and 3 evaluate the same state; the oracle agrees
    assert out[True][0][0] == out[True][1][0] == out[True][3][0] != out[True][2][0]
    _, osl, osc = _oracle_eval(oracle, counts, tree, sf, PARAMS, mats, props, "felsenstein_mean", 100.0)
    _close(out[True][3][2], osl, SITE_RTOL)
    _close(out[True][3][3], osc, SITE_RTOL)


def test_reference_call_sequence_and_store_restore(core_mod, oracle):
    """The traverse() call sequence, node by node, then a proposal that is rejected and one that is accepted."""
    rng = np.random.default_rng(3)
    tree = fixture_9_leaves()
    d = synthetic(9, 257, seed=4, tree=tree)
    counts = d["counts"]
    sf, _ = oracle.size_factors(oracle.counts5(counts))
    rates, props = substmodel.gamma_category_rates(1.0, 4)
    mats = substmodel.node_matrices(tree, rates)
    n = tree.n
    c = core_mod.ScsCudaLikelihoodCore(n, 257, 4, 4, True, True, True)
    c.set_counts(counts)
    c.set_size_factors(sf)
    c.set_leaf_params(*PARAMS)
    c.update_leaves(for_update=False)
    c.set_root(tree.root, props, 0)
    for node in range(tree.node_count):
        if node != tree.root:
            c.set_node_matrix_for_update(node)
            for k in range(4):
                c.set_node_matrix(node, k, mats[node, k])
    for p, c1, c2 in tree.ops():
        c.set_node_partials_for_update(p)
        c.calculate_log_partials(c1, c1 < n, c2, 0 <= c2 < n, p, 0)
    r0 = c.evaluate()
    logp0, sl0, _ = _oracle_eval(oracle, counts, tree, sf, PARAMS, mats, props, "felsenstein_mean", 1000.0)
    got0 = core_mod.combine_shards([r0], "felsenstein_mean", 1000.0)
    assert abs(got0 - logp0) <= TOTAL_RTOL * abs(logp0)
    site0 = c.site_log_likelihoods()[0]
    # proposal: rescale the branch above leaf 4 (dirty chain 4 -> root)
    c.store()
    tree2 = tree.copy()
    mats2 = mats.copy()
    for k in range(4):
        mats2[4, k] = substmodel.transition_probabilities(tree.branch_length(4) * 3.0 * rates[k]).reshape(16)
    chain = tree.ops(tree.dirty_closure([tree.parent[4]]))
    r1 = c.step([4], mats2[[4]], chain)
    logp1, sl1, _ = _oracle_eval(oracle, counts, tree2, sf, PARAMS, mats2, props, "felsenstein_mean", 1000.0)
    _close(c.site_log_likelihoods()[0], sl1, SITE_RTOL)
    assert abs(core_mod.combine_shards([r1], "felsenstein_mean", 1000.0) - logp1) <= TOTAL_RTOL * abs(logp1)
    # reject: index swap only; re-deriving the root from the restored buffers gives the old answer bit for bit
    c.restore()
    c.update_nodes(tree.ops([tree.root]), flip=False)
    r2 = c.evaluate()
    np.testing.assert_array_equal(c.site_log_likelihoods()[0], site0)
    assert r2.sum_log_l == r0.sum_log_l
    # accept path: same proposal again, then a leaf-parameter proposal on top
    c.store()
    c.step([4], mats2[[4]], chain)
    c.store()
    p2 = (0.25, 45.0, 60.0, 0.01, 120.0, 2.5)
    c.set_leaf_params(*p2)
    r3 = c.step([], np.zeros((0, 4, 16)), tree.ops(), update_leaves=True)
    logp3, sl3, _ = _oracle_eval(oracle, counts, tree2, sf, p2, mats2, props, "felsenstein_mean", 1000.0)
    _close(c.site_log_likelihoods()[0], sl3, SITE_RTOL)
    # reject the parameter proposal: leaves flip back too
    c.restore()
    c.set_leaf_params(*PARAMS)
    c.update_nodes(tree.ops([tree.root]), flip=False)
    c.evaluate()
    _close(c.site_log_likelihoods()[0], sl1, SITE_RTOL)
    c.close()


def test_incremental_updates_equal_full_recompute(core_mod, oracle):
    from sieve_b200.treelikelihood import ScsTreeLikelihood
    rng = np.random.default_rng(21)
    d = synthetic(60, 500, seed=22)
    counts, tree = d["counts"], d["tree"]
    sf, _ = oracle.size_factors(oracle.counts5(counts))
    tl = ScsTreeLikelihood(counts, tree.copy(), size_factors=sf, asc_mode="felsenstein_mean", n_background_sites=3000.0)
    full = ScsTreeLikelihood(counts, tree.copy(), size_factors=sf, asc_mode="felsenstein_mean", n_background_sites=3000.0)
    tl.calculate_logp()
    for it in range(12):
        t2 = tl.tree.copy()
        node = int(rng.integers(tree.n, tree.node_count - 1))       # an internal non-root node: move its height
        lo = max(t2.height[t2.left[node]], t2.height[t2.right[node]])
        hi = t2.height[t2.parent[node]]
        t2.height[node] = lo + (hi - lo) * rng.random()
        tl.store()
        tl.set_tree(t2, [node])
        a = tl.calculate_logp()
        full.has_dirt = 2
        full.tree = t2.copy()
        b = full.calculate_logp()
        assert abs(a - b) <= 1e-12 * abs(b), (it, a, b)
        np.testing.assert_allclose(tl.pattern_log_likelihoods()[0], full.pattern_log_likelihoods()[0], rtol=1e-13)
        if it % 3 == 2:
            tl.restore()
            assert tl.calculate_logp() is not None
    tl.close()
    full.close()


def test_site_shards_combine_like_threads(core_mod, oracle):
    from sieve_b200.treelikelihood import ScsTreeLikelihood, ThreadedScsTreeLikelihood
    d = synthetic(17, 403, seed=31)
    counts, tree = d["counts"], d["tree"]
    M = 7777.0
    for asc in ("none", "felsenstein_mean", "lewis"):
        one = ScsTreeLikelihood(counts, tree.copy(), asc_mode=asc, n_background_sites=M)
        a = one.calculate_logp()
        for T in (2, 3, 8):
            many = ThreadedScsTreeLikelihood(counts, tree.copy(), T, asc_mode=asc, n_background_sites=M)
            b = many.calculate_logp()
            assert abs(a - b) <= 1e-11 * abs(a), (asc, T, a, b)
            np.testing.assert_array_equal(many.pattern_log_likelihoods()[0], one.pattern_log_likelihoods()[0])
            many.close()
        sf = one.size_factors
        logp, _, _ = _oracle_eval(oracle, counts, tree, sf, PARAMS, substmodel.node_matrices(tree, one.rates),
                                  one.proportions, asc, M, T=3)
        assert abs(a - logp) <= TOTAL_RTOL * abs(logp)
        one.close()


def test_weights_and_device_size_factors(core_mod, oracle):
    rng = np.random.default_rng(41)
    d = synthetic(12, 150, seed=42)
    counts, tree = d["counts"], d["tree"]
    w = rng.integers(1, 5, size=150).astype(np.int32)
    c5 = oracle.counts5(counts)
    sf_ref, st_ref = oracle.size_factors(c5, w)
    rates, props = substmodel.gamma_category_rates(1.0, 4)
    mats = substmodel.node_matrices(tree, rates)
    c = core_mod.ScsCudaLikelihoodCore(12, 150, 4, 4, True, True, True)
    c.set_counts(counts)
    c.set_pattern_weights(w)
    sf, st = c.compute_size_factors(0)
    np.testing.assert_allclose(sf, sf_ref, rtol=1e-13)
    np.testing.assert_allclose(st, st_ref, rtol=1e-10)
    c.set_leaf_params(*PARAMS)
    c.update_leaves(for_update=False)
    c.set_root(tree.root, props, 0)
    nodes = [i for i in range(tree.node_count) if i != tree.root]
    c.set_matrices(nodes, mats[nodes], for_update=False)
    c.update_nodes(tree.ops(), flip=False)
    r = c.evaluate()
    for asc in ("felsenstein_mean", "felsenstein_geometric", "lewis"):
        logp, _, _ = _oracle_eval(oracle, counts, tree, sf_ref, PARAMS, mats, props, asc, 900.0, weights=w)
        got = core_mod.combine_shards([r], asc, 900.0)
        assert abs(got - logp) <= TOTAL_RTOL * abs(logp), asc
    # zeroCovMode = 1 (SeqCovModelInterface.java:308-311)
    sf1, _ = c.compute_size_factors(1)
    np.testing.assert_allclose(sf1, oracle.size_factors(c5, w, 1)[0], rtol=1e-13)
    c.close()


def test_argument_errors_are_status_codes(core_mod):
    from sieve_b200 import _lib
    with pytest.raises(_lib.SieveError) as e:
        core_mod.ScsCudaLikelihoodCore(4, 10, 4, 3)
    assert e.value.code == _lib.ERR_INVALID
    c = core_mod.ScsCudaLikelihoodCore(4, 10)
    with pytest.raises(_lib.SieveError) as e:
        c.update_leaves()
    assert e.value.code == _lib.ERR_STATE
    with pytest.raises(_lib.SieveError):
        c.calculate_partials(0, False, 1, True, 5)    # is_leaf contradicts the numbering
    with pytest.raises(_lib.SieveError):
        c.set_node_matrix(99, 0, np.eye(4))
    c.close()


def test_mid_size_properties(core_mod):
    """Size-independent checks at a size the oracle is not asked to follow: shard additivity, walk == per-level,
    incremental == full, restore is exact."""
    from sieve_b200.treelikelihood import ScsTreeLikelihood, ThreadedScsTreeLikelihood
    d = synthetic(300, 6000, seed=77)
    counts, tree = d["counts"], d["tree"]
    one = ScsTreeLikelihood(counts, tree.copy(), asc_mode="felsenstein_mean", n_background_sites=1e5)
    a = one.calculate_logp()
    many = ThreadedScsTreeLikelihood(counts, tree.copy(), 4, asc_mode="felsenstein_mean", n_background_sites=1e5,
                                     size_factors=one.size_factors)
    b = many.calculate_logp()
    assert np.isfinite(a) and abs(a - b) <= 1e-11 * abs(a)
    site = one.pattern_log_likelihoods()[0]
    one.store()
    t2 = tree.copy()
    node = tree.n + 5
    lo = max(t2.height[t2.left[node]], t2.height[t2.right[node]])
    t2.height[node] = 0.5 * (lo + t2.height[t2.parent[node]])
    one.set_tree(t2, [node])
    c = one.calculate_logp()
    assert c != a
    one.restore()
    one.core.update_nodes(tree.ops([tree.root]), flip=False)
    one.core.evaluate()
    np.testing.assert_array_equal(one.pattern_log_likelihoods()[0], site)
    one.close()
    many.close()


@pytest.mark.parametrize("asc,K,per_level", [("felsenstein_mean", 4, False), ("none", 4, False), ("felsenstein_mean", 3, False),
                                            ("felsenstein_mean", 4, True)])
def test_distance_fast_path_matches_oracle(core_mod, oracle, asc, K, per_level):
    """sieve_set_distances: P = exp(Q d) formed on the device for SIEVE's fixed rate matrix, Q-specialised walk."""
    d = synthetic(31, 555, seed=15)
    counts, tree = d["counts"], d["tree"]
    tree.height[tree.n + 3] = max(tree.height[tree.left[tree.n + 3]], tree.height[tree.right[tree.n + 3]])  # a zero-length branch
    sf, _ = oracle.size_factors(oracle.counts5(counts))
    rates, props = substmodel.gamma_category_rates(0.6, K)
    br = np.random.default_rng(1).lognormal(0, 0.3, tree.node_count)
    mats = substmodel.node_matrices(tree, rates, branch_rates=br)
    dist = substmodel.node_distances(tree, rates, branch_rates=br)
    assert (dist[:tree.root] == 0).any()
    S, n = counts.shape[:2]
    out = {}
    for mode in ("distances", "matrices"):
        c = core_mod.ScsCudaLikelihoodCore(n, S, K, 4, True, asc != "none", True)
        c.set_counts(counts)
        c.set_size_factors(sf)
        c.set_leaf_params(*PARAMS)
        c.update_leaves(for_update=False)
        c.set_root(tree.root, props, 0)
        nodes = [i for i in range(tree.node_count) if i != tree.root]
        if mode == "distances":
            c.set_distances(nodes, dist[nodes], for_update=False)
        else:
            c.set_matrices(nodes, mats[nodes], for_update=False)
        if per_level:
            ops, off = tree.ops_by_level()
            c.update_nodes(ops, off, flip=False, per_level=True)
        else:
            c.update_nodes(tree.ops(), flip=False)
        r = c.evaluate()
        out[mode] = (r, c.site_log_likelihoods(), c.get_node_partials(tree.root - 1), c)
    logp, osl, osc = _oracle_eval(oracle, counts, tree, sf, PARAMS, mats, props, asc, 3000.0)
    for mode in out:
        r, (sl, sc), _, _ = out[mode]
        _close(sl, osl, SITE_RTOL)
        if asc != "none":
            _close(sc, osc, SITE_RTOL)
        got = core_mod.combine_shards([r], asc, 3000.0)
        assert abs(got - logp) <= TOTAL_RTOL * abs(logp)
    np.testing.assert_allclose(out["distances"][2], out["matrices"][2], rtol=1e-12, atol=1e-11)
    # an incremental step through the fast path: new distance on one branch, dirty chain only
    c = out["distances"][3]
    c.store()
    node = 5
    dist2 = dist.copy()
    dist2[node] *= 2.5
    mats2 = mats.copy()
    for k in range(K):
        mats2[node, k] = substmodel.transition_probabilities(dist2[node, k]).reshape(16)
    chain = tree.ops(tree.dirty_closure([tree.parent[node]]))
    r = c.step_distances([node], dist2[[node]], chain)
    logp2, osl2, _ = _oracle_eval(oracle, counts, tree, sf, PARAMS, mats2, props, asc, 3000.0)
    _close(c.site_log_likelihoods()[0], osl2, SITE_RTOL)
    assert abs(core_mod.combine_shards([r], asc, 3000.0) - logp2) <= TOTAL_RTOL * abs(logp2)
    c.restore()
    c.update_nodes(tree.ops([tree.root]), flip=False)
    c.evaluate()
    np.testing.assert_array_equal(c.site_log_likelihoods()[0], out["distances"][1][0])
    for mode in out:
        out[mode][3].close()


@pytest.mark.parametrize("mix,asc", [("stage1", "none"), ("stage2", "felsenstein_mean"), ("relaxed2", "felsenstein_mean")])
def test_mcmc_replay_keeps_the_incremental_state_exact(core_mod, oracle, mix, asc):
    """Hundreds of proposals (all operator classes, accept and reject) through store / step / restore; afterwards the
    incrementally maintained log-likelihood must equal a from-scratch evaluation of the final state on the device AND
    the oracle's evaluation of that state (tree, branch rates, parameters read back from the driver)."""
    from sieve_b200.mcmc import McmcReplay, gamma_category_rates
    d = synthetic(26, 400, seed=51)
    counts, tree = d["counts"], d["tree"]
    M = 6000.0
    S, n = counts.shape[:2]
    c = core_mod.ScsCudaLikelihoodCore(n, S, 4, 4, True, asc != "none", True)
    c.set_counts(counts)
    sf, _ = c.compute_size_factors(0)
    mc = McmcReplay(c, tree, mix=mix, asc_mode=asc, n_background_sites=M, seed=7)
    st = mc.run(400)
    assert st["states"] == 400 and 0 < st["accepted"] < 400 and st["likelihood_evaluations"] > 100
    assert len(st["by_op"]) >= 7
    state = mc.state()
    t2 = state["tree"]
    assert (t2.branch_lengths()[:t2.root] >= 0).all() and sorted(t2.post_order()) == list(range(n, 2 * n))
    again = mc.full_recompute()
    assert abs(again - state["log_p"]) <= 1e-11 * abs(again)
    rates = gamma_category_rates(state["gamma_shape"], 4)
    mats = substmodel.node_matrices(t2, rates, branch_rates=state["branch_rates"])
    ref, _, _ = _oracle_eval(oracle, counts, t2, sf, state["leaf"], mats, np.full(4, 0.25), asc, M)
    assert abs(state["log_p"] - ref) <= TOTAL_RTOL * abs(ref), (state["log_p"], ref)
    mc.close()
    c.close()


def test_streamed_evaluation_equals_resident(core_mod, oracle, monkeypatch):
    """SIEVE_FLAG_EPHEMERAL: site chunks, leaves recomputed per chunk, internal partials on a few reused scratch slots.
    Same arithmetic per site => bit-identical site log-likelihoods; also through store / propose / restore."""
    from sieve_b200.treelikelihood import ScsTreeLikelihood
    monkeypatch.setenv("SIEVE_B200_CHUNK_SITES", "96")      # 5 chunks, the last one ragged
    d = synthetic(45, 437, seed=61)
    counts, tree = d["counts"], d["tree"]
    M = 2500.0
    res = ScsTreeLikelihood(counts, tree.copy(), asc_mode="felsenstein_mean", n_background_sites=M)
    eph = ScsTreeLikelihood(counts, tree.copy(), asc_mode="felsenstein_mean", n_background_sites=M, ephemeral=True,
                            size_factors=res.size_factors)
    a, b = res.calculate_logp(), eph.calculate_logp()
    np.testing.assert_array_equal(eph.pattern_log_likelihoods()[0], res.pattern_log_likelihoods()[0])
    np.testing.assert_array_equal(eph.pattern_log_likelihoods()[1], res.pattern_log_likelihoods()[1])
    assert a == b
    rates, props = substmodel.gamma_category_rates(1.0, 4)
    ref, _, _ = _oracle_eval(oracle, counts, tree, res.size_factors, PARAMS, substmodel.node_matrices(tree, rates), props,
                             "felsenstein_mean", M)
    assert abs(b - ref) <= TOTAL_RTOL * abs(ref)
    rng = np.random.default_rng(5)
    for it in range(6):
        t2 = eph.tree.copy()
        node = int(rng.integers(tree.n, tree.node_count - 1))
        lo = max(t2.height[t2.left[node]], t2.height[t2.right[node]])
        t2.height[node] = lo + (t2.height[t2.parent[node]] - lo) * rng.random()
        for tl in (res, eph):
            tl.store()
            tl.set_tree(t2.copy(), [node])
        a, b = res.calculate_logp(), eph.calculate_logp()
        assert a == b
        if it % 2:
            for tl in (res, eph):
                tl.restore()
    # a leaf-parameter proposal and its rejection
    from sieve_b200.treelikelihood import LeafModelParams
    for tl in (res, eph):
        tl.store()
        tl.set_leaf_params(LeafModelParams(0.2, 40.0, 70.0, 0.02, 100.0, 3.0))
    assert res.calculate_logp() == eph.calculate_logp()
    for tl in (res, eph):
        tl.restore()
        tl.has_dirt = 2
    assert res.calculate_logp() == eph.calculate_logp()
    with pytest.raises(Exception):
        eph.core.get_node_partials(tree.n)
    res.close()
    eph.close()


def test_streamed_mcmc_replay(core_mod, oracle, monkeypatch):
    from sieve_b200.mcmc import McmcReplay, gamma_category_rates
    monkeypatch.setenv("SIEVE_B200_CHUNK_SITES", "128")
    d = synthetic(20, 300, seed=71)
    counts, tree = d["counts"], d["tree"]
    c = core_mod.ScsCudaLikelihoodCore(20, 300, 4, 4, True, True, True, ephemeral=True)
    c.set_counts(counts)
    sf, _ = c.compute_size_factors(0)
    mc = McmcReplay(c, tree, mix="stage1", asc_mode="felsenstein_mean", n_background_sites=900.0, seed=3)
    st = mc.run(150)
    state = mc.state()
    rates = gamma_category_rates(state["gamma_shape"], 4)
    mats = substmodel.node_matrices(state["tree"], rates, branch_rates=state["branch_rates"])
    ref, _, _ = _oracle_eval(oracle, counts, state["tree"], sf, state["leaf"], mats, np.full(4, 0.25), "felsenstein_mean", 900.0)
    assert 0 < st["accepted"] < 150 and abs(state["log_p"] - ref) <= TOTAL_RTOL * abs(ref)
    mc.close()
    c.close()


@pytest.mark.parametrize("tag", ["cells10", "cells40"])
def test_background_likelihood_matches_oracle(core_mod, oracle, tag):
    g = json.load(open(os.path.join(GOLD, "golden_" + tag + ".json")))
    hists = [{int(k): v for k, v in h.items()} for h in g["background_hists"]]
    got = core_mod.background_log_likelihood(hists, int(g["M"]), g["n"], 0.008, 150.0)
    # the terms (up to ~1e9 each) cancel to ~1e5: the summation order alone moves the 12th digit; budget 1e-9
    assert abs(got - g["background_logP"]) <= 1e-10 * abs(got)
    for eps, w1 in ((0.02, 80.0), (0.001, 400.0)):
        ref = oracle.background_log_likelihood(hists, int(g["M"]), g["n"], eps, w1)
        got = core_mod.background_log_likelihood(hists, int(g["M"]), g["n"], eps, w1)
        assert abs(got - ref) <= 1e-10 * abs(ref)


def test_transient_nodes_do_not_change_results(core_mod, oracle, monkeypatch):
    """The planner leaves small subtrees unsaved (SV_OP_NOSTORE) and recomputes them on demand: same numbers as with
    every node saved, through proposals, rejections and partial read-backs."""
    from sieve_b200.treelikelihood import ScsTreeLikelihood
    d = synthetic(50, 300, seed=81)
    counts, tree = d["counts"], d["tree"]
    runs = {}
    for T in ("0", "3", "8", "64"):
        monkeypatch.setenv("SIEVE_B200_TRANSIENT_LEAVES", T)
        tl = ScsTreeLikelihood(counts, tree.copy(), asc_mode="felsenstein_mean", n_background_sites=4000.0)
        out = [tl.calculate_logp()]
        rng = np.random.default_rng(9)
        for it in range(10):
            t2 = tl.tree.copy()
            node = int(rng.integers(tree.n, tree.node_count - 1))
            lo = max(t2.height[t2.left[node]], t2.height[t2.right[node]])
            t2.height[node] = lo + (t2.height[t2.parent[node]] - lo) * rng.random()
            tl.store()
            tl.set_tree(t2, [node])
            out.append(tl.calculate_logp())
            if it % 3 == 1:
                tl.restore()
                out.append(tl.calculate_logp())
        parts = [tl.core.get_node_partials(v) for v in (tree.n, tree.n + 11, tree.n + 30, tree.root - 1)]
        cparts = [tl.core.get_node_partials(v, const=True) for v in (tree.n + 3, tree.root)]
        out.append(tl.calculate_logp())
        runs[T] = (out, tl.pattern_log_likelihoods()[0], parts, cparts)
        tl.close()
    for T in ("3", "8", "64"):
        assert runs[T][0] == runs["0"][0]
        np.testing.assert_array_equal(runs[T][1], runs["0"][1])
        for a, b in zip(runs[T][2] + runs[T][3], runs["0"][2] + runs["0"][3]):
            np.testing.assert_array_equal(a, b)


def test_change_tracking_skips_unchanged_nodes_bit_identically(core_mod, oracle, monkeypatch):
    """A caller that dirties more than changed (SIEVE under a relaxed clock with Felsenstein's correction dirties the
    whole tree for every branch-rate proposal, ScsTreeLikelihood.java:2814-2821): the core recomputes only what has new
    inputs.  Every number must equal the run with the short-cut off, through accepts, rejects, tree moves, leaf
    parameter moves and read-backs; and most of the dirty nodes must indeed have been skipped."""
    from sieve_b200.treelikelihood import ScsTreeLikelihood, LeafModelParams
    d = synthetic(60, 280, seed=101)
    counts, tree = d["counts"], d["tree"]
    runs, elided = {}, {}
    for T in ("8", "0", "8m"):
        monkeypatch.setenv("SIEVE_B200_TRANSIENT_LEAVES", T[0])
        for on in (True, False):
            rng = np.random.default_rng(17)
            br0 = rng.lognormal(0.0, 0.2, size=tree.node_count)
            # "8m": the plugin hands over 16-double matrices instead of distances (content compare of whole matrices)
            tl = ScsTreeLikelihood(counts, tree.copy(), asc_mode="felsenstein_mean", n_background_sites=4000.0,
                                   branch_rates=br0, use_distances=not T.endswith("m"))
            tl.core.set_elision(on)
            out = [tl.calculate_logp()]
            for it in range(24):
                tl.store()
                kind = it % 8
                if kind == 5:      # tree move: one node height
                    t2 = tl.tree.copy()
                    node = int(rng.integers(tree.n, tree.node_count - 1))
                    lo = max(t2.height[t2.left[node]], t2.height[t2.right[node]])
                    t2.height[node] = lo + (t2.height[t2.parent[node]] - lo) * rng.random()
                    tl.set_tree(t2, [node])
                elif kind == 7:    # leaf parameter move
                    p = tl.params.copy()
                    p.eps *= float(np.exp(0.1 * rng.normal()))
                    tl.set_leaf_params(p)
                else:              # branch-rate move: ONE rate changes, the whole tree is flagged
                    br = tl.branch_rates.copy()
                    br[int(rng.integers(0, tree.node_count - 1))] *= float(np.exp(0.3 * rng.normal()))
                    tl.set_branch_rates(br)
                out.append(tl.calculate_logp())
                if rng.random() < 0.6:
                    tl.restore()
                if it % 6 == 3:
                    out.append(tl.core.get_node_partials(tree.n + 5))
                    out.append(tl.core.get_node_partials(tree.root - 1, const=True))
            out.append(tl.pattern_log_likelihoods()[0].copy())
            out.append(tl.pattern_log_likelihoods()[1].copy())
            runs[(T, on)] = out
            elided[(T, on)] = tl.core.elided_count()
            tl.close()
    for T in ("8", "0", "8m"):
        for a, b in zip(runs[(T, True)], runs[(T, False)]):
            np.testing.assert_array_equal(a, b)
        assert elided[(T, False)] == 0
        # 18 branch-rate proposals x ~59 dirty internal nodes each; only the path to the root has new inputs
        assert elided[(T, True)] > 18 * 40, elided
    for a, b in zip(runs[("8", True)], runs[("0", True)]):
        np.testing.assert_array_equal(a, b)


def test_sparse_residency_is_bit_identical_to_dense(core_mod, oracle, monkeypatch):
    """SIEVE_FLAG_SPARSE with a pool far smaller than 2 * n_internal slots: only the larger subtrees keep their partials,
    small ones are recomputed through recycled temporary slots.  Same ops per site => the same bits as the dense core,
    through branch-rate, tree and leaf-parameter proposals, accepts, rejects and read-backs; the MCMC driver's
    incrementally maintained state must still equal a from-scratch evaluation."""
    from sieve_b200.treelikelihood import ScsTreeLikelihood
    from sieve_b200.mcmc import McmcReplay
    d = synthetic(150, 200, seed=111)
    counts, tree = d["counts"], d["tree"]
    runs = {}
    for tag, sparse, slots in (("dense", False, None), ("sparse", True, "110"), ("sparse_tight", True, "96"),
                               ("sparse_single_leaf", True, "110"), ("dense_single_leaf", False, None)):
        if slots:
            monkeypatch.setenv("SIEVE_B200_POOL_SLOTS", slots)
        rng = np.random.default_rng(23)
        br0 = rng.lognormal(0.0, 0.2, size=tree.node_count)
        tl = ScsTreeLikelihood(counts, tree.copy(), asc_mode="felsenstein_mean", n_background_sites=4000.0,
                               branch_rates=br0, sparse=sparse, single_leaf_buffer=tag.endswith("single_leaf"))
        out = [tl.calculate_logp()]
        for it in range(30):
            tl.store()
            kind = it % 10
            if kind in (2, 5, 8):      # tree move: one node height
                t2 = tl.tree.copy()
                node = int(rng.integers(tree.n, tree.node_count - 1))
                lo = max(t2.height[t2.left[node]], t2.height[t2.right[node]])
                t2.height[node] = lo + (t2.height[t2.parent[node]] - lo) * rng.random()
                tl.set_tree(t2, [node])
            elif kind == 7:            # leaf parameter move
                p = tl.params.copy()
                p.eps *= float(np.exp(0.1 * rng.normal()))
                tl.set_leaf_params(p)
            elif kind == 9:            # everything changes
                tl.set_gamma_shape(tl.gamma_shape * float(np.exp(0.2 * rng.normal())))
            else:                      # branch-rate move (whole tree flagged, one branch changed)
                br = tl.branch_rates.copy()
                br[int(rng.integers(0, tree.node_count - 1))] *= float(np.exp(0.3 * rng.normal()))
                tl.set_branch_rates(br)
            out.append(tl.calculate_logp())
            if rng.random() < 0.6:
                tl.restore()
            if it % 7 == 3:
                out.append(tl.core.get_node_partials(tree.n + 5))          # a small node: materialised on demand
                out.append(tl.core.get_node_partials(tree.root - 1))
                out.append(tl.core.get_node_partials(tree.root - 2, const=True))
        out.append(tl.pattern_log_likelihoods()[0].copy())
        out.append(tl.pattern_log_likelihoods()[1].copy())
        runs[tag] = out
        tl.close()
    for tag in ("sparse", "sparse_tight", "sparse_single_leaf", "dense_single_leaf"):
        for a, b in zip(runs[tag], runs["dense"]):
            np.testing.assert_array_equal(a, b)
    # the native MCMC driver on a sparse core (topology moves included)
    S, n = counts.shape[:2]
    monkeypatch.setenv("SIEVE_B200_POOL_SLOTS", "120")
    for mix, states in (("relaxed2", 300), ("stage1", 500)):
        finals = {}
        for sparse, elide, single in ((False, False, False), (False, True, False), (True, True, False), (True, True, True)):
            c = core_mod.ScsCudaLikelihoodCore(n, S, 4, 4, True, True, True, sparse=sparse, single_leaf_buffer=single)
            c.set_elision(elide)
            c.set_counts(counts)
            c.compute_size_factors(0)
            mc = McmcReplay(c, tree, mix=mix, asc_mode="felsenstein_mean", n_background_sites=5000.0, seed=5)
            st = mc.run(states)
            again = mc.full_recompute()
            assert abs(again - st["log_p"]) <= 1e-11 * abs(again)
            finals[(sparse, elide, single)] = (st["log_p"], st["accepted"])
            mc.close()
            c.close()
        # the chain is a deterministic function of the likelihood values: any difference anywhere changes its course
        assert len(set(finals.values())) == 1, (mix, finals)
    # misuse is loud
    c = core_mod.ScsCudaLikelihoodCore(n, S, 4, 4, True, False, True, sparse=True)
    with pytest.raises(RuntimeError):
        c.ml_trace()
    c.close()


def test_walk_kernels_are_bit_identical(core_mod, oracle, monkeypatch):
    """The lean default walk (k_walk, walk.cuh) and the general kernel (k_prune, SIEVE_B200_WALK=0) run the same op list with the
    same arithmetic, and the L2 prefetch annotations (SIEVE_B200_PF_DIST) only move data: every number must be identical --
    resident with proposals / rejections, and streamed with ragged chunks.  (The rejected cp.async / register / TMA pipelines
    this test used to cover live under profiles/experiments/r01_rejected_walks, outside the product library.)"""
    from sieve_b200.treelikelihood import ScsTreeLikelihood
    d = synthetic(41, 334, seed=83)
    counts, tree = d["counts"], d["tree"]
    runs = {}
    for P in ("default", "k_prune", "pf0", "pf7"):
        monkeypatch.setenv("SIEVE_B200_WALK", "0" if P == "k_prune" else "1")
        if P.startswith("pf"):
            monkeypatch.setenv("SIEVE_B200_PF_DIST", P[2:])
        else:
            monkeypatch.delenv("SIEVE_B200_PF_DIST", raising=False)
        out = []
        for eph in (False, True):
            if eph:
                monkeypatch.setenv("SIEVE_B200_CHUNK_SITES", "100")
            tl = ScsTreeLikelihood(counts, tree.copy(), asc_mode="felsenstein_mean", n_background_sites=4000.0,
                                   ephemeral=eph)
            out.append(tl.calculate_logp())
            rng = np.random.default_rng(9)
            for it in range(6):
                t2 = tl.tree.copy()
                node = int(rng.integers(tree.n, tree.node_count - 1))
                lo = max(t2.height[t2.left[node]], t2.height[t2.right[node]])
                t2.height[node] = lo + (t2.height[t2.parent[node]] - lo) * rng.random()
                tl.store()
                tl.set_tree(t2, [node])
                out.append(tl.calculate_logp())
                if it % 2 == 1:
                    tl.restore()
                    out.append(tl.calculate_logp())
            out.append(tl.pattern_log_likelihoods()[0].copy())
            tl.close()
            monkeypatch.delenv("SIEVE_B200_CHUNK_SITES", raising=False)
        runs[P] = out
    for P in ("k_prune", "pf0", "pf7"):
        for a, b in zip(runs["default"], runs[P]):
            np.testing.assert_array_equal(a, b, err_msg=P)


def test_algebraic_constant_site_path_equals_per_site_twin(core_mod, oracle, monkeypatch):
    """The constant-site partials factorise into (a site-independent 4-vector per node and category) x (product over the
    leaves below of their constant-genotype likelihood); the default path uses that.  SIEVE_B200_CONST_TWIN=1 prunes them
    per site like the reference: both must agree with each other and with the oracle, through proposals and rejections."""
    from sieve_b200.treelikelihood import ScsTreeLikelihood, LeafModelParams
    d = synthetic(37, 260, seed=91)
    counts, tree = d["counts"], d["tree"]
    M = 5000.0
    out = {}
    for twin in ("0", "1"):
        monkeypatch.setenv("SIEVE_B200_CONST_TWIN", twin)
        for asc in ("felsenstein_mean", "lewis"):
            tl = ScsTreeLikelihood(counts, tree.copy(), asc_mode=asc, n_background_sites=M)
            vals = [tl.calculate_logp()]
            consts = [tl.pattern_log_likelihoods()[1].copy()]
            rng = np.random.default_rng(4)
            for it in range(6):
                t2 = tl.tree.copy()
                node = int(rng.integers(tree.n, tree.node_count - 1))
                lo = max(t2.height[t2.left[node]], t2.height[t2.right[node]])
                t2.height[node] = lo + (t2.height[t2.parent[node]] - lo) * rng.random()
                tl.store()
                tl.set_tree(t2, [node])
                vals.append(tl.calculate_logp())
                if it % 2:
                    tl.restore()
                    # the per-site arrays are not double buffered (in the reference neither): re-derive the root first
                    tl.core.update_nodes(tl.tree.ops([tl.tree.root]), flip=False)
                    tl.core.evaluate()
                    consts.append(tl.pattern_log_likelihoods()[1].copy())
            tl.store()
            tl.set_leaf_params(LeafModelParams(0.25, 45.0, 60.0, 0.012, 130.0, 2.4))
            vals.append(tl.calculate_logp())
            consts.append(tl.pattern_log_likelihoods()[1].copy())
            tl.restore()
            tl.has_dirt = 2
            vals.append(tl.calculate_logp())
            cp = tl.core.get_node_partials(tree.n + 9, const=True)
            out[(twin, asc)] = (vals, consts, cp)
            tl.close()
    for asc in ("felsenstein_mean", "lewis"):
        a, b = out[("0", asc)], out[("1", asc)]
        np.testing.assert_allclose(a[0], b[0], rtol=1e-12)
        for x, y in zip(a[1], b[1]):
            np.testing.assert_allclose(x, y, rtol=1e-13)
        # entries that are O(d^2) in a short branch carry a relative error of ~eps/d in either path
        np.testing.assert_allclose(a[2], b[2], rtol=1e-11, atol=1e-10)


def test_ingest_shard_loads_like_set_counts(core_mod, tmp_path):
    from sieve_b200 import ingest
    d = synthetic(9, 130, seed=33)
    counts = d["counts"]
    outs = []
    for mode in ("host_site_major", "shard_u16", "shard_i32"):
        c = core_mod.ScsCudaLikelihoodCore(9, 130, 4, 4, True, False, True)
        if mode == "host_site_major":
            c.set_counts(counts)
        else:
            p = str(tmp_path / (mode + ".sieve"))
            ingest.write_shard(p, counts, 0, 130, dtype="uint16" if mode == "shard_u16" else "int32")
            meta, arr = ingest.open_shard(p)
            c.set_counts_cellmajor(np.ascontiguousarray(arr))
        sf, st = c.compute_size_factors(0)
        c.set_leaf_params(*PARAMS)
        c.update_leaves(for_update=False)
        outs.append((sf, st, c.get_node_partials(3)))
        c.close()
    for o in outs[1:]:
        np.testing.assert_array_equal(o[0], outs[0][0])
        np.testing.assert_array_equal(o[2], outs[0][2])


@pytest.mark.parametrize("n", [1, 2, 3])
def test_smallest_trees(core_mod, oracle, n):
    """n = 1: the unary root sits directly on the only leaf; n = 2: one cherry under the trunk."""
    rng = np.random.default_rng(n)
    tree = caterpillar(n, 0.0, 0.07)
    S = 45
    counts = np.zeros((S, n, 4), dtype=np.int32)
    cov = rng.integers(0, 70, size=(S, n))
    counts[..., 0] = (cov * rng.random((S, n)) * (rng.random((S, n)) < 0.5)).astype(int)
    counts[..., 3] = cov
    counts[3] = 0                                    # a site nobody covers
    c5 = oracle.counts5(counts)
    sf = np.ones(n) if n == 1 else oracle.size_factors(c5)[0]
    rates, props = substmodel.gamma_category_rates(1.3, 4)
    mats = substmodel.node_matrices(tree, rates)
    for mode in ("distances", "matrices", "sparse"):
        c = core_mod.ScsCudaLikelihoodCore(n, S, 4, 4, True, True, True, sparse=mode == "sparse",
                                           single_leaf_buffer=mode == "sparse")
        c.set_counts(counts)
        c.set_size_factors(sf)
        c.set_leaf_params(*PARAMS)
        c.update_leaves(for_update=False)
        c.set_root(tree.root, props, 0)
        nodes = [i for i in range(tree.node_count) if i != tree.root]
        if mode != "matrices":
            c.set_distances(nodes, substmodel.node_distances(tree, rates)[nodes], for_update=False)
        else:
            c.set_matrices(nodes, mats[nodes], for_update=False)
        c.update_nodes(tree.ops(), flip=False)
        r = c.evaluate()
        logp, osl, osc = _oracle_eval(oracle, counts, tree, sf, PARAMS, mats, props, "felsenstein_mean", 100.0)
        sl, sc = c.site_log_likelihoods()
        _close(sl, osl, SITE_RTOL)
        _close(sc, osc, SITE_RTOL)
        assert abs(core_mod.combine_shards([r], "felsenstein_mean", 100.0) - logp) <= TOTAL_RTOL * abs(logp)
        c.close()


def test_native_comm_single_rank(core_mod):
    """A one-rank NCCL communicator bound to the handle: every evaluation all-gathers the result struct on the handle's
    stream (the N > 1 path is the same call with more ranks; covered at N = 2 by bench.py under torchrun)."""
    from sieve_b200.mcmc import McmcReplay
    d = synthetic(12, 200, seed=13)
    counts, tree = d["counts"], d["tree"]
    c = core_mod.ScsCudaLikelihoodCore(12, 200, 4, 4, True, True, True)
    c.set_counts(counts)
    c.compute_size_factors(0)
    c.comm_init(0, 1, core_mod.comm_unique_id())
    assert c.comm_size() == 1
    mc = McmcReplay(c, tree, mix="stage2", asc_mode="felsenstein_mean", n_background_sites=500.0, seed=2)
    st = mc.run(60)
    g = c.last_global_results()
    assert len(g) == 1 and np.isfinite(g[0].sum_log_l) and g[0].weight_sum == 200.0
    again = mc.full_recompute()
    assert abs(again - st["log_p"]) <= 1e-11 * abs(again)
    mc.close()
    c.close()


def test_streamed_ragged_chunks_are_deterministic(core_mod, monkeypatch):
    """Regression: with reused scratch slots, lanes beyond the last site of a chunk must not store (a duplicate lane in
    another warp could overwrite a slot the real lane had not consumed yet).  Ragged chunk tails, deep tree, repeats."""
    monkeypatch.setenv("SIEVE_B200_CHUNK_SITES", "1000")     # rounded to 1024: chunks 1024 + 1024 + 953 (tail 953 = 59 * 16 + 9)
    rng = np.random.default_rng(3)
    n, S = 700, 3001
    tree = random_coalescent(n, rng)
    tree.height *= 0.3 / tree.height[tree.root]
    counts = np.zeros((S, n, 4), dtype=np.int32)
    cov = rng.poisson(40, size=(S, n))
    counts[..., 0] = (cov * rng.random((S, n)) * (rng.random((S, n)) < 0.2)).astype(np.int32)
    counts[..., 3] = cov
    rates, props = substmodel.gamma_category_rates(1.0, 4)
    nodes = np.array([i for i in range(tree.node_count) if i != tree.root], dtype=np.int32)
    D = substmodel.node_distances(tree, rates)[nodes]
    out = {}
    for eph in (False, True):
        c = core_mod.ScsCudaLikelihoodCore(n, S, 4, 4, True, True, True, ephemeral=eph)
        c.set_counts(counts)
        c.compute_size_factors(0)
        c.set_leaf_params(*PARAMS)
        c.update_leaves(False)
        c.set_root(tree.root, props, 0)
        vals = []
        for _ in range(5):
            r = c.step_distances(nodes, D, tree.ops(), update_leaves=True)
            vals.append((r.sum_log_l, r.const_lse))
        assert all(v == vals[0] for v in vals), vals
        out[eph] = (vals[0], c.site_log_likelihoods()[0])
        c.close()
    assert out[True][0] == out[False][0]
    np.testing.assert_array_equal(out[True][1], out[False][1])


def test_lazy_renormalisation_is_bit_identical_to_renormalising_every_op(core_mod, oracle, monkeypatch):
    """The walk renormalises a partial only where the host's magnitude bounds ask for it (core_planner.cuh, sv_bound_op).
    Scaling by 2^k is exact and exponents are summed as integers, so the schedule cannot change a bit: site
    log-likelihoods, totals and exported node partials equal those of a core that renormalises after every op
    (SIEVE_B200_RENORM_ALWAYS=1) -- on a deep caterpillar (the bounds run out every dozen ops), on a coalescent tree,
    through a leaf-parameter proposal, a local move and a restore, resident and streamed."""
    from sieve_b200.treelikelihood import LeafModelParams, ScsTreeLikelihood
    for shape in ("caterpillar", "coalescent"):
        if shape == "caterpillar":
            n, S = 260, 700
            tree = caterpillar(n, 0.0, 1.0)
            tree.height *= 0.6 / tree.height[tree.root]
            d = synthetic(n, S, seed=61, tree=tree)
        else:
            n, S = 180, 900
            d = synthetic(n, S, seed=62)
            tree = d["tree"]
        counts = d["counts"]
        sf, _ = oracle.size_factors(oracle.counts5(counts))
        v = tree.n + tree.n // 2
        moved = tree.copy()
        lo = max(moved.height[moved.left[v]], moved.height[moved.right[v]])
        moved.height[v] = lo + 0.41 * (moved.height[moved.parent[v]] - lo)
        out = {}
        for mode in ("lazy", "always", "lazy_streamed"):
            if mode == "always":
                monkeypatch.setenv("SIEVE_B200_RENORM_ALWAYS", "1")
            else:
                monkeypatch.delenv("SIEVE_B200_RENORM_ALWAYS", raising=False)
            tl = ScsTreeLikelihood(counts, tree.copy(), size_factors=sf, asc_mode="felsenstein_mean", n_background_sites=1e4,
                                   ephemeral=(mode == "lazy_streamed"))
            seq = [tl.calculate_logp()]
            seq.append(tl.pattern_log_likelihoods()[0].copy())
            tl.store()
            tl.set_leaf_params(LeafModelParams(0.25, 44.0, 61.0, 0.011, 120.0, 2.5))
            seq.append(tl.calculate_logp())
            tl.restore()
            tl.store()
            tl.set_tree(moved.copy(), [v])
            seq.append(tl.calculate_logp())
            seq.append(tl.pattern_log_likelihoods()[0].copy())
            if mode != "lazy_streamed":
                seq.append(tl.core.get_node_partials(v))
                seq.append(tl.core.get_node_partials(tree.root - 1))
            out[mode] = seq
            tl.close()
        for a, b in zip(out["lazy"], out["always"]):
            np.testing.assert_array_equal(a, b)
        for a, b in zip(out["lazy_streamed"], out["lazy"][:5]):
            np.testing.assert_array_equal(a, b)
        assert np.isfinite(out["lazy"][1]).all()
        # and the numbers are right: against the oracle on the start state
        rates, props = substmodel.gamma_category_rates(1.0, 4)
        mats = substmodel.node_matrices(tree, rates)
        logp, osl, _ = _oracle_eval(oracle, counts, tree, sf, PARAMS, mats, props, "felsenstein_mean", 1e4, T=8)
        _close(out["lazy"][1], osl, SITE_RTOL)
        assert abs(out["lazy"][0] - logp) <= TOTAL_RTOL * abs(logp)
