"""Parity at the shapes BASELINE.json names (cells of configs[2] and configs[3]/[4]), not only on toy trees.

  (i)  1 000 cells x 4 096 sites, dense resident core: bias correction on / off, branch distances and 16-double matrices;
  (ii) 10 000 cells x 1 024 sites: sparse residency, streamed evaluation, sparse + single leaf buffer.

Each case walks the sequence an MCMC chain produces (likelihood/ScsTreeLikelihood.java:2561-2616, 2827-2880): full
evaluation -> store -> leaf-parameter proposal (configs[4]: every read-count likelihood recomputed) -> reject (restore)
-> a local tree move (ScsUniform-like: one internal node's height) -> evaluation.  Every evaluation is compared with the
CPU oracle on the same inputs: per-site log-likelihood and logConstRoot <= 1e-12 relative, total <= 1e-9 relative
(the tolerances of BASELINE.json's north_star).  The oracle runs with the reference's thread split over the host cores.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from sieve_b200 import substmodel  # noqa: E402
from sieve_b200.data import synthetic  # noqa: E402
from sieve_b200.tree import random_coalescent  # noqa: E402
from sieve_b200.treelikelihood import LeafModelParams, ScsTreeLikelihood  # noqa: E402

SITE_RTOL = 1e-12
TOTAL_RTOL = 1e-9
M_BACKGROUND = 2.0e5
PARAMS = LeafModelParams()                                              # template start values
PARAMS_PROPOSED = LeafModelParams(0.27, 46.5, 58.0, 0.0093, 141.0, 2.3)   # all three sub-models move


class _Case:
    """Data + the oracle's answers for the three states the sequence visits (computed once per shape, on demand)."""

    def __init__(self, oracle, n, S, seed):
        rng = np.random.default_rng(seed)
        tree = random_coalescent(n, rng)
        tree.height *= 0.3 / tree.height[tree.root]
        d = synthetic(n, S, seed=seed + 1, tree=tree)
        self.oracle, self.n, self.S = oracle, n, S
        self.counts, self.tree = d["counts"], tree
        self.c5 = oracle.counts5(self.counts)
        self.sf, _ = oracle.size_factors(self.c5)
        self.rates, self.props = substmodel.gamma_category_rates(1.0, 4)
        # the local move: an internal node with an internal parent slides between its taller child and its parent
        v = next(x for x in range(n + n // 2, 2 * n - 2) if tree.parent[x] != tree.root)
        lo = max(tree.height[tree.left[v]], tree.height[tree.right[v]])
        hi = tree.height[tree.parent[v]]
        self.moved = tree.copy()
        self.moved.height[v] = lo + 0.37 * (hi - lo)
        self.moved_node = v
        self._cache = {}

    def expect(self, state, asc):
        """state: 'start' | 'proposed' (leaf parameters) | 'moved' (tree move at the start parameters)."""
        key = (state, asc)
        if key not in self._cache:
            so = self.oracle
            tree = self.moved if state == "moved" else self.tree
            p = PARAMS_PROPOSED if state == "proposed" else PARAMS
            mats = substmodel.node_matrices(tree, self.rates)
            self._cache[key] = so.full_eval(self.c5, self.sf, so.leaf_params(*p.as_tuple()), tree.ops(), mats, self.props,
                                            tree.root, asc, M_BACKGROUND, so.max_threads())
        return self._cache[key]


@pytest.fixture(scope="module")
def case_1000(oracle):
    return _Case(oracle, 1000, 4096, seed=4101)


@pytest.fixture(scope="module")
def case_10000(oracle):
    return _Case(oracle, 10000, 1024, seed=4202)


def _check(tl, case, state, asc):
    logp = tl.calculate_logp()
    ref, ref_site, ref_const = case.expect(state, asc)
    site, const = tl.pattern_log_likelihoods()
    assert np.isfinite(site).all()
    err = np.max(np.abs(site - ref_site) / np.abs(ref_site))
    assert err <= SITE_RTOL, (state, "site logL", err)
    if asc != "none":
        cerr = np.max(np.abs(const - ref_const) / np.abs(ref_const))
        assert cerr <= SITE_RTOL, (state, "logConstRoot", cerr)
    assert abs(logp - ref) <= TOTAL_RTOL * abs(ref), (state, logp, ref)
    return logp


def _mcmc_sequence(case, asc, **kw):
    tl = ScsTreeLikelihood(case.counts, case.tree.copy(), size_factors=case.sf, leaf_params=PARAMS, asc_mode=asc,
                           n_background_sites=M_BACKGROUND, **kw)
    try:
        first = _check(tl, case, "start", asc)
        tl.store()
        tl.set_leaf_params(PARAMS_PROPOSED)            # configs[4]: all leaves + the whole tree
        proposed = _check(tl, case, "proposed", asc)
        assert proposed != first
        tl.restore()                                   # rejected
        tl.store()
        tl.set_tree(case.moved.copy(), [case.moved_node])
        moved = _check(tl, case, "moved", asc)         # only the path to the root is recomputed; the leaves are the old ones
        assert moved != first
        tl.restore()
        # back at the start state: nothing is dirty, a forced full re-evaluation reproduces the first number bit for bit
        tl.has_dirt = 1
        assert tl.calculate_logp() == first
    finally:
        tl.close()


@pytest.mark.parametrize("asc,use_distances", [("felsenstein_mean", True), ("felsenstein_mean", False), ("none", True),
                                               ("none", False)])
def test_1000_cells_dense(case_1000, asc, use_distances):
    _mcmc_sequence(case_1000, asc, use_distances=use_distances)


@pytest.mark.parametrize("mode", ["sparse", "streamed", "sparse_single_leaf"])
def test_10000_cells(case_10000, mode, monkeypatch):
    kw = {}
    if mode == "streamed":
        monkeypatch.setenv("SIEVE_B200_CHUNK_SITES", "416")      # 1 024 sites = two whole chunks and a ragged one
        kw["ephemeral"] = True
    else:
        # fewer slots than (buffer, node) pairs, so that small subtrees really are recomputed on the way
        monkeypatch.setenv("SIEVE_B200_POOL_SLOTS", "2600")
        kw["sparse"] = True
        kw["single_leaf_buffer"] = mode == "sparse_single_leaf"
    _mcmc_sequence(case_10000, "felsenstein_mean", **kw)
    if mode == "sparse":
        # the pool really was smaller than a dense core's
        tl = ScsTreeLikelihood(case_10000.counts[:64], case_10000.tree.copy(), size_factors=case_10000.sf,
                               asc_mode="felsenstein_mean", n_background_sites=M_BACKGROUND, sparse=True)
        tl.calculate_logp()
        info = tl.core.sparse_info()
        tl.close()
        assert info["pool_slots"] == 2600 and info["keep_leaves"] >= 8
