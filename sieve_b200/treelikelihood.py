"""Host-side driver with the call sequence of the reference's `ScsTreeLikelihood` / `ThreadedScsTreeLikelihood`.

In production this logic stays in the Java plugin (BEAST owns the tree, the operators and the dirty flags); the classes
here replay the same sequence against the C-ABI so that tests and the benchmark exercise the device path the way the
plugin would:  requiresRecalculation (likelihood/ScsTreeLikelihood.java:2786-2824) -> traverse (:545-935) ->
calcLogP (:2621-2651), store / restore (:2827-2880), and the site split + combine of
likelihood/ThreadedScsTreeLikelihood.java:95-269, 366-389.
"""
import numpy as np

from . import substmodel
from .core import ScsCudaLikelihoodCore, combine_shards, pattern_points

IS_CLEAN, IS_DIRTY, IS_FILTHY = 0, 1, 2


class LeafModelParams:
    """The six parameters of RawReadCountsModelFiniteMuDM with the template start values
    (examples/templates/smc_stage_1.xml:30-38)."""

    def __init__(self, theta=0.3, t=50.0, v=50.0, eps=0.008, w1=150.0, w2=2.0):
        self.theta, self.t, self.v, self.eps, self.w1, self.w2 = theta, t, v, eps, w1, w2

    def as_tuple(self):
        return (self.theta, self.t, self.v, self.eps, self.w1, self.w2)

    def copy(self):
        return LeafModelParams(*self.as_tuple())


class ScsTreeLikelihood:
    """One shard: one core on one device."""

    def __init__(self, counts, tree, size_factors=None, leaf_params=None, gamma_shape=1.0, n_categories=4,
                 clock_rate=1.0, branch_rates=None, use_log_partials=True, asc_mode="none", n_background_sites=0.0,
                 weights=None, single_ado=True, device=-1, return_const_sum=False, use_distances=True, ephemeral=False,
                 sparse=False, single_leaf_buffer=False, private_seq_cov=False, zero_cov_mode=0):
        counts = np.asarray(counts)
        self.S, self.n = counts.shape[0], counts.shape[1]
        assert tree.n == self.n, "The number of nodes in the tree does not match the number of sequences"
        self.tree = tree
        self.K = n_categories
        self.gamma_shape = gamma_shape
        self.clock_rate = clock_rate
        self.branch_rates = None if branch_rates is None else np.asarray(branch_rates, dtype=np.float64).copy()
        self.params = leaf_params.copy() if leaf_params is not None else LeafModelParams()
        self.asc_mode = asc_mode
        self.M = float(n_background_sites)
        self.return_const_sum = return_const_sum
        # True: hand the core branch distances (fast path for the fixed rate matrix); False: 16-double matrices
        self.use_distances = use_distances
        # PrivateAllelicSeqCovModel instead of ExploredSharedAllelicSeqCovModel: (t, v) per (category, pattern)
        self.private_seq_cov, self.zero_cov_mode, self.single_ado = bool(private_seq_cov), int(zero_cov_mode), bool(single_ado)
        # accept(): re-evaluate only the patterns whose (t, v) moved (as the reference does); False: a whole-shard pass
        self.retraverse_changed_only = bool(private_seq_cov) and not (ephemeral or sparse)
        self.last_retraversed_sites = None
        self._counts = counts          # a reference (tied pairs / tied sites are re-examined on the host side)
        self._ctor = dict(n_categories=n_categories, use_log_partials=use_log_partials, single_ado=single_ado, device=device,
                          use_distances=use_distances)
        self.core = ScsCudaLikelihoodCore(self.n, self.S, n_categories, 4, use_log_partials, asc_mode != "none",
                                          single_ado, device, ephemeral, sparse, single_leaf_buffer,
                                          private_seq_cov=private_seq_cov)
        self.core.set_counts(counts)
        self.core.set_pattern_weights(weights)
        self.branch_lengths = np.full(tree.node_count, -1.0)   # m_branchLengths
        self.size_factors = None
        if isinstance(size_factors, str) and size_factors == "defer":
            pass        # a shard of a larger alignment: the global size factors come later (finish_init)
        elif size_factors is None:
            size_factors, self.cov_stats = self.core.compute_size_factors(0)
            self.finish_init(size_factors, already_set=True)
        else:
            self.finish_init(size_factors)
        self.stored_branch_lengths = self.branch_lengths.copy()
        self.has_dirt = IS_FILTHY
        self.update_leaves = False
        self.dirty_nodes = set()
        self.logP = None
        self.result = None
        self._refresh_site_model()

    def finish_init(self, size_factors, already_set=False):
        """initCore (:430-468) once the size factors are known: leaves into both buffers, everything filthy."""
        if not already_set:
            self.core.set_size_factors(size_factors)
        self.size_factors = np.asarray(size_factors)
        self.core.set_leaf_params(*self.params.as_tuple())
        if self.private_seq_cov:
            # every pair starts at the shared values (PrivateAllelicSeqCovModel.java:159-161)
            self.core.set_private_seq_cov(np.full((self.K, self.S), self.params.t), np.full((self.K, self.S), self.params.v))
        self.core.update_leaves(for_update=False)

    def accept(self):
        """ScsTreeLikelihood.accept -> postProcess (:2454-2520, 2893-2916) for the per-pattern coverage model: max-sum trace
        of the accepted state, re-estimation of every (category, pattern) pair's (t, v) from the ML numbers of sequenced
        alleles, then the state is recomputed with the new pairs.  -> number of pairs that moved."""
        if not self.private_seq_cov:
            return 0
        from .private_cov import resolve_tied_pairs
        changed, tied = self.core.private_update_from_ml(self.zero_cov_mode)
        if tied.any():
            changed += resolve_tied_pairs(self.core, self.tree, self._counts, self.size_factors, tied, self.zero_cov_mode,
                                          self.single_ado)
        if changed:
            if self.retraverse_changed_only:
                # step 3 (:2493-2498): the whole tree for the changed patterns only, in place; the accepted state is the
                # stored state afterwards
                self.last_retraversed_sites = self.core.private_retraverse_changed(self.tree.ops())
                self.result = self.core.evaluate()
                if not self.return_const_sum:
                    self.logP = combine_shards([self.result], self.asc_mode, self.M)
                self.store()
            else:
                # a whole-shard pass: unchanged pairs reproduce their values bit for bit
                self.update_leaves = True
                self.has_dirt = max(self.has_dirt, IS_DIRTY)
                self.calculate_logp()
        return changed

    def call_variants(self):
        """VariantCaller's core step on the current state (ScsTreeLikelihood.callVariants, :2444-2449): ML genotypes of every
        node, dropout states of the leaves, ML rate category.  On a sparse core the tied sites (rare) are re-evaluated on a
        small dense core over just those sites, which has the per-site read-backs the tie-break needs."""
        from . import variants
        made = []

        def tied_core(sites):
            sub = ScsTreeLikelihood(self._counts[sites], self.tree.copy(), size_factors=self.size_factors, leaf_params=self.params,
                                    gamma_shape=self.gamma_shape, clock_rate=self.clock_rate, branch_rates=self.branch_rates,
                                    asc_mode=self.asc_mode, n_background_sites=self.M, **self._ctor)
            sub.calculate_logp()
            made.append(sub)
            return sub.core

        out = variants.call_variants(self.core, self.tree, substmodel.ROOT_GENOTYPE, tied_core)
        for sub in made:
            sub.close()
        return out

    def close(self):
        self.core.close()

    # ---- what BEAST's state nodes would tell requiresRecalculation (:2786-2824) ----
    def _refresh_site_model(self):
        self.rates, self.proportions = substmodel.gamma_category_rates(self.gamma_shape, self.K)
        self.core.set_root(self.tree.root, self.proportions, substmodel.ROOT_GENOTYPE)

    def set_leaf_params(self, params):
        self.params = params.copy()
        self.has_dirt = max(self.has_dirt, IS_DIRTY)
        self.update_leaves = True

    def set_gamma_shape(self, shape):
        self.gamma_shape = shape
        self._refresh_site_model()
        self.has_dirt = max(self.has_dirt, IS_DIRTY)

    def set_branch_rates(self, branch_rates=None, clock_rate=None):
        if branch_rates is not None:
            self.branch_rates = np.asarray(branch_rates, dtype=np.float64).copy()
        if clock_rate is not None:
            self.clock_rate = clock_rate
        # isForcingTreeDirtyRMC (:2814-2821): with Felsenstein's correction a rate change dirties the whole tree
        if self.asc_mode.startswith("felsenstein"):
            self.has_dirt = max(self.has_dirt, IS_DIRTY)

    def set_tree(self, tree, dirty_nodes):
        """A tree operator moved: `dirty_nodes` are the nodes it flagged (Node.makeDirty)."""
        self.tree = tree
        self.dirty_nodes |= set(int(x) for x in dirty_nodes)

    # ---- calculateLogP (:2561-2616) ----
    def calculate_logp(self):
        tree = self.tree
        n = self.n
        if self.update_leaves:
            self.core.set_leaf_params(*self.params.as_tuple())
        bl = tree.branch_lengths()
        br = np.full(tree.node_count, self.clock_rate) if self.branch_rates is None else self.clock_rate * self.branch_rates
        branch_time = bl * br
        all_dirty = self.has_dirt != IS_CLEAN
        # traverse, first block (:563-580): which matrices are refreshed
        mat_nodes = [i for i in range(tree.node_count) if i != tree.root and
                     (all_dirty or i in self.dirty_nodes or branch_time[i] != self.branch_lengths[i])]
        for node in mat_nodes:
            self.branch_lengths[node] = branch_time[node]
        if self.use_distances:
            D = branch_time[mat_nodes][:, None] * self.rates[None, :] if len(mat_nodes) else np.zeros((0, self.K))
        else:
            P = np.empty((len(mat_nodes), self.K, 16))
            for j, node in enumerate(mat_nodes):
                for k in range(self.K):
                    P[j, k] = substmodel.transition_probabilities(branch_time[node] * self.rates[k]).reshape(16)
        # second block (:615-628): a node is recomputed when something below it changed
        if all_dirty or self.update_leaves:
            ops = tree.ops()
        else:
            # a node whose matrix changed (or that an operator flagged) makes its PARENT recompute (:621-628)
            touched = set(mat_nodes) | self.dirty_nodes
            ops = tree.ops(tree.dirty_closure([tree.parent[x] for x in touched if tree.parent[x] >= 0]))
        if len(ops) == 0 and not self.update_leaves:
            return self.result if self.return_const_sum else self.logP
        if self.use_distances:
            self.result = self.core.step_distances(mat_nodes, D, ops, update_leaves=self.update_leaves)
        else:
            self.result = self.core.step(mat_nodes, P, ops, update_leaves=self.update_leaves)
        self.has_dirt = IS_CLEAN
        self.update_leaves = False
        self.dirty_nodes = set()
        if self.return_const_sum:
            return self.result
        self.logP = combine_shards([self.result], self.asc_mode, self.M)
        return self.logP

    def pattern_log_likelihoods(self):
        return self.core.site_log_likelihoods()

    # ---- store / restore (:2827-2880) ----
    def store(self):
        self.core.store()
        self.stored_branch_lengths = self.branch_lengths.copy()
        self._stored = (self.tree.copy(), self.params.copy(), self.gamma_shape, self.clock_rate,
                        None if self.branch_rates is None else self.branch_rates.copy(), self.logP)

    def restore(self):
        self.core.restore()
        self.branch_lengths, self.stored_branch_lengths = self.stored_branch_lengths, self.branch_lengths
        self.tree, self.params, self.gamma_shape, self.clock_rate, self.branch_rates, self.logP = self._stored
        self._refresh_site_model()
        self.core.set_leaf_params(*self.params.as_tuple())
        self.has_dirt = IS_CLEAN
        self.update_leaves = False
        self.dirty_nodes = set()


class ThreadedScsTreeLikelihood:
    """Site-sharded evaluation: contiguous site ranges (calcPatternPoints, :227-238), one core per range, global size
    factors computed before the split (:119-131), per-shard (logP, constSum) combined in shard order (:371-379)."""

    def __init__(self, counts, tree, n_shards, devices=None, asc_mode="none", n_background_sites=0.0, **kw):
        counts = np.asarray(counts)
        S, n = counts.shape[:2]
        self.points = pattern_points(S, n_shards)
        self.asc_mode, self.M = asc_mode, float(n_background_sites)
        sf = kw.pop("size_factors", None)
        self.shards = []
        for r in range(n_shards):
            dev = -1 if devices is None else devices[r % len(devices)]
            self.shards.append(ScsTreeLikelihood(counts[self.points[r]:self.points[r + 1]], tree.copy(),
                                                 size_factors="defer" if sf is None else sf,
                                                 asc_mode=asc_mode, n_background_sites=n_background_sites, device=dev,
                                                 return_const_sum=True, **kw))
        if sf is None:
            # deeplyInitialize on the full alignment (:119-131): the size factors are global.  No device holds the full
            # alignment here, so the per-cell medians are found by exact selection over the shards' own count tensors
            # (sieve_sf_*): bit-identical to one handle over all sites, and nothing but the shards is ever allocated.
            from .sizefactors import SizeFactorPart, solve_parts
            parts = [SizeFactorPart(core=s.core) for s in self.shards]
            sf, self.cov_stats = solve_parts(parts)
            for p in parts:
                p.close()
            for s in self.shards:
                s.finish_init(sf)
        self.size_factors = sf

    def close(self):
        for s in self.shards:
            s.close()

    def calculate_logp(self):
        res = [s.calculate_logp() for s in self.shards]
        self.results = res
        self.logP = combine_shards(res, self.asc_mode, self.M)
        return self.logP

    def pattern_log_likelihoods(self):
        parts = [s.pattern_log_likelihoods() for s in self.shards]
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])

    def __getattr__(self, name):
        # broadcast of the state-changing calls to every shard
        if name in ("set_leaf_params", "set_gamma_shape", "set_branch_rates", "store", "restore"):
            def f(*a, **k):
                for s in self.shards:
                    getattr(s, name)(*a, **k)
            return f
        if name == "set_tree":
            def g(tree, dirty):
                for s in self.shards:
                    s.set_tree(tree.copy(), dirty)
            return g
        raise AttributeError(name)
