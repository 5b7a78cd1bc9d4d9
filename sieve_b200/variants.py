"""Variant calling on top of the device's max-sum pass: the host half of
`ScsTreeLikelihood.collectMLGenotypesAndLogLikelihoods` (likelihood/ScsTreeLikelihood.java:2013-2281).

The device returns, per site, the ML rate category, the unique ML genotype of every node and the unique dropout state of
every leaf (`ScsCudaLikelihoodCore.ml_call`).  Sites where several combinations are tied (rare: the reference prints a
warning for them, :2262-2266) are resolved here the reference's way: enumerate the combinations that follow the arg-max
lists from the root genotype, and keep those with the largest logSumExp over the nodes of the node's log-likelihood at
its genotype (`getMLGenotypesLogLikelihood`, :2388-2413).

A combination is the reference's int array: entry [leaf] = number of dropped alleles of that leaf's ML component,
entry [n_leaves + node] = genotype of `node` (:140-146).
"""
import itertools

import numpy as np


def _bits(mask):
    return [i for i in range(8) if (int(mask) >> i) & 1]


def _log_sum_exp(v):
    """math/util/MathFunctions.java:101-112"""
    m = max(v)
    return m + float(np.log(sum(np.exp(x - m) for x in v)))


def enumerate_combinations(children, root, n_leaves, back, root_genotype=0, limit=1 << 16):
    """back: uint8 [node][4] of one (category, site) as `ml_site` returns it.  -> sorted list of tuples."""
    n = int(n_leaves)

    def below(node, g):
        if node < n:
            return [{node: a, n + node: g} for a in _bits(back[node][g])]
        c1, c2 = children[node]
        b = int(back[node][g])
        opts = [d for cg in _bits(b & 15) for d in below(c1, cg)]
        if c2 >= 0:
            opts2 = [d for cg in _bits(b >> 4) for d in below(c2, cg)]
            if len(opts) * len(opts2) > limit:
                raise OverflowError("too many tied ML genotype combinations at one site")
            opts = [{**x, **y} for x, y in itertools.product(opts, opts2)]
        return [{**d, n + node: g} for d in opts]

    out = []
    for d in below(root, root_genotype):
        arr = [-1] * (3 * n)
        for k, v in d.items():
            arr[k] = v
        out.append(tuple(arr))
    return sorted(out)


class VariantCalls:
    """genotypes int8 [node][S], ado int8 [leaf][S], category uint8 [S]; `tied` maps the sites with several equally
    good results to the list of (category, combination) the reference would report for them."""

    def __init__(self, genotypes, ado, category, tied):
        self.genotypes, self.ado, self.category, self.tied = genotypes, ado, category, tied


def call_variants(core, tree, root_genotype=0, tied_core=None):
    """Run the max-sum pass on `core`'s current state and resolve tied sites.

    A sparse core traces in site chunks and keeps only the calls; the per-site read-backs that the tie-break needs
    (back-pointers and sum-product partials of EVERY node) exist on dense cores only.  `tied_core(sites)` must then return a
    dense core over exactly those sites in the same state (ScsTreeLikelihood.call_variants builds one); without it the tied
    sites of a sparse core stay unresolved (-1 entries, listed in `.tied` with an empty candidate list)."""
    core.ml_trace()
    g, a, cat, amb = core.ml_call()
    n = core.n
    children = {p: (int(tree.left[p]), int(tree.right[p])) for p in range(n, 2 * n)}
    tied = {}
    if amb.any():
        sites = [int(s) for s in np.nonzero(amb)[0]]
        src, index = core, {s: s for s in sites}
        try:
            core.ml_site(sites[0])
            chunked = False
        except RuntimeError:      # a core that traces in site chunks has no per-site read-back
            chunked = True
        if chunked:
            if tied_core is None:
                return VariantCalls(g, a, cat, {s: [] for s in sites})
            src = tied_core(sites)
            src.ml_trace()
            index = {s: i for i, s in enumerate(sites)}
        cats = src.ml_categories()
        for s in sites:
            back, logp = src.ml_site(index[s])
            cand = []
            for k in _bits(cats[index[s]]):
                for comb in enumerate_combinations(children, tree.root, n, back[:, k, :], root_genotype):
                    cand.append((k, comb))
            if len(cand) > 1:
                best, keep = -np.finfo(float).max, []
                for k, comb in cand:
                    v = _log_sum_exp([float(logp[node, k, comb[n + node]]) for node in range(2 * n)])
                    if best < v:
                        best, keep = v, [(k, comb)]
                    elif best == v:
                        keep.append((k, comb))
                cand = keep
            k0, c0 = cand[0]
            if len(cand) == 1:
                cat[s] = k0
                g[:, s] = c0[n:]
                a[:, s] = c0[:n]
            else:
                tied[int(s)] = cand
    return VariantCalls(g, a, cat, tied)
