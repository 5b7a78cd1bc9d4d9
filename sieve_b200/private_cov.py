"""Host half of the per-pattern coverage model, `PrivateAllelicSeqCovModel`
(rawreadcountsmodel/seqcovmodel/PrivateAllelicSeqCovModel.java), for the (category, pattern) pairs the device leaves to
the host: pairs whose maximum-likelihood genotype combination is not unique.

The device re-estimates (allelicSeqCov, allelicSeqCovRawVar) of every pair with ONE combination
(`sieve_private_update_from_ml`, csrc/leaf_private.cuh k_private_update).  For a tied pair the reference weighs all
combinations (`updateAllelicSeqCovAndRawVar`, :312-432, fed by ScsTreeLikelihood.java:1066-1135): the combinations are
enumerated here from the back-pointers of that site (`ScsCudaLikelihoodCore.ml_site`), turned into sorted, de-duplicated
patterns of numbers of sequenced alleles with their multiplicities, and the grouped method of moments is evaluated per
pattern.  Ties are rare (exactly equal doubles), so this is plain numpy on a handful of pairs.
"""
import numpy as np

from .variants import enumerate_combinations

NUMERICAL_STABILIZER = 1e-6   # SeqCovModelInterface.java:218


def moments_of_combination(x, inv_sf, alleles, modeled_alleles):
    """One combination (:329-412): x = processedSeqCov / sizeFactor per cell, alleles = ML number of sequenced alleles per
    cell.  -> (sumT / N[0], sumV / N[1] or None when no group has a positive variance estimate)."""
    sum_t = sum_v = 0.0
    n0 = n1 = 0
    for a in modeled_alleles:
        sel = alleles == a
        h = int(sel.sum())
        q = z = 0.0
        if h > 0:
            n0 += 1
            q = float(np.cumsum(x[sel])[-1]) / h              # sequential sum in cell order, like the reference
        if h > 1:
            n1 += 1
            w = float(np.cumsum((x[sel] - q) ** 2)[-1]) / (h - 1)
            z = w - q * float(np.cumsum(inv_sf[sel])[-1]) / h
            if z <= 0:
                z = 0.0
                n1 -= 1
        sum_t += q / (a + NUMERICAL_STABILIZER)
        sum_v += z / (a + NUMERICAL_STABILIZER) ** 2
    return sum_t / n0, (sum_v / n1 if n1 > 0 else None)


def update_pair(processed_cov, size_factors, combinations, single_ado, prev_v):
    """(t, v) of one (category, pattern) pair from ALL its ML combinations.  combinations: tuples as
    `enumerate_combinations` returns them (entry [leaf] = dropped alleles of the leaf's ML component)."""
    n = len(size_factors)
    modeled = [1, 2] if single_ado else [0, 1, 2]
    pats = sorted(tuple(2 - c[j] for j in range(n)) for c in combinations)   # getNrOfAlleles() == 2 (ScsTreeLikelihood.java:1080)
    uniq, weights = [], []
    for p in pats:
        if uniq and uniq[-1] == p:
            weights[-1] += 1
        else:
            uniq.append(p)
            weights.append(1)
    sf = np.asarray(size_factors, dtype=np.float64)
    x = np.asarray(processed_cov, dtype=np.float64) / sf
    inv = 1.0 / sf
    s_t = s_v = 0.0
    w_t = w_v = 0
    for p, w in zip(uniq, weights):
        t_c, v_c = moments_of_combination(x, inv, np.asarray(p), modeled)
        s_t += w * t_c
        w_t += w
        if v_c is not None:
            s_v += w * v_c
            w_v += w
    return s_t / w_t, (s_v / w_v if s_v > 0 else prev_v)


def resolve_tied_pairs(core, tree, counts, size_factors, tied, zero_cov_mode=0, single_ado=True):
    """For every tied (category, site) pair: enumerate, re-estimate, and write the pairs back to the core.
    -> number of pairs whose (t, v) moved."""
    ks, ss = np.nonzero(tied)
    if len(ks) == 0:
        return 0
    n = core.n
    children = {p: (int(tree.left[p]), int(tree.right[p])) for p in range(n, 2 * n)}
    t, v = core.get_private_seq_cov()
    moved = 0
    site_cache = {}
    for k, s in zip(ks, ss):
        if s not in site_cache:
            site_cache[s] = core.ml_site(int(s))[0]
        combos = enumerate_combinations(children, tree.root, n, site_cache[s][:, k, :], 0)
        cov = np.asarray(counts[s, :, -1], dtype=np.int64) + (1 if zero_cov_mode == 1 else 0)
        nt, nv = update_pair(cov, size_factors, combos, single_ado, v[k, s])
        if nt != t[k, s] or nv != v[k, s]:
            moved += 1
            t[k, s], v[k, s] = nt, nv
    if moved:
        core.set_private_seq_cov(t, v)
    return moved
