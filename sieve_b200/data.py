"""Read-count data: the reference's `read_counts.tsv` grammar, pattern compression, and the synthetic generator.

Count tuples are kept as int32 [site][cell][4] = (alt1, alt2, alt3, coverage); the reference's fifth value
(ref = cov - sum(alt), datatype/FullSupsCov.java:74-77) is redundant and never stored.
"""
import re

import numpy as np

from . import substmodel
from .tree import random_coalescent


# ------------------------------------------------------------------------------------------------------------------
# read_counts.tsv  (app/datacollector/DataCollector.java:213-364, 409-423, 477-502, 541-610)
# ------------------------------------------------------------------------------------------------------------------
def _chrom_key(label):
    s = label.strip()
    s = s[3:] if s.lower().startswith("chr") else s
    if s.isdigit():
        return (0, int(s), "")
    order = {"X": 1, "Y": 2, "M": 3, "MT": 3}
    return (1, order.get(s.upper(), 9), s.upper())


def load_read_counts_tsv(path, cell_names_path=None):
    """Parse a DataCollector input file.

    Returns dict(counts=int32[S][n][4] with sites sorted by locus and cells sorted by name (the order in which
    DataCollector writes the <sequence> elements, :541-583), cell_names, loci, n_background_sites,
    background=list of {value: count} histograms)."""
    section = None
    n_cells = n_sites = n_bg = None
    rows = []
    background = []
    with open(path) as f:
        for line in f:
            s = line.strip()
            if not s:
                continue
            m = re.fullmatch(r"=(\w+)=", s)
            if m:
                section = m.group(1)
                continue
            if section == "numSamples":
                n_cells = int(s)
            elif section == "numCandidateMutatedSites":
                n_sites = int(s)
            elif section == "numBackgroundSites":
                n_bg = int(s)
            elif section == "mutations":
                rows.append(s.split("\t"))
            elif section == "background":
                background.append({int(a): int(b) for a, b in (x.split(",") for x in s.split("\t"))})
    rows.sort(key=lambda r: (_chrom_key(r[0]), int(r[1]), r[3].strip().upper()))
    if n_sites is not None and n_sites != len(rows):
        raise ValueError("Unmatched number of candidate mutated sites: %d specified, %d found" % (n_sites, len(rows)))
    n = len(rows[0]) - 4
    if n_cells is not None and n != n_cells:
        raise ValueError("The number of cells and the data columns do not match: %d vs %d" % (n_cells, n))
    counts = np.zeros((len(rows), n, 4), dtype=np.int32)
    for i, r in enumerate(rows):
        for j in range(n):
            nums = r[4 + j].strip().split(";")[1].split(",")
            counts[i, j] = [int(x) for x in nums]
    if (counts < 0).any():
        raise ValueError("Read counts should be non-negative integers")
    names = None
    if cell_names_path is not None:
        with open(cell_names_path) as f:
            names = [ln.split()[0] for ln in f if ln.strip()]
        if len(names) != n:
            raise ValueError("The number of cells and the number of cell names provided do not match")
        order = np.argsort(np.array(names))     # String.compareTo order for ASCII names
        counts = counts[:, order, :]
        names = [names[k] for k in order]
    if n_bg is None and background:
        n_bg = sum(background[0].values()) // n
    loci = [(r[0], int(r[1]), r[2], r[3]) for r in rows]
    return dict(counts=counts, cell_names=names, loci=loci, n_background_sites=n_bg, background=background)


def sort_patterns(counts4):
    """alignment/ScsAlignment.java:400-446: sort sites lexicographically over [cell][alt1,alt2,alt3,ref,cov], merge
    duplicates into weighted patterns. Returns (patterns int32[P][n][4], weights int32[P], site_to_pattern int32[S])."""
    c = np.asarray(counts4, dtype=np.int64)
    S, n, _ = c.shape
    key = np.empty((S, n, 5), dtype=np.int64)
    key[..., :3] = c[..., :3]
    key[..., 3] = c[..., 3] - c[..., :3].sum(-1)
    key[..., 4] = c[..., 3]
    flat = key.reshape(S, n * 5)
    order = np.lexsort(flat.T[::-1])
    sorted_flat = flat[order]
    new = np.ones(S, dtype=bool)
    new[1:] = (sorted_flat[1:] != sorted_flat[:-1]).any(axis=1)
    pat_id = np.cumsum(new) - 1
    first = np.nonzero(new)[0]
    patterns = np.asarray(counts4, dtype=np.int32)[order[first]]
    weights = np.bincount(pat_id).astype(np.int32)
    site_to_pattern = np.empty(S, dtype=np.int32)
    site_to_pattern[order] = pat_id
    return patterns, weights, site_to_pattern


# ------------------------------------------------------------------------------------------------------------------
# synthetic data (SURVEY.md section 8d "Concrete inputs")
# ------------------------------------------------------------------------------------------------------------------
def synthetic(n_cells, n_sites, seed=20260101, theta=0.3, t=50.0, v=50.0, eps=0.008, w1=150.0, w2=2.0,
              gamma_shape=1.0, K=4, zero_cov_frac=0.05, tree=None, tree_scale=None, max_cov=None):
    """Seeded generator following the model that is evaluated: coalescent trunk tree, genotypes evolved from a 0/0 root
    under the finite-sites Q with one of K gamma categories per site, single allelic dropout, NB coverage with per-cell
    size factors, Dirichlet-multinomial nucleotide counts, ~5 % zero-coverage entries.

    Returns dict(counts int32[S][n][4], tree, size_factors_true, params)."""
    rng = np.random.default_rng(seed)
    n, S = n_cells, n_sites
    if tree is None:
        tree = random_coalescent(n, rng)
        # scale heights so that the expected number of mutations per site along a root-to-tip path is ~0.3
        scale = (0.3 if tree_scale is None else tree_scale) / max(tree.height[tree.root], 1e-12)
        tree.height *= scale
    rates, _ = substmodel.gamma_category_rates(gamma_shape, K)
    cat = rng.integers(0, K, size=S)
    # genotypes, pre-order
    geno = np.zeros((tree.node_count, S), dtype=np.int8)
    bl = tree.branch_lengths()
    order = tree.post_order()[::-1]
    for p in order:
        for c in (tree.left[p], tree.right[p]):
            if c < 0:
                continue
            u = rng.random(S)
            new = np.zeros(S, dtype=np.int8)
            for k in range(K):
                P = substmodel.transition_probabilities(bl[c] * rates[k])
                cum = np.cumsum(P, axis=1)
                sel = cat == k
                rows = cum[geno[p, sel]]
                new[sel] = (u[sel, None] > rows[:, :3]).sum(axis=1)
            geno[c] = new
    g = geno[:n].T                                     # [S][n]
    sf = rng.lognormal(0.0, 0.3, size=n)
    # allelic dropout: each cell-site loses one allele with prob theta
    ado = rng.random((S, n)) < theta
    alleles = np.where(ado, 1, 2)
    mean = alleles * t * sf[None, :]
    r = t * t / v
    p = r / (r + mean)
    cov = rng.negative_binomial(r, p).astype(np.int64)
    cov[rng.random((S, n)) < zero_cov_frac] = 0
    if max_cov is not None:
        cov = np.minimum(cov, max_cov)
    # which DM parameter vector: genotype + dropout -> (alt1, alt2, alt3, ref) concentration
    f = eps / 3.0
    a_ref = np.array([f, f, f, 1 - eps]) * w1          # 0/0
    a_alt = np.array([1 - eps, f, f, f]) * w1          # 1/1
    a_het2 = np.array([0.5 - f, 0.5 - f, f, f]) * w2   # 1/1'
    a_het = np.array([0.5 - f, f, f, 0.5 - f]) * w2    # 0/1
    a_alt2 = np.array([f, 1 - eps, f, f]) * w1         # 1'/1' after dropout of allele 1 (approximation)
    alpha = np.empty((S, n, 4))
    coin = rng.random((S, n)) < 0.5
    alpha[...] = a_ref
    alpha[(g == 1) & ~ado] = a_het
    alpha[(g == 1) & ado & coin] = a_alt
    alpha[(g == 2)] = a_alt
    alpha[(g == 3) & ~ado] = a_het2
    alpha[(g == 3) & ado & coin] = a_alt
    alpha[(g == 3) & ado & ~coin] = a_alt2
    probs = rng.standard_gamma(alpha)
    probs /= probs.sum(-1, keepdims=True)
    # multinomial via sequential binomials
    counts = np.zeros((S, n, 4), dtype=np.int32)
    remaining = cov.copy()
    rest = np.ones((S, n))
    for i in range(3):
        q = np.clip(probs[..., i] / np.maximum(rest, 1e-300), 0.0, 1.0)
        x = rng.binomial(remaining, q)
        counts[..., i] = x
        remaining -= x
        rest -= probs[..., i]
    counts[..., 3] = cov
    params = dict(theta=theta, t=t, v=v, eps=eps, w1=w1, w2=w2, gamma_shape=gamma_shape, K=K)
    return dict(counts=counts, tree=tree, size_factors_true=sf, params=params, genotypes=geno, dropout=ado)
