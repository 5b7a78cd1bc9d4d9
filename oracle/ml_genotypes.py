"""ORACLE (test infrastructure, never on the product path): the max-sum / maximum-likelihood-genotype half of the
reference, restated on the CPU.

What is restated (reference paths relative to /root/reference/src/beast/evolution):
  * leaf dropout-component arg-max        rawreadcountsmodel/RawReadCountsModelFiniteMuDM.java:54-77  (C: so_leaf_likelihood_ml)
  * max-sum pruning with back-pointers    likelihood/ScsBeerLikelihoodCore.java:2918-3486             (C: so_ml_log_pruning)
  * ML rate category per pattern          likelihood/ScsBeerLikelihoodCore.java:278-296               (C: so_pattern_categories)
  * the set of ML genotype combinations   likelihood/ScsTreeLikelihood.java:947-1460 (getMLGenotypes, first run)
  * updateMaxLikelihoodGenotypes          likelihood/ScsTreeLikelihood.java:1471-1590
  * tie-break among combinations          likelihood/ScsTreeLikelihood.java:2013-2281, 2388-2413

Pinning: `update_max_likelihood_genotypes` is checked against the four golden cases of the reference's own unit test
(test/beast/evolution/likelihood/ScsTreeLikelihoodTest.java:153-292) in tests/test_oracle_ml.py -- the only part of
this path the reference pins.  The floating-point parts (component values, max-sum partials) have no golden values in
the reference: parity unpinned for those, like the sum-product oracle.

A genotype combination is stored the reference's way: an int array of length n_nodes + n_leaves, entry [leaf] for
leaf < n_leaves = index of the ML dropout component of that leaf, entry [n_leaves + node] = genotype of `node`
(ScsTreeLikelihood.java:140-146); -1 = not set.
"""
import ctypes as C
import itertools

import numpy as np

from . import sieve_oracle as so

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_bp = C.POINTER(C.c_ubyte)
_bound = False


def _lib():
    global _bound
    L = so.lib()
    if not _bound:
        L.so_leaf_likelihood_ml.argtypes = [_ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                            C.POINTER(so.LeafParams), _dp, _bp]
        L.so_leaf_likelihood_ml.restype = None
        L.so_ml_log_pruning.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _bp]
        L.so_ml_log_pruning.restype = None
        L.so_pattern_categories.argtypes = [C.c_int, C.c_int, _dp, C.c_int, _bp]
        L.so_pattern_categories.restype = None
        _bound = True
    return L


def leaf_likelihood_ml(counts_s_n_5, taxon, size_factor, params):
    """-> (log leaf partials [S][4], dropout-component tie masks uint8 [S][4])."""
    c = np.ascontiguousarray(counts_s_n_5, dtype=np.int32)
    S, n = c.shape[0], c.shape[1]
    out = np.empty((S, 4))
    ado = np.empty((S, 4), dtype=np.uint8)
    _lib().so_leaf_likelihood_ml(c.ctypes.data_as(_ip), n, 0, S, int(taxon), float(size_factor), C.byref(params),
                                 out.ctypes.data_as(_dp), ado.ctypes.data_as(_bp))
    return out, ado


def ml_log_pruning(ch1, m1, ch2, m2):
    """ch* [K][S][4] (leaf partials replicated over K, or a child's ML partials), m* [K][16]; ch2 None = unary root.
    -> (ML partials [K][S][4], back-pointers uint8 [K][S][4]: low nibble child 1, high nibble child 2)."""
    ch1 = np.ascontiguousarray(ch1, dtype=np.float64)
    K, S = ch1.shape[0], ch1.shape[1]
    m1 = np.ascontiguousarray(m1, dtype=np.float64)
    p2 = q2 = None
    if ch2 is not None:
        ch2 = np.ascontiguousarray(ch2, dtype=np.float64)
        m2 = np.ascontiguousarray(m2, dtype=np.float64)
        p2, q2 = ch2.ctypes.data_as(_dp), m2.ctypes.data_as(_dp)
    par = np.empty((K, S, 4))
    back = np.empty((K, S, 4), dtype=np.uint8)
    _lib().so_ml_log_pruning(K, S, ch1.ctypes.data_as(_dp), m1.ctypes.data_as(_dp), p2, q2, par.ctypes.data_as(_dp),
                             back.ctypes.data_as(_bp))
    return par, back


def pattern_categories(root_partials, root_genotype=0):
    rp = np.ascontiguousarray(root_partials, dtype=np.float64)
    K, S = rp.shape[0], rp.shape[1]
    out = np.empty(S, dtype=np.uint8)
    _lib().so_pattern_categories(K, S, rp.ctypes.data_as(_dp), int(root_genotype), out.ctypes.data_as(_bp))
    return out


def _argmax_masks(cand):
    """cand [..., 4 parent genotypes, 4 child genotypes] -> (max [..., 4], masks uint8 [..., 4]): the reference's sequential
    rule (first index seeds, '<' replaces, '==' appends) lists exactly the indices that attain the maximum."""
    mx = cand.max(axis=-1)
    eq = cand == mx[..., None]
    masks = (eq * (1 << np.arange(4))).sum(axis=-1).astype(np.uint8)
    return mx, masks


def ml_lin_pruning(ch1, leaf1, m1, ch2, leaf2, m2):
    """The max-sum part of the LINEAR-space twin (useLogPartials = false), ScsBeerLikelihoodCore.java:1759-2335:
    calculateLeafLeafPruning :1895-2016, calculateLeafPartialPruning :2018-2166, calculatePartialPartialPruning :2168-2335.
    A leaf child contributes log(P[p][c] * leaf[c]) with its LINEAR likelihood and the unclamped P (:1948-1949, 1964-1978);
    an internal child log(P > 0 ? P : Double.MIN_VALUE) + ML[c] with its (log-valued) ML partial (:2260-2263).
    ch* [K][S][4]: linear leaf partials (leaf* True) or log ML partials; m* [K][16]; ch2 None = unary root.
    -> (ML partials [K][S][4], back-pointers uint8 [K][S][4])."""
    def side(ch, is_leaf, m):
        ch = np.asarray(ch, dtype=np.float64)
        P = np.asarray(m, dtype=np.float64).reshape(-1, 1, 4, 4)           # [K][1][p][c]
        with np.errstate(divide="ignore"):
            if is_leaf:
                cand = np.log(P * ch[:, :, None, :])
            else:
                cand = np.log(np.where(P > 0.0, P, 4.9406564584124654e-324)) + ch[:, :, None, :]
        return _argmax_masks(cand)
    mx1, b1 = side(ch1, leaf1, m1)
    if ch2 is None:
        return mx1, b1
    mx2, b2 = side(ch2, leaf2, m2)
    return mx1 + mx2, (b1 | (b2 << 4)).astype(np.uint8)


def ml_trace(counts_s_n_5, size_factors, params, ops, matrices, root_genotype=0):
    """The whole traceMLGenotypes evaluation on one tree (ScsTreeLikelihood.traverse :700-722), log partials or -- with
    params.use_log == 0 -- the linear-space twin (the leaf partials and sum-product partials are then linear likelihoods).

    ops: [(parent, c1, c2|-1)] post-order; matrices [2n][K][16].
    -> dict(leaf [n][S][4], ado [n][S][4] u8, ml {node: [K][S][4]}, back {node: [K][S][4] u8},
            partials {node: [K][S][4]} sum-product partials of every node, categories [S] u8 mask)."""
    c = np.ascontiguousarray(counts_s_n_5, dtype=np.int32)
    S, n = c.shape[0], c.shape[1]
    matrices = np.asarray(matrices, dtype=np.float64)
    K = matrices.shape[1]
    use_log = bool(params.use_log)
    leaf = np.empty((n, S, 4))
    ado = np.empty((n, S, 4), dtype=np.uint8)
    for j in range(n):
        leaf[j], ado[j] = leaf_likelihood_ml(c, j, size_factors[j], params)
    core = so.Core(2 * n, n, S, K, 4, use_log=use_log)
    for j in range(n):
        core.set_node_partials(j, leaf[j])
    for node in range(2 * n):
        for k in range(K):
            core.set_node_matrix(node, k, matrices[node, k])
    ml, back, root = {}, {}, None
    for parent, c1, c2 in ops:
        parent, c1, c2 = int(parent), int(c1), int(c2)
        core.calculate_partials(c1, c1 < n, c2, 0 <= c2 < n, parent, 0)
        a1 = np.broadcast_to(leaf[c1], (K, S, 4)) if c1 < n else ml[c1]
        a2 = None
        if c2 >= 0:
            a2 = np.broadcast_to(leaf[c2], (K, S, 4)) if c2 < n else ml[c2]
        if use_log:
            ml[parent], back[parent] = ml_log_pruning(a1, matrices[c1], a2, matrices[c2] if c2 >= 0 else None)
        else:
            ml[parent], back[parent] = ml_lin_pruning(a1, c1 < n, matrices[c1], a2, 0 <= c2 < n,
                                                      matrices[c2] if c2 >= 0 else None)
        root = parent
    partials = {j: np.broadcast_to(leaf[j], (K, S, 4)) for j in range(n)}
    for parent, _, _ in ops:
        partials[int(parent)] = core.get_node_partials(int(parent)).reshape(K, S, 4)
    cats = pattern_categories(partials[root], root_genotype)
    return dict(leaf=leaf, ado=ado, ml=ml, back=back, partials=partials, categories=cats, root=root)


def _bits(mask):
    return [i for i in range(8) if (int(mask) >> i) & 1]


def enumerate_ml_genotypes(ops, n_leaves, back_at, ado_at, root_genotype=0, limit=1 << 16):
    """All ML genotype combinations of one (category, pattern): what getMLGenotypes leaves in
    maxLikelihoodGenotypes[mpIndex] after its first run (ScsTreeLikelihood.java:947-1460) -- every assignment that
    follows the back-pointers from the root genotype, each leaf paired with each arg-max dropout component of its
    genotype.

    back_at: {node: uint8[4]} back-pointer bytes of that (category, pattern); ado_at: uint8 [n_leaves][4].
    -> sorted list of tuples of length n_nodes + n_leaves."""
    n = int(n_leaves)
    children = {int(p): (int(c1), int(c2)) for p, c1, c2 in ops}
    root = int(ops[-1][0])
    n_nodes = 2 * n

    def below(node, g):
        """list of {index: value} dicts for the subtree under `node` given its genotype g"""
        if node < n:
            return [{node: a, n + node: g} for a in _bits(ado_at[node][g])]
        c1, c2 = children[node]
        b = int(back_at[node][g])
        opts1 = [d for cg in _bits(b & 15) for d in below(c1, cg)]
        if c2 < 0:
            combos = opts1
        else:
            opts2 = [d for cg in _bits(b >> 4) for d in below(c2, cg)]
            if len(opts1) * len(opts2) > limit:
                raise OverflowError("too many tied ML genotype combinations")
            combos = [{**a, **bb} for a, bb in itertools.product(opts1, opts2)]
        return [{**d, n + node: g} for d in combos]

    out = []
    for d in below(root, root_genotype):
        arr = [-1] * (n_nodes + n)
        for k, v in d.items():
            arr[k] = v
        out.append(tuple(arr))
    return sorted(out)


def ml_genotypes_log_likelihood(partials_at, genotypes, n_leaves):
    """getMLGenotypesLogLikelihood (ScsTreeLikelihood.java:2388-2413): logSumExp over all nodes of the node's
    sum-product log partial at its ML genotype.  partials_at: [n_nodes][4] of one (category, pattern)."""
    n = int(n_leaves)
    vals = [float(partials_at[node][genotypes[n + node]]) for node in range(len(partials_at))]
    return so.log_sum_exp(vals)


def call_pattern(ops, n_leaves, trace, pattern, root_genotype=0):
    """collectMLGenotypesAndLogLikelihoods for one pattern (ScsTreeLikelihood.java:2013-2281): candidates = the
    combinations of every ML category; more than one -> keep those with the largest getMLGenotypesLogLikelihood.
    -> list of (category, combination)."""
    n = int(n_leaves)
    cats = _bits(trace["categories"][pattern])
    cand = []
    for k in cats:
        back_at = {node: trace["back"][node][k, pattern] for node in trace["back"]}
        for g in enumerate_ml_genotypes(ops, n, back_at, trace["ado"][:, pattern, :], root_genotype):
            cand.append((k, g))
    if len(cand) <= 1:
        return cand
    best, keep = -np.finfo(float).max, []
    for k, g in cand:
        pa = np.stack([trace["partials"][node][k, pattern] for node in range(2 * n)])
        v = ml_genotypes_log_likelihood(pa, g, n)
        if best < v:
            best, keep = v, [(k, g)]
        elif best == v:
            keep.append((k, g))
    return keep


# ---------------------------------------------------------------------------------------------------------------
# updateMaxLikelihoodGenotypes, ScsTreeLikelihood.java:1471-1590 (+ traverseSubtree :1601-1613,
# getMLGenotypeIndices :1933-1952).  `children`: {node: [child, ...]}; arrays are lists of ints of length
# n_nodes + n_ext.  Mutates and returns ml_genotypes (a list of lists), like the reference.
# ---------------------------------------------------------------------------------------------------------------
def _subtree_labels(children, node, out):
    out.append(node)
    for c in children.get(node, []):
        _subtree_labels(children, c, out)


def update_max_likelihood_genotypes(ml_genotypes, children, node, genotype, n_ext, n_nodes):
    if not isinstance(genotype, int):
        for g in genotype:  # the List<Integer> overload, :1462-1469
            update_max_likelihood_genotypes(ml_genotypes, children, node, int(g), n_ext, n_nodes)
        return ml_genotypes
    kids = children.get(node, [])
    width = n_nodes + n_ext
    labels = []
    for ch in kids:
        lab = []
        _subtree_labels(children, ch, lab)
        labels.append(lab)
    p_indices = [i for i, a in enumerate(ml_genotypes) if a[node + n_ext] == genotype]
    rest = len(ml_genotypes) - len(p_indices)
    unique_root = []
    for i in p_indices:
        pat = list(ml_genotypes[i][:width])
        for lab in labels:
            for k in lab:
                pat[k + n_ext] = -1
                if k < n_ext:
                    pat[k] = -1
        if pat not in unique_root:
            unique_root.append(pat)
    unique_sub = [[] for _ in kids]
    for i in p_indices:
        for j, lab in enumerate(labels):
            pat = [-1] * width
            for k in lab:
                pat[k + n_ext] = ml_genotypes[i][k + n_ext]
                if k < n_ext:
                    pat[k] = ml_genotypes[i][k]
            if pat not in unique_sub[j]:
                unique_sub[j].append(pat)
    comb = len(unique_root)
    for u in unique_sub:
        comb *= len(u)
    for root_p in unique_root:
        # the reference walks a mixed-radix counter over the children's unique patterns (:1551-1585)
        for choice in itertools.product(*[range(len(u)) for u in reversed(unique_sub)]):
            choice = choice[::-1]
            merged = list(root_p)
            for j, m in enumerate(choice):
                sub = unique_sub[j][m]
                for k in range(width):
                    if merged[k] < 0:
                        merged[k] = sub[k]
            if merged not in ml_genotypes:
                ml_genotypes.append(merged)
    assert len(ml_genotypes) - rest == comb
    return ml_genotypes
