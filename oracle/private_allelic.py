"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the per-pattern sequencing-coverage model of the reference,
`PrivateAllelicSeqCovModel` (rawreadcountsmodel/seqcovmodel/PrivateAllelicSeqCovModel.java), SURVEY.md section 8 row f-4.

Nothing of the product imports this file, and there is no CUDA counterpart yet: the model makes the leaf likelihoods
depend on the rate category (one (t, v) pair per (matrix, pattern) instead of one shared pair) and re-estimates the
pairs from the maximum-likelihood numbers of sequenced alleles after every accepted state.  This module is the first
step the task prescribes for a new row -- the oracle -- so that the device side can be written against it.

Parity: unpinned like the rest of the floating-point path (no JVM here, the reference holds no vector for this class);
`tests/test_oracle_private_allelic.py` pins it by independent derivations (sample mean / variance identities, a
vectorised re-implementation, scipy's negative binomial).

Plain Python / numpy loops in the reference's own operation order; sizes are those of a unit test.
"""
import math

import numpy as np

NUMERICAL_STABILIZER = 1e-6   # SeqCovModelInterface.java:218


def processed_seq_cov(cov, zero_cov_mode=0):
    """SeqCovModelInterface.java:296-314: coverage as the model sees it (zeroCovMode 1 adds one to every entry)."""
    cov = np.asarray(cov, dtype=np.int64)
    if zero_cov_mode == 0:
        return cov.copy()
    if zero_cov_mode == 1:
        return cov + 1
    raise ValueError("ERROR! Illegal value for zeroCovMode. Only 0 or 1 is supported.")


def seq_cov_mean_var(n_alleles, size_factor, t, v):
    """PrivateAllelicSeqCovModel.java:505-506 (same as ExploredSharedAllelicSeqCovModel.java:266-277 with the pattern's
    own t, v): mean and variance of the coverage of a cell with `n_alleles` sequenced alleles."""
    ka = n_alleles + NUMERICAL_STABILIZER
    mean = ka * t * size_factor
    var = mean + math.pow(size_factor, 2) * math.pow(ka, 2) * v
    return mean, var


def seq_cov_log_likelihood_per_allele(cov, n_alleles, size_factor, t, v, nb_log_density):
    """PrivateAllelicSeqCovModel.java:499-511, log mode.  `nb_log_density(x, r, p)` is the restated
    math/distributions/NegativeBinomial.java:76-79 (oracle.sieve_oracle.nb_log_density)."""
    mean, var = seq_cov_mean_var(n_alleles, size_factor, t, v)
    p = mean / var
    r = math.pow(mean, 2) / (var - mean)
    return nb_log_density(cov, r, p)


def deterministic_allelic_seq_cov_and_var(processed_cov, size_factors, n_alleles, prev_v=0.0):
    """PrivateAllelicSeqCovModel.java:458-489 (`computeDeterministicAllelicSeqCovAndVar`), one pattern: every cell has
    the same, known number of sequenced alleles.  Returns (t, v); v keeps `prev_v` when the moment estimate is not
    positive (:485-486)."""
    n = len(size_factors)
    q = 0.0
    for j in range(n):
        q += processed_cov[j] / size_factors[j]
    q /= n
    w = z = 0.0
    if n > 1:
        for j in range(n):
            w += math.pow(processed_cov[j] / size_factors[j] - q, 2)
            z += 1 / size_factors[j]
        w /= (n - 1)
        z = w - q * z / n
    t = q / (n_alleles + NUMERICAL_STABILIZER)
    v = prev_v
    if z > 0:
        v = z / math.pow(n_alleles + NUMERICAL_STABILIZER, 2)
    return t, v


def update_allelic_seq_cov_and_raw_var(processed_cov, size_factors, modeled_alleles, combinations, weights, prev_v):
    """PrivateAllelicSeqCovModel.java:312-432 (`updateAllelicSeqCovAndRawVar`) for ONE (matrix, pattern) pair.

    processed_cov   [n] coverage of the pattern per cell (processed_seq_cov)
    size_factors    [n]
    modeled_alleles sorted numbers of sequenced alleles the evolutionary model allows, e.g. [1, 2]
    combinations    list of [n] integer arrays: the maximum-likelihood number of sequenced alleles of every cell, one
                    array per ML genotype combination (MLNrOfSequencedAllelesPatterns)
    weights         list of ints, how often each combination occurred (MLNrOfSequencedAllelesWeights)
    prev_v          the raw variance the pair held before: kept when no combination yields a positive estimate (:423-424)

    Returns (t, v).  Quirks kept: a group of cells with a non-positive moment estimate z contributes 0 and does not count
    (:389-402); the coverage estimate divides by the number of NON-EMPTY groups N[0] (:414), the variance estimate by
    the number of groups with a positive z, N[1] (:416-419), and only combinations with N[1] > 0 enter its weighted mean.
    """
    n = len(size_factors)
    A = len(modeled_alleles)
    sum_seq_cov = 0.0
    sum_seq_cov_var = 0.0
    sum_weights = [0, 0]
    for comb, weight in zip(combinations, weights):
        h = [0] * A
        indices = [[0] * n for _ in range(A)]
        N = [0, 0]
        q = [0.0] * A
        w = [0.0] * A
        z = [0.0] * A
        for j in range(n):
            idx = modeled_alleles.index(int(comb[j]))     # Arrays.binarySearch on the sorted modeledAlleles (:350)
            indices[idx][h[idx]] = j
            h[idx] += 1
        for a in range(A):
            if h[a] > 0:
                N[0] += 1
                sum_q = 0.0
                for i in range(h[a]):
                    j = indices[a][i]
                    sum_q += processed_cov[j] / size_factors[j]
                q[a] = sum_q / h[a]
            if h[a] > 1:
                N[1] += 1
                sum_w = 0.0
                sum_z = 0.0
                for i in range(h[a]):
                    j = indices[a][i]
                    sum_w += math.pow(processed_cov[j] / size_factors[j] - q[a], 2)
                    sum_z += 1 / size_factors[j]
                w[a] = sum_w / (h[a] - 1)
                z[a] = w[a] - q[a] * sum_z / h[a]
                if z[a] <= 0:
                    z[a] = 0.0
                    N[1] -= 1
        sum_t = 0.0
        sum_v = 0.0
        for a in range(A):
            sum_t += q[a] / (modeled_alleles[a] + NUMERICAL_STABILIZER)
            sum_v += z[a] / math.pow(modeled_alleles[a] + NUMERICAL_STABILIZER, 2)
        sum_seq_cov += weight * sum_t / N[0]
        sum_weights[0] += weight
        if N[1] > 0:
            sum_seq_cov_var += weight * sum_v / N[1]
            sum_weights[1] += weight
    t = sum_seq_cov / sum_weights[0]
    v = prev_v
    if sum_seq_cov_var > 0:
        v = sum_seq_cov_var / sum_weights[1]
    return t, v


def leaf_seq_cov_log_likelihoods(cov, size_factors, modeled_alleles, t_kp, v_kp, nb_log_density):
    """PrivateAllelicSeqCovModel.java:272-296 (`computeSeqCovLikelihood`) for all cells of all patterns:
    out[k][p][cell][a] = log NB(cov[p][cell]; alleles a, t_kp[k][p], v_kp[k][p]) -- the category-dependent coverage term
    that replaces the shared-model term in the leaf mixture (RawReadCountsModelFiniteMuDM.java:121-260)."""
    cov = np.asarray(cov)
    K, P = np.asarray(t_kp).shape
    n = cov.shape[1]
    out = np.empty((K, P, n, len(modeled_alleles)))
    for k in range(K):
        for p in range(P):
            for j in range(n):
                for a, al in enumerate(modeled_alleles):
                    out[k, p, j, a] = seq_cov_log_likelihood_per_allele(int(cov[p, j]), al, size_factors[j], t_kp[k][p],
                                                                        v_kp[k][p], nb_log_density)
    return out


def leaf_likelihood_private(oracle, counts_s_n_5, taxon, size_factor, params, t_kp, v_kp):
    """The leaf log-partials of one cell under the private model, [K][S][4]: RawReadCountsModelInterface.java:458-506
    with `nrOfMatrices = K` -- the nucleotide (Dirichlet-multinomial) terms are those of the shared model, the coverage
    terms come from the (matrix, pattern) pair's own t, v (PrivateAllelicSeqCovModel.java:272-296), and the dropout
    mixture (RawReadCountsModelFiniteMuDM.java:121-260) is taken per matrix.  Evaluated with the restated shared-model
    leaf likelihood, one (matrix, pattern) at a time with that pair's t, v.  `oracle` is the oracle.sieve_oracle module."""
    t_kp, v_kp = np.asarray(t_kp, dtype=np.float64), np.asarray(v_kp, dtype=np.float64)
    K, S = t_kp.shape
    out = np.empty((K, S, 4))
    for k in range(K):
        for p in range(S):
            P = oracle.LeafParams(params.ado_rate, float(t_kp[k, p]), float(v_kp[k, p]), params.eff_seq_err, params.shape1,
                                  params.shape2, params.single_ado, params.use_log)
            out[k, p] = oracle.leaf_likelihood(counts_s_n_5, taxon, size_factor, P, p, p + 1)[0]
    return out


# ---------------------------------------------------------------------------------------------------------------
# Whole evaluations under the private model (test infrastructure for the device side, sieve_b200/csrc/leaf_private.cuh)

def full_eval_private(oracle, counts_s_n_5, size_factors, params, t_kp, v_kp, ops, matrices, proportions, root):
    """Site log-likelihoods and logConstRoot with category-dependent leaves: the restated likelihood core
    (oracle.Core = ScsBeerLikelihoodCore) fed the [K][S][4] leaf partials of `leaf_likelihood_private`
    (ScsTreeLikelihood.java:583-601 with nrOfMatrices = K), then traverse / integratePartials / calculateLogLikelihoods.
    -> (site_logl [S], site_log_const [S], {node: partials [K][S][4]})."""
    c = np.ascontiguousarray(counts_s_n_5, dtype=np.int32)
    S, n = c.shape[0], c.shape[1]
    matrices = np.asarray(matrices, dtype=np.float64)
    K = matrices.shape[1]
    core = oracle.Core(2 * n, n, S, K, 4, use_log=params.use_log)
    leaves = {}
    for j in range(n):
        leaves[j] = leaf_likelihood_private(oracle, c, j, size_factors[j], params, t_kp, v_kp)
        core.set_node_partials(j, leaves[j])
    for node in range(2 * n):
        if node == root:
            continue
        for k in range(K):
            core.set_node_matrix(node, k, matrices[node, k])
    for parent, c1, c2 in ops:
        parent, c1, c2 = int(parent), int(c1), int(c2)
        core.calculate_partials(c1, c1 < n, c2, 0 <= c2 < n, parent, 0)
    rp, cr = core.integrate_partials(root, proportions, 0)
    sl, lc = core.calculate_log_likelihoods(rp, cr)
    partials = dict(leaves)
    for parent, _, _ in ops:
        partials[int(parent)] = core.get_node_partials(int(parent)).reshape(K, S, 4)
    return sl, lc, partials


def ml_trace_private(oracle, mlo, counts_s_n_5, size_factors, params, t_kp, v_kp, ops, matrices):
    """oracle.ml_genotypes.ml_trace with category-dependent leaves: leaf [n][K][S][4], ado [n][K][S][4]."""
    import ctypes as C
    c = np.ascontiguousarray(counts_s_n_5, dtype=np.int32)
    S, n = c.shape[0], c.shape[1]
    matrices = np.asarray(matrices, dtype=np.float64)
    K = matrices.shape[1]
    L = mlo._lib()
    leaf = np.empty((n, K, S, 4))
    ado = np.empty((n, K, S, 4), dtype=np.uint8)
    _dp, _ip, _bp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_ubyte)
    o = np.empty(4)
    a = np.empty(4, dtype=np.uint8)
    for j in range(n):
        for k in range(K):
            for p in range(S):
                P = oracle.LeafParams(params.ado_rate, float(t_kp[k][p]), float(v_kp[k][p]), params.eff_seq_err,
                                      params.shape1, params.shape2, params.single_ado, params.use_log)
                L.so_leaf_likelihood_ml(c.ctypes.data_as(_ip), n, p, p + 1, j, float(size_factors[j]), C.byref(P),
                                        o.ctypes.data_as(_dp), a.ctypes.data_as(_bp))
                leaf[j, k, p] = o
                ado[j, k, p] = a
    ml, back = {}, {}
    for parent, c1, c2 in ops:
        parent, c1, c2 = int(parent), int(c1), int(c2)
        a1 = leaf[c1] if c1 < n else ml[c1]
        a2 = None
        if c2 >= 0:
            a2 = leaf[c2] if c2 < n else ml[c2]
        ml[parent], back[parent] = mlo.ml_log_pruning(a1, matrices[c1], a2, matrices[c2] if c2 >= 0 else None)
    return dict(leaf=leaf, ado=ado, ml=ml, back=back)


def post_process(oracle, mlo, counts_s_n_5, size_factors, params, t_kp, v_kp, ops, matrices, zero_cov_mode=0):
    """ScsTreeLikelihood.postProcess steps 1-2 (likelihood/ScsTreeLikelihood.java:2454-2475) from scratch: the ML genotype
    combinations of every (matrix, pattern) pair (getMLGenotypes, first run), their numbers of sequenced alleles sorted and
    de-duplicated into weighted patterns (:1066-1101), then updateAllelicSeqCovAndRawVar for every pair.
    -> (t [K][S], v [K][S], n_combinations [K][S])."""
    c = np.ascontiguousarray(counts_s_n_5, dtype=np.int32)
    S, n = c.shape[0], c.shape[1]
    K = np.asarray(matrices).shape[1]
    tr = ml_trace_private(oracle, mlo, c, size_factors, params, t_kp, v_kp, ops, matrices)
    modeled = [1, 2] if params.single_ado else [0, 1, 2]
    cov = processed_seq_cov(c[:, :, 4] if c.shape[2] == 5 else c[:, :, 3], zero_cov_mode)   # [S][n]
    t_new = np.array(t_kp, dtype=np.float64).copy()
    v_new = np.array(v_kp, dtype=np.float64).copy()
    ncomb = np.zeros((K, S), dtype=np.int64)
    for k in range(K):
        for p in range(S):
            back_at = {node: tr["back"][node][k, p] for node in tr["back"]}
            combos = mlo.enumerate_ml_genotypes(ops, n, back_at, tr["ado"][:, k, p, :], 0)
            ncomb[k, p] = len(combos)
            pats = sorted(tuple(2 - comb[j] for j in range(n)) for comb in combos)   # getNrOfAlleles() == 2 (:1080)
            uniq, weights = [], []
            for q in pats:
                if uniq and uniq[-1] == q:
                    weights[-1] += 1
                else:
                    uniq.append(q)
                    weights.append(1)
            t_new[k, p], v_new[k, p] = update_allelic_seq_cov_and_raw_var(cov[p], size_factors, modeled,
                                                                          [np.array(q) for q in uniq], weights, v_new[k, p])
    return t_new, v_new, ncomb
