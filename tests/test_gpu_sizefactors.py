"""Global size factors over site shards on the device (include/sieve_b200.h, sieve_sf_*): the exact distributed selection
must reproduce, bit for bit, what one handle over ALL sites computes with its sort (sieve_compute_size_factors), which the
oracle pins (seqcovmodel/SeqCovModelInterface.java:291-351)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from sieve_b200.core import ScsCudaLikelihoodCore, pattern_points  # noqa: E402
from sieve_b200.data import synthetic  # noqa: E402
from sieve_b200.sizefactors import SizeFactorPart, solve_parts  # noqa: E402
from sieve_b200.treelikelihood import ScsTreeLikelihood, ThreadedScsTreeLikelihood  # noqa: E402


def _bits(a):
    return np.asarray(a, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("n,S,n_parts,zero_cov_mode,with_weights", [(70, 2300, 3, 0, False), (33, 999, 2, 0, True),
                                                                    (33, 999, 5, 1, False), (300, 6000, 8, 0, False)])
def test_sharded_selection_equals_single_handle_sort(oracle, n, S, n_parts, zero_cov_mode, with_weights):
    d = synthetic(n, S, seed=90 + n_parts, zero_cov_frac=0.12)
    counts = d["counts"]
    counts[5, :, :] = 0            # a site nobody covers
    w = np.random.default_rng(3).integers(1, 5, size=S).astype(np.int32) if with_weights else None
    whole = ScsCudaLikelihoodCore(n, S, 1, 4, True, False, True)
    whole.set_counts(counts)
    whole.set_pattern_weights(w)
    ref, ref_st = whole.compute_size_factors(zero_cov_mode)
    osf, ost = oracle.size_factors(oracle.counts5(counts), w, zero_cov_mode)
    np.testing.assert_allclose(ref, osf, rtol=1e-13)
    pts = pattern_points(S, n_parts)
    cores, parts = [], []
    for r in range(n_parts):
        c = ScsCudaLikelihoodCore(n, int(pts[r + 1] - pts[r]), 1, 4, True, False, True)
        c.set_counts(counts[pts[r]:pts[r + 1]])
        c.set_pattern_weights(None if w is None else w[pts[r]:pts[r + 1]])
        cores.append(c)
        parts.append(SizeFactorPart(core=c, zero_cov_mode=zero_cov_mode))
    sf, st = solve_parts(parts)
    assert np.array_equal(_bits(sf), _bits(ref))                       # the split cannot change a bit
    assert np.array_equal(_bits(st), _bits(ref_st))
    # one part over the whole alignment too (the pool-free entry point on an unsharded device tensor)
    one = SizeFactorPart(core=whole, zero_cov_mode=zero_cov_mode)
    sf1, _ = solve_parts([one])
    assert np.array_equal(_bits(sf1), _bits(ref))
    for p in parts + [one]:
        p.close()
    for c in cores + [whole]:
        c.close()


def test_threaded_likelihood_uses_global_size_factors(oracle):
    """The sharded driver: same logP as one shard over everything (the reference's thread-split invariance), which only
    holds when every shard uses the size factors of the FULL alignment."""
    d = synthetic(40, 1500, seed=17)
    counts, tree = d["counts"], d["tree"]
    M = 30000.0
    one = ScsTreeLikelihood(counts, tree.copy(), asc_mode="felsenstein_mean", n_background_sites=M)
    ref = one.calculate_logp()
    thr = ThreadedScsTreeLikelihood(counts, tree.copy(), 3, asc_mode="felsenstein_mean", n_background_sites=M)
    got = thr.calculate_logp()
    assert np.array_equal(_bits(thr.size_factors), _bits(one.size_factors))
    assert abs(got - ref) <= 1e-11 * abs(ref)
    osf, _ = oracle.size_factors(oracle.counts5(counts))
    np.testing.assert_allclose(thr.size_factors, osf, rtol=1e-13)
    a, _ = one.pattern_log_likelihoods()
    b, _ = thr.pattern_log_likelihoods()
    np.testing.assert_array_equal(a, b)
    one.close()
    thr.close()
