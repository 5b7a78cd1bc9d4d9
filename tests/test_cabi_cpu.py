"""CPU-side checks of the boundary: the shared library loads, exports every symbol include/sieve_b200.h declares,
fails loudly without a device, and its pure-host helpers (shard combine, site split) follow the reference."""
import ctypes as C
import re

import numpy as np
import pytest

from sieve_b200 import _lib


@pytest.fixture(scope="module")
def L():
    _lib.build()
    return _lib.lib()


def test_every_declared_symbol_is_exported_and_bound(L):
    header = open(_lib.HEADER_PATH).read()
    declared = set(re.findall(r"\b(sieve_[a-z_0-9]+)\s*\(", header))
    declared -= {"sieve_b200"}
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(L, name), "missing export: " + name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_mcmc_header_symbols_are_exported(L):
    import os
    header = open(os.path.join(os.path.dirname(_lib.HEADER_PATH), "sieve_b200_mcmc.h")).read()
    declared = set(re.findall(r"\b(sieve_[a-z_0-9]+)\s*\(", header))
    assert {"sieve_mcmc_create", "sieve_mcmc_run", "sieve_gamma_category_rates"} <= declared
    for name in sorted(declared):
        assert hasattr(L, name), "missing export: " + name
    # the discrete-gamma rates of the driver (no device needed) equal the host mirror's
    from sieve_b200 import substmodel
    from sieve_b200.mcmc import gamma_category_rates
    for shape in (0.3, 1.0, 2.5):
        np.testing.assert_allclose(gamma_category_rates(shape, 4), substmodel.gamma_category_rates(shape, 4)[0], rtol=1e-9)


def test_no_device_is_a_loud_error_not_a_fallback(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert L.sieve_device_count() == 0
    cfg = _lib.Config(4, 8, 4, 4, _lib.FLAG_LOG_PARTIALS, -1)
    h = C.c_void_p()
    rc = L.sieve_create(C.byref(cfg), C.byref(h))
    assert rc == _lib.ERR_CUDA
    assert b"cuda" in L.sieve_last_error().lower() or b"device" in L.sieve_last_error().lower()


def test_pattern_points_follow_calcPatternPoints(L, oracle):
    from sieve_b200.core import pattern_points
    for S, T in [(10, 3), (273, 8), (100000, 7), (5, 5)]:
        np.testing.assert_array_equal(pattern_points(S, T), oracle.pattern_points(S, T))


def test_combine_shards_matches_oracle_formulas(L, oracle):
    from sieve_b200.core import combine_shards
    rng = np.random.default_rng(0)
    T, per = 4, 25
    ll = rng.uniform(-900, -100, size=(T, per))
    lc = rng.uniform(-50, -1, size=(T, per))
    M = 4321.0
    shards = []
    for t in range(T):
        shards.append(_lib.ShardResult(ll[t].sum(), oracle.log_sum_exp(lc[t]), lc[t].sum(),
                                       oracle.asc_bias(lc[t], None, "lewis", 0.0), float(per)))
    base = ll.sum()
    # several workers: alignment/ScsAlignment.java:811-820
    got = combine_shards(shards, "felsenstein_mean", M)
    ref = base + oracle.asc_bias_threaded([s.const_lse for s in shards], "felsenstein_mean", T * per, M)
    assert abs(got - ref) < 1e-9 * abs(ref)
    got = combine_shards(shards, "felsenstein_geometric", M)
    ref = base + oracle.asc_bias_threaded([s.const_sum for s in shards], "felsenstein_geometric", T * per, M)
    assert abs(got - ref) < 1e-9 * abs(ref)
    got = combine_shards(shards, "lewis", M)
    assert abs(got - (base + sum(s.lewis_sum for s in shards))) < 1e-9 * abs(base)
    # one worker: alignment/ScsAlignment.java:764-782
    one = [shards[0]]
    for mode in ("felsenstein_mean", "felsenstein_geometric", "lewis", "none"):
        ref = ll[0].sum() + oracle.asc_bias(lc[0], None, mode, M)
        assert abs(combine_shards(one, mode, M) - ref) < 1e-9 * abs(ref)


def test_walk_block_count_follows_the_wave_model(L):
    """100 000 sites are 1.76 / 1.51 / 1.32 waves at 6 / 7 / 8 blocks per SM and stay at 6, 125 000 sites are
    2.20 / 1.89 / 1.65 and go to 7; launches that fit into one wave are never squeezed."""
    assert L.sieve_walk_blocks_per_sm(100000) == 6
    assert L.sieve_walk_blocks_per_sm(125000) == 7
    assert L.sieve_walk_blocks_per_sm(113664) == 6          # exactly two waves at 6
    for s in (1, 273, 762, 56832):
        assert L.sieve_walk_blocks_per_sm(s) == 6
    assert all(L.sieve_walk_blocks_per_sm(s) in (6, 7, 8) for s in range(57000, 2000000, 7919))
