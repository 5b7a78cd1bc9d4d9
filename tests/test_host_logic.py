"""Host-side pieces: the TSV grammar, pattern compression, tree bookkeeping, the rate matrix."""
import os

import numpy as np
import pytest

from sieve_b200 import data, substmodel
from sieve_b200.tree import caterpillar, fixture_9_leaves, random_coalescent

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_rate_matrix_is_the_reference_q():
    # Q = 0.3 * [[-1,1,0,0],[1/6,-2/3,1/6,1/3],[0,1/3,-1,2/3],[0,1/3,1/3,-2/3]]
    Q = substmodel.rate_matrix()
    ref = 0.3 * np.array([[-1, 1, 0, 0], [1 / 6, -2 / 3, 1 / 6, 1 / 3], [0, 1 / 3, -1, 2 / 3], [0, 1 / 3, 1 / 3, -2 / 3]])
    np.testing.assert_allclose(Q, ref, atol=1e-15)
    assert abs(-np.trace(Q) - 1.0) < 1e-15


def test_transition_probabilities_against_expm():
    from scipy.linalg import expm
    Q = substmodel.rate_matrix()
    for d in (0.0, 1e-6, 0.01, 0.3, 2.5, 40.0):
        P = substmodel.transition_probabilities(d)
        np.testing.assert_allclose(P, expm(Q * d), atol=2e-15)
        np.testing.assert_allclose(P.sum(1), 1.0, atol=1e-14)


def test_gamma_categories_have_mean_one():
    for a in (0.1, 1.0, 7.3):
        r, p = substmodel.gamma_category_rates(a, 4)
        assert abs(r.mean() - 1.0) < 1e-14 and np.all(np.diff(r) > 0) and np.allclose(p, 0.25)


def test_tsv_fixture_round_trip(tmp_path):
    z = np.load(os.path.join(GOLD, "cells10.npz"))
    counts = z["counts"]
    S, n = counts.shape[:2]
    p = tmp_path / "rc.tsv"
    names = tmp_path / "names"
    cells = ["cell%02d" % i for i in range(n)]
    with open(p, "w") as f:
        f.write("=numSamples=\n%d\n=numCandidateMutatedSites=\n%d\n=numBackgroundSites=\n9727\n=mutations=\n" % (n, S))
        # written in reverse locus order and reverse cell order: the loader sorts both (DataCollector.java:477-502, 541-583)
        for s in reversed(range(S)):
            cols = ["T,C,G;" + ",".join(str(x) for x in counts[s, j]) for j in reversed(range(n))]
            f.write("chr1\t%d\tA\tT\t%s\n" % (10 + s, "\t".join(cols)))
        f.write("=background=\n0,5\t1,2\n")
    with open(names, "w") as f:
        for c in reversed(cells):
            f.write(c + "\tCT\n")
    d = data.load_read_counts_tsv(str(p), str(names))
    np.testing.assert_array_equal(d["counts"], counts)
    assert d["cell_names"] == cells and d["n_background_sites"] == 9727 and d["background"][0] == {0: 5, 1: 2}


def test_tsv_errors(tmp_path):
    p = tmp_path / "bad.tsv"
    p.write_text("=numSamples=\n2\n=numCandidateMutatedSites=\n3\n=mutations=\nchr1\t5\tA\tT\tT,C,G;0,0,0,3\tT,C,G;1,0,0,3\n")
    with pytest.raises(ValueError):
        data.load_read_counts_tsv(str(p))


def test_sort_patterns_merges_duplicates():
    rng = np.random.default_rng(2)
    base = rng.integers(0, 5, size=(6, 3, 4)).astype(np.int32)
    base[..., 3] += base[..., :3].sum(-1)
    sites = base[[0, 1, 2, 1, 0, 3, 4, 5, 0]]
    pat, w, idx = data.sort_patterns(sites)
    assert w.sum() == 9 and len(pat) == len(np.unique(sites.reshape(9, -1), axis=0))
    for s in range(9):
        np.testing.assert_array_equal(pat[idx[s]], sites[s])
    # sorted lexicographically over (cell, [alt1, alt2, alt3, ref, cov])
    key = np.concatenate([pat[..., :3], (pat[..., 3] - pat[..., :3].sum(-1))[..., None], pat[..., 3:]], -1).reshape(len(pat), -1)
    assert all(tuple(key[i]) < tuple(key[i + 1]) for i in range(len(pat) - 1))


def test_tree_shapes_and_ops():
    t = caterpillar(5)
    assert t.node_count == 10 and t.root == 9 and t.right[9] == -1 and t.left[9] == 8
    np.testing.assert_allclose(t.height[5:], [1, 2, 3, 4, 5])
    ops = t.ops()
    assert len(ops) == 5 and tuple(ops[-1]) == (9, 8, -1)
    seen = set(range(5))
    for p, a, b in ops:
        assert a in seen and (b < 0 or b in seen)
        seen.add(p)
    t9 = fixture_9_leaves()
    lv_ops, off = t9.ops_by_level()
    lv = t9.levels()
    assert off[0] == 0 and off[-1] == len(lv_ops) == 9
    for i in range(len(off) - 1):
        assert len(set(lv[lv_ops[off[i]:off[i + 1], 0]])) == 1
    assert t9.dirty_closure([t9.parent[4]]) == [12, 13, 15, 16, 17]
    rt = random_coalescent(30, np.random.default_rng(0))
    assert (rt.branch_lengths()[:-1] > 0).all() and len(rt.ops()) == 30


def test_synthetic_generator_is_seeded_and_consistent():
    a = data.synthetic(8, 64, seed=3)
    b = data.synthetic(8, 64, seed=3)
    np.testing.assert_array_equal(a["counts"], b["counts"])
    c = a["counts"]
    assert c.shape == (64, 8, 4) and (c[..., :3].sum(-1) <= c[..., 3]).all() and (c >= 0).all()
    assert 0.0 < (c[..., 3] == 0).mean() < 0.2


def test_native_gamma_rates_match_scipy():
    from sieve_b200.mcmc import gamma_category_rates
    for a in (0.05, 0.5, 1.0, 3.7, 50.0):
        ref, _ = substmodel.gamma_category_rates(a, 4)
        np.testing.assert_allclose(gamma_category_rates(a, 4), ref, rtol=1e-9)


def test_ingest_shards_round_trip(tmp_path):
    from sieve_b200 import ingest
    rng = np.random.default_rng(5)
    counts = rng.integers(0, 200, size=(53, 7, 4)).astype(np.int32)
    counts[..., 3] += counts[..., :3].sum(-1)
    names = ["c%02d" % i for i in range(7)]
    sf = rng.random(7) + 0.5
    paths = ingest.write_shards(str(tmp_path / "d"), counts, 3, names, sf, 9727)
    got, begin = [], 0
    for r, p in enumerate(paths):
        meta, arr = ingest.open_shard(p)
        assert meta["site_begin"] == begin and meta["n_sites_total"] == 53 and meta["dtype"] == "uint16"
        assert meta["cell_names"] == names and meta["n_background_sites"] == 9727
        np.testing.assert_allclose(meta["size_factors"], sf)
        assert arr.shape == (7, meta["n_sites"], 4)
        got.append(np.asarray(arr).transpose(1, 0, 2))
        begin += meta["n_sites"]
    np.testing.assert_array_equal(np.concatenate(got), counts)
    big = counts.copy()
    big[0, 0, 3] = 70000
    meta = ingest.write_shard(str(tmp_path / "big.sieve"), big, 0, 53)
    assert meta["dtype"] == "int32"
    with pytest.raises(ValueError):
        ingest.write_shard(str(tmp_path / "bad.sieve"), -counts, 0, 53)
    with pytest.raises(ValueError):
        (tmp_path / "junk").write_bytes(b"hello")
        ingest.open_shard(str(tmp_path / "junk"))


def test_private_cov_tied_pair_update_matches_the_oracle():
    """sieve_b200.private_cov.update_pair (host half of PrivateAllelicSeqCovModel.updateAllelicSeqCovAndRawVar for tied
    pairs) against the loop restatement of oracle/private_allelic.py, incl. repeated and degenerate combinations."""
    from oracle import private_allelic as pa
    from sieve_b200 import private_cov
    rng = np.random.default_rng(3)
    n = 11
    for single in (True, False):
        modeled = [1, 2] if single else [0, 1, 2]
        for trial in range(40):
            sf = rng.uniform(0.4, 2.5, n)
            cov = rng.integers(0, 120, n)
            n_comb = int(rng.integers(1, 5))
            combos = []
            for _ in range(n_comb):
                ado = rng.integers(0, len(modeled), n)           # dropped alleles of each leaf's ML component
                if trial % 7 == 0:
                    ado[:] = 0                                   # one group only
                combos.append(tuple(int(a) for a in ado) + (0,) * (2 * n))
            if trial % 5 == 0:
                combos.append(combos[0])                          # a duplicate: weight 2
            pats = sorted(tuple(2 - c[j] for j in range(n)) for c in combos)
            uniq, weights = [], []
            for p in pats:
                if uniq and uniq[-1] == p:
                    weights[-1] += 1
                else:
                    uniq.append(p)
                    weights.append(1)
            want = pa.update_allelic_seq_cov_and_raw_var(cov, sf, modeled, [np.array(p) for p in uniq], weights, 7.5)
            got = private_cov.update_pair(cov, sf, combos, single, 7.5)
            assert abs(got[0] - want[0]) <= 1e-13 * abs(want[0]) + 1e-300
            assert abs(got[1] - want[1]) <= 1e-13 * abs(want[1]) + 1e-300


def test_ncu_traffic_file_belongs_to_the_sources_in_the_tree():
    """profiles/ncu_traffic.json (the per-launch DRAM bytes bench.py quotes as roofline.traffic) is stamped with the hash of
    the library's sources at capture time; bench.py refuses the figure when the stamp differs.  The committed file must be
    the capture of the committed sources (profiles/make_ncu_traffic.py regenerates it after any change under sieve_b200/csrc
    or include/)."""
    import json
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    tr = json.load(open(os.path.join(root, "profiles", "ncu_traffic.json")))
    assert tr["library_build_id"] == bench.library_build_id()
    k = tr["kernels"]["k_prune"]
    assert k["kernel"].startswith("k_walk") and 1.5 < k["duration_ms_under_ncu"] < 3.0
    assert 8e9 < k["dram_bytes_read"] + k["dram_bytes_write"] < 11e9
