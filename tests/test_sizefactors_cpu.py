"""The distributed size-factor protocol (include/sieve_b200.h: sieve_sf_solve) on CPU.

`sieve_sf_solve` is host logic over callbacks: per round every process reports, per cell, the weight of its coverage
ratios below a pivot; the counts are summed over processes; 63 rounds of bisection on the ratio's bit pattern give the
exact weighted median (seqcovmodel/SeqCovModelInterface.java:291-351, computed over the FULL alignment before the split,
likelihood/ThreadedScsTreeLikelihood.java:119-131).  On a GPU box the local primitives are device kernels
(tests/test_gpu_sizefactors.py); here they are stand-ins built from the oracle's own geometric mean, so the protocol --
targets, bisection, termination, the all-reduce wiring over gloo with 2 and 3 ranks -- is checked bit for bit against the
oracle's sort-based size factors of the unsharded data."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _HostPart:
    """Stand-in for the device primitives of one part (a site range): same definitions, evaluated with numpy."""

    def __init__(self, so, counts4, weights, zero_cov_mode):
        c5 = so.counts5(counts4)
        cov = c5[..., 4].astype(np.int64) + (1 if zero_cov_mode == 1 else 0)      # [S][n]
        S, n = cov.shape
        L = so.lib()
        L.so_geometric_mean.restype = C.c_double
        L.so_geometric_mean.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int]
        geo = np.empty(S)
        for s in range(S):
            row = np.ascontiguousarray(cov[s], dtype=np.int32)
            geo[s] = L.so_geometric_mean(row.ctypes.data_as(C.POINTER(C.c_int)), 0, n)
        keep = np.ones((S, n), dtype=bool) if zero_cov_mode == 1 else (cov > 0) & (geo[:, None] > 0)
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = cov / geo[:, None]
        self.keys = np.where(keep, ratio, np.inf).view(np.uint64)                 # bit patterns; +inf never counts
        w = np.ones(S, dtype=np.int64) if weights is None else np.asarray(weights, dtype=np.int64)
        self.w = np.where(keep, w[:, None], 0)
        self.mom = (int(cov.sum()), int((cov * cov).sum()), int(S * n))
        self.n = n

    def moments(self):
        return self.mom

    def weight_le(self, piv):
        return ((self.keys[:, :, None] <= piv[None, :, :]) * self.w[:, :, None]).sum(axis=0)


def _data(seed, n, S, with_weights):
    from sieve_b200.data import synthetic
    d = synthetic(n, S, seed=seed, zero_cov_frac=0.15)
    counts = d["counts"]
    counts[3, :, :] = 0                    # a site nobody covers: geometric mean 0, skipped everywhere
    counts[:, 2, :] = 0                    # a cell without any coverage: no ratio at all -> NaN
    w = None
    if with_weights:
        w = np.random.default_rng(seed).integers(1, 4, size=S).astype(np.int32)
    return counts, w


@pytest.mark.parametrize("zero_cov_mode,with_weights,n_parts", [(0, False, 1), (0, True, 3), (1, False, 2), (0, False, 4)])
def test_protocol_single_process_matches_oracle_sort(oracle, zero_cov_mode, with_weights, n_parts):
    from sieve_b200 import sizefactors as sfm
    from sieve_b200.core import pattern_points
    counts, w = _data(31, 12, 157, with_weights)
    ref, ref_st = oracle.size_factors(oracle.counts5(counts), w, zero_cov_mode)
    pts = pattern_points(counts.shape[0], n_parts)
    parts = [_HostPart(oracle, counts[pts[i]:pts[i + 1]], None if w is None else w[pts[i]:pts[i + 1]], zero_cov_mode)
             for i in range(n_parts)]
    calls = [0]

    def moments():
        m = [p.moments() for p in parts]
        return [sum(x[i] for x in m) for i in range(3)]

    def weight_le(piv):
        calls[0] += 1
        return sum(p.weight_le(piv) for p in parts)
    sf, st = sfm.solve_with_ops(12, moments, weight_le)
    assert np.array_equal(np.isnan(sf), np.isnan(ref))
    assert np.isnan(sf[2]) == (zero_cov_mode == 0)
    ok = ~np.isnan(ref)
    assert np.array_equal(sf[ok].view(np.uint64), ref[ok].view(np.uint64)), (sf, ref)      # bit for bit
    np.testing.assert_allclose(st, ref_st, rtol=1e-12)
    assert calls[0] <= 65                                                                     # one totals pass + <= 64 rounds


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from oracle import sieve_oracle as so
    from sieve_b200 import sizefactors as sfm
    from sieve_b200.distributed import shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counts, w = _data(47, 9, 203, True)
    b, e = shard_range(rank, world, counts.shape[0])
    part = _HostPart(so, counts[b:e], w[b:e], 0)
    sf, st = sfm.solve_with_ops(9, part.moments, part.weight_le, sfm.torch_allreduce())
    ref, ref_st = so.size_factors(so.counts5(counts), w, 0)
    q.put((rank, sf.tobytes(), ref.tobytes(), st.tolist(), ref_st.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_protocol_over_gloo_ranks_is_bit_identical_to_unsharded(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, sf, ref, st, ref_st in res:
        a, b = np.frombuffer(sf), np.frombuffer(ref)
        ok = ~np.isnan(b)
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.array_equal(a[ok].view(np.uint64), b[ok].view(np.uint64)), rank    # every rank: the unsharded answer
        np.testing.assert_allclose(st, ref_st, rtol=1e-12)
    assert len({r[1] for r in res}) == 1
