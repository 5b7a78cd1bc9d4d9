"""The N > 1 path on CPU: world_size 2 and 3 with gloo.  Each rank derives its site range like calcPatternPoints, produces
its shard's result struct (here with the oracle, since there is no GPU), all-gathers 5 doubles and combines; the answer
must equal the unsharded evaluation and be identical on every rank."""
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


def _worker(rank, world, port, asc, q):
    import torch.distributed as dist

    from oracle import sieve_oracle as so
    from sieve_b200 import _lib, substmodel
    from sieve_b200.data import synthetic
    from sieve_b200.distributed import global_log_p, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = synthetic(9, 101, seed=5)
    counts, tree = d["counts"], d["tree"]
    c5 = so.counts5(counts)
    sf, _ = so.size_factors(c5)                     # global size factors, computed before the split
    rates, props = substmodel.gamma_category_rates(1.0, 4)
    mats = substmodel.node_matrices(tree, rates)
    b, e = shard_range(rank, world, counts.shape[0])
    _, sl, sc = so.full_eval(c5[b:e], sf, so.leaf_params(), tree.ops(), mats, props, tree.root, "none", 0.0, 1)
    local = _lib.ShardResult(sl.sum(), so.log_sum_exp(sc), sc.sum(), so.asc_bias(sc, None, "lewis", 0.0), float(e - b))
    M = 4000.0
    got = global_log_p(local, asc, M)
    ref, _, _ = so.full_eval(c5, sf, so.leaf_params(), tree.ops(), mats, props, tree.root, asc, M, 1)
    q.put((rank, got, ref, b, e))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,asc", [(2, "felsenstein_mean"), (2, "none"), (3, "lewis")])
def test_gloo_shards_combine(world, asc):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, asc, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert [r[3] for r in res][0] == 0 and res[-1][4] == 101
    assert all(res[i][4] == res[i + 1][3] for i in range(world - 1))      # contiguous, no overlap
    assert len({r[1] for r in res}) == 1                                  # identical on every rank
    assert abs(res[0][1] - res[0][2]) <= 1e-11 * abs(res[0][2])
