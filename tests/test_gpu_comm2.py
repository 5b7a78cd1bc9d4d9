"""The per-evaluation exchange on two GPUs, one process per GPU: fused into k_reduce over peer memory (CUDA IPC mailboxes,
reduce.cuh SvMail) and, forced with SIEVE_B200_P2P=0, as an ncclAllGather -- both must hand every rank the two shard
results bit for bit, through several evaluations (the mailboxes alternate between two parities).  Needs two devices: skipped
on a one-GPU box (bench.py --gpus N covers the same path under torchrun)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, time, json
    import numpy as np
    sys.path.insert(0, %r)
    from sieve_b200 import core as core_mod, substmodel
    from sieve_b200.data import synthetic
    rank, world, idfile, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    d = synthetic(14, 600, seed=21)
    counts, tree = d["counts"], d["tree"]
    lo, hi = (0, 300) if rank == 0 else (300, 600)
    c = core_mod.ScsCudaLikelihoodCore(14, hi - lo, 4, 4, True, True, True, device=rank)
    c.set_counts(counts[lo:hi])
    c.set_size_factors(np.linspace(0.7, 1.4, 14))
    if rank == 0:
        uid = core_mod.comm_unique_id()
        with open(idfile + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(idfile + ".tmp", idfile)
    else:
        t0 = time.time()
        while not os.path.exists(idfile):
            assert time.time() - t0 < 120
            time.sleep(0.05)
        uid = open(idfile, "rb").read()
    c.comm_init(rank, world, uid)
    c.set_leaf_params(0.3, 50.0, 50.0, 0.008, 150.0, 2.0)
    c.update_leaves(for_update=False)
    rates, props = substmodel.gamma_category_rates(1.0, 4)
    c.set_root(tree.root, props, 0)
    nodes = np.array([i for i in range(tree.node_count) if i != tree.root], dtype=np.int32)
    rows = []
    for it in range(5):
        D = substmodel.node_distances(tree, rates * (1.0 + 0.01 * it))[nodes]
        r = c.step_distances(nodes, np.ascontiguousarray(D), np.ascontiguousarray(tree.ops(), dtype=np.int32), update_leaves=False)
        g = c.last_global_results()
        rows.append({"own": [r.sum_log_l, r.const_lse, r.const_sum, r.lewis_sum, r.weight_sum],
                     "world": [[x.sum_log_l, x.const_lse, x.const_sum, x.lewis_sum, x.weight_sum] for x in g]})
    json.dump({"mode": c.comm_exchange_mode(), "rows": rows}, open(out, "w"))
    c.close()
""") % ROOT


def _run_pair(tmp_path, tag, env_extra):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    idfile = str(tmp_path / ("id_" + tag))
    env = dict(os.environ, **env_extra)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", idfile, str(tmp_path / ("out_%s_%d.json" % (tag, r)))],
                              env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    import json
    return [json.load(open(tmp_path / ("out_%s_%d.json" % (tag, r)))) for r in range(2)]


@pytest.mark.gpu
def test_two_rank_exchange_fused_and_nccl_agree(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    fused = _run_pair(tmp_path, "p2p", {})
    nccl = _run_pair(tmp_path, "nccl", {"SIEVE_B200_P2P": "0"})
    assert [x["mode"] for x in nccl] == [1, 1]
    assert fused[0]["mode"] == fused[1]["mode"] and fused[0]["mode"] in (1, 2)   # 2 wherever the peers can be mapped
    for res in (fused, nccl):
        for it in range(5):
            w0, w1 = res[0]["rows"][it]["world"], res[1]["rows"][it]["world"]
            assert w0 == w1                                   # every rank holds the same table ...
            assert w0[0] == res[0]["rows"][it]["own"] and w0[1] == res[1]["rows"][it]["own"]   # ... of the ranks' own results
    assert [r["world"] for r in fused[0]["rows"]] == [r["world"] for r in nccl[0]["rows"]]
