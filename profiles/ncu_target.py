"""One full evaluation + one max-sum trace at the bench workload (1 000 cells x 100 000 sites), nothing else: the
program `ncu --set full` is pointed at (profiles/README.md has the command).  Run it plain first; it prints the
per-phase wall times as a sanity check."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sieve_b200 import substmodel  # noqa: E402
from sieve_b200.core import ScsCudaLikelihoodCore  # noqa: E402


def main():
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    # default: the bench workload; `ncu_target.py 10000 125000 sparse` = the per-GPU share of the north-star data set
    n, S = bench.N_CELLS, bench.N_SITES
    mode = "dense"
    if len(sys.argv) >= 3:
        n, S = int(sys.argv[1]), int(sys.argv[2])
        mode = sys.argv[3] if len(sys.argv) > 3 else "dense"
    tree, rates, props, br, mats = bench.make_tree_and_matrices(bench.SEED, n)
    counts_dev = bench.generate_counts_device(tree, rates, br, n, S, bench.SEED, dev)
    torch.cuda.synchronize()
    from sieve_b200.sizefactors import SizeFactorPart, solve_parts
    part = SizeFactorPart(n, S, counts_dev.data_ptr(), None, 0, 0)   # pool-free size factors, before the core takes the HBM
    sf, _ = solve_parts([part], None)
    part.close()
    core = ScsCudaLikelihoodCore(n, S, bench.K_CAT, 4, True, True, True, 0, sparse=(mode == "sparse"), external_counts=True)
    core.set_counts_device(counts_dev.data_ptr())   # external counts: the tensor stays alive until the core is closed
    core.set_size_factors(sf)
    core.set_leaf_params(*bench.PARAMS)
    core.update_leaves(for_update=False)
    core.set_root(tree.root, props, 0)
    nodes = np.array([i for i in range(tree.node_count) if i != tree.root], dtype=np.int32)
    D = np.ascontiguousarray(substmodel.node_distances(tree, rates, branch_rates=br)[nodes])
    ops = np.ascontiguousarray(tree.ops(), dtype=np.int32)
    r = core.step_distances(nodes, D, ops, update_leaves=False)
    for rep in range(2):   # the second pass is the one to look at (first-touch effects gone)
        t0 = time.perf_counter()
        pr = list(bench.PARAMS)
        pr[3] *= 1.0 + 1e-9 * (rep + 1)
        core.set_leaf_params(*pr)
        r = core.step_distances(nodes, D, ops, update_leaves=True)
        t1 = time.perf_counter()
        if mode == "dense" and n <= 2000:
            core.ml_trace()
        t2 = time.perf_counter()
        print("pass %d: evaluation %.2f ms, max-sum trace %.2f ms, sum logL %.6f" % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), r.sum_log_l))
    core.close()


if __name__ == "__main__":
    main()
