"""Debug: the walk's time in different kinds of steps at the bench workload."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
W = bench.Workload(bench.N_CELLS, bench.N_SITES, bench.SEED, dev, 0, 1, mode="dense", want_matrices=False, n_sample=64)
core = W.core
core.enable_timing(True)
def show(tag):
    t = core.last_timing()
    print("%-40s prune %.3f leaf %.3f total %.3f launches %d elided %d" % (tag, t["prune_ms"], t["leaf_ms"], t["total_ms"], t["launches"], core.elided_count()))
W.exact_eval(); show("exact_eval")
for i in range(3):
    W.proposal_step(); show("proposal_step (leaf params)")
for i in range(3):
    core.step_distances(W.nodes, W.D_host * (1.0 + 1e-12 * (i + 1)), W.ops, update_leaves=False); show("tree-only, all matrices new")
for i in range(2):
    core.step_distances(W.nodes, W.D_host, W.ops, update_leaves=False); show("tree-only, D_host")
core.set_elision(False)
for i in range(3):
    core.step_distances(W.nodes, W.D_host, W.ops, update_leaves=False); show("tree-only, same D, elision off")
for i in range(2):
    W.proposal_step(); show("proposal_step, elision off")
core.set_elision(True)
W.close()
