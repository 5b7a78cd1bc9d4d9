import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_CELLS
S = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
W = bench.Workload(n, S, bench.SEED, dev, 0, 1, mode="dense", want_matrices=False, n_sample=64)
core = W.core
W.exact_eval()
for i in range(3):
    t0 = time.perf_counter(); W.proposal_step(); t1 = time.perf_counter()
    print("proposal_step wall %.3f ms" % (1e3 * (t1 - t0)), file=sys.stderr)
for i in range(3):
    t0 = time.perf_counter(); core.step_distances(W.nodes, W.D_host * (1.0 + 1e-12 * (i + 1)), W.ops, update_leaves=False); t1 = time.perf_counter()
    print("tree-only step wall %.3f ms" % (1e3 * (t1 - t0)), file=sys.stderr)
W.close()
