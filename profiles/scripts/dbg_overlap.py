"""Debug: is leaf + walk slower than the sum of the two alone?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
W = bench.Workload(bench.N_CELLS, bench.N_SITES, bench.SEED, dev, 0, 1, mode="dense", want_matrices=False, n_sample=64)
core = W.core
st = torch.cuda.ExternalStream(core.stream(), device=dev)
def timed(fn, n=20):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(st): e0.record()
    fn(n)
    with torch.cuda.stream(st): e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
W.exact_eval()
for rep in range(2):
    W.proposal_step()
    full = timed(lambda n: core.replay(True, n))
    core.step_distances(W.nodes, W.D_host * (1 + 1e-12), W.ops, update_leaves=False)
    tree = timed(lambda n: core.replay(False, n))
    def leaves(n):
        for _ in range(n): core.update_leaves(True)
    leaf = timed(leaves)
    print("rep %d: leaf+walk %.3f ms   walk alone %.3f ms   leaf alone (host loop) %.3f ms   sum %.3f" % (rep, full, tree, leaf, tree + leaf))
# alternate: walk, walk, leaf, walk ... using replay granularity 1
def alt(n):
    for _ in range(n):
        core.replay(False, 1); core.update_leaves(True)
print("alternating walk / leaf via separate calls: %.3f ms per pair" % timed(alt))
W.close()
