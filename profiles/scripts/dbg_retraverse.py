"""debug: chain A (partial retraverse) vs chain B (whole-shard pass) after accept()"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import sieve_oracle as so
from sieve_b200.data import synthetic
from sieve_b200.treelikelihood import ScsTreeLikelihood

n, S = 19, 333
d = synthetic(n, S, seed=11)
counts, tree = d["counts"], d["tree"]
sf, _ = so.size_factors(so.counts5(counts))
mk = lambda: ScsTreeLikelihood(counts, tree.copy(), size_factors=sf, asc_mode="felsenstein_mean", n_background_sites=5e3, private_seq_cov=True)
A, B = mk(), mk()
B.retraverse_changed_only = False
print("init", A.calculate_logp() == B.calculate_logp())
for it in range(4):
    tA0 = A.core.get_private_seq_cov()
    ma, mb = A.accept(), B.accept()
    tA, tB = A.core.get_private_seq_cov(), B.core.get_private_seq_cov()
    ch = ((tA[0] != tA0[0]) | (tA[1] != tA0[1])).any(axis=0)
    la, ca = A.pattern_log_likelihoods(); lb, cb = B.pattern_log_likelihoods()
    dl = la != lb; dc = ca != cb
    print("it", it, "moved", ma, mb, "sites listed", A.last_retraversed_sites, "changed sites", int(ch.sum()), "tv equal", np.array_equal(tA[0], tB[0]) and np.array_equal(tA[1], tB[1]),
          "logP", A.logP, B.logP, "site diffs", int(dl.sum()), "const diffs", int(dc.sum()),
          "diff sites in changed", int((dl & ch).sum()), "max |d|", float(np.max(np.abs(la - lb))), float(np.max(np.abs(ca - cb))))
    if dl.any():
        i = np.flatnonzero(dl)[:5]
        print("   sites", i, la[i], lb[i])
    for node in (0, tree.n, tree.n + 3, tree.root - 1, tree.root):
        pa_, pb_ = A.core.get_node_partials(node), B.core.get_node_partials(node)
        print("   node", node, "partials equal", np.array_equal(pa_, pb_), float(np.nanmax(np.abs(pa_ - pb_))))

print("---- with tree moves in between ----")
from oracle import private_allelic as pa
from sieve_b200 import substmodel
A.close(); B.close()
A, B = mk(), mk()
B.retraverse_changed_only = False
A.calculate_logp(); B.calculate_logp()
rng = np.random.default_rng(3)
c5 = so.counts5(counts)

def vs_oracle(tl, tag):
    t_fin, v_fin = tl.core.get_private_seq_cov()
    P = so.leaf_params(*tl.params.as_tuple(), single_ado=True)
    mats = substmodel.node_matrices(tl.tree, tl.rates)
    sl, lc, _ = pa.full_eval_private(so, c5, sf, P, t_fin, v_fin, tl.tree.ops(), mats, tl.proportions, tl.tree.root)
    gl, gc = tl.pattern_log_likelihoods()
    bad = np.flatnonzero(np.abs(gl - sl) > 1e-10 * np.abs(sl))
    print("   ", tag, "vs oracle: max rel site", float(np.max(np.abs(gl - sl) / np.abs(sl))), "const", float(np.max(np.abs(gc - lc) / np.abs(lc))), "bad sites", bad[:8], len(bad))
    return bad

for it in range(5):
    ma, mb = A.accept(), B.accept()
    print("it", it, "accept moved", ma, mb, "listed", A.last_retraversed_sites, "logP equal", A.logP == B.logP, A.logP, B.logP)
    badA = vs_oracle(A, "A"); vs_oracle(B, "B")
    t2 = A.tree.copy()
    node = int(rng.integers(tree.n, tree.node_count - 1))
    lo = max(t2.height[t2.left[node]], t2.height[t2.right[node]])
    t2.height[node] = lo + (t2.height[t2.parent[node]] - lo) * rng.random()
    for tl in (A, B):
        tl.store(); tl.set_tree(t2.copy(), [node]); tl.calculate_logp()
    print("   move node", node, "logP equal", A.logP == B.logP)
    vs_oracle(A, "A after move"); vs_oracle(B, "B after move")
    if it % 2 == 1:
        for tl in (A, B):
            tl.restore(); tl.calculate_logp()
        print("   restored, logP equal", A.logP == B.logP)
        vs_oracle(A, "A restored"); vs_oracle(B, "B restored")
