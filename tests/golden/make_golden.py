"""Regenerates the committed fixtures under tests/golden/ (run in the build container, where /root/reference exists):

  cells10.npz / cells40.npz   the example read-count tensors of the reference (examples/data/{10,40}_cells),
                              parsed by sieve_b200.data.load_read_counts_tsv: counts int32 [S][n][4], n_background
  golden_cells10.json / golden_cells40.json
                              ORACLE outputs at the template start values (examples/templates/smc_stage_1.xml:30-38:
                              t=50, v=50, eps=0.008, w1=150, w2=2, theta=0.3, gammaShape=1, K=4) on the deterministic
                              caterpillar tree of ScsTree.makeCaterpillar(0, 1) (tree/ScsTree.java:96, 220-260) scaled by
                              0.01, with the P matrices recorded.

The reference's own tests hold no numeric vectors for this path (parity unpinned); these files freeze the
oracle's answers so that GPU tests need neither /root/reference nor a rebuild of the fixtures.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import sieve_oracle as so  # noqa: E402
from sieve_b200 import data, substmodel  # noqa: E402
from sieve_b200.tree import caterpillar  # noqa: E402

REF = "/root/reference/examples/data"


def main():
    for name in ("10_cells", "40_cells"):
        d = data.load_read_counts_tsv(os.path.join(REF, name, "read_counts.tsv"), os.path.join(REF, name, "cell_names"))
        counts = d["counts"]
        # the alignment evaluates sorted, weighted patterns (alignment/ScsAlignment.java:400-446)
        patterns, weights, _ = data.sort_patterns(counts)
        tag = "cells" + name.split("_")[0]
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), counts=counts, patterns=patterns, weights=weights,
                            n_background=np.int64(d["n_background_sites"]))
        S, n = patterns.shape[:2]
        tree = caterpillar(n, 0.0, 1.0)
        tree.height *= 0.01
        rates, props = substmodel.gamma_category_rates(1.0, 4)
        mats = substmodel.node_matrices(tree, rates)
        c5 = so.counts5(patterns)
        sf, stats = so.size_factors(c5, weights)
        P = so.leaf_params()
        M = float(d["n_background_sites"])
        out = dict(n=n, S=S, M=M, size_factors=sf.tolist(), cov_stats=stats.tolist(), matrices=mats.tolist(),
                   proportions=props.tolist(), rates=rates.tolist(), tree_scale=0.01)
        for asc in ("none", "felsenstein_mean", "felsenstein_geometric", "lewis"):
            logp, sl, sc = so.full_eval(c5, sf, P, tree.ops(), mats, props, tree.root, asc, M, 1, weights=weights)
            out["logP_" + asc] = logp
        out["site_logl"] = sl.tolist()
        out["site_log_const"] = sc.tolist()
        out["leaf0_log_partials"] = so.leaf_likelihood(c5, 0, sf[0], P).tolist()
        # threaded combine (2 workers) for the Felsenstein path
        out["logP_felsenstein_mean_T2"] = so.full_eval(c5, sf, P, tree.ops(), mats, props, tree.root,
                                                       "felsenstein_mean", M, 2, weights=weights)[0]
        # stage-1 background term (likelihood/ScsBackgroundLikelihood.java:169-268) from the file's five histograms
        out["background_hists"] = [{str(k): v for k, v in h.items()} for h in d["background"]]
        out["background_logP"] = so.background_log_likelihood(d["background"], d["n_background_sites"], n, 0.008, 150.0)
        with open(os.path.join(HERE, "golden_" + tag + ".json"), "w") as f:
            json.dump(out, f)
        print(tag, "S=%d patterns (of %d sites), n=%d, logP=%.10f" % (S, counts.shape[0], n, out["logP_none"]))


if __name__ == "__main__":
    main()
