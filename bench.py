#!/usr/bin/env python
"""bench.py -- full-tree log-likelihood evaluations per second of the SIEVE hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the largest single-GPU configuration): synthetic 1,000 cells x 100,000 candidate
sites, K = 4 gamma categories, acquisition-bias correction on (rmc_stage_2 semantics).  One step = one complete
evaluation of the hot path: leaf read-count likelihoods of every (cell, site) (a parameter proposal), Felsenstein pruning
of every internal node with fresh transition matrices (a relaxed-clock / scaler proposal dirties the whole tree), root
integration, per-site log and reduction, acquisition-bias term.  At N > 1 every rank holds a shard of that same size
(weak scaling, sites are independent); the only exchange is an all-gather of 5 doubles per rank per step.

value  : steps/s with matrices and ops already staged in HBM (kernels only, CUDA events on the handle's stream).
e2e    : the same through sieve_step() with HOST buffers: host->device copy of matrices + ops and device->host read of
         the result inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_CELLS = int(os.environ.get("SIEVE_BENCH_CELLS", 1000))
N_SITES = int(os.environ.get("SIEVE_BENCH_SITES", 100000))
K_CAT = 4
N_BACKGROUND = 1.0e7
PARAMS = (0.3, 50.0, 50.0, 0.008, 150.0, 2.0)   # theta, t, v, eps, w1, w2 (examples/templates/smc_stage_1.xml:30-38)
SEED = 20260101
METRIC = "loglik_evals_per_sec"
UNIT = ("evals/s (one eval = leaf + full-tree pruning + root + reduce over {:,} cells x {:,} sites; at N GPUs each rank "
        "evaluates one such shard per step)").format(N_CELLS, N_SITES)


# ----------------------------------------------------------------------------------------------------------------
# synthetic data on the device (same generative model as sieve_b200.data.synthetic, vectorised in torch)
# ----------------------------------------------------------------------------------------------------------------
def make_tree_and_matrices(seed, n):
    from sieve_b200 import substmodel
    from sieve_b200.tree import random_coalescent
    rng = np.random.default_rng(seed)
    tree = random_coalescent(n, rng)
    tree.height *= 0.3 / tree.height[tree.root]
    rates, props = substmodel.gamma_category_rates(1.0, K_CAT)
    # relaxed clock: per-branch rate multipliers LogNormal(mean 1, sigma 0.2) folded into P on the host
    br = rng.lognormal(-0.02, 0.2, size=tree.node_count)
    mats = substmodel.node_matrices(tree, rates, branch_rates=br)
    return tree, rates, props, br, mats


def generate_counts_device(tree, rates, br, n, S, seed, device, chunk=None):
    """int32 [cell][site][4] on `device`, generated in site chunks so that the temporaries stay small."""
    import torch

    from sieve_b200 import substmodel
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    torch.manual_seed(seed)          # _standard_gamma draws from the global generator
    torch.cuda.manual_seed_all(seed)
    theta, t, v, eps, w1, w2 = PARAMS
    bl = tree.branch_lengths()
    order = tree.post_order()[::-1]
    cums = {}
    for p in order:
        for c in (tree.left[p], tree.right[p]):
            if c >= 0:
                cums[c] = torch.tensor(np.stack([np.cumsum(substmodel.transition_probabilities(bl[c] * br[c] * rates[k]),
                                                           axis=1) for k in range(K_CAT)]), device=device)   # [K][4][4]
    sf = torch.exp(0.3 * torch.randn(n, device=device, generator=g, dtype=torch.float64))
    f = eps / 3.0
    table = torch.tensor([[f * w1, f * w1, f * w1, (1 - eps) * w1],          # 0: 0/0
                          [(1 - eps) * w1, f * w1, f * w1, f * w1],          # 1: 1/1
                          [(.5 - f) * w2, (.5 - f) * w2, f * w2, f * w2],    # 2: 1/1'
                          [(.5 - f) * w2, f * w2, f * w2, (.5 - f) * w2],    # 3: 0/1
                          [f * w1, (1 - eps) * w1, f * w1, f * w1]],         # 4: 1'/1' (one allele of 1/1' left)
                         device=device, dtype=torch.float32)
    out = torch.zeros((n, S, 4), dtype=torch.int32, device=device)
    if chunk is None:
        chunk = max(256, min(S, int(2.0e8 // n)))
    for s0 in range(0, S, chunk):
        ns = min(chunk, S - s0)
        cat = torch.randint(0, K_CAT, (ns,), device=device, generator=g)
        geno = torch.zeros((tree.node_count, ns), dtype=torch.int64, device=device)
        for p in order:
            for c in (tree.left[p], tree.right[p]):
                if c < 0:
                    continue
                rows = cums[c][cat, geno[p]]                                              # [ns][4]
                u = torch.rand(ns, device=device, generator=g, dtype=torch.float64)
                geno[c] = (u[:, None] > rows[:, :3]).sum(1)
        gl = geno[:n]                                                                     # [n][ns]
        del geno
        ado = torch.rand((n, ns), device=device, generator=g) < theta
        mean = torch.where(ado, 1.0, 2.0).to(torch.float64) * t * sf[:, None]
        r = t * t / v
        lam = torch._standard_gamma(torch.full((n, ns), r, device=device, dtype=torch.float64)) * (mean / r)
        cov = torch.poisson(lam, generator=g)
        cov[torch.rand((n, ns), device=device, generator=g) < 0.05] = 0.0
        del lam, mean
        coin = torch.rand((n, ns), device=device, generator=g) < 0.5
        idx = torch.zeros((n, ns), dtype=torch.int64, device=device)
        idx[(gl == 1) & ~ado] = 3
        idx[(gl == 1) & ado & coin] = 1
        idx[gl == 2] = 1
        idx[(gl == 3) & ~ado] = 2
        idx[(gl == 3) & ado & coin] = 1
        idx[(gl == 3) & ado & ~coin] = 4
        del gl, ado, coin
        # Dirichlet-multinomial by sequential beta-binomials
        gam = torch._standard_gamma(table[idx].clamp_min(1e-3))                           # [n][ns][4] float32
        del idx
        rest = gam.sum(-1).to(torch.float64)
        remaining = cov.clone()
        o = out[:, s0:s0 + ns, :]
        for i in range(3):
            gi = gam[..., i].to(torch.float64)
            q = (gi / rest.clamp_min(1e-300)).clamp(0.0, 1.0)
            x = torch.binomial(remaining, q, generator=g)
            o[..., i] = x.to(torch.int32)
            remaining -= x
            rest -= gi
        o[..., 3] = cov.to(torch.int32)
        del gam, rest, remaining, cov
    return out


# ----------------------------------------------------------------------------------------------------------------
def algorithmic_bytes_prune(tree, S, with_const):
    """per (internal node, site) write K*G*8 = 128 B (+128 B constant-site twin), read 128 B per
    internal child (+128 B twin), 32 B per leaf child (category independent; the constant twin adds nothing)."""
    total = 0
    twin = 2 if with_const else 1
    for p, c1, c2 in tree.ops():
        b = 128 * twin
        for c in (c1, c2):
            if c < 0:
                continue
            b += 32 if c < tree.n else 128 * twin
        total += b
    return total * S


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(o[0]))
                self.max_mhz = float(o[1])
                for nm, val in zip(names, o[2:]):
                    if val.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_eval_sample(tree, mats, props, counts_sample, sf, threads, target_s=12.0):
    """Times the restated reference core (oracle) on a bounded site sample with the reference's own thread split."""
    from oracle import sieve_oracle as so
    c5 = so.counts5(counts_sample)
    P = so.leaf_params(*PARAMS)
    ops = tree.ops()
    probe = min(c5.shape[0], max(threads, 2 * threads))
    t0 = time.perf_counter()
    so.full_eval(c5[:probe], sf, P, ops, mats, props, tree.root, "felsenstein_mean", N_BACKGROUND, threads, want_sites=False)
    dt = time.perf_counter() - t0
    n_sites = int(min(c5.shape[0], max(probe, probe * target_s / max(dt, 1e-6))))
    n_sites = max(threads, n_sites // threads * threads)
    t0 = time.perf_counter()
    logp, _, _ = so.full_eval(c5[:n_sites], sf, P, ops, mats, props, tree.root, "felsenstein_mean", N_BACKGROUND, threads,
                              want_sites=False)
    dt = time.perf_counter() - t0
    return n_sites, dt, logp


def run_reference(args):
    """Reference arm: the restated reference CPU core (the Java reference cannot run: no JVM/BEAST in the image) on a
    bounded site sample of the same workload, all host threads; throughput scaled to the full 100,000 sites."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import sieve_oracle as so
    from sieve_b200.data import synthetic
    so.build()
    threads = so.max_threads()
    tree, rates, props, br, mats = make_tree_and_matrices(SEED, N_CELLS)
    P = so.leaf_params(*PARAMS)
    ops = tree.ops()
    per_thread = int(os.environ.get("SIEVE_BENCH_REF_SITES_PER_THREAD", 0))
    if per_thread <= 0:
        # calibrate: a probe of 16 sites per thread, then a sample sized for ~1.5 s per step and <= ~150 s for the run
        probe = synthetic(N_CELLS, threads * 16, seed=SEED + 1, tree=tree)
        pc5 = so.counts5(probe["counts"])
        psf, _ = so.size_factors(pc5)
        t0 = time.perf_counter()
        so.full_eval(pc5, psf, P, ops, mats, props, tree.root, "felsenstein_mean", N_BACKGROUND, threads, want_sites=False)
        dt = max(time.perf_counter() - t0, 1e-3)
        budget = min(1.5, 150.0 / max(1, args.steps + args.warmup))
        per_thread = int(min(4096, max(16, 16 * budget / dt)))
    n_sample = threads * per_thread
    d = synthetic(N_CELLS, n_sample, seed=SEED, tree=tree)
    c5 = so.counts5(d["counts"])
    sf, _ = so.size_factors(c5)

    def step():
        return so.full_eval(c5, sf, P, ops, mats, props, tree.root, "felsenstein_mean", N_BACKGROUND, threads,
                            want_sites=False)[0]
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    sample_steps_per_s = args.steps / dt
    value = sample_steps_per_s * n_sample / N_SITES
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps * N_SITES / n_sample, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic %d cells x %d sites, K=4, felsenstein bias correction (configs[2])" % (N_CELLS, N_SITES)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d of %d sites per step (all %d cells), site ranges over %d threads like "
                                   "ThreadedScsTreeLikelihood; evals/s scaled by %d/%d (sites are independent)"
                                   % (n_sample, N_SITES, N_CELLS, threads, n_sample, N_SITES)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--per-level", action="store_true", help="issue the pruning as one launch per tree level")
    ap.add_argument("--mcmc-states", type=int, default=400, help="proposals per operator mix in the MCMC replay (0 = skip)")
    ap.add_argument("--streamed", action="store_true", help="SIEVE_FLAG_EPHEMERAL: no resident partials (shapes that exceed HBM, e.g. SIEVE_BENCH_CELLS=10000 SIEVE_BENCH_SITES=125000)")
    ap.add_argument("--sparse", action="store_true", help="SIEVE_FLAG_SPARSE: leaves resident, only the larger subtrees keep their partials (10 000-cell shapes)")
    ap.add_argument("--single-leaf-buffer", action="store_true", help="SIEVE_FLAG_SINGLE_LEAF_BUFFER: leaves are not double-buffered (rejected leaf-parameter proposals recompute them)")
    ap.add_argument("--no-variants", action="store_true", help="skip the max-sum (variant calling) timing")
    ap.add_argument("--matrices", action="store_true", help="hand the core 16-double matrices (generic kernel) instead of branch distances")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("SIEVE_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    from sieve_b200 import _lib
    from sieve_b200.core import ScsCudaLikelihoodCore, combine_shards
    _lib.lib()

    n, S = N_CELLS, N_SITES
    tree, rates, props, br, mats = make_tree_and_matrices(SEED, n)
    counts_dev = generate_counts_device(tree, rates, br, n, S, SEED + 17 * rank, dev)
    torch.cuda.synchronize()

    counts_host = None
    via_host = args.sparse or n * S * 16 > 40e9   # (also when two copies of the counts would not fit next to the pools)
    if via_host:
        # a sparse core sizes its partial pool from the memory that is free when it is created: give the generator's
        # memory back first and load the counts from the host (as a production run would, from the ingest shard file)
        counts_host = counts_dev.cpu().numpy()
        del counts_dev
        torch.cuda.empty_cache()
    core = ScsCudaLikelihoodCore(n, S, K_CAT, 4, True, True, True, local, ephemeral=args.streamed, sparse=args.sparse,
                                 single_leaf_buffer=args.single_leaf_buffer)
    if via_host:
        core.set_counts_cellmajor(counts_host)
        counts_dev = torch.from_numpy(counts_host[:, :min(S, 4096), :].copy()).to(dev)   # only the sample below is read
        del counts_host
    else:
        core.set_counts_device(counts_dev.data_ptr())
    # a host sample of this rank's sites for the parity spot check and the CPU baseline
    n_keep = min(S, 64 if (args.no_cpu_baseline or world > 1) else max(64, int(2.4e7 // n)))
    sample = counts_dev[:, :n_keep, :].permute(1, 0, 2).contiguous().cpu().numpy()
    del counts_dev
    torch.cuda.empty_cache()
    # global size factors: every rank computes them over its own shard here (weak scaling: independent data sets)
    sf, stats = core.compute_size_factors(0)
    core.set_leaf_params(*PARAMS)
    core.update_leaves(for_update=False)
    core.set_root(tree.root, props, 0)
    nodes = np.array([i for i in range(tree.node_count) if i != tree.root], dtype=np.int32)
    P_host = np.ascontiguousarray(mats[nodes])
    from sieve_b200 import substmodel
    D_host = np.ascontiguousarray(substmodel.node_distances(tree, rates, branch_rates=br)[nodes])
    if args.per_level:
        ops, level_off = tree.ops_by_level()
    else:
        ops, level_off = tree.ops(), None
    ops = np.ascontiguousarray(ops, dtype=np.int32)

    step_no = [0]

    def one_step_e2e():
        # what the plugin does per MCMC state: leaf parameters changed + every matrix changed -> full evaluation.
        # The proposal alternates between a coverage parameter (NB tables rebuilt) and a nucleotide parameter (DM tables).
        step_no[0] += 1
        pr = list(PARAMS)
        if step_no[0] % 2:
            pr[1] = PARAMS[1] * (1.0 + 1e-9 * (step_no[0] % 7))
        else:
            pr[3] = PARAMS[3] * (1.0 + 1e-9 * (step_no[0] % 7))
        core.set_leaf_params(*pr)
        if args.per_level:
            core.update_leaves(for_update=True)
            if args.matrices:
                core.set_matrices(nodes, P_host, for_update=True)
            else:
                core.set_distances(nodes, D_host, for_update=True)
            core.update_nodes(ops, level_off, flip=True, per_level=True)
            return core.evaluate()
        if args.matrices:
            return core.step(nodes, P_host, ops, update_leaves=True)
        return core.step_distances(nodes, D_host, ops, update_leaves=True)

    gather_buf = torch.zeros((world, 5), dtype=torch.float64, device=dev)
    res_ptr = _lib.C.c_void_p()
    _lib.check(core.L.sieve_get_result_device_ptr(core.h, _lib.C.byref(res_ptr)))
    lib_stream = torch.cuda.ExternalStream(core.stream(), device=dev)

    class _DevView:
        """zero-copy torch view of the library's 5-double result struct in HBM"""
        def __init__(self, ptr):
            self.__cuda_array_interface__ = {"shape": (5,), "typestr": "<f8", "data": (ptr, False), "version": 2}
    res_view = torch.as_tensor(_DevView(res_ptr.value), device=dev)

    def exchange():
        """the per-step collective: all-gather of the 5 result doubles of every rank (NCCL over NVLink), ordered after
        the reduce kernel on the library's stream; the next step's kernels are ordered after the collective."""
        if world == 1:
            return
        with torch.cuda.stream(lib_stream):
            dist.all_gather_into_tensor(gather_buf.view(-1), res_view)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def exact_eval():
        core.set_leaf_params(*PARAMS)
        if args.matrices or args.per_level:
            core.update_leaves(for_update=True)
            if args.matrices:
                core.set_matrices(nodes, P_host, for_update=True)
            else:
                core.set_distances(nodes, D_host, for_update=True)
            core.update_nodes(ops, level_off, flip=True, per_level=args.per_level)
            return core.evaluate()
        return core.step_distances(nodes, D_host, ops, update_leaves=True)

    logp_first = combine_shards([exact_eval()], "felsenstein_mean", N_BACKGROUND)
    if os.environ.get("SIEVE_BENCH_DEBUG"):
        r_ = exact_eval()
        print("debug first again", logp_first, combine_shards([r_], "felsenstein_mean", N_BACKGROUND), r_.sum_log_l, r_.const_lse, file=sys.stderr)
    # ---------------- warm-up (also stages matrices + ops in HBM for the device-only loop) ----------------
    for _ in range(max(args.warmup, 3)):
        r = one_step_e2e()
        exchange()
    barrier()

    # ---------------- timed region 1: device path, inputs resident in HBM ----------------
    launches0 = core.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(lib_stream):
        ev0.record()
    if world == 1:
        core.replay(update_leaves=True, n_times=args.steps)      # K steps enqueued by one native call: no host in the loop
    else:
        for _ in range(args.steps):
            core.replay(update_leaves=True, n_times=1)
            exchange()
    with torch.cuda.stream(lib_stream):
        ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = core.launch_count() - launches0
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())

    # ---------------- timed region 2: end to end through the C-ABI with host buffers ----------------
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = one_step_e2e()
        exchange()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s_max = float(t.item())

    # ---------------- per-kernel timing (CUDA events on the handle's stream, inside the library) ----------------
    core.enable_timing(True)
    tm = []
    for _ in range(5):
        one_step_e2e()
        tm.append(core.last_timing())
    core.enable_timing(False)
    leaf_ms = float(np.median([x["leaf_ms"] for x in tm]))
    prune_ms = float(np.median([x["prune_ms"] for x in tm]))
    reduce_ms = float(np.median([x["reduce_ms"] for x in tm]))
    # tree-only evaluation (stage-2 style proposal: matrices change, leaves do not)
    if args.matrices or args.per_level:
        core.set_matrices(nodes, P_host, for_update=True) if args.matrices else core.set_distances(nodes, D_host, for_update=True)
        core.update_nodes(ops, level_off, flip=True, per_level=args.per_level)
        core.evaluate()
    else:
        core.step_distances(nodes, D_host, ops, update_leaves=False)
    torch.cuda.synchronize()
    ev0.record(lib_stream)
    core.replay(update_leaves=False, n_times=args.steps)
    ev1.record(lib_stream)
    torch.cuda.synchronize()
    tree_only_ms = ev0.elapsed_time(ev1) / args.steps

    r_last = exact_eval()
    logp_last = combine_shards([r_last], "felsenstein_mean", N_BACKGROUND)
    if os.environ.get("SIEVE_BENCH_DEBUG"):
        r_ = exact_eval()
        print("debug last again", logp_last, combine_shards([r_], "felsenstein_mean", N_BACKGROUND), r_last.sum_log_l, r_last.const_lse, r_.sum_log_l, r_.const_lse, file=sys.stderr)
    # ---------------- parity spot check of this very workload against the oracle (outside the timed regions) ----------
    parity = None
    cpu = None
    if rank == 0:
        from oracle import sieve_oracle as so
        so.build()
        site_gpu, const_gpu = core.site_log_likelihoods()
        n_chk = 16
        c5 = so.counts5(sample[:n_chk])
        _, osl, osc = so.full_eval(c5, sf, so.leaf_params(*PARAMS), tree.ops(), mats, props, tree.root, "none", 0.0,
                                   min(so.max_threads(), n_chk))
        parity = {"sites_checked": n_chk,
                  "max_rel_site_logl": float(np.max(np.abs(site_gpu[:n_chk] - osl) / np.abs(osl))),
                  "max_rel_site_const": float(np.max(np.abs(const_gpu[:n_chk] - osc) / np.abs(osc))),
                  "logP": logp_last, "logP_first": logp_first, "logP_repeatable": bool(logp_first == logp_last)}
        if world == 1 and not args.no_cpu_baseline:
            threads = so.max_threads()
            ns, dt, _ = cpu_eval_sample(tree, mats, props, sample, sf, threads)
            cpu = {"value": (1.0 / dt) * ns / S, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "one evaluation of %d of the %d sites (all %d cells) in %.1f s, site ranges over %d threads "
                             "like ThreadedScsTreeLikelihood; evals/s scaled by %d/%d" % (ns, S, n, dt, threads, ns, S)}

    # ---------------- MCMC states/s: replay of the templates' operator schedules (native driver) ----------------
    mcmc = {}
    if args.mcmc_states > 0:
        from sieve_b200.distributed import bind_native_comm
        from sieve_b200.mcmc import McmcReplay
        if world > 1:
            bind_native_comm(core)      # the chain's per-state exchange is then a native NCCL all-gather inside sieve_step
        # relaxed2 twice: the product default (change tracking: a branch-rate proposal, for which the reference dirties
        # the whole tree, recomputes only the path with new inputs -- same bits) and with every dirty node recomputed
        for mix, elide in (("stage1", True), ("stage2", True), ("relaxed2", True), ("relaxed2", False)):
            core.set_elision(elide)
            key = mix if elide else mix + "_every_dirty_node_recomputed"
            mc = McmcReplay(core, tree, mix=mix, asc_mode="felsenstein_mean", n_background_sites=N_BACKGROUND,
                            gamma_shape=1.0, leaf=PARAMS, branch_rates=br, seed=SEED)
            mc.run(20)                      # warm-up
            barrier()
            st = mc.run(args.mcmc_states, reset=True)
            barrier()
            check = mc.full_recompute() if world == 1 else None
            mcmc[key] = {"states_per_s": st["states"] / st["seconds"], "states": st["states"],
                         "acceptance": st["accepted"] / st["states"],
                         "mean_nodes_recomputed": st["ops_total"] / max(1, st["likelihood_evaluations"]),
                         "leaf_updates": st["leaf_updates"],
                         "incremental_vs_full_recompute_rel": (abs(check - st["log_p"]) / abs(check)) if check else None,
                         "ms_by_op": {k: 1000.0 * v["seconds"] / v["proposed"] for k, v in st["by_op"].items()}}
            mc.close()
        core.set_elision(True)

    # ---------------- variant calling: the max-sum pass on the final state (one-off per analysis) ----------------
    variant_calling = None
    if not args.streamed and not args.sparse and not args.no_variants:
        core.ml_trace()                     # first call allocates the pools
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        core.ml_trace()
        t_trace = time.perf_counter() - t0
        t0 = time.perf_counter()
        g_ml, a_ml, cat_ml, amb_ml = core.ml_call()
        t_call = time.perf_counter() - t0
        variant_calling = {"ml_trace_ms": 1000.0 * t_trace, "ml_call_ms_incl_d2h": 1000.0 * t_call,
                           "sites": int(S), "tied_sites": int((amb_ml != 0).sum()),
                           "non_reference_leaf_calls": int((g_ml[:n] > 0).sum()),
                           "leaf_dropout_calls": int((a_ml > 0).sum())}

    sampler.stop_flag = True
    sampler.join(timeout=2)
    if rank == 0:
        peak, peak_src = measured_peak()
        # the default core eliminates the constant-site twin algebraically: the walk then moves the
        # variant partials only, so its algorithmic bytes are the 384 B figure; SIEVE_B200_CONST_TWIN=1 restores 768 B
        with_const = os.environ.get("SIEVE_B200_CONST_TWIN", "0") == "1"
        alg_prune = algorithmic_bytes_prune(tree, S, with_const)
        alg_leaf = 48 * n * S
        value = world * args.steps / (dev_ms_max / 1000.0)
        e2e_value = world * args.steps / e2e_s_max
        stage_bytes = core.last_stage_bytes() + 48
        achieved = alg_prune / (prune_ms / 1000.0) / 1e9
        traffic = None
        try:   # per-launch DRAM bytes of the dominant kernel from the committed ncu capture of this very workload
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            w = tr["workload"]
            if (w["cells"], w["sites"]) == (n, S) and not (args.per_level or args.matrices or args.streamed or args.sparse or with_const):
                k = next(v for name, v in tr.items() if name.startswith("k_prune<0, 1, 1"))
                traffic = k["dram_bytes_read"] + k["dram_bytes_write"]
        except Exception:
            traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic %d cells x %d sites per GPU, K=4 gamma categories, relaxed-clock branch rates, "
                                   "felsenstein acquisition-bias correction (BASELINE.json configs[2])" % (n, S),
                       "step": "leaf recompute + all %d internal nodes + root + reduce" % (tree.n),
                       "schedule": ("streamed: site chunks, leaves + walk on scratch slots" if args.streamed else
                                    "sparse residency: leaves resident, partials of the larger subtrees only; single-launch walk" if args.sparse else
                                    "per-level launches" if args.per_level else "single-launch walk"),
                       "transition_matrices": "16-double matrices from the host" if args.matrices else "branch distances, P formed on the device (fixed Q)",
                       "l2": "inputs larger than L2 (%.1f GB of partials per evaluation)" % (alg_prune / 1e9),
                       "parallelism": "site shards x%d, all-gather of 5 doubles per rank" % world},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(stage_bytes), "d2h_bytes_per_step": 40,
                    "ms_per_step": 1000.0 * e2e_s_max / args.steps},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": "k_prune<%s, geno0, %s>" % ("const twin" if with_const else "variant only", "generic" if args.matrices else "qfast"), "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(alg_prune), "launch_ms": prune_ms},
            "kernels": {"leaf_ms": leaf_ms, "leaf_GBps": alg_leaf / (leaf_ms / 1000.0) / 1e9 if leaf_ms > 0 else None,
                        "leaf_frac": alg_leaf / (leaf_ms / 1000.0) / 1e9 / peak if leaf_ms > 0 else None,
                        "prune_ms": prune_ms, "reduce_ms": reduce_ms, "tree_only_eval_ms": tree_only_ms,
                        "tree_only_evals_per_s": 1000.0 / tree_only_ms},
            "mcmc": mcmc,
            "mcmc_states_per_s": mcmc.get("stage1", {}).get("states_per_s"),
            "variant_calling": variant_calling,
            "device_bytes": core.device_bytes(),
            "residency": core.sparse_info(),
            "parity": parity,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    core.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
