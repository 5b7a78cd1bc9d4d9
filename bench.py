#!/usr/bin/env python
"""bench.py -- full-tree log-likelihood evaluations per second of the SIEVE hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Headline workload (BASELINE.json configs[2], the largest configuration that is quoted for one GPU): synthetic 1,000 cells
x 100,000 candidate sites, K = 4 gamma categories, acquisition-bias correction on (rmc_stage_2 semantics).  One step = one
complete evaluation of the hot path: leaf read-count likelihoods of every (cell, site) (a parameter proposal), Felsenstein
pruning of every internal node with fresh transition matrices (a relaxed-clock / scaler proposal dirties the whole tree),
root integration, per-site log and reduction, acquisition-bias term.  At N > 1 the data set is N x 100,000 sites, one
contiguous site range per rank (weak scaling, sites are independent) with GLOBAL size factors (exact distributed
selection, sieve_sf_*); the only per-step exchange is the all-gather of 5 doubles per rank, fused into the reduction
kernel over peer memory (ncclAllGather where the mailboxes cannot be mapped).

value  : steps/s with matrices and ops already staged in HBM (kernels only, CUDA events on the handle's stream; the same
         native K-step loop at every N).
e2e    : the same through sieve_step_distances() with HOST buffers: host->device copy of distances + ops and device->host
         read of the result inside the timed region.

north_star block (every N): the 10,000-cell x 1,000,000-site data set of BASELINE.json configs[3]/[4], site-sharded over
the N GPUs (strong scaling; at N = 1 the 1/8 share an 8-GPU run gives each GPU): ms per evaluation, evaluations/s of the
whole data set, the leaf stress of configs[4], the three MCMC operator mixes and per-rank parity against the oracle.
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
NS_CELLS = int(os.environ.get("SIEVE_BENCH_NS_CELLS", 10000))          # north star: configs[3] / configs[4]
NS_SITES_TOTAL = int(os.environ.get("SIEVE_BENCH_NS_SITES", 1000000))
K_CAT = 4
N_BACKGROUND = 1.0e7
PARAMS = (0.3, 50.0, 50.0, 0.008, 150.0, 2.0)   # theta, t, v, eps, w1, w2 (examples/templates/smc_stage_1.xml:30-38)
SEED = 20260101
METRIC = "loglik_evals_per_sec"
UNIT = ("evals/s (one eval = leaf + full-tree pruning + root + reduce over {:,} cells x {:,} sites; at N GPUs each rank "
        "evaluates one such shard per step)").format(N_CELLS, N_SITES)
WORKLOAD = ("synthetic %d cells x %d sites per GPU, K=4 gamma categories, relaxed-clock branch rates, felsenstein "
            "acquisition-bias correction (BASELINE.json configs[2])" % (N_CELLS, N_SITES))


# ----------------------------------------------------------------------------------------------------------------
# synthetic data (same generative model as sieve_b200.data.synthetic, vectorised in torch; runs on cuda or cpu)
# ----------------------------------------------------------------------------------------------------------------
def make_tree_and_matrices(seed, n, want_matrices=True):
    from sieve_b200 import substmodel
    from sieve_b200.tree import random_coalescent
    rng = np.random.default_rng(seed)
    tree = random_coalescent(n, rng)
    tree.height *= 0.3 / tree.height[tree.root]
    rates, props = substmodel.gamma_category_rates(1.0, K_CAT)
    # relaxed clock: per-branch rate multipliers LogNormal(mean 1, sigma 0.2) folded into P on the host
    br = rng.lognormal(-0.02, 0.2, size=tree.node_count)
    mats = substmodel.node_matrices(tree, rates, branch_rates=br) if want_matrices else None
    return tree, rates, props, br, mats


def _depth_levels(tree):
    """children grouped by depth below the root: the genotypes of one level only need their parents' (one level up)"""
    depth = np.zeros(tree.node_count, dtype=np.int64)
    for p in tree.post_order()[::-1]:
        for c in (tree.left[p], tree.right[p]):
            if c >= 0:
                depth[c] = depth[p] + 1
    levels = []
    for d in range(1, int(depth.max()) + 1):
        c = np.nonzero(depth == d)[0]
        levels.append((c, tree.parent[c]))
    return levels


def generate_counts_device(tree, rates, br, n, S, seed, device, chunk=None):
    """int32 [cell][site][4] on `device`, generated in site chunks so that the temporaries stay small.  The genotypes are
    evolved level by level (all nodes of one depth in one batched draw)."""
    import torch

    from sieve_b200 import substmodel
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    torch.manual_seed(seed)          # _standard_gamma draws from the global generator
    if torch.device(device).type == "cuda":
        torch.cuda.manual_seed_all(seed)
    theta, t, v, eps, w1, w2 = PARAMS
    bl = tree.branch_lengths()
    d_all = (bl * br)[:, None] * np.asarray(rates)[None, :]                                  # [node][K]
    P_all = (np.eye(4)[None, None] + np.expm1(-0.2 * d_all)[..., None, None] * substmodel.PI1[None, None]
             + np.expm1(-0.4 * d_all)[..., None, None] * substmodel.PI2[None, None])
    cums = torch.tensor(np.cumsum(np.clip(P_all, 0.0, None), axis=-1)[..., :3], device=device, dtype=torch.float32)  # [node][K][4][3]
    levels = [(torch.tensor(c, device=device), torch.tensor(p, device=device)) for c, p in _depth_levels(tree)]
    widest = max(len(c) for c, _ in levels)
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
        chunk = max(256, min(S, int(1.2e8 // max(n, widest))))
    for s0 in range(0, S, chunk):
        ns = min(chunk, S - s0)
        cat = torch.randint(0, K_CAT, (ns,), device=device, generator=g)
        geno = torch.zeros((tree.node_count, ns), dtype=torch.int64, device=device)
        for c, p in levels:
            rows = cums[c[:, None], cat[None, :], geno[p]]                                   # [len(c)][ns][3]
            u = torch.rand((c.numel(), ns), device=device, generator=g, dtype=torch.float32)
            geno[c] = (u[..., None] > rows).sum(-1)
            del rows, u
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


def library_build_id():
    """sha256 (first 16 hex digits) of everything the library is built from -- sieve_b200/csrc/* and include/*.h, in name
    order -- so that a rebuild of the same sources on another box keeps the stamp: ncu-derived figures are tied to the code
    they were taken from.  A kernel-variant library (SIEVE_B200_LIB, experiments) adds the hash of its binary."""
    import glob
    import hashlib
    h = hashlib.sha256()
    files = sorted(glob.glob(os.path.join(ROOT, "sieve_b200", "csrc", "*")) + glob.glob(os.path.join(ROOT, "include", "*.h")))
    for fn in files:
        if not os.path.isfile(fn):
            continue
        h.update(os.path.basename(fn).encode())
        with open(fn, "rb") as f:
            h.update(f.read())
    if os.environ.get("SIEVE_B200_LIB"):
        with open(os.environ["SIEVE_B200_LIB"], "rb") as f:
            h.update(hashlib.sha256(f.read()).digest())
    return h.hexdigest()[:16]


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
    bounded site sample of the same workload -- same tree, same generator and seed as the GPU arm's rank 0 (torch's CPU
    and CUDA random streams differ, so the draws are not identical) -- all host threads; throughput scaled to the full
    100,000 sites (sites are independent and every site costs the same)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch

    from oracle import sieve_oracle as so
    so.build()
    threads = so.max_threads()
    tree, rates, props, br, mats = make_tree_and_matrices(SEED, N_CELLS)
    P = so.leaf_params(*PARAMS)
    ops = tree.ops()

    def sample(n_sites):
        c = generate_counts_device(tree, rates, br, N_CELLS, n_sites, SEED, "cpu")
        return so.counts5(np.ascontiguousarray(c.permute(1, 0, 2).numpy()))
    per_thread = int(os.environ.get("SIEVE_BENCH_REF_SITES_PER_THREAD", 0))
    if per_thread <= 0:
        # calibrate: a probe of 16 sites per thread, then a sample sized for ~1.5 s per step and <= ~150 s for the run
        pc5 = sample(threads * 16)
        psf, _ = so.size_factors(pc5)
        t0 = time.perf_counter()
        so.full_eval(pc5, psf, P, ops, mats, props, tree.root, "felsenstein_mean", N_BACKGROUND, threads, want_sites=False)
        dt = max(time.perf_counter() - t0, 1e-3)
        budget = min(1.5, 150.0 / max(1, args.steps + args.warmup))
        per_thread = int(min(4096, max(16, 16 * budget / dt)))
    n_sample = threads * per_thread
    c5 = sample(n_sample)
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
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d of %d sites per step (all %d cells), site ranges over %d threads like "
                                   "ThreadedScsTreeLikelihood; evals/s scaled by %d/%d (sites are independent); ONE shard "
                                   "on rank 0 at every N" % (n_sample, N_SITES, N_CELLS, threads, n_sample, N_SITES)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# one resident workload on this rank's GPU
# ----------------------------------------------------------------------------------------------------------------
class Workload:
    def __init__(self, n, S, seed_data, dev, local, world, mode="dense", want_matrices=True, n_sample=1024):
        import torch

        from sieve_b200 import substmodel
        from sieve_b200.core import ScsCudaLikelihoodCore
        from sieve_b200.sizefactors import SizeFactorPart, solve_parts, torch_allreduce
        self.n, self.S, self.mode, self.world = n, S, mode, world
        self.tree, self.rates, self.props, self.br, self.mats = make_tree_and_matrices(SEED, n, want_matrices)
        t0 = time.perf_counter()
        self.counts = generate_counts_device(self.tree, self.rates, self.br, n, S, seed_data, dev)   # [cell][site][4]
        torch.cuda.synchronize()
        self.gen_s = time.perf_counter() - t0
        # host samples of this rank's sites: a strided one for the parity check, a contiguous one for the CPU baseline
        self.sample_idx = np.unique(np.linspace(0, S - 1, min(S, n_sample)).astype(np.int64))
        idx_t = torch.as_tensor(self.sample_idx, device=dev)
        self.sample = self.counts[:, idx_t, :].permute(1, 0, 2).contiguous().cpu().numpy()
        torch.cuda.empty_cache()
        # GLOBAL size factors over the sites of all ranks, before any pool exists (sieve_sf_*: exact distributed selection)
        t0 = time.perf_counter()
        part = SizeFactorPart(n, S, self.counts.data_ptr(), None, 0, local)
        self.sf, self.cov_stats = solve_parts([part], torch_allreduce() if world > 1 else None)
        part.close()
        self.sf_s = time.perf_counter() - t0
        # the core adopts the generator's tensor (SIEVE_FLAG_EXTERNAL_COUNTS): no second copy of the counts
        self.core = ScsCudaLikelihoodCore(n, S, K_CAT, 4, True, True, True, local, ephemeral=(mode == "streamed"),
                                          sparse=mode.startswith("sparse"), single_leaf_buffer=(mode == "sparse_single_leaf"),
                                          external_counts=True)
        self.core.set_counts_device(self.counts.data_ptr())
        self.core.set_size_factors(self.sf)
        self.core.set_leaf_params(*PARAMS)
        self.core.update_leaves(for_update=False)
        self.core.set_root(self.tree.root, self.props, 0)
        self.nodes = np.array([i for i in range(self.tree.node_count) if i != self.tree.root], dtype=np.int32)
        self.D_host = np.ascontiguousarray(substmodel.node_distances(self.tree, self.rates, branch_rates=self.br)[self.nodes])
        self.ops = np.ascontiguousarray(self.tree.ops(), dtype=np.int32)
        self.step_no = 0

    def close(self):
        import torch
        self.core.close()
        self.counts = None
        torch.cuda.empty_cache()

    def exact_eval(self):
        self.core.set_leaf_params(*PARAMS)
        return self.core.step_distances(self.nodes, self.D_host, self.ops, update_leaves=True)

    def proposal_step(self):
        """what the plugin does per MCMC state of this workload: leaf parameters changed + every matrix changed -> full
        evaluation.  The proposal alternates between a coverage parameter (NB tables) and a nucleotide parameter (DM tables)."""
        self.step_no += 1
        pr = list(PARAMS)
        i = 1 if self.step_no % 2 else 3
        pr[i] = PARAMS[i] * (1.0 + 1e-9 * (self.step_no % 7))
        self.core.set_leaf_params(*pr)
        return self.core.step_distances(self.nodes, self.D_host, self.ops, update_leaves=True)

    def parity(self, so, threads):
        """per-site parity of this rank's strided sample against the oracle (needs the state of exact_eval)"""
        site_gpu, const_gpu = self.core.site_log_likelihoods()
        mats = self.mats
        if mats is None:
            from sieve_b200 import substmodel
            mats = substmodel.node_matrices(self.tree, self.rates, branch_rates=self.br)
        c5 = so.counts5(self.sample)
        _, osl, osc = so.full_eval(c5, self.sf, so.leaf_params(*PARAMS), self.tree.ops(), mats, self.props, self.tree.root,
                                   "none", 0.0, max(1, min(threads, c5.shape[0])))
        g, gc = site_gpu[self.sample_idx], const_gpu[self.sample_idx]
        return {"sites": int(len(self.sample_idx)),
                "max_rel_site_logl": float(np.max(np.abs(g - osl) / np.abs(osl))),
                "max_rel_site_const": float(np.max(np.abs(gc - osc) / np.abs(osc))),
                "rel_sampled_total": float(abs(g.sum() - osl.sum()) / abs(osl.sum()))}


def time_device_loop(core, lib_stream, steps, update_leaves, barrier):
    """K evaluations enqueued by ONE native call (sieve_replay: no host in the loop, inputs resident in HBM; with a bound
    communicator every evaluation ends with the native NCCL all-gather), timed with CUDA events on the handle's stream."""
    import torch
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(lib_stream):
        ev0.record()
    core.replay(update_leaves=update_leaves, n_times=steps)
    with torch.cuda.stream(lib_stream):
        ev1.record()
    barrier()
    return ev0.elapsed_time(ev1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mcmc-states", type=int, default=400, help="proposals per operator mix in the MCMC replay (0 = skip)")
    ap.add_argument("--mode", default="dense", choices=["dense", "sparse", "sparse_single_leaf", "streamed"],
                    help="residency of the headline workload (SIEVE_BENCH_CELLS / SIEVE_BENCH_SITES change its shape)")
    ap.add_argument("--no-variants", action="store_true", help="skip the max-sum (variant calling) timing")
    ap.add_argument("--no-dropin", action="store_true", help="skip the call-by-call traverse sequence (e2e_dropin)")
    ap.add_argument("--no-north-star", action="store_true", help="skip the 10,000-cell x 1,000,000-site block")
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
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # never overridden: a caller's NCCL_DEBUG=INFO stays in force
        dist.init_process_group("nccl", device_id=dev)

    from oracle import sieve_oracle as so
    from sieve_b200 import _lib
    from sieve_b200.core import combine_shards
    from sieve_b200.distributed import bind_native_comm
    _lib.lib()
    if rank == 0:
        so.build()
    if world > 1:
        dist.barrier()
    so.lib()
    threads = so.max_threads()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rank_max(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def global_results(core, r):
        """the per-rank result structs of the last evaluation, rank order (native all-gather when a communicator is bound)"""
        return core.last_global_results() if world > 1 else [r]

    def combine_check(results):
        """the combine of the workers (ThreadedScsTreeLikelihood.java:371-379, ScsAlignment.java:811-815) recomputed with
        the oracle's functions from the same per-rank numbers"""
        loci = int(round(sum(x.weight_sum for x in results)))
        want = sum(x.sum_log_l for x in results) + so.asc_bias_threaded([x.const_lse for x in results], "felsenstein_mean",
                                                                         loci, N_BACKGROUND)
        got = combine_shards(results, "felsenstein_mean", N_BACKGROUND)
        return got, float(abs(got - want) / abs(want))

    sampler = ClockSampler(local)
    sampler.start()

    # =====================================================================================================================
    # headline: configs[2] shard per rank
    # =====================================================================================================================
    n, S = N_CELLS, N_SITES
    W = Workload(n, S, SEED + 17 * rank, dev, local, world, mode=args.mode)
    core, tree = W.core, W.tree
    if world > 1:
        bind_native_comm(core)      # every evaluation now ends with the exchange of the 5-double results of all ranks
    exchange_desc = {0: "no exchange (one rank)", 1: "ncclAllGather of 5 doubles per rank behind every reduction",
                     2: "exchange of 5 doubles per rank fused into the reduction kernel over NVLink peer memory (CUDA IPC "
                        "mailboxes; NCCL only at set-up)"}[core.comm_exchange_mode()]
    lib_stream = torch.cuda.ExternalStream(core.stream(), device=dev)

    r_first = W.exact_eval()
    logp_first, _ = combine_check(global_results(core, r_first))
    # ---------------- warm-up (also stages distances + ops in HBM for the device-only loop) ----------------
    warm = max(args.warmup, 3)
    for _ in range(warm):
        W.proposal_step()
    barrier()
    # ---------------- timed region 1: device path, inputs resident in HBM ----------------
    launches0 = core.launch_count()
    dev_ms_max = rank_max(time_device_loop(core, lib_stream, args.steps, True, barrier))
    launches = core.launch_count() - launches0
    # ---------------- timed region 2: end to end through the C-ABI with host buffers ----------------
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        W.proposal_step()
    barrier()
    e2e_s_max = rank_max(time.perf_counter() - t0)
    stage_bytes = core.last_stage_bytes() + 48
    # ---------------- the no-extra-edit integration: traverse's own call sequence, natively, call by call ----------------
    dropin = None
    if not args.no_dropin and args.mode != "streamed":
        for _ in range(2):
            W.core.set_leaf_params(*PARAMS)
            core.traverse_full(tree.root, W.ops, W.mats, update_leaves=True, want_sites=True)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            pr = list(PARAMS)
            pr[1] = PARAMS[1] * (1.0 + 1e-9 * (1 + i % 7))
            core.set_leaf_params(*pr)
            core.traverse_full(tree.root, W.ops, W.mats, update_leaves=True, want_sites=True)
        barrier()
        dropin_s = rank_max(time.perf_counter() - t0)
        dropin = {"value": world * args.steps / dropin_s, "unit": UNIT, "ms_per_step": 1000.0 * dropin_s / args.steps,
                  "h2d_bytes_per_step": int(core.last_stage_bytes() + 48), "d2h_bytes_per_step": int(40 + 8 * S),
                  "calls": "per step: %d x setNodeMatrixForUpdate, %d x setNodeMatrix (16 doubles), updateLeaves, %d x "
                           "(setNodePartialsForUpdate + calculateLogPartials), evaluate, getPatternLogLikelihoods (%d doubles "
                           "back) -- ScsTreeLikelihood.traverse unchanged; 16-double matrices, generic walk kernel"
                           % (tree.node_count - 1, (tree.node_count - 1) * K_CAT, tree.n, S)}
    # ---------------- per-kernel timing (CUDA events on the handle's stream, inside the library) ----------------
    core.enable_timing(True)
    tm = []
    for _ in range(5):
        W.proposal_step()
        tm.append(core.last_timing())
    core.enable_timing(False)
    leaf_ms = float(np.median([x["leaf_ms"] for x in tm]))
    prune_ms = float(np.median([x["prune_ms"] for x in tm]))
    reduce_ms = float(np.median([x["reduce_ms"] for x in tm]))
    # tree-only evaluation (stage-2 style proposal: matrices change, leaves do not)
    # (every distance perturbed in the last bits: with the distances of the previous step the change tracking would find
    # part of the tree unchanged and the replayed op list would not be the whole tree)
    core.step_distances(W.nodes, W.D_host * (1.0 + 1e-12), W.ops, update_leaves=False)
    tree_only_ms = time_device_loop(core, lib_stream, args.steps, False, barrier) / args.steps
    # the walk of a tree-only step on its own (events inside the library), for comparison with the walk after k_leaf
    core.enable_timing(True)
    tm2 = []
    for i in range(5):
        core.step_distances(W.nodes, W.D_host * (1.0 + 1e-12 * (i + 1)), W.ops, update_leaves=False)   # every matrix new
        tm2.append(core.last_timing())
    core.enable_timing(False)
    prune_tree_only_ms = float(np.median([x["prune_ms"] for x in tm2]))

    r_last = W.exact_eval()
    results_last = global_results(core, r_last)
    logp_last, combine_rel = combine_check(results_last)
    # ---------------- parity of this very workload against the oracle: every rank, strided sites ----------------
    par = W.parity(so, threads)
    parity = {"sites_checked_per_rank": par["sites"], "ranks_checked": world,
              "max_rel_site_logl": rank_max(par["max_rel_site_logl"]), "max_rel_site_const": rank_max(par["max_rel_site_const"]),
              "max_rel_sampled_total": rank_max(par["rel_sampled_total"]), "combine_rel_vs_oracle": combine_rel,
              "logP": logp_last, "logP_first": logp_first, "logP_repeatable": bool(logp_first == logp_last),
              "size_factors": "global over all %d ranks (sieve_sf_*), %.2f s" % (world, W.sf_s)}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_keep = min(S, max(64, int(2.4e7 // n)))
        contiguous = W.counts[:, :n_keep, :].permute(1, 0, 2).contiguous().cpu().numpy()
        ns, dt, _ = cpu_eval_sample(tree, W.mats, W.props, contiguous, W.sf, threads)
        cpu = {"value": (1.0 / dt) * ns / S, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "one evaluation of %d of the %d sites (all %d cells) in %.1f s, site ranges over %d threads "
                         "like ThreadedScsTreeLikelihood; evals/s scaled by %d/%d" % (ns, S, n, dt, threads, ns, S)}

    # ---------------- MCMC states/s: replay of the templates' operator schedules (native driver) ----------------
    def run_mcmc(wl, n_states, mixes):
        from sieve_b200.mcmc import McmcReplay
        out = {}
        for mix, elide in mixes:
            wl.core.set_elision(elide)
            key = mix if elide else mix + "_every_dirty_node_recomputed"
            mc = McmcReplay(wl.core, wl.tree, mix=mix, asc_mode="felsenstein_mean", n_background_sites=N_BACKGROUND,
                            gamma_shape=1.0, leaf=PARAMS, branch_rates=wl.br, seed=SEED)
            mc.run(min(20, n_states))                      # warm-up
            barrier()
            st = mc.run(n_states, reset=True)
            barrier()
            check = mc.full_recompute() if world == 1 else None
            out[key] = {"states_per_s": st["states"] / st["seconds"], "states": st["states"],
                        "acceptance": st["accepted"] / st["states"],
                        "mean_nodes_recomputed": st["ops_total"] / max(1, st["likelihood_evaluations"]),
                        "leaf_updates": st["leaf_updates"],
                        "incremental_vs_full_recompute_rel": (abs(check - st["log_p"]) / abs(check)) if check else None,
                        "ms_by_op": {k: 1000.0 * v["seconds"] / v["proposed"] for k, v in st["by_op"].items()}}
            mc.close()
        wl.core.set_elision(True)
        return out

    mcmc = {}
    if args.mcmc_states > 0:
        # relaxed2 twice: the product default (change tracking: a branch-rate proposal, for which the reference dirties
        # the whole tree, recomputes only the path with new inputs -- same bits) and with every dirty node recomputed
        mcmc = run_mcmc(W, args.mcmc_states, (("stage1", True), ("stage2", True), ("relaxed2", True), ("relaxed2", False)))

    # ---------------- variant calling: the max-sum pass on the final state (one-off per analysis) ----------------
    variant_calling = None
    if args.mode == "dense" and not args.no_variants:
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
        del g_ml, a_ml, cat_ml, amb_ml
    device_bytes = core.device_bytes() + 16 * n * S
    residency = core.sparse_info()
    with_const_twin = os.environ.get("SIEVE_B200_CONST_TWIN", "0") == "1"
    alg_prune = algorithmic_bytes_prune(tree, S, with_const_twin)
    clocks_headline = sampler.summary()
    W.close()

    # =====================================================================================================================
    # north star: the 10,000 x 1,000,000 data set, site-sharded over the ranks (strong scaling)
    # =====================================================================================================================
    north = None
    if not args.no_north_star:
        share = NS_SITES_TOTAL // (world if world > 1 else 8)
        per_gpu_counts = 16.0 * NS_CELLS * share
        # what fits in 180 GB next to 16 B/cell-site of counts: both leaf buffers (64 B) up to ~1.6e9 cell-sites, one leaf
        # buffer (32 B) up to ~3e9, beyond that nothing but the counts is resident
        cs = float(NS_CELLS) * share
        mode = "sparse" if cs <= 1.6e9 else ("sparse_single_leaf" if cs <= 3.0e9 else "streamed")
        # the residency that fits is tried first; should a rank run out of memory (a pool that does not get its slots), ALL
        # ranks step down together -- the main line must not be lost to the north-star block
        order = ["sparse", "sparse_single_leaf", "streamed"]
        NSW, ns_fallback = None, []
        for mode in order[order.index(mode):]:
            try:
                NSW = Workload(NS_CELLS, share, SEED + 1000 + 17 * rank, dev, local, world, mode=mode, want_matrices=False,
                               n_sample=256)
                ok = 1.0
            except RuntimeError as exc:      # SIEVE_ERR_OOM from sieve_create, or torch's allocator
                ok, NSW = 0.0, None
                ns_fallback.append("%s: %s" % (mode, str(exc)[:160]))
                torch.cuda.empty_cache()
            if -rank_max(-ok) > 0.5:
                break
            if NSW is not None:
                NSW.close()
                NSW = None
        if NSW is None:
            raise RuntimeError("north star: no residency fits: %s" % ns_fallback)
        if world > 1:
            bind_native_comm(NSW.core)
        ns_stream = torch.cuda.ExternalStream(NSW.core.stream(), device=dev)
        ns_steps = max(3, min(args.steps, 8 if mode != "streamed" else 3))
        r0 = NSW.exact_eval()
        ns_logp_first, _ = combine_check(global_results(NSW.core, r0))
        for _ in range(2):
            NSW.proposal_step()
        barrier()
        ns_ms = rank_max(time_device_loop(NSW.core, ns_stream, ns_steps, True, barrier)) / ns_steps
        barrier()
        t0 = time.perf_counter()
        for _ in range(ns_steps):
            NSW.proposal_step()
        barrier()
        ns_e2e_ms = 1000.0 * rank_max(time.perf_counter() - t0) / ns_steps
        NSW.core.enable_timing(True)
        tm = []
        for _ in range(3):
            NSW.proposal_step()
            tm.append(NSW.core.last_timing())
        NSW.core.enable_timing(False)
        NSW.core.step_distances(NSW.nodes, NSW.D_host * (1.0 + 1e-12), NSW.ops, update_leaves=False)   # the whole tree, see above
        ns_tree_ms = rank_max(time_device_loop(NSW.core, ns_stream, ns_steps, False, barrier)) / ns_steps
        r1 = NSW.exact_eval()
        ns_logp, ns_comb = combine_check(global_results(NSW.core, r1))
        npar = NSW.parity(so, threads)
        ns_mcmc = {}
        if args.mcmc_states > 0:
            ns_states = max(24, args.mcmc_states // (4 if mode != "streamed" else 16))
            ns_mcmc = run_mcmc(NSW, ns_states, (("stage1", True), ("stage2", True), ("relaxed2", True)))
        # variant calling on the share: the max-sum pass in site chunks (sparse residency keeps only the calls)
        ns_variants = None
        if mode.startswith("sparse") and not args.no_variants:
            t0 = time.perf_counter()
            NSW.core.ml_trace()
            g_ml, a_ml, cat_ml, amb_ml = NSW.core.ml_call()
            t_vc = time.perf_counter() - t0
            ns_variants = {"ml_trace_and_call_s_incl_d2h": rank_max(t_vc), "sites": int(share), "nodes": int(2 * NS_CELLS),
                           "tied_sites": int((amb_ml != 0).sum()), "non_reference_leaf_calls": int((g_ml[:NS_CELLS] > 0).sum()),
                           "d2h_bytes": int(g_ml.nbytes + a_ml.nbytes + cat_ml.nbytes + amb_ml.nbytes)}
            del g_ml, a_ml, cat_ml, amb_ml
        peak, _ = measured_peak()
        ns_leaf_ms = rank_max(float(np.median([x["leaf_ms"] for x in tm])))
        ns_prune_ms = rank_max(float(np.median([x["prune_ms"] for x in tm])))
        ns_alg = algorithmic_bytes_prune(NSW.tree, share, False)
        north = {
            "dataset": "synthetic %d cells x %d sites (BASELINE.json configs[3] / configs[4])" % (NS_CELLS, NS_SITES_TOTAL),
            "sites_per_gpu": int(share),
            "gpus_for_the_whole_dataset": world if world > 1 else 8,
            "note": ("strong scaling: the data set is split over the %d ranks" % world) if world > 1 else
                    "one GPU holding the 1/8 share that an 8-GPU run gives each rank (no exchange)",
            "residency": mode, "residency_fallbacks": ns_fallback, "slots": NSW.core.sparse_info(),
            "ms_per_evaluation": ns_ms,
            "dataset_evals_per_s": (1000.0 / ns_ms) if world > 1 else None,
            "share_evals_per_s": 1000.0 / ns_ms,
            "e2e_ms_per_evaluation": ns_e2e_ms,
            "tree_only_ms": ns_tree_ms,
            "leaf_stress": {"what": "configs[4]: every step proposes a sequencing-error / coverage parameter, all %d x %d read-count "
                                    "likelihoods are recomputed" % (NS_CELLS, share),
                            # streamed residency: leaf and walk kernels alternate per site chunk and are timed together
                            # (prune_ms below is then the whole chunk loop); the leaf share is not separable
                            "leaf_ms": ns_leaf_ms if ns_leaf_ms > 0 else None,
                            "leaf_GBps_algorithmic": (48.0 * cs / (ns_leaf_ms / 1000.0) / 1e9) if ns_leaf_ms > 0 else None,
                            "leaf_frac_of_hbm_peak": (48.0 * cs / (ns_leaf_ms / 1000.0) / 1e9 / peak) if ns_leaf_ms > 0 else None},
            "prune_ms": ns_prune_ms,
            "prune_frac_of_hbm_peak_algorithmic": (ns_alg / (ns_prune_ms / 1000.0) / 1e9 / peak) if ns_prune_ms > 0 else None,
            "mcmc": ns_mcmc,
            "variant_calling": ns_variants,
            "parity": {"sites_checked_per_rank": npar["sites"], "ranks_checked": world,
                       "max_rel_site_logl": rank_max(npar["max_rel_site_logl"]),
                       "max_rel_site_const": rank_max(npar["max_rel_site_const"]),
                       "max_rel_sampled_total": rank_max(npar["rel_sampled_total"]), "combine_rel_vs_oracle": ns_comb,
                       "logP": ns_logp, "logP_repeatable": bool(ns_logp == ns_logp_first),
                       "size_factors": "global over all %d ranks (sieve_sf_*), %.2f s" % (world, NSW.sf_s)},
            "setup_s": {"generate": NSW.gen_s, "size_factors": NSW.sf_s},
            "device_bytes": NSW.core.device_bytes() + int(per_gpu_counts),
        }
        NSW.close()

    sampler.stop_flag = True
    sampler.join(timeout=2)
    if rank == 0:
        peak, peak_src = measured_peak()
        alg_leaf = 48 * n * S
        value = world * args.steps / (dev_ms_max / 1000.0)
        e2e_value = world * args.steps / e2e_s_max
        achieved = alg_prune / (prune_ms / 1000.0) / 1e9
        traffic, traffic_note = None, None
        build_id = library_build_id()
        try:   # per-launch DRAM bytes of the dominant kernel from the committed ncu capture: only for THIS build and workload
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            w = tr["workload"]
            k = tr["kernels"]["k_prune"]
            if tr.get("library_build_id") != build_id:
                traffic_note = "profiles/ncu_traffic.json was captured from build %s, this is %s: not quoted" % (tr.get("library_build_id"), build_id)
            elif (w["cells"], w["sites"]) != (n, S) or args.mode != "dense" or with_const_twin:
                traffic_note = "profiles/ncu_traffic.json is for another workload: not quoted"
            elif abs(k["duration_ms_under_ncu"] - prune_ms) > 0.15 * prune_ms:
                traffic_note = "capture duration %.3f ms disagrees with this run's %.3f ms: not quoted" % (k["duration_ms_under_ncu"], prune_ms)
            else:
                traffic = k["dram_bytes_read"] + k["dram_bytes_write"]
                traffic_note = "ncu --set full capture of this build (profiles/ncu_traffic.json)"
        except Exception as e:
            traffic_note = "no usable profiles/ncu_traffic.json (%s)" % type(e).__name__
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "step": "leaf recompute + all %d internal nodes + root + reduce" % (tree.n),
                       "schedule": {"dense": "single-launch walk", "streamed": "streamed: site chunks, leaves + walk on scratch slots",
                                    "sparse": "sparse residency: leaves resident, partials of the larger subtrees only; single-launch walk",
                                    "sparse_single_leaf": "sparse residency, one leaf buffer"}[args.mode],
                       "transition_matrices": "branch distances, P formed on the device (fixed Q)",
                       "l2": "inputs larger than L2 (%.1f GB of partials per evaluation)" % (alg_prune / 1e9),
                       "parallelism": "site shards x%d, global size factors, %s" % (world, exchange_desc)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(stage_bytes), "d2h_bytes_per_step": 40,
                    "ms_per_step": 1000.0 * e2e_s_max / args.steps},
            "e2e_dropin": dropin,
            "gpu_launches": int(launches),
            "clocks": clocks_headline,
            "roofline": {"bound": "hbm", "kernel": ("k_prune<const twin, geno0, qfast>" if with_const_twin else "k_walk (single-launch walk, Q-specialised, variant partials)"),
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(alg_prune), "launch_ms": prune_ms,
                         "frac_dram_real": (traffic / (prune_ms / 1000.0) / 1e9 / peak) if traffic else None},
            "library_build_id": build_id,
            "kernels": {"leaf_ms": leaf_ms, "leaf_GBps": alg_leaf / (leaf_ms / 1000.0) / 1e9 if leaf_ms > 0 else None,
                        "leaf_frac": alg_leaf / (leaf_ms / 1000.0) / 1e9 / peak if leaf_ms > 0 else None,
                        "prune_ms": prune_ms, "prune_ms_in_tree_only_step": prune_tree_only_ms, "reduce_ms": reduce_ms, "tree_only_eval_ms": tree_only_ms,
                        "tree_only_evals_per_s": 1000.0 / tree_only_ms},
            "mcmc": mcmc,
            "mcmc_states_per_s": mcmc.get("stage1", {}).get("states_per_s"),
            "variant_calling": variant_calling,
            "device_bytes": device_bytes,
            "residency": residency,
            "parity": parity,
            "cpu_baseline": cpu,
            "north_star": north,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
