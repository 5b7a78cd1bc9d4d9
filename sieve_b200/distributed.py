"""Multi-GPU plumbing: one process per GPU, sites sharded contiguously, one tiny exchange per evaluation.

The reference shards sites over threads of one JVM (likelihood/ThreadedScsTreeLikelihood.java:176-219) and combines two
doubles per worker in worker order (:371-379).  Here each rank owns one shard on its GPU; the per-rank result struct
(5 doubles) is all-gathered (NCCL over NVLink on GPUs, gloo in CPU tests) and every rank applies the same fixed-order
combine, so all ranks hold bit-identical logP.  There is no other data-path collective: sites are independent.
"""
import numpy as np
import torch
import torch.distributed as dist

from ._lib import ShardResult
from .core import combine_shards, pattern_points

RESULT_DOUBLES = 5


def shard_range(rank, world, n_sites):
    """calcPatternPoints (likelihood/ThreadedScsTreeLikelihood.java:227-238): [begin, end) of `rank`."""
    pts = pattern_points(n_sites, world)
    return int(pts[rank]), int(pts[rank + 1])


def result_to_tensor(r, device="cpu"):
    return torch.tensor([r.sum_log_l, r.const_lse, r.const_sum, r.lewis_sum, r.weight_sum], dtype=torch.float64,
                        device=device)


def tensor_to_results(t):
    a = t.detach().cpu().numpy().reshape(-1, RESULT_DOUBLES)
    return [ShardResult(*[float(x) for x in row]) for row in a]


def all_gather_results(local, group=None):
    """`local` is a ShardResult or a [5] float64 tensor (on the GPU for NCCL). Returns the list over ranks, rank order."""
    t = local if torch.is_tensor(local) else result_to_tensor(local)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tensor_to_results(t)
    world = dist.get_world_size(group)
    out = torch.empty(world * RESULT_DOUBLES, dtype=torch.float64, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return tensor_to_results(out)


def global_log_p(local, asc_mode="none", n_background_sites=0.0, group=None):
    """logP of the whole data set from this rank's shard result (identical on every rank)."""
    return combine_shards(all_gather_results(local, group), asc_mode, n_background_sites)


def bind_native_comm(core, group=None):
    """Binds an NCCL communicator to `core` (include/sieve_b200.h: sieve_comm_init) so that every evaluation all-gathers
    the shard results natively on the core's stream.  The 128-byte id is drawn on rank 0 and broadcast with
    torch.distributed (any backend)."""
    from .core import comm_unique_id
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, src=0, group=group)
    core.comm_init(rank, world, bytes(t.cpu().numpy().tobytes()))
    return world
