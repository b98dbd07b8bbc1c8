"""Multi-GPU plumbing for the update operator.  The path shards by independent sequences (one sequence per
GPU, like the reference's DDP ranks, train.py:90-93) with no data-path collective; torch.distributed is used
only for the barrier around the timed region and for max-over-ranks timing."""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None):
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_sequences(n_sequences, rank, world):
    """contiguous, balanced assignment of independent sequences to ranks"""
    base, rem = divmod(n_sequences, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def patch_owner(n_patches, world):
    """owner rank of every patch: contiguous, balanced blocks (= source-frame blocks when patches are frame-major,
    SURVEY 8e).  int64 [n_patches]"""
    base, rem = divmod(n_patches, world)
    sizes = torch.tensor([base + (1 if r < rem else 0) for r in range(world)], dtype=torch.int64)
    return torch.repeat_interleave(torch.arange(world, dtype=torch.int64), sizes)


def patch_range(n_patches, rank, world):
    base, rem = divmod(n_patches, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_edges_by_patch(kk, n_patches, rank, world):
    """indices (ascending, so the edge order inside a patch is kept) of the edges owned by `rank`: every edge of a
    patch lives on the patch's owner, which keeps the depth blocks C,u and the columns of E rank-local."""
    lo, hi = patch_range(n_patches, rank, world)
    return torch.nonzero((kk >= lo) & (kk < hi)).reshape(-1)


def gather_patch_depths(patches, n_patches, group=None):
    """after a sharded BA: every rank broadcasts the depth channel of the patches it owns (in place on `patches`
    [1,Np,3,P,P] / [Np,3,P,P]) so the replicas agree again.  Bit-exact: owners' values are copied, not summed."""
    if not (dist.is_available() and dist.is_initialized()):
        return patches
    world = dist.get_world_size(group)
    if world == 1:
        return patches
    p = patches.view(-1, 3, patches.shape[-2], patches.shape[-1])
    for r in range(world):
        lo, hi = patch_range(n_patches, r, world)
        if hi > lo:
            buf = p[lo:hi, 2].contiguous()
            dist.broadcast(buf, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
            p[lo:hi, 2] = buf
    return patches


def max_over_ranks(value, device=None):
    """device-timed durations are combined as the max over ranks (never wall clock)"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank, seconds_local, device=None):
    """whole-job throughput = units processed by all ranks / max-over-ranks time"""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    return world * units_per_rank / max_over_ranks(seconds_local, device)
