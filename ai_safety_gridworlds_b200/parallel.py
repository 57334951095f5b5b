"""Multi-GPU plumbing: one process per GPU, the environment batch sharded by global index.

Environments are independent, so stepping needs no collective.  The only exchange
is the end-of-rollout SUM all-reduce of the raw statistics vector (NCCL over NVLink on the GPU box;
gloo in the CPU tests).  The vector holds integer-valued doubles, so the reduced result is exact
and independent of how the batch was sharded.
"""
import torch
import torch.distributed as dist


def shard_range(total_envs, rank, world):
    """Global environment indices [lo, hi) owned by `rank`: contiguous blocks, remainder spread
    over the first ranks.  `lo` is the shard's env_index_base (keys the Philox streams)."""
    per, rem = divmod(int(total_envs), int(world))
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def all_reduce_raw_stats(raw, group=None):
    """SUM all-reduce of a raw statistics vector (float64 [GW_STATS_RAW_LEN]); returns a new tensor."""
    out = raw.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def make_sharded_env(env, total_envs, device=None, **kwargs):
    """VectorEnv for this rank's shard of a `total_envs`-wide job."""
    from .vector_env import VectorEnv
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(total_envs, rank, world)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return VectorEnv(env, hi - lo, device=device, env_index_base=lo, **kwargs)
