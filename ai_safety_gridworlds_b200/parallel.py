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


class MultiAgentStatsMixin(object):
    """End-of-rollout statistics of the multi-agent kernels (include/gwsim_fm.h, GW_MA_STATS_*): the raw vector holds
    exact integer sums, so the SUM all-reduce over shards equals the unsharded batch's vector bit for bit.  The host
    class provides `_stats_fns = (device_fn, clear_fn)`, `_stats_columns` = [(agent chr, [reward keys])] and `_raw_dev`."""

    def stats_raw_device(self):
        from . import _abi
        from .vector_env import _ptr
        _abi.check(self._stats_fns[0](self._h, _ptr(self._raw_dev), self._stream()))
        return self._raw_dev

    def clear_stats(self):
        from . import _abi
        _abi.check(self._stats_fns[1](self._h, self._stream()))

    def finalize_stats(self, raw_host):
        from . import _abi
        raw = [float(x) for x in raw_host]
        episodes = raw[1]
        out = dict(env_steps=int(raw[0]), episodes=int(episodes), length_sum=int(raw[2]), agent_finishes=int(raw[3]),
                   mean_length=(raw[2] / episodes) if episodes else float("nan"), return_sum={}, mean_return={})
        k = _abi.GW_MA_STATS_RETURN0
        for agent, keys in self._stats_columns:
            sums = [raw[k + j] / _abi.GW_MA_STATS_SCALE for j in range(len(keys))]
            k += len(keys)
            if agent is None:                                  # columns of an agent this game does not have
                continue
            out["return_sum"][agent] = dict(zip(keys, sums))
            out["mean_return"][agent] = dict(zip(keys, ((v / episodes) if episodes else float("nan") for v in sums)))
        return out

    def stats(self, group=None):
        """Rollout statistics since construction / clear_stats; with a torch.distributed `group` (True = the default
        group) the raw vector is all-reduced (SUM; NCCL over NVLink on the GPU box) first."""
        raw = self.stats_raw_device()
        if group is not None:
            raw = raw.clone()
            dist.all_reduce(raw, op=dist.ReduceOp.SUM, group=None if group is True else group)
        return self.finalize_stats(raw.cpu().numpy())
