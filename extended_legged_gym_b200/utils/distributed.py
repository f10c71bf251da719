"""Sharding of independent envs across the GPUs of one box and the two genuine reductions of the path.

The reference env is single-GPU (no collective anywhere on the env path); envs are independent, so
they shard with no data-path collective.  What does reduce: the episode statistics of ``reset_idx``
(envs/base/legged_robot.py:200-213: means of ``episode_sums`` over the envs that reset) and the MPPI update
(``utils/mppi.py``).  One process per GPU, ``torch.distributed`` (NCCL on the box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous block of ``total`` units owned by ``rank`` (first ``total % world`` ranks get one more)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedEpisodeStats:
    """(sum, count) accumulators of the per-term episode returns of the envs that reset, kept on the device with no
    host sync; ``reduce()`` -- called by every rank at the same point, e.g. once per rollout iteration -- all-reduces
    one small vector and returns exactly the means ``extras['episode']`` would hold on a single GPU owning all envs."""

    def __init__(self, names, device, group=None):
        self.names = list(names)
        self.group = group
        self.buf = torch.zeros(len(self.names) + 3, dtype=torch.float64, device=device)   # term sums | n_resets, sum_levels, n_envs

    def accumulate(self, episode_sums, env_ids, terrain_levels=None):
        if len(env_ids) == 0:
            return
        k = len(self.names)
        for i, name in enumerate(self.names):
            self.buf[i] += episode_sums[name][env_ids].sum(dtype=torch.float64)
        self.buf[k] += len(env_ids)
        if terrain_levels is not None:
            self.buf[k + 1] = terrain_levels.sum(dtype=torch.float64)
            self.buf[k + 2] = terrain_levels.numel()

    def reduce(self, max_episode_length_s):
        v = self.buf.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(v, op=dist.ReduceOp.SUM, group=self.group)
        k = len(self.names)
        n = v[k]
        out = {}
        if float(n) > 0:
            for i, name in enumerate(self.names):
                out["rew_" + name] = (v[i] / n / max_episode_length_s).to(torch.float)
        if float(v[k + 2]) > 0:
            out["terrain_level"] = (v[k + 1] / v[k + 2]).to(torch.float)
        out["num_resets"] = int(n)
        self.buf.zero_()
        return out
