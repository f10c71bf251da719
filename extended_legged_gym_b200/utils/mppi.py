"""MPPI cost-weighted control update with the rollout (sample) dimension sharded across ranks.

In-tree statement of the update in the reference: legged_gym/tests/score_sampling/cmp_mppi_wbfo.py:216-233; glue that
produces its inputs: envs/batch_rollout/robot_traj_grad_sampling.py:249-280 (``rollout_batch``).  One process per GPU;
the only exchanges are an all-gather of the per-sample costs (4 bytes per sample) and an all-reduce of
``[sum_e, sum_e * sample]`` per main env -- a few KB per optimisation iteration.
"""
import torch
import torch.distributed as dist

from .. import _lib


class _CudaOps:
    """The three local stages on the GPU (csrc/elg_mppi.cu)."""

    @staticmethod
    def costs(step_rewards):
        r = step_rewards if (step_rewards.dtype == torch.float and step_rewards.is_contiguous()) else step_rewards.float().contiguous()
        if not r.is_cuda:
            raise _lib.ElgError("mppi_update has no CPU path: tensors must be CUDA tensors")
        M, S, T = r.shape
        out = torch.empty(M, S, device=r.device)
        _lib.check(_lib.load().elg_mppi_costs(r.data_ptr(), M, S, T, out.data_ptr(), torch.cuda.current_stream(r.device).cuda_stream),
                   "elg_mppi_costs")
        return out

    @staticmethod
    def partials(costs_all, first, samples, temp):
        c = costs_all.contiguous()
        M, S_local = samples.shape[0], samples.shape[1]
        flat = samples.reshape(M, S_local, -1)
        flat = flat if (flat.dtype == torch.float and flat.is_contiguous()) else flat.float().contiguous()
        out = torch.empty(M, 1 + flat.shape[2], device=c.device)
        rc = _lib.load().elg_mppi_partials(c.data_ptr(), M, c.shape[1], first, S_local, flat.data_ptr(), flat.shape[2], float(temp),
                                           out.data_ptr(), torch.cuda.current_stream(c.device).cuda_stream)
        _lib.check(rc, "elg_mppi_partials")
        return out

    @staticmethod
    def finish(partial, traj_shape):
        M, KD = partial.shape[0], partial.shape[1] - 1
        out = torch.empty(M, KD, device=partial.device)
        _lib.check(_lib.load().elg_mppi_finish(partial.data_ptr(), M, KD, out.data_ptr(), torch.cuda.current_stream(partial.device).cuda_stream),
                   "elg_mppi_finish")
        return out.reshape(M, *traj_shape)


_WORKSPACE = {}


def mppi_update_native(step_rewards, samples, temperature, comm=None):
    """The whole update as ONE extension call on the current stream (``elg_mppi_update``): local costs -> ncclAllGather ->
    weights + partial sums -> ncclAllReduce -> mean trajectories.  ``comm``: utils.distributed.ElgComm or None (one rank).
    No host synchronisation, CUDA-graph capturable."""
    r = step_rewards if (step_rewards.dtype == torch.float and step_rewards.is_contiguous()) else step_rewards.float().contiguous()
    if not r.is_cuda:
        raise _lib.ElgError("mppi_update has no CPU path: tensors must be CUDA tensors")
    M, S_local, T = r.shape
    flat = samples.reshape(M, S_local, -1)
    flat = flat if (flat.dtype == torch.float and flat.is_contiguous()) else flat.float().contiguous()
    KD = flat.shape[2]
    world = comm.world if comm is not None else 1
    # scratch (costs of all ranks, partial sums) is kept per shape: the call is a handful of microseconds of kernels, two
    # allocator round trips would double its host time; the result is a fresh tensor
    key = (world, M, S_local, KD, r.device)
    ws = _WORKSPACE.get(key)
    if ws is None or torch.cuda.is_current_stream_capturing():
        ws = (torch.empty(world, M, S_local, device=r.device), torch.empty(M, 1 + KD, device=r.device))
        if not torch.cuda.is_current_stream_capturing():
            _WORKSPACE[key] = ws
    costs, partial = ws
    out = torch.empty(M, KD, device=r.device)
    rc = _lib.load().elg_mppi_update(r.data_ptr(), flat.data_ptr(), M, S_local, T, KD, float(temperature), costs.data_ptr(), partial.data_ptr(),
                                     out.data_ptr(), comm.handle if comm is not None else None, torch.cuda.current_stream(r.device).cuda_stream)
    _lib.check(rc, "elg_mppi_update")
    return out.reshape(M, *samples.shape[2:])


def mppi_update(step_rewards, samples, temperature, group=None, ops=None, comm=None):
    """step_rewards [M, S_local, T] and samples [M, S_local, K, D] are THIS rank's share of the rollouts of every main
    env (ranks hold equal shares, in rank order); returns the updated mean trajectories [M, K, D], identical on all
    ranks.  With ``comm`` (an ``ElgComm``) or on a single rank the update is one extension call with its collectives on
    the compute stream; ``ops`` replaces the local GPU stages and routes the exchange through ``torch.distributed``
    (tests run the collective plumbing on CPU over gloo with the oracle's stages)."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if ops is None and (comm is not None or world == 1):
        return mppi_update_native(step_rewards, samples, temperature, comm)
    ops = ops or _CudaOps
    rank = dist.get_rank(group) if world > 1 else 0
    costs = ops.costs(step_rewards)                                    # [M, S_local]
    S_local = costs.shape[1]
    if world > 1:
        parts = [torch.empty_like(costs) for _ in range(world)]
        dist.all_gather(parts, costs.contiguous(), group=group)
        costs_all = torch.cat(parts, dim=1)                            # [M, world * S_local], rank-major like the shards
    else:
        costs_all = costs
    partial = ops.partials(costs_all, rank * S_local, samples, temperature)
    if world > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return ops.finish(partial, tuple(samples.shape[2:]))


def rollout_batch(env, all_us):
    """RobotTrajGradSampling.rollout_batch (robot_traj_grad_sampling.py:249-280): roll every rollout env through the
    horizon, rewards [num_rollout_envs, horizon]; rollouts are re-synchronised with their mains before and after.
    Envs that have the method (RobotBatchRollout) run the loop as one CUDA graph; this is the step-by-step public-API form."""
    if hasattr(env, "rollout_batch"):
        return env.rollout_batch(all_us)
    batch, horizon = all_us.shape[0], all_us.shape[1]
    rewards = torch.zeros((batch, horizon), device=all_us.device)
    env._sync_main_to_rollout()
    for i in range(horizon):
        _, _, r, _, _ = env.step_rollout(all_us[:, i, :])
        rewards[:, i] = r
    env._sync_main_to_rollout()
    return rewards
