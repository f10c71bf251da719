"""TEST INFRASTRUCTURE ONLY (oracle) -- the checker, never the product path.

torch-CPU restatement of the MPPI cost-weighted update, batched over main envs.  The production optimiser of the
reference is the external ``traj_sampling`` package (PegasusFlow; unpinned, absent from /root/reference and from this
image -- call sites envs/batch_rollout/robot_traj_grad_sampling.py:62-69, :222-280).  The only in-tree statement of
the update is legged_gym/tests/score_sampling/cmp_mppi_wbfo.py:216-233, which this follows line by line;
``tests/test_mppi.py`` pins it against that unmodified method (container only) and tests/golden/mppi.npz.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.
"""
import torch


def mppi_update(step_rewards, samples, temp):
    """step_rewards [M, S, T], samples [M, S, K, D] -> mean trajectories [M, K, D] (cmp_mppi_wbfo.py:220-233 per main)."""
    out = []
    for m in range(step_rewards.shape[0]):
        costs = torch.sum(step_rewards[m], dim=1)
        cost_mean = costs.mean()
        cost_std = costs.std() + 1e-6
        normalized_costs = (costs - cost_mean) / cost_std
        weights = torch.softmax(normalized_costs / temp, dim=0)
        out.append(torch.sum(weights.view(-1, 1, 1) * samples[m], dim=0))
    return torch.stack(out)


# the three local stages of the sharded update (csrc/elg_mppi.cu), used by the gloo tests as stand-ins for the kernels
def local_costs(step_rewards):
    return step_rewards.sum(dim=2)


def local_partials(costs_all, first, samples, temp):
    M, S_local = samples.shape[0], samples.shape[1]
    mean = costs_all.mean(dim=1, keepdim=True)
    std = costs_all.std(dim=1, keepdim=True) + 1e-6
    n = (costs_all - mean) / std
    mx = n.max(dim=1, keepdim=True).values
    e = torch.exp((n[:, first:first + S_local] - mx) / temp)
    flat = samples.reshape(M, S_local, -1)
    return torch.cat([e.sum(dim=1, keepdim=True), (e.unsqueeze(-1) * flat).sum(dim=1)], dim=1)


def finish(partial, traj_shape):
    return (partial[:, 1:] / partial[:, :1]).reshape(partial.shape[0], *traj_shape)
