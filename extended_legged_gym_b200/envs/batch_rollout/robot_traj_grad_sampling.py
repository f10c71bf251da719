"""``RobotTrajGradSampling`` -- env-side glue of the sampling-based trajectory optimiser, host side.

Mirrors envs/batch_rollout/robot_traj_grad_sampling.py of the reference for everything that lives in the tree: the action
(de)normalisation of :282-345, ``step`` / ``step_rollout`` with it (:347-373), ``rollout_batch`` (:249-280) and
``optimize_all_trajectories`` (:226-247).  The optimiser itself is the external ``traj_sampling`` package (PegasusFlow;
unpinned, absent from the reference tree and from this image): what is built here is its in-tree statement, the MPPI
cost-weighted update of tests/score_sampling/cmp_mppi_wbfo.py:216-233 (``utils/mppi.py``: sampling around the node
trajectories, ``rollout_batch``, ``elg_mppi_update``), with the rollout dimension sharded over ranks when an ``ElgComm`` is set.
"""
import torch

from ...utils import mppi as _mppi
from .robot_batch_rollout import RobotBatchRollout


class RobotTrajGradSampling(RobotBatchRollout):
    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless)
        self._init_action_normalization()
        self.comm = None                       # utils.distributed.ElgComm when the rollouts of every main env are sharded over ranks
        to = getattr(cfg, "trajectory_opt", None)
        self.traj_opt_enabled = bool(to is not None and getattr(to, "enable_traj_opt", False))
        if self.traj_opt_enabled:
            self.horizon_samples = int(to.horizon_samples)
            self.horizon_nodes = int(to.horizon_nodes)
            self.temp_sample = float(to.temp_sample)
            self.noise_scaling = float(getattr(to, "noise_scaling", 1.0))
            self.num_diffuse_steps = int(getattr(to, "num_diffuse_steps", 1))
            self.num_diffuse_steps_init = int(getattr(to, "num_diffuse_steps_init", self.num_diffuse_steps))
            # node trajectories [num_main, horizon_nodes + 1, A]; control sequences are their linear interpolation
            self.node_trajectories = torch.zeros(self.num_main_envs, self.horizon_nodes + 1, self.num_actions, device=self.device)
            t = torch.linspace(0, self.horizon_nodes, self.horizon_samples + 1, device=self.device)
            lo = t.floor().clamp(max=self.horizon_nodes - 1).long()
            self._interp = (lo, (t - lo).view(1, -1, 1))

    # ------------------------------------------------------------------------------------------
    # action (de)normalisation (robot_traj_grad_sampling.py:282-345)
    # ------------------------------------------------------------------------------------------
    def _init_action_normalization(self):
        self.use_action_normalization = bool(getattr(self.cfg.control, "jointpos_action_normalization", False))
        self._action_denorm = (None, None)
        if self.use_action_normalization:
            spec = self.sim.spec
            q0 = self.default_dof_pos.view(-1)
            self.joint_lower_limits = torch.tensor(spec.dof_lower, dtype=torch.float, device=self.device) - q0
            self.joint_upper_limits = torch.tensor(spec.dof_upper, dtype=torch.float, device=self.device) - q0
            self.joint_ranges = self.joint_upper_limits - self.joint_lower_limits
            self.joint_mid_points = (self.joint_upper_limits + self.joint_lower_limits) / 2.0
            self._action_denorm = (self.joint_lower_limits.contiguous(), self.joint_ranges.contiguous())

    def _normalize_actions(self, joint_targets):
        if not self.use_action_normalization:
            return joint_targets
        return torch.clamp(2.0 * (joint_targets - self.joint_lower_limits) / self.joint_ranges - 1.0, -1.0, 1.0)

    def _denormalize_actions(self, normalized_actions):
        if not self.use_action_normalization:
            return normalized_actions
        normalized_actions = torch.clamp(normalized_actions, -1.0, 1.0)
        return self.joint_lower_limits + (normalized_actions + 1.0) * self.joint_ranges / 2.0

    def step(self, actions):
        if self.use_action_normalization:
            actions = self._denormalize_actions(actions)
        out = super().step(actions)
        if self.traj_opt_enabled:
            self.shift_trajectory_batch()
        return out

    def step_rollout(self, actions, action_noise=None):
        if self.use_action_normalization:
            actions = self._denormalize_actions(actions)
        return super().step_rollout(actions, action_noise)

    # ------------------------------------------------------------------------------------------
    # the optimiser loop around rollout_batch
    # ------------------------------------------------------------------------------------------
    def node2u(self, nodes):
        """[..., horizon_nodes + 1, A] node trajectories -> [..., horizon_samples + 1, A] control sequences (linear)."""
        lo, w = self._interp
        return nodes[..., lo, :] * (1 - w) + nodes[..., lo + 1, :] * w

    def shift_trajectory_batch(self):
        """advance the node trajectories by one control step (the first node drops out, the last one is repeated)"""
        u = self.node2u(self.node_trajectories)
        shifted = torch.cat([u[:, 1:], u[:, -1:]], dim=1)
        idx = torch.linspace(0, self.horizon_samples, self.horizon_nodes + 1, device=self.device).round().long()
        self.node_trajectories = shifted[:, idx]

    def optimize_all_trajectories(self, n_diffuse=None, initial=False):
        """``n_diffuse`` MPPI iterations for all main envs in batch (:226-247): sample node trajectories around the means
        (one sample per local rollout env), roll them out, cost-weighted update (all-gather / all-reduce over ``self.comm``)."""
        if not self.traj_opt_enabled:
            return []
        n = n_diffuse if n_diffuse is not None else (self.num_diffuse_steps_init if initial else self.num_diffuse_steps)
        M, R = self.num_main_envs, self.num_rollout_per_main
        for _ in range(n):
            eps = torch.randn(M, R, self.horizon_nodes + 1, self.num_actions, device=self.device) * self.noise_scaling
            eps[:, 0] = 0.0                                           # the first rollout of every main env carries the mean itself
            samples = self.node_trajectories.unsqueeze(1) + eps
            us = self.node2u(samples)[:, :, :self.horizon_samples]    # [M, R, horizon, A]
            rew = self.rollout_batch(us.reshape(M * R, self.horizon_samples, self.num_actions))
            self.node_trajectories = _mppi.mppi_update(rew.view(M, R, self.horizon_samples), samples, self.temp_sample, comm=self.comm)
        return []
