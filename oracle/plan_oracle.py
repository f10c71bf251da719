"""CPU oracle of the kinematic state integration -- TEST INFRASTRUCTURE, never on the product path.

Restates ``RobotPlanGradSampling._integrate_state_velocities`` and ``_sync_integration_to_sim``
(envs/batch_rollout/robot_plan_grad_sampling.py:103-195, :197-225 in /root/reference/legged_gym/legged_gym) as functions on a
plain state object ``o`` carrying the reference's attributes (``integration_*``, ``root_states``, ``dof_pos``, ``dof_vel``,
``base_lin_vel``, ``base_ang_vel``, ``dof_pos_limits`` and the ``max_*`` / ``integration_method`` / ``enforce_joint_limits``
settings).  Pinned by ``tests/test_plan_integration.py`` against the UNMODIFIED reference methods (container only) and
``tests/golden/plan_integration.npz`` (``tests/golden/make_plan_golden.py``).
"""
import numpy as np
import torch

from . import torch_utils as tu


def integrate_state_velocities(o, state_vels, dt, env_indices):
    lin = torch.clamp(state_vels[:, :3], -o.max_base_lin_vel, o.max_base_lin_vel)
    ang = torch.clamp(state_vels[:, 3:6], -o.max_base_ang_vel, o.max_base_ang_vel)
    jv = torch.clamp(state_vels[:, 6:], -o.max_joint_vel, o.max_joint_vel)
    n_steps = int(np.ceil(dt / min(dt, o.max_integration_step)))
    h = dt / n_steps
    for _ in range(n_steps):
        if o.integration_method == "euler":
            o.integration_base_pos[env_indices] += lin * h
        else:   # the reference's "rk4": four identical stages
            o.integration_base_pos[env_indices] += (lin + 2 * lin + 2 * lin + lin) * h / 6
        angle = torch.norm(ang, dim=1, keepdim=True) * h
        axis = ang / (torch.norm(ang, dim=1, keepdim=True) + 1e-8)
        rot = tu.quat_from_angle_axis(angle.squeeze(-1), axis)
        q = tu.quat_mul(o.integration_base_quat[env_indices], rot)
        o.integration_base_quat[env_indices] = q / torch.norm(q, dim=1, keepdim=True)
        o.integration_dof_pos[env_indices] += jv * h
    if o.enforce_joint_limits:
        o.integration_dof_pos[env_indices] = torch.clamp(o.integration_dof_pos[env_indices], o.dof_pos_limits[:, 0].unsqueeze(0),
                                                         o.dof_pos_limits[:, 1].unsqueeze(0))
    o.integration_base_lin_vel[env_indices] = lin
    o.integration_base_ang_vel[env_indices] = ang
    o.integration_dof_vel[env_indices] = jv


def sync_integration_to_sim(o, env_indices):
    o.root_states[env_indices, :3] = o.integration_base_pos[env_indices]
    o.root_states[env_indices, 3:7] = o.integration_base_quat[env_indices]
    o.root_states[env_indices, 7:10] = o.integration_base_lin_vel[env_indices]
    o.root_states[env_indices, 10:13] = o.integration_base_ang_vel[env_indices]
    o.dof_pos[env_indices] = o.integration_dof_pos[env_indices]
    o.dof_vel[env_indices] = o.integration_dof_vel[env_indices]
    q = o.integration_base_quat[env_indices]
    o.base_lin_vel[env_indices] = tu.quat_rotate_inverse(q, o.integration_base_lin_vel[env_indices])
    o.base_ang_vel[env_indices] = tu.quat_rotate_inverse(q, o.integration_base_ang_vel[env_indices])


KEYS = ("integration_base_pos", "integration_base_quat", "integration_dof_pos", "integration_base_lin_vel", "integration_base_ang_vel",
        "integration_dof_vel", "root_states", "dof_pos", "dof_vel", "base_lin_vel", "base_ang_vel")


def make_state(n, d, seed, method="euler", enforce=False, max_step=0.01):
    from types import SimpleNamespace
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(n, 4, generator=g)
    o = SimpleNamespace(total_num_envs=n, num_dof=d, device="cpu", integration_method=method, enforce_joint_limits=enforce,
                        max_base_lin_vel=3.0, max_base_ang_vel=2.0, max_joint_vel=10.0, max_integration_step=max_step)
    o.integration_base_pos = torch.randn(n, 3, generator=g)
    o.integration_base_quat = q / q.norm(dim=1, keepdim=True)
    o.integration_dof_pos = torch.randn(n, d, generator=g) * 0.5
    o.integration_base_lin_vel = torch.zeros(n, 3)
    o.integration_base_ang_vel = torch.zeros(n, 3)
    o.integration_dof_vel = torch.zeros(n, d)
    o.dof_pos_limits = torch.stack([-0.6 * torch.ones(d), 0.8 * torch.ones(d)], dim=1)
    o.root_states = torch.randn(n, 13, generator=g)
    o.dof_state = torch.randn(n * d, 2, generator=g)
    o.dof_pos = o.dof_state.view(n, d, 2)[..., 0]
    o.dof_vel = o.dof_state.view(n, d, 2)[..., 1]
    o.base_pos, o.base_quat = o.root_states[:, 0:3], o.root_states[:, 3:7]
    o.base_lin_vel = torch.randn(n, 3, generator=g)
    o.base_ang_vel = torch.randn(n, 3, generator=g)
    return o


def snapshot(o):
    return {k: getattr(o, k).clone() for k in KEYS}
