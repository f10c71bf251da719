"""TEST INFRASTRUCTURE ONLY (oracle) -- the checker, never the product path.

torch-CPU restatement of the main/rollout state clone of the reference
(/root/reference/legged_gym/legged_gym/envs/batch_rollout/robot_batch_rollout.py):

  init_env_indices        :119-164   main env k at row k (1 + R), its rollouts behind it
  sync_main_to_rollout    :1447-1535 14 gather -> scatter pairs (+ optional position drift :1493-1497)
  cache_main_env_states   :1537-1583
  restore_main_env_states :1585-1640

``tests/test_rollout_oracle.py`` pins it against the unmodified reference methods bound to a synthetic ``self``
(container only) and against ``tests/golden/rollout_clone.npz`` (generated from the reference).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.
"""
from types import SimpleNamespace

import torch

SYNC_ORDER = ("root_states", "dof_pos", "dof_vel", "actions", "last_actions", "last_dof_vel", "last_root_vel", "base_pos",
              "base_quat", "base_lin_vel", "base_ang_vel", "projected_gravity")          # then drift, then the feet state
SYNC_FEET = ("feet_air_time", "feet_contact_time", "last_contacts")
CACHE_KEYS = ("root_states", "dof_pos", "dof_vel", "actions", "last_actions", "last_dof_vel", "last_root_vel", "base_pos",
              "base_quat", "base_lin_vel", "base_ang_vel", "base_lin_acc", "base_ang_acc", "projected_gravity", "feet_air_time",
              "feet_contact_time", "last_contacts")


def make_rollout_state(num_main, rollouts, num_dof=12, num_feet=4, seed=0):
    """Synthetic per-env tensors with the aliasing of the reference (_init_buffers, legged_robot.py:575-580)."""
    g = torch.Generator().manual_seed(seed)
    n = num_main * (1 + rollouts)
    o = SimpleNamespace()
    o.num_main_envs, o.num_rollout_per_main, o.total_num_envs, o.device = num_main, rollouts, n, "cpu"
    o.root_states = torch.randn(n, 13, generator=g)
    o.dof_state = torch.randn(n * num_dof, 2, generator=g)
    o.dof_pos = o.dof_state.view(n, num_dof, 2)[..., 0]
    o.dof_vel = o.dof_state.view(n, num_dof, 2)[..., 1]
    o.base_pos = o.root_states[:, :3]
    o.base_quat = o.root_states[:, 3:7]
    for name, w in (("actions", num_dof), ("last_actions", num_dof), ("last_dof_vel", num_dof), ("last_root_vel", 6),
                    ("base_lin_vel", 3), ("base_ang_vel", 3), ("base_lin_acc", 3), ("base_ang_acc", 3), ("projected_gravity", 3),
                    ("feet_air_time", num_feet), ("feet_contact_time", num_feet)):
        setattr(o, name, torch.randn(n, w, generator=g))
    o.last_contacts = torch.rand(n, num_feet, generator=g) > 0.5
    return o


def init_env_indices(o):
    R, total = o.num_rollout_per_main, o.total_num_envs
    o.main_env_indices = torch.arange(0, total, 1 + R)
    o.rollout_to_main_map = torch.zeros(total, dtype=torch.long)
    o.is_main_env = torch.zeros(total, dtype=torch.bool)
    o.is_main_env[o.main_env_indices] = True
    o.is_rollout_env = ~o.is_main_env
    for i in range(o.num_main_envs):
        m = int(o.main_env_indices[i])
        o.rollout_to_main_map[m:min(m + 1 + R, total)] = m
    o.rollout_env_indices = torch.nonzero(o.is_rollout_env).flatten()
    o.main_to_rollout_indices = []
    for i in range(o.num_main_envs):
        m = o.main_env_indices[i]
        idx = torch.nonzero(o.rollout_to_main_map == m).flatten()
        o.main_to_rollout_indices.append(idx[idx != m])


def sync_main_to_rollout(o, drift=0.0, drift_u=None):
    """drift_u: the [num_rollout, 3] sample torch.rand_like(base_pos[rollout]) would return (None: draw it)."""
    ri = o.rollout_env_indices
    if len(ri) == 0:
        return
    src = o.rollout_to_main_map[ri]
    for name in SYNC_ORDER:
        t = getattr(o, name)
        t[ri] = t[src]
    if drift > 0.0:
        u = torch.rand_like(o.base_pos[ri]) if drift_u is None else drift_u
        o.base_pos[ri] += (u - 0.5) * drift
    for name in SYNC_FEET:
        t = getattr(o, name)
        t[ri] = t[src]


def cache_main_env_states(o):
    mi = o.main_env_indices
    o.main_env_cache = {k: getattr(o, k)[mi].clone() for k in CACHE_KEYS}


def restore_main_env_states(o):
    mi = o.main_env_indices
    for k in CACHE_KEYS:
        getattr(o, k)[mi] = o.main_env_cache[k].clone()


STATE_KEYS = ("root_states", "dof_state", "actions", "last_actions", "last_dof_vel", "last_root_vel", "base_lin_vel", "base_ang_vel",
              "base_lin_acc", "base_ang_acc", "projected_gravity", "feet_air_time", "feet_contact_time", "last_contacts")


def snapshot(o):
    return {k: getattr(o, k).clone() for k in STATE_KEYS}
