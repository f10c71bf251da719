"""Generate tests/golden/nav_commands.npz from the UNMODIFIED reference (container only):
RobotBatchRolloutNav._update_navigation_commands / _check_goal_reached
(envs/batch_rollout/robot_batch_rollout_nav.py:135-247) bound to a synthetic ``self``, three consecutive callbacks
(prev_commands None on the first, as after construction), 2-D and 3-D navigation.

    python tests/golden/make_nav_golden.py
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402


def make_state(m, r, seed):
    g = torch.Generator().manual_seed(seed)
    n = m * (1 + r)
    root = torch.zeros(n, 13)
    root[:, 0:2] = torch.rand(n, 2, generator=g) * 8 - 1
    root[:, 2] = 0.5 + torch.randn(n, generator=g) * 0.05
    q = torch.randn(n, 4, generator=g) * torch.tensor([0.1, 0.1, 1.0, 1.0])
    root[:, 3:7] = q / q.norm(dim=1, keepdim=True)
    root[:, 7:13] = torch.randn(n, 6, generator=g)
    goals = torch.rand(m, 3, generator=g) * 6
    goals[:, 2] = 0.5
    # a few envs sit on / next to their goal: the tolerance test and the zero-velocity limit
    for k in range(min(m, 3)):
        root[k * (1 + r) + 1, 0:3] = goals[k] + torch.tensor([0.3 * k, 0.1, 0.0])
    return root, goals


def main():
    ref_harness.install()
    from legged_gym.envs.batch_rollout.robot_batch_rollout_nav import RobotBatchRolloutNav as Ref
    out = {}
    for tag, (m, r, use2d, seed) in {"a": (5, 3, True, 0), "b": (4, 6, False, 1)}.items():
        opt = SimpleNamespace(use_2d_nav=use2d, kp_linear=1.0, kp_angular=2.0, max_linear_vel=1.0, max_angular_vel=1.0,
                              cmd_smooth_factor=0.1, tolerance_rad=0.5)
        root, goals = make_state(m, r, seed)
        n = root.shape[0]
        o = SimpleNamespace(cfg=SimpleNamespace(navi_opt=opt), device="cpu", total_num_envs=n, num_rollout_per_main=r, root_states=root,
                            goal_positions=goals, commands=torch.randn(n, 4, generator=torch.Generator().manual_seed(5)),
                            prev_commands=None, goal_reached=None)
        out[f"{tag}__meta"] = np.array([m, r, int(use2d)], dtype=np.int64)
        out[f"{tag}__goals"] = goals.numpy()
        out[f"{tag}__commands0"] = o.commands.clone().numpy()
        g = torch.Generator().manual_seed(100 + seed)
        for s in range(3):
            out[f"{tag}__s{s}__root_states"] = o.root_states.clone().numpy()
            Ref._update_navigation_commands(o)
            Ref._check_goal_reached(o)
            out[f"{tag}__s{s}__commands"] = o.commands.clone().numpy()
            out[f"{tag}__s{s}__prev_commands"] = o.prev_commands.clone().numpy()
            out[f"{tag}__s{s}__goal_reached"] = o.goal_reached.clone().numpy()
            # the robots move a little towards wherever they are sent
            o.root_states = o.root_states.clone()
            o.root_states[:, 0:2] += 0.3 * torch.rand(n, 2, generator=g) * torch.sign(o.commands[:, 0:2])
    path = os.path.join(ROOT, "tests", "golden", "nav_commands.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
