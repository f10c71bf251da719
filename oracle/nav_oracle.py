"""CPU oracle of the navigation command update -- TEST INFRASTRUCTURE, never on the product path.

Restates ``RobotBatchRolloutNav._update_navigation_commands`` and ``_check_goal_reached``
(envs/batch_rollout/robot_batch_rollout_nav.py:135-222, :224-247 in /root/reference/legged_gym/legged_gym) without the
reference's Python loops over ``total_num_envs`` (the goal of env i is the goal of main env i // (1 + rollouts), :144-147).
Pinned by ``tests/test_nav_commands.py``: against the UNMODIFIED reference methods bound to a synthetic ``self`` (container
only) and against ``tests/golden/nav_commands.npz`` generated from them (``tests/golden/make_nav_golden.py``).
"""
import torch

from . import torch_utils as tu


def goal_per_env(goal_positions, rollouts_per_main):
    return goal_positions.repeat_interleave(1 + rollouts_per_main, dim=0)


def update_navigation_commands(root_states, goal_positions, rollouts_per_main, opt, commands, prev_commands, goal_reached):
    """-> prev_commands' (the smoothed commands); ``commands`` is updated in place like the reference does.
    opt: object with use_2d_nav, kp_linear, kp_angular, max_linear_vel, max_angular_vel, cmd_smooth_factor."""
    n = root_states.shape[0]
    pos, quat = root_states[:, 0:3], root_states[:, 3:7]
    goal = goal_per_env(goal_positions, rollouts_per_main)
    desired = torch.zeros(n, 3)
    if opt.use_2d_nav:
        err = goal[:, 0:2] - pos[:, 0:2]
        desired[:, 0:2] = opt.kp_linear * err
        mag = torch.norm(desired[:, 0:2], dim=1)
        scale = torch.clamp(opt.max_linear_vel / (mag + 1e-8), max=1.0)
        desired[:, 0:2] *= scale.unsqueeze(1)
    else:
        err = goal - pos
        desired = opt.kp_linear * err
        mag = torch.norm(desired, dim=1)
        scale = torch.clamp(opt.max_linear_vel / (mag + 1e-8), max=1.0)
        desired = desired * scale.unsqueeze(1)
    robot = tu.quat_rotate_inverse(quat, desired)
    if opt.use_2d_nav:
        yaw = torch.atan2(2 * (quat[:, 3] * quat[:, 2] + quat[:, 0] * quat[:, 1]), 1 - 2 * (quat[:, 1] ** 2 + quat[:, 2] ** 2))
        want = torch.atan2(err[:, 1], err[:, 0])
        d = want - yaw
        d = torch.atan2(torch.sin(d), torch.cos(d))
        ang = torch.clamp(opt.kp_angular * d, -opt.max_angular_vel, opt.max_angular_vel)
    else:
        ang = torch.zeros(n)
    new = torch.stack([robot[:, 0], robot[:, 1], ang], dim=1)
    if prev_commands is not None:
        a = opt.cmd_smooth_factor
        smoothed = a * prev_commands + (1 - a) * new
    else:
        smoothed = new
    commands[:, 0:3] = smoothed
    if goal_reached is not None:
        commands[goal_reached] = 0.0
    return smoothed.clone()


def check_goal_reached(root_states, goal_positions, rollouts_per_main, opt):
    goal = goal_per_env(goal_positions, rollouts_per_main)
    pos = root_states[:, 0:3]
    dist = torch.norm(goal[:, 0:2] - pos[:, 0:2], dim=1) if opt.use_2d_nav else torch.norm(goal - pos, dim=1)
    return dist < opt.tolerance_rad, dist
