"""Navigation command update of the batch-rollout nav task (SURVEY section 8f-4;
RobotBatchRolloutNav._update_navigation_commands / _check_goal_reached, robot_batch_rollout_nav.py:135-247).

not gpu: the oracle against the golden vectors generated from the UNMODIFIED reference methods and (container only) against
those methods themselves; ABI checks.  gpu: elg_nav_commands against the golden vectors / the oracle, and the
RobotBatchRolloutNav class (commands reach this step's observations, goal flags stop the commands one step later)."""
import ctypes as C
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402
from extended_legged_gym_b200 import _lib, synthetic  # noqa: E402
from oracle import nav_oracle, ref_harness  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "nav_commands.npz")
DEV = "cuda:0"
RTOL, ATOL = 1e-5, 1e-6


def opt_for(use2d):
    return SimpleNamespace(use_2d_nav=use2d, kp_linear=1.0, kp_angular=2.0, max_linear_vel=1.0, max_angular_vel=1.0,
                           cmd_smooth_factor=0.1, tolerance_rad=0.5)


def load_case(tag):
    z = np.load(GOLDEN)
    m, r, use2d = (int(x) for x in z[f"{tag}__meta"])
    steps = [{k: torch.from_numpy(z[f"{tag}__s{s}__{k}"]) for k in ("root_states", "commands", "prev_commands", "goal_reached")} for s in range(3)]
    return m, r, bool(use2d), torch.from_numpy(z[f"{tag}__goals"]), torch.from_numpy(z[f"{tag}__commands0"]), steps


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_matches_reference_fixture(tag):
    m, r, use2d, goals, cmd, steps = load_case(tag)
    opt, prev, reached = opt_for(use2d), None, None
    cmd = cmd.clone()
    for st in steps:
        prev = nav_oracle.update_navigation_commands(st["root_states"], goals, r, opt, cmd, prev, reached)
        reached, _ = nav_oracle.check_goal_reached(st["root_states"], goals, r, opt)
        torch.testing.assert_close(cmd, st["commands"], rtol=RTOL, atol=ATOL)
        torch.testing.assert_close(prev, st["prev_commands"], rtol=RTOL, atol=ATOL)
        assert torch.equal(reached, st["goal_reached"])


@pytest.mark.skipif(not ref_harness.available(), reason="the reference checkout is only present in the build container")
def test_oracle_matches_live_reference_methods():
    ref_harness.install()
    from legged_gym.envs.batch_rollout.robot_batch_rollout_nav import RobotBatchRolloutNav as Ref
    g = torch.Generator().manual_seed(9)
    m, r = 6, 4
    n = m * (1 + r)
    for use2d in (True, False):
        opt = opt_for(use2d)
        root = torch.randn(n, 13, generator=g) * 2
        root[:, 3:7] /= root[:, 3:7].norm(dim=1, keepdim=True)
        goals = torch.randn(m, 3, generator=g) * 2
        o = SimpleNamespace(cfg=SimpleNamespace(navi_opt=opt), device="cpu", total_num_envs=n, num_rollout_per_main=r, root_states=root,
                            goal_positions=goals, commands=torch.randn(n, 4, generator=g), prev_commands=torch.randn(n, 3, generator=g),
                            goal_reached=torch.rand(n, generator=g) < 0.3)
        cmd, prev, reached = o.commands.clone(), o.prev_commands.clone(), o.goal_reached.clone()
        Ref._update_navigation_commands(o)
        Ref._check_goal_reached(o)
        prev2 = nav_oracle.update_navigation_commands(root, goals, r, opt, cmd, prev, reached)
        reached2, _ = nav_oracle.check_goal_reached(root, goals, r, opt)
        torch.testing.assert_close(cmd, o.commands, rtol=RTOL, atol=ATOL)
        torch.testing.assert_close(prev2, o.prev_commands, rtol=RTOL, atol=ATOL)
        assert torch.equal(reached2, o.goal_reached)


def test_nav_abi_argument_checks():
    lib = _lib.load()
    assert lib.elg_sizeof_nav_params() == C.sizeof(_lib.ElgNavParams)
    p = _lib.ElgNavParams()
    p.num_commands = 4
    assert lib.elg_nav_commands(2, 3, None, 16, 16, 16, 16, 16, None, None) == -4
    assert lib.elg_nav_commands(2, 3, C.byref(p), None, 16, 16, 16, 16, None, None) == -4
    p.num_commands = 2
    assert lib.elg_nav_commands(2, 3, C.byref(p), 16, 16, 16, 16, 16, None, None) == -1
    assert lib.elg_nav_commands(0, 3, C.byref(_lib.ElgNavParams(num_commands=4)), 16, 16, 16, 16, 16, None, None) == 0      # nothing to do


# ---------------------------------------------------------------------------------------------------------------
def run_kernel(m, r, opt, root, goals, cmd, prev, reached, use_prev, zero_reached):
    lib = _lib.load()
    p = _lib.ElgNavParams()
    p.use_2d_nav, p.use_prev, p.zero_reached, p.num_commands = int(opt.use_2d_nav), int(use_prev), int(zero_reached), cmd.shape[1]
    p.kp_linear, p.kp_angular, p.max_linear_vel, p.max_angular_vel = opt.kp_linear, opt.kp_angular, opt.max_linear_vel, opt.max_angular_vel
    p.smooth, p.smooth_c, p.tolerance_rad = opt.cmd_smooth_factor, 1 - opt.cmd_smooth_factor, opt.tolerance_rad
    dist = torch.zeros(root.shape[0], device=DEV)
    _lib.check(lib.elg_nav_commands(m, r, C.byref(p), root.data_ptr(), goals.data_ptr(), cmd.data_ptr(), prev.data_ptr(), reached.data_ptr(),
                                    dist.data_ptr(), None))
    torch.cuda.synchronize()
    return dist


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_nav_kernel_matches_reference_fixture(tag):
    m, r, use2d, goals, cmd0, steps = load_case(tag)
    opt = opt_for(use2d)
    cmd = cmd0.to(DEV).contiguous()
    prev = torch.zeros(cmd.shape[0], 3, device=DEV)
    reached = torch.zeros(cmd.shape[0], dtype=torch.bool, device=DEV)
    gd = goals.to(DEV).contiguous()
    for s, st in enumerate(steps):
        run_kernel(m, r, opt, st["root_states"].to(DEV).contiguous(), gd, cmd, prev, reached, use_prev=s > 0, zero_reached=s > 0)
        torch.testing.assert_close(cmd.cpu(), st["commands"], rtol=RTOL, atol=ATOL)
        torch.testing.assert_close(prev.cpu(), st["prev_commands"], rtol=RTOL, atol=ATOL)
        assert torch.equal(reached.cpu(), st["goal_reached"])


@pytest.mark.gpu
@pytest.mark.parametrize("use2d", [True, False])
def test_nav_kernel_matches_oracle_full_size(use2d):
    """BASELINE config 5 layout: 64 mains x 512 rollouts; robots on and next to the tolerance circle keep the flag bit-exact"""
    m, r = 64, 512
    n = m * (1 + r)
    g = torch.Generator().manual_seed(4)
    opt = opt_for(use2d)
    root = torch.randn(n, 13, generator=g) * 3
    root[:, 3:7] /= root[:, 3:7].norm(dim=1, keepdim=True)
    goals = torch.randn(m, 3, generator=g) * 3
    goal_env = nav_oracle.goal_per_env(goals, r)
    ring = torch.arange(0, n, 7)
    ang = torch.rand(len(ring), generator=g) * 6.28
    root[ring, 0] = goal_env[ring, 0] + 0.5 * torch.cos(ang)      # on the tolerance circle, up to rounding
    root[ring, 1] = goal_env[ring, 1] + 0.5 * torch.sin(ang)
    if not use2d:
        root[ring, 2] = goal_env[ring, 2]
    root[1::97, 0:3] = goal_env[1::97]                            # exactly at the goal: zero error, zero velocity
    cmd = torch.randn(n, 4, generator=g)
    prev = torch.randn(n, 3, generator=g)
    reached = torch.rand(n, generator=g) < 0.2
    cmd_d, prev_d, reached_d = cmd.to(DEV), prev.to(DEV), reached.to(DEV)
    dist = run_kernel(m, r, opt, root.to(DEV), goals.to(DEV), cmd_d, prev_d, reached_d, True, True)
    prev2 = nav_oracle.update_navigation_commands(root, goals, r, opt, cmd, prev, reached)
    reached2, dist2 = nav_oracle.check_goal_reached(root, goals, r, opt)
    torch.testing.assert_close(cmd_d.cpu(), cmd, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(prev_d.cpu(), prev2, rtol=RTOL, atol=ATOL)
    assert torch.equal(dist.cpu(), dist2), "distance to goal is not bit-identical (torch.norm chain)"
    assert torch.equal(reached_d.cpu(), reached2)


@pytest.mark.gpu
def test_nav_class_commands_feed_the_same_step_and_goal_flags_stop_them():
    from extended_legged_gym_b200.envs import RobotBatchRolloutNav, RobotBatchRolloutNavCfg
    from extended_legged_gym_b200.envs.anymal_c.anymal_c_config import AnymalCRoughCfg
    from extended_legged_gym_b200.sim_backend import SyntheticSim

    class Cfg(AnymalCRoughCfg, RobotBatchRolloutNavCfg):
        class env(AnymalCRoughCfg.env):
            num_envs = 4
            rollout_envs = 3
            episode_length_s = 30

        class navi_opt(RobotBatchRolloutNavCfg.navi_opt):
            goal_pos = [[5.0, 5.0, 0.5], [0.0, 0.0, 0.5], [20.0, 3.0, 0.5]]

        class commands(RobotBatchRolloutNavCfg.commands):
            pass

        class domain_rand(AnymalCRoughCfg.domain_rand):
            rollout_envs_sync_pos_drift = 0.0
            push_robots = False

    cfg = Cfg()
    n = 4 * 4
    _, spec, st = common.make_case_state("anymal_c_rough", n, seed=8)
    hf = synthetic.make_height_field(seed=0)
    env = RobotBatchRolloutNav(cfg, None, SyntheticSim(cfg, n, DEV, spec=spec, height_samples=hf, state=st), DEV, True)
    env.set_env_state(st)
    assert env.goal_positions.shape == (4, 3) and torch.equal(env.goal_positions[3], env.goal_positions[2])     # padded with the last goal
    env.add_noise = False
    env.root_states[4:8, 0:2] = torch.tensor([0.1, 0.1], device=DEV)          # main 1 and its rollouts sit on their goal
    env.post_physics_step()
    torch.cuda.synchronize()
    opt = cfg.navi_opt
    cmd = st["commands"].clone()
    nav_oracle.update_navigation_commands(env.root_states.cpu(), env.goal_positions.cpu(), 3, opt, cmd, None, None)
    not_reset = ~env.reset_buf.cpu()
    torch.testing.assert_close(env.commands.cpu()[not_reset][:, :3], cmd[not_reset][:, :3], rtol=RTOL, atol=ATOL)
    scale = torch.tensor([2.0, 2.0, 0.25])
    torch.testing.assert_close(env.obs_buf.cpu()[not_reset][:, 9:12], cmd[not_reset][:, :3] * scale, rtol=RTOL, atol=ATOL)
    assert env.get_goal_reached_status(main_env_only=False)[4:8].all() and not env.get_goal_reached_status()[0]
    env.post_physics_step()                                                   # the flags of the previous check stop the commands now
    torch.cuda.synchronize()
    assert float(env.commands[4:8].abs().max()) == 0.0
    assert env.get_distance_to_goal().shape == (4,)
