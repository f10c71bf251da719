"""Shared test fixtures: the BASELINE configs as (cfg factory, robot spec) cases, seeded synthetic
state, an 'everything enabled' stress config, and comparison helpers with the tolerances of
BASELINE.json north_star (bit-exact integer/bool, 1e-5 rel / 1e-6 abs for fp32)."""
import copy

import numpy as np
import torch

from extended_legged_gym_b200 import synthetic
from extended_legged_gym_b200.envs import robot_specs as rs
from extended_legged_gym_b200.envs.a1.a1_config import A1RoughCfg
from extended_legged_gym_b200.envs.anymal_c.anymal_c_config import AnymalCFlatCfg, AnymalCRoughCfg
from extended_legged_gym_b200.envs.go2.go2_config import Go2RoughCfg
from extended_legged_gym_b200.envs.elspider_air.elspider_air_config import ElSpiderAirRoughCfg

RTOL, ATOL = 1e-5, 1e-6

ALL_TERM_SCALES = dict(
    action_rate=-0.01, ang_vel_xy=-0.05, base_foot_height=-1.5, base_height=-2.0, collision=-1.0, dof_acc=-2.5e-7,
    dof_pos_limits=-10.0, dof_vel=-1e-3, dof_vel_limits=-0.3, feet_air_time=1.0, feet_contact_forces=-0.01, feet_slip=-0.1,
    feet_stumble=-0.5, feet_stumble_liftup=0.3, four_footup=-1.0, gait_2_step=-0.2, jump_air=-0.4, lin_vel_z=-2.0,
    orientation=-0.5, stand_still=-0.1, termination=-3.0, torque_limits=-0.02, torques=-1e-5, tracking_ang_vel=0.5,
    tracking_lin_vel=1.0)


def _all_terms(base_cls, heading=False, only_positive=False):
    class Cfg(base_cls):
        class rewards(base_cls.rewards):
            only_positive_rewards = only_positive
            soft_dof_vel_limit = 0.05
            soft_torque_limit = 0.3
            max_contact_force = 60.0

            class scales(base_cls.rewards.scales):
                pass

        class commands(base_cls.commands):
            heading_command = heading
    for k, v in ALL_TERM_SCALES.items():
        setattr(Cfg.rewards.scales, k, v)
    Cfg.__name__ = base_cls.__name__ + "AllTerms"
    return Cfg


class ElSpiderPDCfg(ElSpiderAirRoughCfg):
    """the hexapod config on the PD torque path (the actuator-network path has its own tests) at reward stage 1"""
    class control(ElSpiderAirRoughCfg.control):
        use_actuator_network = False

    class terrain(ElSpiderAirRoughCfg.terrain):
        border_size = 25          # the synthetic height field's border (the reference config's 100 m would clip every scan point)

    class rewards(ElSpiderAirRoughCfg.rewards):
        reward_min_stage = 1


def _control(base_cls, mode):
    """the other two controller types of _compute_torques (legged_robot.py:425-448): velocity targets (its damping term divides
    by sim_params.dt, not the env dt -- SURVEY App. A-13) and direct torques"""
    class Cfg(base_cls):
        class control(base_cls.control):
            control_type = mode
    Cfg.__name__ = f"{base_cls.__name__}Control{mode}"
    return Cfg


A1ControlVCfg = _control(A1RoughCfg, "V")
Go2ControlTCfg = _control(Go2RoughCfg, "T")
A1AllTermsCfg = _all_terms(A1RoughCfg)
ElSpiderAllTermsCfg = _all_terms(ElSpiderPDCfg, heading=True)
ElSpiderAllTermsCfg.rewards.multi_stage_rewards = False
Go2AllTermsHeadingCfg = _all_terms(Go2RoughCfg, heading=True, only_positive=True)

CASES = {
    "anymal_c_flat": (AnymalCFlatCfg, rs.anymal_c, "AnymalCFlatCfg"),
    "anymal_c_rough": (AnymalCRoughCfg, rs.anymal_c, "AnymalCRoughCfg"),
    "a1_rough": (A1RoughCfg, rs.a1, "A1RoughCfg"),
    "go2_rough": (Go2RoughCfg, rs.go2, "Go2RoughCfg"),
    "a1_all_terms": (A1AllTermsCfg, rs.a1, None),
    "go2_all_terms_heading": (Go2AllTermsHeadingCfg, rs.go2, None),
    "elspider_air_rough": (ElSpiderPDCfg, rs.elspider_air, None),
    "elspider_all_terms": (ElSpiderAllTermsCfg, rs.elspider_air, None),
    "a1_control_V": (A1ControlVCfg, rs.a1, None),
    "go2_control_T": (Go2ControlTCfg, rs.go2, None),
}
# reference class the golden generator runs for a case (default: LeggedRobot)
REF_ENV_CLASS = {"elspider_air_rough": "ElSpider", "elspider_all_terms": "ElSpider"}

STATE_FIELDS = ["obs_buf", "rew_buf", "reset_buf", "time_out_buf", "episode_length_buf", "base_lin_vel", "base_ang_vel",
                "base_lin_acc", "base_ang_acc", "projected_gravity", "foot_positions", "foot_velocities", "torques", "commands",
                "feet_air_time", "feet_contact_time", "last_contacts", "last_actions", "last_dof_vel", "last_root_vel",
                "root_states", "dof_state", "measured_heights"]
EXACT_FIELDS = {"reset_buf", "time_out_buf", "episode_length_buf", "last_contacts"}


def make_case_state(case, num_envs, seed=0, adversarial=False):
    cfg_cls, spec_fn, _ = CASES[case]
    cfg, spec = cfg_cls(), spec_fn()
    q0 = [cfg.init_state.default_joint_angles[n] for n in spec.dof_names]
    st = synthetic.make_state(num_envs, spec.num_dof, spec.num_bodies, spec.indices_matching(cfg.asset.foot_name),
                              spec.indices_matching(cfg.asset.penalize_contacts_on),
                              spec.indices_matching(cfg.asset.terminate_after_contacts_on), q0, spec.foot_offsets,
                              num_commands=cfg.commands.num_commands, seed=seed)
    if adversarial:
        make_adversarial(st, cfg, spec, seed)
    return cfg, spec, st


def make_adversarial(st, cfg, spec, seed=0):
    """Push inputs onto the decision thresholds: |F| = 1 and 0.1 within a few ulp, Fz = 1 +- ulp,
    stumble ratio 5 within an ulp -- the masks must still be bit-exact."""
    g = torch.Generator().manual_seed(seed + 77)
    N = st["root_states"].shape[0]
    B = spec.num_bodies
    cf = st["contact_forces"].view(N, B, 3)

    def near(shape, radius):
        v = torch.randn(*shape, 3, generator=g)
        v = v / v.norm(dim=-1, keepdim=True)
        k = torch.randint(-4, 5, shape, generator=g).float()
        return v * (radius * (1.0 + k * 2.0 ** -23)).unsqueeze(-1)

    term = spec.indices_matching(cfg.asset.terminate_after_contacts_on)
    pen = spec.indices_matching(cfg.asset.penalize_contacts_on)
    feet = spec.indices_matching(cfg.asset.foot_name)
    half = N // 2
    cf[:half, term] = near((half, len(term)), 1.0)
    cf[:half, pen] = near((half, len(pen)), 0.1)
    fz = 1.0 + torch.randint(-3, 4, (half, len(feet)), generator=g).float() * 2.0 ** -23
    cf[:half, feet, 2] = fz
    ang = torch.rand(half, len(feet), generator=g) * 6.28
    r = 5.0 * fz * (1.0 + torch.randint(-3, 4, (half, len(feet)), generator=g).float() * 2.0 ** -23)
    cf[:half, feet, 0] = r * torch.cos(ang)
    cf[:half, feet, 1] = r * torch.sin(ang)
    # episode lengths straddling the time-out and the command-resampling boundaries
    st["episode_length_buf"][:half] = torch.randint(997, 1003, (half,), generator=g)
    st["episode_length_buf"][half:half + half // 2] = torch.randint(497, 502, (half // 2,), generator=g)


def assert_state_close(got, want, fields=None, what=""):
    """got/want: dict name -> tensor (CPU)."""
    bad = []
    for f in fields or sorted(want.keys()):
        if f not in want or want[f] is None or not torch.is_tensor(want[f]):
            continue
        a, b = got[f].cpu(), want[f].cpu()
        if a.shape != b.shape:
            bad.append(f"{f}: shape {tuple(a.shape)} vs {tuple(b.shape)}")
            continue
        if f in EXACT_FIELDS or not b.dtype.is_floating_point:
            n = int((a.to(b.dtype) != b).sum())
            if n:
                bad.append(f"{f}: {n} of {b.numel()} entries differ (must be bit-exact)")
        else:
            ok = torch.isclose(a, b, rtol=RTOL, atol=ATOL, equal_nan=True)
            if not bool(ok.all()):
                i = int((~ok).flatten().nonzero()[0])
                bad.append(f"{f}: {int((~ok).sum())} of {b.numel()} outside rtol={RTOL} atol={ATOL}; first idx {i}: "
                           f"{a.flatten()[i].item()!r} vs {b.flatten()[i].item()!r}")
    assert not bad, what + "\n  " + "\n  ".join(bad)


def snapshot(obj, fields=STATE_FIELDS):
    out = {}
    for f in fields:
        v = getattr(obj, f, None)
        if torch.is_tensor(v):
            out[f] = v.detach().cpu().clone()
    for k, v in getattr(obj, "episode_sums", {}).items():
        out["sum_" + k] = v.detach().cpu().clone()
    return out
