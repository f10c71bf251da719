"""Robot-specific main / rollout classes (envs/anymal_c/batch_rollout/anymal_c_batch_rollout.py:49-225,
envs/go2/batch_rollout/go2_batch_rollout.py:49-230 in /root/reference/legged_gym/legged_gym): ``RobotBatchRolloutPercept`` plus the
actuator-network torque path, the gait scheduler on the env clock and the reset of upside-down MAIN robots.
not-gpu: ``RobotBatchRolloutOracle`` bit-identical to tests/golden/rollout_step_anymal.npz (generated from the unmodified
``AnymalCBatchRollout``: tests/golden/make_rollout_step_golden.py --robot); the configs against the reference's.
gpu: ``AnymalCBatchRollout.post_physics_step`` against the same fixture (host-driven and fused reset paths), ``step`` /
``step_rollout`` / ``rollout_batch`` with the actuator network."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402
from oracle.rollout_oracle import RobotBatchRolloutOracle  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
from extended_legged_gym_b200 import _lib, synthetic  # noqa: E402
import test_rollout_step as trs  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rollout_step_anymal.npz")
ACTNET = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "actuator_net.npz")
DEV = "cuda:0"
CASE = "anymal_c_rough"


def load(tag="c"):
    old = trs.GOLDEN
    trs.GOLDEN = GOLDEN
    try:
        return trs.load(tag)
    finally:
        trs.GOLDEN = old


ROBOT_TAGS = {"c": ("anymal_c_rough", dict(upside_down_rows="main", gait_period=1.0)),
              "d": ("elspider_air_rough", dict(upside_down_rows="all", gait_period=None))}


@pytest.mark.parametrize("tag", list(ROBOT_TAGS))
def test_robot_rollout_oracle_matches_reference_fixture(tag):
    m, r, seed, c0, inputs, outs, names = load(tag)
    case, kw = ROBOT_TAGS[tag]
    cfg_cls, spec_fn, _ = common.CASES[case]
    ora = RobotBatchRolloutOracle(cfg_cls(), spec_fn(), {k: v.clone() for k, v in inputs.items()}, synthetic.make_height_field(seed=0), m, r, **kw)
    ora.common_step_counter = c0
    flipped_rollout_reset = flipped_rollout_kept = flipped_main_reset = False
    for s, want in enumerate(outs):
        torch.manual_seed(5000 + 17 * s + seed)
        ora.torques = ora.compute_torques(ora.actions).view(ora.torques.shape)
        up_before = ora.root_states[:, 3] == 1.0                 # the rows the generator turned over (quaternion (1, 0, 0, 0))
        ora.post_physics_step(noise_u=want["noise_u"])
        snap = trs.snapshot(ora)
        if tag == "c":
            snap["gait_idx"] = ora.gait_idx
        trs.check(snap, want, f"{tag} step {s}", exact=True)
        sums = torch.stack([ora.episode_sums[k] for k in names])
        assert torch.equal(sums, want["episode_sums"]), f"{tag} step {s}: episode sums differ"
        ora.t_main += ora.dt
        rb, up = want["reset_buf"].bool(), want["projected_gravity"][:, 2] > 0
        to = want["time_out_buf"].bool()
        flipped_rollout_kept |= bool((up & ~rb)[ora.rollout_env_indices].any())
        flipped_rollout_reset |= bool((up_before & rb & ~to)[ora.rollout_env_indices].any())
        flipped_main_reset |= bool((up_before & rb)[ora.main_env_indices].any())
    assert flipped_main_reset
    # ANYmal: an upside-down rollout robot is left alone; hexapod: it is reset like every other row
    assert flipped_rollout_kept if tag == "c" else flipped_rollout_reset


@pytest.mark.skipif(not rh.available(), reason="the reference checkout is only present in the build container")
def test_robot_rollout_configs_match_the_reference():
    rh.install()
    from legged_gym.envs.anymal_c.batch_rollout.anymal_c_batch_rollout_config import AnymalCBatchRolloutCfg as RefA
    from legged_gym.envs.go2.batch_rollout.go2_batch_rollout_config import Go2BatchRolloutCfg as RefG
    from legged_gym.envs.elspider_air.batch_rollout.elspider_air_batch_rollout_config import ElSpiderAirBatchRolloutCfg as RefE
    from extended_legged_gym_b200.envs import AnymalCBatchRolloutCfg, Go2BatchRolloutCfg, ElSpiderAirBatchRolloutCfg
    from extended_legged_gym_b200.utils.helpers import class_to_dict
    for ours, ref in ((AnymalCBatchRolloutCfg, RefA), (Go2BatchRolloutCfg, RefG), (ElSpiderAirBatchRolloutCfg, RefE)):
        a, b = class_to_dict(ours), class_to_dict(ref)
        for block in ("gait_scheduler", "control", "init_state", "commands"):
            for k, v in b[block].items():
                assert a[block][k] == v, f"{ours.__name__}.{block}.{k}: {a[block][k]} != {v}"
        assert a["rewards"]["scales"] == b["rewards"]["scales"], ours.__name__
        for k in ("max_contact_force", "base_height_target", "only_positive_rewards"):
            assert a["rewards"][k] == b["rewards"][k], k
        for k in ("name", "foot_name", "penalize_contacts_on", "terminate_after_contacts_on", "self_collisions"):
            assert a["asset"][k] == b["asset"][k], k
        for k in ("num_observations", "num_actions", "episode_length_s"):
            assert a["env"][k] == b["env"][k], k
        for k in ("mesh_type", "measure_heights", "curriculum"):
            assert a["terrain"][k] == b["terrain"][k], k


ROLLOUT_MODE_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rollout_mode_step.npz")


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_rollout_mode_step_oracle_matches_reference_fixture(tag):
    """``post_physics_step_rollout`` (envs/batch_rollout/robot_batch_rollout.py:763-817 and the robot classes' scheduler call) -- the
    step the horizon loop of the sampling-based optimiser repeats -- as the oracles restate it, against the unmodified reference
    methods' output after one main step (tests/golden/make_rollout_step_golden.py --rollout-mode): rollout rows, bit for bit.  (The
    reference refreshes the derived state of the rollout rows only; main rows are put back by _restore_main_env_states.)"""
    from oracle.rollout_oracle import BatchRolloutOracle
    z = np.load(ROLLOUT_MODE_GOLDEN)
    m, r, seed, steps, c0 = (int(x) for x in z[f"{tag}__meta"])
    inputs = {k[len(tag) + 6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"{tag}__in__")}
    want = {k[len(tag) + 7:]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith(f"{tag}__out__")}
    case = {**trs.TAGS, **{t: c for t, (c, _) in ROBOT_TAGS.items()}}[tag]
    cfg_cls, spec_fn, _ = common.CASES[case]
    hf = synthetic.make_height_field(seed=0)
    st = {k: v.clone() for k, v in inputs.items()}
    if tag in ROBOT_TAGS:
        ora = RobotBatchRolloutOracle(cfg_cls(), spec_fn(), st, hf, m, r, **ROBOT_TAGS[tag][1])
    else:
        ora = BatchRolloutOracle(cfg_cls(), spec_fn(), st, hf, m, r)
    ora.common_step_counter = c0
    torch.manual_seed(5000 + seed)
    ora.torques = ora.compute_torques(ora.actions).view(ora.torques.shape)
    ora.post_physics_step(noise_u=want["noise_u_main"])
    t_rollout = 0.0
    if tag == "c":
        ora.t_main += ora.dt
        t_rollout = ora.t_main
    ora.actions[:] = want["actions_rollout"]
    ora.torques = ora.compute_torques(ora.actions).view(ora.torques.shape)
    if tag in ROBOT_TAGS:
        ora.post_physics_step_rollout(noise_u=want["noise_u_rollout"], t_rollout=t_rollout)
    else:
        ora.post_physics_step_rollout(noise_u=want["noise_u_rollout"])
    snap = trs.snapshot(ora)
    if tag in ROBOT_TAGS:
        snap["gait_idx"] = ora.gait_idx
    rows = ora.rollout_env_indices
    checked = 0
    for k, w in want.items():
        if k in ("noise_u_main", "noise_u_rollout", "actions_rollout", "episode_sums") or k.startswith("extras__"):
            continue
        g = snap[k]
        if g.dim() > 0 and g.shape[0] == ora.num_envs * ora.num_dof:       # dof_state [N * D, 2]
            g, w = g.view(ora.num_envs, -1), w.view(ora.num_envs, -1)
        if g.dim() > 0 and g.shape[0] == ora.num_envs:
            g, w = g[rows], w[rows]
        assert torch.equal(g.to(w.dtype), w), f"{tag}: {k} differs from the reference's rollout-mode step"
        checked += 1
    assert checked >= 20


def test_gait_clock_phase_equals_the_schedulers_expression():
    """GaitClockMixin.clock_phase against GaitScheduler.step(..., t) (utils/gait_scheduler.py:63-72): the torch expression of the
    reference on 80 000 clock values accumulated like t_main / t_rollout (+= dt), and -- in the container -- the reference object"""
    from extended_legged_gym_b200.envs.anymal_c.batch_rollout.anymal_c_batch_rollout import GaitClockMixin
    sched = None
    if rh.available():
        rh.install()
        from legged_gym.utils.gait_scheduler import GaitScheduler, GaitSchedulerCfg
        sched = GaitScheduler(None, *([None] * 8), 3, "cpu", gait_cfg=GaitSchedulerCfg())
    for period, dt in ((1.0, 0.02), (0.6, 0.005 * 4), (1.4, 0.02), (0.37, 0.017)):
        t = 0.0
        for i in range(20000):
            want = torch.remainder(t / period * torch.ones(3, dtype=torch.float), 1.0)
            assert GaitClockMixin.clock_phase(t, period) == float(want[0]), (period, i)
            if sched is not None and i % 997 == 0:
                sched.gait_cfg.period = period
                sched.step(None, None, None, t)
                assert torch.equal(sched.gait_idx, want)
            t += dt


def make_env(fused, inputs, m, r, tag="c"):
    from extended_legged_gym_b200.envs import AnymalCBatchRollout, ElSpiderAirBatchRollout
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    case, kw = ROBOT_TAGS[tag]
    cfg_cls, spec_fn, _ = common.CASES[case]
    cfg, spec = cfg_cls(), spec_fn()
    cfg.env.num_envs, cfg.env.rollout_envs = m, r
    cfg.control.use_actuator_network = False          # the fixtures' torques are the PD controller's (the classes' configs: off)
    if tag == "c":
        from extended_legged_gym_b200.envs import AnymalCBatchRolloutCfg
        cfg.gait_scheduler = AnymalCBatchRolloutCfg.gait_scheduler      # the scheduler config the fixture's reference object was given
    n = m * (1 + r)
    hf = synthetic.make_height_field(seed=0)
    ora = RobotBatchRolloutOracle(cfg_cls(), spec_fn(), {k: v.clone() for k, v in inputs.items()}, hf, m, r, **kw)    # terrain bookkeeping only
    cls = AnymalCBatchRollout if tag == "c" else ElSpiderAirBatchRollout
    env = cls(cfg, None, SyntheticSim(cfg, n, DEV, spec=spec, height_samples=hf, state={k: v.clone() for k, v in inputs.items()}), DEV, True)
    env.set_env_state(inputs)
    env.fused_reset = fused
    env._rand = lambda lo, hi, shape: ((hi - lo) * torch.rand(*shape) + lo).to(DEV)
    env._randint_like = lambda t, high: torch.randint_like(t.cpu(), high).to(DEV)
    if getattr(ora, "custom_origins", False):
        env.terrain_levels = ora.terrain_levels.clone().to(DEV)
        env.terrain_types = ora.terrain_types.clone().to(DEV)
        env.terrain_origins = ora.terrain_origins.clone().to(DEV)
        env.env_origins = ora.env_origins.clone().to(DEV)
    return env


@pytest.mark.gpu
def test_anymal_rollout_env_matches_reference_fixture():
    m, r, seed, c0, inputs, outs, names = load()
    env = make_env(False, inputs, m, r)
    env.common_step_counter = c0
    assert env._native_params().terminate_upside_down == 2
    for s, want in enumerate(outs):
        torch.manual_seed(5000 + 17 * s + seed)
        env.noise_u = want["noise_u"].to(DEV)
        env.torques = env._compute_torques(env.actions).view(env.torques.shape)
        env.post_physics_step()
        torch.cuda.synchronize()
        snap = trs.snapshot(env)
        snap["gait_idx"] = env.gait_idx
        trs.check(snap, want, f"c step {s}", exact=False)
        sums = torch.stack([env.episode_sums[k] for k in names]).cpu()
        assert torch.allclose(sums, want["episode_sums"], rtol=1e-5, atol=1e-6), f"c step {s}: episode sums differ"
        env.t_main += env.dt
    rb = env.reset_buf.cpu().bool()
    up = env.projected_gravity[:, 2].cpu() > 0
    assert bool((up & ~rb)[env.rollout_env_indices.cpu()].any())      # the upside-down rollout robot stays


@pytest.mark.gpu
def test_anymal_rollout_fused_reset_equals_host_path():
    """in-kernel termination (upside-down main rows included) + in-kernel reset == the host-driven path under shared uniforms"""
    from test_fused_reset import feed_host_path_from_table
    m, r = 12, 5
    n = m * (1 + r)
    _, _, st = common.make_case_state(CASE, n, seed=9, adversarial=True)
    for row in (0, 7, 6 * 3, 6 * 3 + 2, 6 * 7):
        st["root_states"][row, 3:7] = torch.tensor([1.0, 0.0, 0.0, 0.0])
    U = torch.rand(n, _lib.RESET_UNIFORMS, generator=torch.Generator().manual_seed(13)).to(DEV)
    a, b = make_env(True, st, m, r), make_env(False, st, m, r)
    for e in (a, b):
        e.cfg.domain_rand.push_robots = False
    a.reset_uniforms = U
    feed_host_path_from_table(b, U)
    g = torch.Generator().manual_seed(5)
    for step in range(2):
        u = torch.rand(n, a.num_obs, generator=g).to(DEV)
        for env in (a, b):
            env.noise_u = u
            env.torques = env._compute_torques(env.actions).view(env.torques.shape)
            env._obs_clip_for_step = 100.0
            env.post_physics_step()
        torch.cuda.synchronize()
        sa, sb = trs.snapshot(a), trs.snapshot(b)
        for k in sb:
            if sb[k] is not None:
                assert torch.equal(sa[k].cpu(), sb[k].cpu()), f"step {step}: fused vs host path differ in {k}"
        if step == 0:
            rb = b.reset_buf.cpu().bool()
            assert bool(rb[0]) and bool(rb[18]) and bool(rb[42])        # upside-down mains
    assert torch.equal(a.gait_idx.cpu(), b.gait_idx.cpu())


@pytest.mark.gpu
def test_anymal_rollout_steps_with_the_actuator_network():
    """step / step_rollout / rollout_batch of the robot class with the LSTM torque path on every row (``_compute_torques`` with
    and without ``env_ids``), network state cleared for reset rows, the lean kernel kept for the rollout-mode step"""
    from extended_legged_gym_b200.envs import AnymalCBatchRollout, AnymalCBatchRolloutCfg
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    from extended_legged_gym_b200.envs import robot_specs
    cfg = AnymalCBatchRolloutCfg()
    cfg.env.num_envs, cfg.env.rollout_envs = 6, 7
    cfg.control.use_actuator_network = True
    cfg.control.actuator_net_weights = ACTNET
    m, r = 6, 7
    n = m * (1 + r)
    spec = robot_specs.anymal_c()
    env = AnymalCBatchRollout(cfg, None, SyntheticSim(cfg, n, DEV, spec=spec, seed=4), DEV, True)
    env.root_states[8, 3:7] = torch.tensor([1.0, 0.0, 0.0, 0.0], device=DEV)       # main 1 upside down (its rollouts follow at the sync)
    assert env.sea_hidden_state.shape == (2, n * 12, 8) and env.num_obs == 48
    g = torch.Generator().manual_seed(1)
    a_main = torch.randn(m, 12, generator=g).to(DEV)
    env.sea_hidden_state.normal_(generator=None)
    obs, _, rew, reset, extras = env.step(a_main)
    torch.cuda.synchronize()
    assert obs.shape == (m, 48) and rew.shape == (m,) and reset.shape == (m,)
    assert bool(reset[1])                                               # the upside-down main env terminated ...
    assert float(env.sea_hidden_state_per_env[:, 8].abs().max()) == 0.0   # ... and its network state was cleared
    t_all = env._compute_torques(env.actions).clone()
    ids = torch.tensor([3, 9, 20], device=DEV)
    env.sea_hidden_state.zero_(); env.sea_cell_state.zero_()
    t_ref = env._compute_torques(env.actions).clone()
    env.sea_hidden_state.zero_(); env.sea_cell_state.zero_()
    assert torch.equal(env._compute_torques(env.actions, env_ids=ids), t_ref[ids]) and t_all.shape == (n, 12)
    obs_r, _, rew_r, reset_r, _ = env.step_rollout(torch.randn(m * r, 12, generator=g).to(DEV))
    assert obs_r.shape == (m * r, 48) and rew_r.shape == (m * r,)
    assert abs(float(env.gait_idx[0]) - np.fmod(np.float32(env.t_rollout - env.dt) / np.float32(1.0), 1.0)) < 1e-6
    rewards = env.rollout_batch(torch.randn(m * r, 5, 12, generator=g).to(DEV) * 0.3)
    torch.cuda.synchronize()
    assert rewards.shape == (m * r, 5) and bool(torch.isfinite(rewards).all())


@pytest.mark.gpu
def test_go2_rollout_class_steps_on_a_plane():
    """Go2BatchRollout (envs/go2/batch_rollout/go2_batch_rollout.py:49-230) with its own config -- multi-stage reward scales, gait
    scheduler period 0.6 s on the env clock, upside-down main robots reset -- on flat ground with the sensors switched off (the
    config's terrain OBJ is a file of the author's machine)"""
    from extended_legged_gym_b200.envs import Go2BatchRollout, Go2BatchRolloutCfg, robot_specs
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    cfg = Go2BatchRolloutCfg()
    cfg.env.num_envs, cfg.env.rollout_envs, cfg.env.num_observations = 5, 3, 48
    cfg.terrain.mesh_type, cfg.terrain.use_terrain_obj = "plane", False
    cfg.raycaster.enable_raycast = False
    cfg.sdf.enable_sdf = False
    n = 5 * 4
    env = Go2BatchRollout(cfg, None, SyntheticSim(cfg, n, DEV, spec=robot_specs.go2(), seed=2), DEV, True)
    assert env._native_params().terminate_upside_down == 2
    assert "feet_slip" not in env.reward_scales and "dof_pos_limits" in env.reward_scales      # stage 0 of [-0.0, -0.4] is off
    env.root_states[4, 3:7] = torch.tensor([1.0, 0.0, 0.0, 0.0], device=DEV)                    # main 1 upside down
    obs, _, rew, reset, _ = env.step(torch.zeros(5, 12, device=DEV))
    torch.cuda.synchronize()
    assert obs.shape == (5, 48) and bool(reset[1]) and bool(torch.isfinite(rew).all())
    assert float(env.gait_idx[0]) == 0.0                                                        # remainder(t_main = 0 / 0.6, 1)
    env.step(torch.zeros(5, 12, device=DEV))
    assert abs(float(env.gait_idx[7]) - float(np.float32(env.dt / 0.6))) < 1e-7


# ---------------------------------------------------------------------------------------------------------------
# AnymalCTrajGradSampling: the DIAL-MPC reward set against the unmodified reference methods (CPU, bit for bit)
# ---------------------------------------------------------------------------------------------------------------
DIAL_TERMS = ("gaits", "air_time", "pos", "upright", "yaw", "vel", "ang_vel", "height", "energy", "alive", "no_fly")
DIAL_GAITS = ("trot", "walk", "gallop", "stand")
DIAL_N = 257
DIAL_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dial_mpc_terms.npz")


def _dial_fake_self(cls, n, seed, heading, flags_dtype):
    """a bare object of ``cls`` carrying the tensors the terms read (the classes' __init__ needs a simulator / a GPU)"""
    from types import SimpleNamespace
    g = torch.Generator().manual_seed(seed)
    o = object.__new__(cls)
    d = o.__dict__
    d["device"], d["total_num_envs"], d["num_envs"] = "cpu", n, n
    d["dt"], d["t_main"], d["t_rollout"] = 0.02, 1.3, 1.46
    d["feet_indices"] = torch.tensor([4, 8, 12, 16])
    q = torch.randn(n, 4, generator=g)
    rs = torch.randn(n, 13, generator=g)
    rs[:, 3:7] = q / q.norm(dim=1, keepdim=True)
    d["root_states"] = rs
    d["base_quat"] = rs[:, 3:7]
    d["commands"] = torch.randn(n, 4, generator=g)
    d["foot_positions"] = torch.rand(n, 4, 3, generator=g) * 0.2
    d["contact_forces"] = torch.randn(n, 17, 3, generator=g) * 2.0
    d["last_contacts"] = torch.rand(n, 4, generator=g) < 0.5
    d["feet_air_time"] = torch.rand(n, 4, generator=g) * (torch.rand(n, 4, generator=g) < 0.7)
    d["projected_gravity"] = torch.randn(n, 3, generator=g)
    d["base_lin_vel"] = torch.randn(n, 3, generator=g)
    d["base_ang_vel"] = torch.randn(n, 3, generator=g)
    d["torques"] = torch.randn(n, 12, generator=g) * 30
    d["dof_vel"] = torch.randn(n, 12, generator=g) * 3
    d["reset_buf"] = (torch.rand(n, generator=g) < 0.3).to(flags_dtype)
    d["cfg"] = SimpleNamespace(commands=SimpleNamespace(heading_command=heading), rewards=SimpleNamespace(base_height_target=0.5))
    return o


@pytest.mark.parametrize("robot", ["anymal_c", "go2"])
@pytest.mark.parametrize("heading", [False, True])
def test_dial_mpc_reward_terms_match_the_reference_fixture(heading, robot):
    """tests/golden/dial_mpc_terms.npz holds the unmodified reference methods' outputs on the seeded tensors of _dial_fake_self
    (tests/golden/make_dial_golden.py): the fixture travels to the GPU box, the reference does not"""
    from extended_legged_gym_b200.envs import AnymalCTrajGradSampling, Go2TrajGradSampling
    cls = AnymalCTrajGradSampling if robot == "anymal_c" else Go2TrajGradSampling
    z = np.load(DIAL_GOLDEN)
    checked = 0
    for gait in DIAL_GAITS:
        ours = _dial_fake_self(cls, DIAL_N, 5, heading, torch.long)
        ours._init_dial_mpc()
        ours._gait = gait
        for name in DIAL_TERMS:
            key = f"{robot}__{int(heading)}__{gait}__{name}"
            if key in z.files:
                assert torch.equal(getattr(ours, "_reward_" + name)().float(), torch.from_numpy(z[key])), key
                checked += 1
        assert torch.equal(ours.feet_air_time, torch.from_numpy(z[f"{robot}__{int(heading)}__{gait}__feet_air_time"]))
        assert torch.equal(ours.last_contacts, torch.from_numpy(z[f"{robot}__{int(heading)}__{gait}__last_contacts"]))
    assert checked == len(DIAL_GAITS) * (11 if robot == "anymal_c" else 10)


@pytest.mark.skipif(not rh.available(), reason="the reference checkout is only present in the build container")
@pytest.mark.parametrize("robot", ["anymal_c", "go2"])
@pytest.mark.parametrize("heading", [False, True])
def test_dial_mpc_reward_terms_equal_the_reference_methods(heading, robot):
    rh.install()
    if robot == "anymal_c":
        from legged_gym.envs.anymal_c.batch_rollout.anymal_c_traj_grad_sampling import AnymalCTrajGradSampling as Ref
        from extended_legged_gym_b200.envs import AnymalCTrajGradSampling
    else:
        from legged_gym.envs.go2.batch_rollout.go2_traj_grad_sampling import Go2TrajGradSampling as Ref
        from extended_legged_gym_b200.envs import Go2TrajGradSampling as AnymalCTrajGradSampling
    from extended_legged_gym_b200.envs.anymal_c.batch_rollout.anymal_c_traj_grad_sampling import DialMpcRewardMixin
    n = DIAL_N
    for gait in DIAL_GAITS:
        ours = _dial_fake_self(AnymalCTrajGradSampling, n, 5, heading, torch.long)
        ref = _dial_fake_self(Ref, n, 5, heading, torch.long)
        ours._init_dial_mpc()
        ours._gait = gait
        # the tables of the reference's __init__ (:40-57), built by its own statements' values
        ref._gait = gait
        ref._gait_phase = {k: (torch.zeros(4) if k == "stand" else torch.tensor(v)) for k, v in DialMpcRewardMixin.GAIT_PHASES.items()}
        ref._gait_params = {k: torch.tensor(v) for k, v in DialMpcRewardMixin.GAIT_PARAMS.items()}
        for name in DIAL_TERMS:
            if not hasattr(ref, "_reward_" + name):      # (no_fly exists in the ANYmal file only)
                assert robot == "go2" and name == "no_fly"
                continue
            got, want = getattr(ours, "_reward_" + name)(), getattr(ref, "_reward_" + name)()
            assert got.shape == want.shape == (n,) and torch.equal(got.float(), want.float()), f"{name} ({gait})"
        # the bookkeeping of air_time went through both objects the same way
        assert torch.equal(ours.feet_air_time, ref.feet_air_time) and torch.equal(ours.last_contacts, ref.last_contacts)
    # flags are bool in this framework: alive is then 1 - flag (the reference's expression raises on a bool tensor)
    ours = _dial_fake_self(AnymalCTrajGradSampling, n, 6, heading, torch.bool)
    assert torch.equal(ours._reward_alive(), (~ours.reset_buf).float())
    if robot == "go2":      # (its reference expression converts the flags: bool works there too)
        ref = _dial_fake_self(Ref, n, 6, heading, torch.bool)
        assert torch.equal(ours._reward_alive(), ref._reward_alive())


@pytest.mark.skipif(not rh.available(), reason="the reference checkout is only present in the build container")
@pytest.mark.parametrize("robot", ["anymal_c", "go2"])
def test_traj_grad_sampling_config_matches_the_reference(robot):
    rh.install()
    if robot == "anymal_c":
        from legged_gym.envs.anymal_c.batch_rollout.anymal_c_traj_grad_sampling_config import AnymalCTrajGradSamplingCfg as Ref
        from extended_legged_gym_b200.envs import AnymalCTrajGradSamplingCfg
    else:
        from legged_gym.envs.go2.batch_rollout.go2_traj_grad_sampling_config import Go2TrajGradSamplingCfg as Ref
        from extended_legged_gym_b200.envs import Go2TrajGradSamplingCfg as AnymalCTrajGradSamplingCfg
    from extended_legged_gym_b200.utils.helpers import class_to_dict
    # (the reference's trajectory_opt / rl_warmstart blocks derive from the absent traj_sampling package -- a stub here: only the
    # blocks defined in the tree are walked, trajectory_opt by attribute)
    a = class_to_dict(AnymalCTrajGradSamplingCfg)
    b = {k: class_to_dict(getattr(Ref, k)) for k in ("gait_scheduler", "control", "init_state", "commands", "rewards", "asset", "env")}
    for block in ("gait_scheduler", "control", "init_state", "commands"):
        for k, v in b[block].items():
            assert a[block][k] == v, f"{block}.{k}: {a[block][k]} != {v}"
    assert a["rewards"]["scales"] == b["rewards"]["scales"]
    for k in ("max_contact_force", "base_height_target", "only_positive_rewards", "tracking_sigma"):
        assert a["rewards"][k] == b["rewards"][k], k
    for k in ("name", "foot_name", "penalize_contacts_on", "terminate_after_contacts_on", "self_collisions"):
        assert a["asset"][k] == b["asset"][k], k
    for k in ("num_envs", "rollout_envs", "num_observations", "num_actions", "episode_length_s"):
        assert a["env"][k] == b["env"][k], k
    # trajectory_opt: its base class lives in traj_sampling, so the class statement yields a stub -- the values are read off the source
    import ast
    import inspect
    import re
    src = inspect.getsource(sys.modules[Ref.__module__])
    block = src[src.index("class trajectory_opt("):src.index("class rl_warmstart(")]
    ref_to = {m.group(1): ast.literal_eval(m.group(2).strip()) for m in re.finditer(r"^\s+(\w+) = ([^#\n]+)", block, re.M)}
    assert len(ref_to) >= 12
    for k, v in ref_to.items():
        assert a["trajectory_opt"][k] == v, f"trajectory_opt.{k}: {a['trajectory_opt'][k]} != {v}"


@pytest.mark.gpu
def test_anymal_traj_grad_sampling_class_on_the_device():
    """AnymalCTrajGradSampling with its default config (stock terms only: one MPPI iteration through the CUDA-graph horizon loop),
    and with two DIAL-MPC terms switched on: the rollout step's reward is then the stock reward plus the scaled Python terms, and
    rollout_batch (eager with Python terms) still fills every column of its reward table"""
    from extended_legged_gym_b200.envs import AnymalCTrajGradSampling, AnymalCTrajGradSamplingCfg, robot_specs
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    m, r = 3, 8
    n = m * (1 + r)

    def make(extra_scales=None):
        cfg = AnymalCTrajGradSamplingCfg()
        cfg.env.num_envs, cfg.env.rollout_envs = m, r
        for k, v in (extra_scales or {}).items():
            setattr(cfg.rewards.scales, k, v)
        return AnymalCTrajGradSampling(cfg, None, SyntheticSim(cfg, n, DEV, spec=robot_specs.anymal_c(), seed=6), DEV, True)

    a = make()
    assert not a._python_terms and a.traj_opt_enabled and a.horizon_samples == 16
    a.optimize_all_trajectories()
    torch.cuda.synchronize()
    assert a.node_trajectories.shape == (m, 5, 12) and bool(torch.isfinite(a.node_trajectories).all())
    a2, b = make(), make({"upright": 0.7, "energy": 0.2})
    assert sorted(b._python_terms) == ["energy", "upright"]
    acts = torch.randn(m * r, 12, generator=torch.Generator().manual_seed(3)).to(DEV)
    rew_a = a2.step_rollout(acts)[2]
    rew_b = b.step_rollout(acts)[2]
    torch.cuda.synchronize()
    want = ((b._reward_upright() * 0.7 + b._reward_energy() * 0.2) * b.dt)[b.rollout_env_indices]
    assert torch.allclose(rew_b - rew_a, want, rtol=1e-4, atol=1e-5), float((rew_b - rew_a - want).abs().max())
    assert abs(float(a2.gait_idx[5]) - float(np.remainder(np.float32((a2.t_rollout - a2.dt) / 1.0), np.float32(1.0)))) < 1e-7
    us = torch.randn(m * r, 4, 12, generator=torch.Generator().manual_seed(4)).to(DEV) * 0.3
    tab = b.rollout_batch(us)
    torch.cuda.synchronize()
    assert tab.shape == (m * r, 4) and bool(torch.isfinite(tab).all()) and bool((tab != 0).all())


@pytest.mark.gpu
def test_go2_traj_grad_sampling_rewards_are_the_python_terms():
    """Go2TrajGradSampling with its default config: six DIAL-MPC terms and no stock term -- the rollout step's reward is exactly
    their scaled sum (evaluated between the DERIVE launch and the registry launch), rollout_batch fills its table"""
    from extended_legged_gym_b200.envs import Go2TrajGradSampling, Go2TrajGradSamplingCfg, robot_specs
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    cfg = Go2TrajGradSamplingCfg()
    cfg.env.num_envs, cfg.env.rollout_envs = 2, 5
    env = Go2TrajGradSampling(cfg, None, SyntheticSim(cfg, 12, DEV, spec=robot_specs.go2(), seed=8), DEV, True)
    scales = {"gaits": 0.1, "upright": 0.5, "yaw": 0.3, "vel": 1.0, "ang_vel": 0.3, "height": 10.0}
    assert not env._kernel_terms and sorted(env._python_terms) == sorted(scales)
    rew = env.step_rollout(torch.randn(10, 12, generator=torch.Generator().manual_seed(2)).to(DEV))[2]
    torch.cuda.synchronize()
    env.t_rollout -= env.dt            # the terms saw the clock before step_rollout advanced it
    want = sum(getattr(env, "_reward_" + k)() * s for k, s in scales.items()) * env.dt
    env.t_rollout += env.dt
    assert torch.allclose(rew, want[env.rollout_env_indices], rtol=1e-5, atol=1e-6), float((rew - want[env.rollout_env_indices]).abs().max())
    tab = env.rollout_batch(torch.randn(10, 3, 12, generator=torch.Generator().manual_seed(3)).to(DEV) * 0.3)
    torch.cuda.synchronize()
    assert tab.shape == (10, 3) and bool(torch.isfinite(tab).all()) and bool((tab != 0).all())


@pytest.mark.gpu
def test_elspider_rollout_env_matches_reference_fixture():
    """ElSpiderAirBatchRollout.post_physics_step (generic kernel, 18 DOF / 6 feet, main / rollout layout) against the fixture of the
    unmodified reference class: every upside-down row resets, the main step leaves the gait scheduler alone"""
    m, r, seed, c0, inputs, outs, names = load("d")
    env = make_env(False, inputs, m, r, tag="d")
    env.common_step_counter = c0
    assert env._native_params().terminate_upside_down == 1 and env.num_dof == 18
    env.gait_idx.fill_(0.37)
    env.gait_prev_foot_z.fill_(0.011)
    saw = False
    for s, want in enumerate(outs):
        torch.manual_seed(5000 + 17 * s + seed)
        env.noise_u = want["noise_u"].to(DEV)
        env.torques = env._compute_torques(env.actions).view(env.torques.shape)
        up_before = (env.root_states[:, 3] == 1.0).cpu()
        env.post_physics_step()
        torch.cuda.synchronize()
        trs.check(trs.snapshot(env), want, f"d step {s}", exact=False)
        sums = torch.stack([env.episode_sums[k] for k in names]).cpu()
        assert torch.allclose(sums, want["episode_sums"], rtol=1e-5, atol=1e-6), f"d step {s}: episode sums differ"
        rb, to = env.reset_buf.cpu().bool(), env.time_out_buf.cpu().bool()
        saw |= bool((up_before & rb & ~to)[env.rollout_env_indices.cpu()].any())
        env.t_main += env.dt
    assert saw                                                            # an upside-down ROLLOUT robot was reset
    assert bool((env.gait_idx == 0.37).all()) and bool((env.gait_prev_foot_z == 0.011).all())


@pytest.mark.skipif(not rh.available(), reason="the reference checkout is only present in the build container")
def test_action_normalisation_equals_the_reference_methods():
    """RobotTrajGradSampling._normalize_actions / _denormalize_actions (robot_traj_grad_sampling.py:307-345) against the unmodified
    reference methods on the same limits, normalisation on and off (CPU, bit for bit)"""
    rh.install()
    from legged_gym.envs.batch_rollout.robot_traj_grad_sampling import RobotTrajGradSampling as Ref
    from extended_legged_gym_b200.envs import RobotTrajGradSampling
    g = torch.Generator().manual_seed(12)
    lower = -torch.rand(12, generator=g) - 0.2
    upper = torch.rand(12, generator=g) + 0.3
    x = torch.randn(300, 12, generator=g) * 1.5
    for on in (True, False):
        objs = []
        for cls in (RobotTrajGradSampling, Ref):
            o = object.__new__(cls)
            o.__dict__.update(use_action_normalization=on, joint_lower_limits=lower.clone(), joint_upper_limits=upper.clone(),
                              joint_ranges=upper - lower, joint_mid_points=(upper + lower) / 2.0)
            objs.append(o)
        ours, ref = objs
        assert torch.equal(ours._normalize_actions(x), ref._normalize_actions(x))
        assert torch.equal(ours._denormalize_actions(x), ref._denormalize_actions(x))
        if on:      # a round trip inside the limits is the identity up to rounding
            inside = lower + (upper - lower) * torch.rand(300, 12, generator=g)
            assert torch.allclose(ours._denormalize_actions(ours._normalize_actions(inside)), inside, atol=1e-6)
