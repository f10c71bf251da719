"""The main / rollout variant of the per-step path (commands, time-outs, curriculum and episode statistics belong to
MAIN envs -- envs/batch_rollout/robot_batch_rollout.py:819-850, :857-866, :876-940, :1366-1413).
not-gpu: ``BatchRolloutOracle`` bit-identical to tests/golden/rollout_step.npz, generated from the unmodified reference class.
gpu: ``RobotBatchRollout.post_physics_step`` (host-driven reset path, shared RNG) against the same fixture; the fused reset /
resample kernels against the host path under shared uniforms."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402
from oracle.rollout_oracle import BatchRolloutOracle  # noqa: E402
from extended_legged_gym_b200 import _lib, synthetic  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rollout_step.npz")
TAGS = {"a": "anymal_c_rough", "b": "go2_all_terms_heading"}
DEV = "cuda:0"


def load(tag):
    z = np.load(GOLDEN)
    m, r, seed, steps, c0 = (int(x) for x in z[f"{tag}__meta"])
    inputs = {k[len(tag) + 6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"{tag}__in__")}
    outs = [{k.split("__", 2)[2]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith(f"{tag}__s{s}__")} for s in range(steps)]
    return m, r, seed, c0, inputs, outs, [str(x) for x in z[f"{tag}__meta__sum_names"]]


def check(got, want, what, exact):
    for k, w in want.items():
        if k == "noise_u" or k.startswith("extras__") or k == "episode_sums":
            continue
        g = got[k].cpu()
        if exact or not w.dtype.is_floating_point or k in common.EXACT_FIELDS:
            assert torch.equal(g.to(w.dtype), w), f"{what}: {k} differs from the reference fixture"
        else:
            assert torch.allclose(g, w, rtol=common.RTOL, atol=common.ATOL, equal_nan=True), f"{what}: {k} outside tolerance"


def snapshot(o):
    s = common.snapshot(o)
    for k in ("env_origins", "terrain_levels"):
        if hasattr(o, k):
            s[k] = getattr(o, k)
    return s


@pytest.mark.parametrize("tag", list(TAGS))
def test_rollout_oracle_matches_reference_fixture(tag):
    m, r, seed, c0, inputs, outs, names = load(tag)
    cfg_cls, spec_fn, _ = common.CASES[TAGS[tag]]
    ora = BatchRolloutOracle(cfg_cls(), spec_fn(), {k: v.clone() for k, v in inputs.items()}, synthetic.make_height_field(seed=0), m, r)
    ora.common_step_counter = c0
    saw_rollout_reset = saw_main_reset = False
    for s, want in enumerate(outs):
        torch.manual_seed(5000 + 17 * s + seed)
        ora.torques = ora.compute_torques(ora.actions).view(ora.torques.shape)
        ora.post_physics_step(noise_u=want["noise_u"])
        check(snapshot(ora), want, f"{tag} step {s}", exact=True)
        sums = torch.stack([ora.episode_sums[k] for k in names])
        assert torch.equal(sums, want["episode_sums"]), f"{tag} step {s}: episode sums differ"
        for k, w in want.items():
            if k.startswith("extras__"):
                assert float(ora.extras["episode"][k[8:]]) == float(w), f"{tag} step {s}: {k}"
        rb = want["reset_buf"].bool()
        saw_main_reset |= bool(rb[ora.main_env_indices].any())
        saw_rollout_reset |= bool(rb[ora.rollout_env_indices].any())
    assert saw_main_reset and saw_rollout_reset      # the fixture exercises both kinds of reset


def make_env(tag, fused, inputs, m, r):
    from extended_legged_gym_b200.envs import RobotBatchRollout
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    cfg_cls, spec_fn, _ = common.CASES[TAGS[tag]]
    cfg, spec = cfg_cls(), spec_fn()
    cfg.env.num_envs, cfg.env.rollout_envs = m, r
    n = m * (1 + r)
    hf = synthetic.make_height_field(seed=0)
    ora = BatchRolloutOracle(cfg_cls(), spec_fn(), {k: v.clone() for k, v in inputs.items()}, hf, m, r)    # terrain bookkeeping only
    env = RobotBatchRollout(cfg, None, SyntheticSim(cfg, n, DEV, spec=spec, height_samples=hf, state={k: v.clone() for k, v in inputs.items()}), DEV, True)
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
@pytest.mark.parametrize("tag", list(TAGS))
def test_rollout_env_matches_reference_fixture(tag):
    m, r, seed, c0, inputs, outs, names = load(tag)
    env = make_env(tag, False, inputs, m, r)
    env.common_step_counter = c0
    for s, want in enumerate(outs):
        torch.manual_seed(5000 + 17 * s + seed)
        env.noise_u = want["noise_u"].to(DEV)
        env.torques = env._compute_torques(env.actions).view(env.torques.shape)
        env.post_physics_step()
        torch.cuda.synchronize()
        check(snapshot(env), want, f"{tag} step {s}", exact=False)
        sums = torch.stack([env.episode_sums[k] for k in names]).cpu()
        assert torch.allclose(sums, want["episode_sums"], rtol=1e-5, atol=1e-6), f"{tag} step {s}: episode sums differ"
        for k, w in want.items():
            if k.startswith("extras__") and k[8:].startswith("rew_"):
                got = float(env.extras["episode"][k[8:]])
                assert abs(got - float(w)) <= 1e-5 * abs(float(w)) + 1e-6, f"{tag} step {s}: {k} {got} vs {float(w)}"


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(TAGS))
def test_fused_rollout_reset_equals_host_path(tag):
    """in-kernel resample / reset with the main / rollout rules == the host-driven path, same uniforms, bit for bit"""
    from test_fused_reset import feed_host_path_from_table
    m, r, seed, c0, inputs, outs, names = load(tag)
    m, r = 24, 9
    cfg_cls, spec_fn, _ = common.CASES[TAGS[tag]]
    n = m * (1 + r)
    _, _, st = common.make_case_state(TAGS[tag], n, seed=7, adversarial=True)
    U = torch.rand(n, _lib.RESET_UNIFORMS, generator=torch.Generator().manual_seed(11)).to(DEV)
    a = make_env(tag, True, st, m, r)
    b = make_env(tag, False, st, m, r)
    for e in (a, b):
        e.cfg.domain_rand.push_robots = False
    a.reset_uniforms = U
    # the host path draws per listed row; rows of the main / rollout layout use the MAIN row's column for commands
    feed_host_path_from_table(b, U)
    g = torch.Generator().manual_seed(5)
    for step in range(3):
        u = torch.rand(n, a.num_obs, generator=g).to(DEV)
        for env in (a, b):
            env.noise_u = u
            env.torques = env._compute_torques(env.actions).view(env.torques.shape)
            env._obs_clip_for_step = 100.0
            env.post_physics_step()
        torch.cuda.synchronize()
        sa, sb = snapshot(a), snapshot(b)
        for k in sb:
            if sb[k] is None:
                continue
            assert torch.equal(sa[k].cpu(), sb[k].cpu()), f"{tag} step {step}: fused vs host path differ in {k}"
        assert torch.equal(a._episode_sums_all.cpu(), b._episode_sums_all.cpu()), f"{tag} step {step}: episode sums"
        for k, v in b.extras.get("episode", {}).items():
            if k.startswith("rew_") and bool(b.reset_buf[b.main_env_indices].any()):
                assert abs(float(a.extras["episode"][k]) - float(v)) <= 1e-5 * abs(float(v)) + 1e-6, k
    assert bool(b.reset_buf.any())
