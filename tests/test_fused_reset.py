"""In-kernel reset path (reset_idx + _resample_commands + terrain curriculum as predicated
kernels, no nonzero() / host sync).  gpu only: with the SAME uniform numbers fed to both, the fused path must leave
every tensor bit-identical to the reference-structured host path (which the other parity tests pin to the oracle);
with in-kernel Philox the distributions and invariants of reset_idx must hold."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402
from extended_legged_gym_b200 import _lib, synthetic  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build(case, n, seed, fused):
    from extended_legged_gym_b200.envs import LeggedRobot
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    cfg, spec, st = common.make_case_state(case, n, seed=seed, adversarial=True)
    cfg.env.num_envs = n
    cfg.domain_rand.push_robots = False
    hf = synthetic.make_height_field(seed=0)
    torch.manual_seed(1234)                                   # terrain levels are drawn at construction
    env = LeggedRobot(cfg, None, SyntheticSim(cfg, n, DEV, spec=spec, height_samples=hf, state={k: v.clone() for k, v in st.items()}), DEV, True)
    env.set_env_state(st)
    env.fused_reset = fused
    return env, st


def feed_host_path_from_table(env, U):
    """Make the host reset path consume column c of the per-env uniform table exactly where the kernel does."""
    D = env.num_dof
    ctx = {"name": None, "ids": None, "k": 0}

    def wrap(name, fn):
        def inner(env_ids, *a, **k):
            prev = dict(ctx)
            ctx.update(name=name, ids=env_ids, k=0)
            try:
                return fn(env_ids, *a, **k)
            finally:
                ctx.update(prev)
        return inner
    env._resample_commands = wrap("cmd", env._resample_commands)
    env._reset_dofs = wrap("dofs", env._reset_dofs)
    env._reset_root_states = wrap("root", env._reset_root_states)
    env._update_terrain_curriculum = wrap("terrain", env._update_terrain_curriculum)

    def rand(lo, hi, shape):
        ids, k = ctx["ids"], ctx["k"]
        ctx["k"] += 1
        if ctx["name"] == "root":          # xy jitter only with custom origins, then the 6 root velocities
            cols = slice(D, D + 2) if shape[1] == 2 else slice(D + 2, D + 8)
        else:
            cols = {"dofs": [slice(0, D)], "cmd": [slice(D + 8, D + 9), slice(D + 9, D + 10), slice(D + 10, D + 11)]}[ctx["name"]][k]
        u = U[ids][:, cols]
        assert tuple(u.shape) == tuple(shape)
        return (hi - lo) * u + lo
    env._rand = rand
    env._randint_like = lambda t, high: torch.clamp((U[ctx["ids"], D + 11] * high).long(), max=high - 1)


@pytest.mark.parametrize("case", ["anymal_c_rough", "go2_all_terms_heading", "anymal_c_flat"])
def test_fused_reset_equals_host_path_with_same_uniforms(case):
    n = 3000
    U = torch.rand(n, _lib.RESET_UNIFORMS, generator=torch.Generator().manual_seed(9)).to(DEV)
    a, st = build(case, n, 3, fused=True)
    b, _ = build(case, n, 3, fused=False)
    a.reset_uniforms = U
    feed_host_path_from_table(b, U)
    if getattr(a, "custom_origins", False):
        for name in ("terrain_levels", "terrain_types", "env_origins"):
            getattr(b, name).copy_(getattr(a, name))
    g = torch.Generator().manual_seed(5)
    for step in range(3):
        u = torch.rand(n, a.num_obs, generator=g).to(DEV)
        for env in (a, b):
            env.noise_u = u
            env._obs_clip_for_step = 100.0
            env.torques = env._compute_torques(env.actions).view(env.torques.shape)
            env.post_physics_step()
        torch.cuda.synchronize()
        assert int(a.reset_buf.sum()) > 0
        sa, sb = common.snapshot(a), common.snapshot(b)
        for k in sb:
            assert torch.equal(sa[k], sb[k]), f"{case} step {step}: {k} differs ({int((sa[k] != sb[k]).sum())} entries)"
        if getattr(a, "custom_origins", False):
            assert torch.equal(a.terrain_levels, b.terrain_levels) and torch.equal(a.env_origins, b.env_origins)
        for k, v in b.extras["episode"].items():
            assert abs(float(a.extras["episode"][k]) - float(v)) <= 1e-5 * abs(float(v)) + 1e-6, k
        assert torch.equal(a.extras["time_outs"], b.extras["time_outs"])


def test_fused_reset_philox_invariants_and_no_resets_keep_extras():
    n = 4096
    env, st = build("anymal_c_rough", n, 1, fused=True)
    twin, _ = build("anymal_c_rough", n, 1, fused=False)
    for e in (env, twin):
        e.torques = e._compute_torques(e.actions).view(e.torques.shape)
        e.post_physics_step()
    torch.cuda.synchronize()
    r = env.reset_buf
    assert torch.equal(r, twin.reset_buf) and 5 < int(r.sum()) < n // 2
    interval = int(env.cfg.commands.resampling_time / env.dt)
    resampled = ((st["episode_length_buf"].to(DEV) + 1) % interval) == 0      # these drew new commands from a different RNG
    assert int(resampled.sum()) > 0
    keep = ~r & ~resampled
    for k in ("root_states", "dof_state", "commands", "obs_buf", "last_dof_vel", "last_root_vel", "feet_air_time", "episode_length_buf"):
        assert torch.equal(getattr(env, k).view(n, -1)[keep], getattr(twin, k).view(n, -1)[keep]), k
    nzd = (env.default_dof_pos != 0).flatten()
    q = env.dof_pos[r][:, nzd] / env.default_dof_pos[:, nzd]
    assert float(q.min()) >= 0.5 - 1e-6 and float(q.max()) <= 1.5 + 1e-6 and 0.2 < float(q.std()) < 0.35
    assert bool((env.dof_pos[r][:, ~nzd] == 0).all())
    assert bool((env.dof_vel[r] == 0).all()) and bool((env.episode_length_buf[r] == 0).all())
    d = env.root_states[r, :2] - env.env_origins[r, :2]
    assert float(d.abs().max()) <= 0.5 + 1e-5
    assert float(env.root_states[r, 7:13].abs().max()) <= 0.5 and torch.equal(env.last_root_vel[r], env.root_states[r, 7:13])
    assert torch.equal(env.root_states[r, 3:7], env.base_init_state[3:7].expand(int(r.sum()), -1))
    for name, row in env.episode_sums.items():
        assert bool((row[r] == 0).all()), name
    small = torch.norm(env.commands[r, :2], dim=1)
    assert bool(((small == 0) | (small > 0.2)).all())
    ep = {k: float(v) for k, v in env.extras["episode"].items()}
    assert all(v == v for v in ep.values())                    # no NaNs
    # a step without resets leaves extras["episode"] as it was
    env.episode_length_buf.fill_(3)
    env.contact_forces.zero_()
    env.post_physics_step()
    torch.cuda.synchronize()
    assert int(env.reset_buf.sum()) == 0
    assert {k: float(v) for k, v in env.extras["episode"].items()} == ep
