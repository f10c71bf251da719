"""GPU parity of the fused post-physics step (through the Python host class and the C ABI) against
the oracle on identical seeded inputs, and against the fixtures generated from the reference.
Bit-exact: reset / time-out / contact masks / episode counters / terrain cell indices.
fp32: rewards, observations, torques, heights, derived state within 1e-5 rel / 1e-6 abs."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402
from test_oracle_pinned import load_golden, step_seed  # noqa: E402
from oracle.legged_oracle import LeggedOracle  # noqa: E402
from extended_legged_gym_b200 import _lib, synthetic  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_env(cfg, spec, st, hf, oracle=None):
    from extended_legged_gym_b200.envs import ElSpider, LeggedRobot
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    n = st["root_states"].shape[0]
    cfg.env.num_envs = n
    sim = SyntheticSim(cfg, n, DEV, spec=spec, height_samples=hf, state={k: v.clone() for k, v in st.items()})
    cls = ElSpider if cfg.asset.name == "elspider_air" else LeggedRobot       # the class the task registry binds the config to
    env = cls(cfg, None, sim, DEV, True)
    env.set_env_state(st)
    env.fused_reset = False      # the parity tests pin the reference's host-driven reset path and its torch RNG stream
    # host RNG hooks draw from the CPU generator so the sparse paths consume the oracle's numbers
    env._rand = lambda lo, hi, shape: ((hi - lo) * torch.rand(*shape) + lo).to(DEV)
    env._randint_like = lambda t, high: torch.randint_like(t.cpu(), high).to(DEV)
    if oracle is not None and getattr(oracle, "custom_origins", False):
        env.terrain_levels = oracle.terrain_levels.clone().to(DEV)
        env.terrain_types = oracle.terrain_types.clone().to(DEV)
        env.terrain_origins = oracle.terrain_origins.clone().to(DEV)
        env.env_origins = oracle.env_origins.clone().to(DEV)
    return env


def run_pair(case, n, seed, steps, adversarial=True):
    cfg, spec, st = common.make_case_state(case, n, seed=seed, adversarial=adversarial)
    hf = synthetic.make_height_field(seed=0)
    ora = LeggedOracle(common.CASES[case][0](), spec, {k: v.clone() for k, v in st.items()}, hf)
    env = make_env(cfg, spec, st, hf, ora)
    g = torch.Generator().manual_seed(4242 + seed)
    for s in range(steps):
        u = torch.rand(n, env.num_obs, generator=g)
        torch.manual_seed(step_seed(s, seed))
        ora.torques = ora.compute_torques(ora.actions).view(ora.torques.shape)
        ora.post_physics_step(noise_u=u)
        torch.manual_seed(step_seed(s, seed))
        env.noise_u = u.to(DEV)
        env.torques = env._compute_torques(env.actions).view(env.torques.shape)
        env.post_physics_step()
        torch.cuda.synchronize()
        common.assert_state_close(common.snapshot(env), common.snapshot(ora), what=f"{case} N={n} seed={seed} step {s}")
        for k, v in ora.extras.get("episode", {}).items():
            got = float(env.extras["episode"][k])
            assert abs(got - float(v)) <= 1e-5 * abs(float(v)) + 1e-6, f"extras {k}: {got} vs {float(v)}"
    return env, ora


@pytest.mark.parametrize("case", list(common.CASES))
def test_step_matches_oracle(case):
    env, ora = run_pair(case, 2048, seed=1, steps=3)
    assert bool(ora.reset_buf.any())


@pytest.mark.parametrize("case", list(common.CASES))
def test_step_matches_reference_fixture(case):
    inputs, outs = load_golden(case)
    cfg_cls, spec_fn, _ = common.CASES[case]
    hf = synthetic.make_height_field(seed=0)
    ora = LeggedOracle(cfg_cls(), spec_fn(), {k: v.clone() for k, v in inputs.items()}, hf)   # only for terrain bookkeeping
    env = make_env(cfg_cls(), spec_fn(), inputs, hf, ora)
    for s, want in enumerate(outs):
        torch.manual_seed(step_seed(s))
        env.noise_u = want["noise_u"].to(DEV)
        env.torques = env._compute_torques(env.actions).view(env.torques.shape)
        env.post_physics_step()
        torch.cuda.synchronize()
        ref = {k: v for k, v in want.items() if not k.startswith("extras__") and k != "noise_u"}
        common.assert_state_close(common.snapshot(env), ref, what=f"{case} fixture step {s}")


@pytest.mark.parametrize("n", [1, 7, 8, 9, 33, 4096, 12289, 40000])
def test_ragged_and_large_env_counts(n):
    """every EPB specialisation (8/16/32 envs per CTA) and ragged tails"""
    run_pair("go2_all_terms_heading", n, seed=2, steps=1)


@pytest.mark.parametrize("case", ["a1_rough", "go2_rough"])
def test_step_parity_at_config3_size(case):
    """BASELINE configs[2]: 65 536 envs -- the persistent, double-buffered many-chunk form of the lean kernel (2 341 chunks
    over 148 CTAs), two steps so that the second one reads what the first one wrote"""
    run_pair(case, 65536, seed=3, steps=2)


def test_terrain_cells_bit_exact():
    n = 4096
    cfg, spec, st = common.make_case_state("anymal_c_rough", n, seed=5)
    # a few robots outside the map and on its border exercise the clip
    st["root_states"][:8, 0] = torch.tensor([-30.0, -25.0, -24.95, 64.9, 65.0, 70.0, 1e6, -1e6])
    st["root_states"][8:12, 1] = torch.tensor([-26.0, 64.95, 65.05, 3e9])
    hf = synthetic.make_height_field(seed=0)
    ora = LeggedOracle(cfg, spec, {k: v.clone() for k, v in st.items()}, hf)
    want = ora.get_heights()
    env = make_env(cfg, spec, st, hf, ora)
    env._sync_native()
    cells = torch.empty(n, env.num_height_points, 2, dtype=torch.int32, device=DEV)
    out = torch.empty(n, env.num_height_points, device=DEV)
    rc = env._lib.elg_get_heights(C.byref(env._dims), C.byref(env._params), env.root_states.data_ptr(), env.height_samples.data_ptr(),
                                  env._height_grid.data_ptr(), out.data_ptr(), cells.data_ptr(), None)
    _lib.check(rc)
    torch.cuda.synchronize()
    px, py = ora.height_cells
    assert torch.equal(cells[..., 0].cpu().long(), px), "terrain row index differs"
    assert torch.equal(cells[..., 1].cpu().long(), py), "terrain col index differs"
    assert torch.equal(out.cpu(), want), "heights differ"
    assert torch.equal(env._get_heights().cpu(), want)


@pytest.mark.parametrize("n", [4096, 4099])
def test_fused_step_heights_bit_exact(n):
    """The height scan inside the fused step (lean kernel for n % 4 == 0, generic otherwise) lands in exactly the reference's
    terrain cells: measured_heights equal the oracle's bit for bit, robots outside the map and on its border included
    (every rounding of the cell chain is kept -- see madd2_unfused in csrc/elg_async.cuh)."""
    cfg, spec, st = common.make_case_state("anymal_c_rough", n, seed=6)
    st["root_states"][:8, 0] = torch.tensor([-30.0, -25.0, -24.95, 64.9, 65.0, 70.0, 1e6, -1e6])
    st["root_states"][8:12, 1] = torch.tensor([-26.0, 64.95, 65.05, 3e9])
    # robots sitting exactly on cell boundaries of the 0.1 m grid
    st["root_states"][12:512, 0] = (torch.arange(500) * 0.1 + 1.0)
    st["root_states"][12:512, 1] = (torch.arange(500) * 0.05 + 3.0)
    hf = synthetic.make_height_field(seed=0)
    ora = LeggedOracle(cfg, spec, {k: v.clone() for k, v in st.items()}, hf)
    want = ora.get_heights()
    env = make_env(cfg, spec, st, hf, ora)
    env.noise_u = torch.rand(n, env.num_obs, generator=torch.Generator().manual_seed(2)).to(DEV)
    env.torques = env._compute_torques(env.actions).view(env.torques.shape)
    env._launch(_lib.PHASE_FUSED)
    torch.cuda.synchronize()
    assert torch.equal(env.measured_heights.cpu(), want), "fused-step heights differ from the oracle's cells"


def test_fused_equals_split_sections():
    """ONE fused launch == derive | termination | reward | obs | history launched one by one."""
    case, n = "go2_all_terms_heading", 1000
    outs = []
    for split in (False, True):
        cfg, spec, st = common.make_case_state(case, n, seed=9)
        hf = synthetic.make_height_field(seed=0)
        env = make_env(cfg, spec, st, hf)
        env.noise_u = torch.rand(n, env.num_obs, generator=torch.Generator().manual_seed(1)).to(DEV)
        env.torques = env._compute_torques(env.actions).view(env.torques.shape)
        if split:
            for ph in (_lib.PHASE_DERIVE, _lib.PHASE_TERMINATION, _lib.PHASE_REWARD, _lib.PHASE_OBS, _lib.PHASE_HISTORY):
                env._launch(ph)
        else:
            env._launch(_lib.PHASE_FUSED)
        env.reset_buf = env._reset_bool
        torch.cuda.synchronize()
        outs.append(common.snapshot(env))
    # the fused launch takes the lean kernel (elg_step_fast.cu), the sections the generic one: per-env sums over DOFs /
    # feet associate differently, so float outputs agree to the fp32 tolerance of the north star; masks, counters and
    # everything not behind such a sum stay bit-identical
    for k in outs[0]:
        a, b = outs[0][k], outs[1][k]
        if a.dtype.is_floating_point and (k.startswith("sum_") or k in ("rew_buf",)):
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6, msg=k)
        else:
            assert torch.equal(a, b), k


@pytest.mark.parametrize("case,n", [("anymal_c_rough", 1000), ("go2_all_terms_heading", 1028), ("a1_all_terms", 1001), ("anymal_c_flat", 512),
                                    ("anymal_c_rough", 8192)])   # 8192 envs: the persistent (multi-chunk) form
def test_rollout_mode_step_matches_oracle(case, n):
    """post_physics_step_rollout (robot_batch_rollout.py:763-817): one launch with ELG_PHASE_DERIVE | REWARD | OBS | HISTORY and
    rollout_mode -- measured heights, reset / time-out flags, commands and the episode counter are inputs and stay untouched,
    episode sums do not accumulate.  n % 4 == 0 takes the lean kernel, 1001 the generic one."""
    cfg, spec, st = common.make_case_state(case, n, seed=21)
    hf = synthetic.make_height_field(seed=0)
    ora = LeggedOracle(common.CASES[case][0](), spec, {k: v.clone() for k, v in st.items()}, hf)
    env = make_env(cfg, spec, st, hf, ora)
    g = torch.Generator().manual_seed(3)
    flags = torch.rand(n, generator=g) < 0.1
    touts = flags & (torch.rand(n, generator=g) < 0.5)
    ora.reset_buf, ora.time_out_buf = flags.clone(), touts.clone()
    env._reset_bool.copy_(flags.to(DEV))
    env.time_out_buf.copy_(touts.to(DEV))
    env.reset_buf = env._reset_bool
    if ora.measure_heights:
        mh = torch.rand(n, env.num_height_points, generator=g) - 0.5
        ora.measured_heights = mh.clone()
        env.measured_heights.copy_(mh.to(DEV))
    sums_before = {k: v.clone() for k, v in env.episode_sums.items()}
    u = torch.rand(n, env.num_obs, generator=g)
    ora.torques = ora.compute_torques(ora.actions).view(ora.torques.shape)
    ora.post_physics_step_rollout(noise_u=u)
    env.noise_u = u.to(DEV)
    env.torques = env._compute_torques(env.actions).view(env.torques.shape)
    env._launch(_lib.PHASE_DERIVE | _lib.PHASE_REWARD | _lib.PHASE_OBS | _lib.PHASE_HISTORY, rollout=True)
    torch.cuda.synchronize()
    common.assert_state_close(common.snapshot(env), common.snapshot(ora), what=f"rollout step {case}")
    for k, v in sums_before.items():
        assert torch.equal(env.episode_sums[k], v), f"episode sum {k} moved in rollout mode"


def test_user_defined_reward_term_and_override():
    """A subclass adds a Python term and overrides a stock one: both are honoured (registry API)."""
    from extended_legged_gym_b200.envs import LeggedRobot
    case, n = "a1_rough", 512

    class MyRobot(LeggedRobot):
        def _reward_lin_vel_z(self):                  # override of a stock term
            return 2.0 * torch.square(self.base_lin_vel[:, 2])

        def _reward_alive(self):                      # brand-new term
            return torch.ones(self.num_envs, device=self.device)

    cfg, spec, st = common.make_case_state(case, n, seed=4)
    cfg.rewards.scales.alive = 0.3
    hf = synthetic.make_height_field(seed=0)
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    cfg.env.num_envs = n
    env = MyRobot(cfg, None, SyntheticSim(cfg, n, DEV, spec=spec, height_samples=hf, state=st), DEV, True)
    env.set_env_state(st)
    assert set(env._python_terms) == {"alive", "lin_vel_z"}
    env.cfg.noise.add_noise = env.add_noise = False
    env.cfg.domain_rand.push_robots = False
    ocfg = common.CASES[case][0]()
    ocfg.noise.add_noise = False
    ora = LeggedOracle(ocfg, spec, {k: v.clone() for k, v in st.items()}, hf)
    ora.torques = ora.compute_torques(ora.actions).view(ora.torques.shape)
    ora.derive(); ora.measured_heights = ora.get_heights(); ora.check_termination(); ora.compute_reward()
    stock = ora.rew_buf.clone()
    env.torques = env._compute_torques(env.actions).view(env.torques.shape)
    env._launch(_lib.PHASE_DERIVE | _lib.PHASE_TERMINATION)
    env.reset_buf = env._reset_bool
    env.compute_reward()
    torch.cuda.synchronize()
    # expected: stock sum - stock lin_vel_z + doubled lin_vel_z + alive, clipped at 0
    lz = torch.square(ora.base_lin_vel[:, 2]) * ora.reward_scales["lin_vel_z"]
    terms = {k: v.clone() for k, v in ora.episode_sums.items()}
    unclipped = sum(terms.values()) + lz + 0.3 * ora.dt
    want = torch.clip(unclipped, min=0.0)
    assert torch.allclose(env.rew_buf.cpu(), want, rtol=1e-5, atol=2e-6)


def _philox4x32_10(c0, c1, c2, c3, k0, k1):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = [np.asarray(x, dtype=np.uint64) for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = np.uint64(k0), np.uint64(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & mask, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & mask]
        k0, k1 = (k0 + np.uint64(W0)) & mask, (k1 + np.uint64(W1)) & mask
    return c


def test_in_kernel_philox_noise_is_exact_and_uniform():
    """ELG_NOISE_PHILOX (csrc/elg_common.cuh): lane l of env's warp draws Philox4x32-10(counter=(env, l | block<<5,
    offset, 0), key=seed) and cuts it into eight 16-bit samples.  Height point p = l + 32 j: sample j%8 of block
    1 + j//8; head entry k = l + 32 m: sample nj%8 + m of block 1 + nj//8 (fits for 12 DOF / 187 points).
    obs = clean + (2 s/65536 - 1) * scale."""
    case, n = "anymal_c_rough", 3000
    cfg, spec, st = common.make_case_state(case, n, seed=6)
    hf = synthetic.make_height_field(seed=0)
    env = make_env(cfg, spec, st, hf)
    env.torques = env._compute_torques(env.actions).view(env.torques.shape)
    env.add_noise = False
    env._launch(_lib.PHASE_DERIVE | _lib.PHASE_OBS)
    clean = env.obs_buf.cpu().clone()
    env.add_noise, env.noise_u, env._noise_step = True, None, 12345
    env._launch(_lib.PHASE_OBS)
    torch.cuda.synchronize()
    noisy = env.obs_buf.cpu()
    O, H = env.num_obs, env.num_height_points
    head = 12 + 3 * env.num_dof
    nj, hm = (H + 31) // 32, (head + 31) // 32
    assert nj % 8 + hm <= 8
    e, k = np.meshgrid(np.arange(n), np.arange(O), indexing="ij")
    is_head = k < head
    p = k - head
    lane = np.where(is_head, k % 32, p % 32)
    block = np.where(is_head, 1 + nj // 8, 1 + (p // 32) // 8)
    samp = np.where(is_head, nj % 8 + k // 32, (p // 32) % 8)
    words = _philox4x32_10(e, lane | (block << 5), 12345, 0, env.noise_seed & 0xFFFFFFFF, env.noise_seed >> 32)
    w = np.choose(samp >> 1, [x.astype(np.uint64) for x in words])
    s16 = ((w >> (np.uint64(16) * (samp & 1).astype(np.uint64))) & np.uint64(0xFFFF)).astype(np.float32)
    want = clean + torch.from_numpy(s16 / np.float32(32768.0) - np.float32(1.0)) * env.noise_scale_vec.cpu()
    assert torch.allclose(noisy, want, rtol=1e-6, atol=1e-6)
    nz = env.noise_scale_vec.cpu() > 0
    uu = torch.from_numpy(s16 / np.float32(65536.0))[:, nz]
    assert abs(float(uu.mean()) - 0.5) < 2e-3 and abs(float(uu.var()) - 1 / 12) < 2e-3
