"""Actuator-network torque path (Anymal._compute_torques, envs/anymal_c/anymal.py:93-105).

not gpu: the oracle restatement of LSTMsea against the golden vectors generated from the UNMODIFIED TorchScript module, and
(container only) against the module itself; ABI checks.  gpu: elg_actuator_net_torques through the Anymal class against
the golden vectors / the oracle, hidden-state clearing on both reset paths, the gait-scheduler reward inside the fused step."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402
from extended_legged_gym_b200 import _lib, synthetic  # noqa: E402
from oracle.actuator_oracle import ActuatorNetOracle, WEIGHT_KEYS  # noqa: E402
from oracle.legged_oracle import LeggedOracle  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "actuator_net.npz")
REF_PT = "/root/reference/legged_gym/resources/actuator_nets/anydrive_v3_lstm.pt"
DEV = "cuda:0"
RTOL, ATOL = 1e-5, 1e-6      # north star: fp32 outputs within 1e-5 relative / 1e-6 absolute
# The torque is out_scale (= 20) times the network's O(1) output, which carries the 1e-6 absolute bar: two float32
# evaluations of the same graph (the TorchScript module vs its eager restatement, both torch CPU) already differ by
# 3.6e-6 on torques that cancel to ~0.05 N m.  The absolute bar for torques is therefore out_scale x 1e-6.
ATOL_TORQUE = 20.0 * ATOL


def golden():
    z = np.load(GOLDEN)
    w = {k: torch.from_numpy(z[k]) for k in WEIGHT_KEYS + ("in_scale", "out_scale")}
    steps = [{k: torch.from_numpy(z[f"s{s}__{k}"]) for k in ("actions", "dof_pos", "dof_vel", "torques", "hidden", "cell")} for s in range(3)]
    return w, torch.from_numpy(z["default_dof_pos"]), float(z["action_scale"][0]), steps


def test_oracle_matches_golden_vectors_of_the_torchscript_module():
    w, q0, scale, steps = golden()
    ora = ActuatorNetOracle(w)
    h = torch.zeros(2, 64 * 12, 8)
    c = torch.zeros(2, 64 * 12, 8)
    for st in steps:
        t, h, c = ora.compute_torques(st["actions"], scale, q0, st["dof_pos"], st["dof_vel"], h, c)
        torch.testing.assert_close(t, st["torques"], rtol=RTOL, atol=ATOL_TORQUE)
        torch.testing.assert_close(h, st["hidden"], rtol=RTOL, atol=ATOL)
        torch.testing.assert_close(c, st["cell"], rtol=RTOL, atol=ATOL)


@pytest.mark.skipif(not os.path.exists(REF_PT), reason="the reference checkout is only present in the build container")
def test_oracle_matches_live_torchscript_module_and_fixture_weights():
    net = torch.jit.load(REF_PT, map_location="cpu")
    w, _, _, _ = golden()
    for k, v in net.state_dict().items():
        assert torch.equal(v, w[k].view(v.shape)), k
    ora = ActuatorNetOracle(w)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(500, 1, 2, generator=g) * torch.tensor([1.0, 5.0])
    h = torch.randn(2, 500, 8, generator=g) * 0.5
    c = torch.randn(2, 500, 8, generator=g)
    with torch.inference_mode():
        t_ref, (h_ref, c_ref) = net(x, (h.clone(), c.clone()))
    t, h2, c2 = ora.forward(x[:, 0], h, c)
    torch.testing.assert_close(t, t_ref, rtol=RTOL, atol=ATOL_TORQUE)
    torch.testing.assert_close(h2, h_ref, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(c2, c_ref, rtol=RTOL, atol=ATOL)


def _emulated_activations(rng):
    """float32 restatement of the kernel's gate activations (csrc/elg_actuator.cu: sigmoid_f / tanh_f) with every MUFU result moved
    by up to its documented error bound (ex2.approx 2 ulp, rcp.approx 1 ulp; the Newton step leaves <= 1 ulp) -- the error budget the
    branch-free forms were admitted on."""
    f32 = np.float32
    l2e = f32(1.4426950408889634)

    def nudge(x, ulps):
        return (x.view(np.int32) + rng.integers(-ulps, ulps + 1, size=x.shape).astype(np.int32)).view(np.float32)

    def ex2(a):
        return nudge(np.exp2(a.astype(np.float64)).astype(f32), 2)

    def rcp(y):
        return nudge((1.0 / y.astype(np.float64)).astype(f32), 1)

    def sigmoid(x):
        y = (f32(1) + ex2((x * -l2e).astype(f32))).astype(f32)
        return rcp(np.minimum(y, f32(1e38)))

    coef = [f32(c) for c in (-0.3333333134651184, 0.13333165645599365, -0.05391916632652283, 0.02136712521314621, -0.006715521216392517)]

    def tanh(x):
        x = x.astype(f32)
        ax, x2 = np.abs(x), (x * x).astype(f32)
        p = coef[4]
        for c in coef[3::-1]:
            p = (p.astype(np.float64) * x2 + c).astype(f32)          # fmaf
        small = ((x * x2).astype(f32).astype(np.float64) * p + x).astype(f32)
        t = ex2((np.minimum(ax, f32(10)) * f32(2.8853900817779268)).astype(f32))
        big = np.copysign((f32(1) - f32(2) * rcp((t + f32(1)).astype(f32))).astype(f32), x)
        return np.where(ax >= f32(0.55), big, small).astype(f32)

    return sigmoid, tanh


def test_kernel_activation_forms_stay_inside_the_tolerance_budget():
    """The kernel replaces libdevice expf / tanhf / __frcp_rn by MUFU-based branch-free forms.  Their worst-case error against float64
    and, through the whole LSTMsea over the three carried steps of the golden fixture, against the TorchScript module's outputs must
    leave at least half of the 1e-5 / 1e-6 bar unused."""
    rng = np.random.default_rng(0)
    sigmoid, tanh = _emulated_activations(rng)
    x = np.linspace(-30, 30, 600001).astype(np.float32)
    x64 = x.astype(np.float64)
    assert np.abs(tanh(x) - np.tanh(x64)).max() < 2.5e-7
    err_s = np.abs(sigmoid(x) - 1 / (1 + np.exp(-x64)))
    assert err_s.max() < 2.5e-7 and (err_s / (1 / (1 + np.exp(-x64)))).max() < 3e-6
    assert np.isnan(tanh(np.array([np.nan], np.float32))).all()
    z = np.load(GOLDEN)
    f32 = np.float32

    def cell(xx, h, c, wih, whh, bih, bhh):
        g = ((xx @ wih.T + bih) + (h @ whh.T + bhh)).astype(f32)
        i, f, gg, o = np.split(g, 4, axis=1)
        c2 = (sigmoid(f) * c + sigmoid(i) * tanh(gg)).astype(f32)
        return (sigmoid(o) * tanh(c2)).astype(f32), c2

    h = np.zeros((2, 768, 8), f32)
    c = np.zeros((2, 768, 8), f32)
    worst = 0.0
    for s in range(3):
        a, q, qd = z[f"s{s}__actions"], z[f"s{s}__dof_pos"], z[f"s{s}__dof_vel"]
        xin = np.stack([(a * z["action_scale"][0] + z["default_dof_pos"] - q).reshape(-1), qd.reshape(-1)], 1).astype(f32) * z["in_scale"][None]
        h0, c0 = cell(xin, h[0], c[0], z["lstm.weight_ih_l0"], z["lstm.weight_hh_l0"], z["lstm.bias_ih_l0"], z["lstm.bias_hh_l0"])
        h1, c1 = cell(h0, h[1], c[1], z["lstm.weight_ih_l1"], z["lstm.weight_hh_l1"], z["lstm.bias_ih_l1"], z["lstm.bias_hh_l1"])
        t = (z["out_scale"] * (h1 @ z["linear.weight"].T + z["linear.bias"])[:, 0]).astype(f32)
        h, c = np.stack([h0, h1]), np.stack([c0, c1])
        for got, ref, atol in ((t.reshape(64, 12), z[f"s{s}__torques"], ATOL_TORQUE), (h, z[f"s{s}__hidden"], ATOL), (c, z[f"s{s}__cell"], ATOL)):
            worst = max(worst, float((np.abs(got - ref) / (atol + RTOL * np.abs(ref))).max()))
    assert worst < 0.5, worst


def test_actuator_abi_layout_and_argument_checks():
    lib = _lib.load()
    assert lib.elg_actuator_net_words() == _lib.ACTNET_WORDS == 976
    d = _lib.ElgDims()
    d.num_envs, d.num_dof = 4, 12
    assert lib.elg_actuator_net_torques(None, 16, 0.5, 16, 16, 16, 16, 16, 16, None) == -4
    assert lib.elg_actuator_net_torques(C.byref(d), None, 0.5, 16, 16, 16, 16, 16, 16, None) == -4
    assert lib.elg_actuator_net_torques(C.byref(d), 16, 0.5, 16, 16, 16, 20, 16, 16, None) == -1     # misaligned hidden state
    assert b"16-byte" in lib.elg_last_error()


def test_anymal_refuses_to_run_without_the_network(monkeypatch):
    from extended_legged_gym_b200.envs.anymal_c.anymal import load_actuator_net_weights
    from extended_legged_gym_b200.envs import AnymalCRoughCfg
    monkeypatch.delenv("LEGGED_GYM_ROOT_DIR", raising=False)
    with pytest.raises(FileNotFoundError):
        load_actuator_net_weights(AnymalCRoughCfg().control)


# ---------------------------------------------------------------------------------------------------------------
def make_anymal(case, n, seed, use_net=True, gait_scale=0.0):
    from extended_legged_gym_b200.envs import Anymal
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    cfg, spec, st = common.make_case_state(case, n, seed=seed)
    cfg.env.num_envs = n
    cfg.control.use_actuator_network = use_net
    cfg.control.actuator_net_weights = GOLDEN
    if gait_scale:
        cfg.rewards.scales.gait_scheduler = gait_scale
    hf = synthetic.make_height_field(seed=0)
    env = Anymal(cfg, None, SyntheticSim(cfg, n, DEV, spec=spec, height_samples=hf, state={k: v.clone() for k, v in st.items()}), DEV, True)
    env.set_env_state(st)
    return env, cfg, spec, st, hf


@pytest.mark.gpu
def test_actuator_kernel_matches_golden_vectors():
    w, q0, scale, steps = golden()
    env, cfg, spec, st, hf = make_anymal("anymal_c_flat", 64, 0)
    assert abs(cfg.control.action_scale - scale) < 1e-9
    env.default_dof_pos.copy_(q0.to(DEV))
    for s in steps:
        env.dof_pos.copy_(s["dof_pos"].to(DEV))
        env.dof_vel.copy_(s["dof_vel"].to(DEV))
        t = env._compute_torques(s["actions"].to(DEV))
        torch.cuda.synchronize()
        torch.testing.assert_close(t.cpu(), s["torques"], rtol=RTOL, atol=ATOL_TORQUE)
        torch.testing.assert_close(env.sea_hidden_state.cpu(), s["hidden"], rtol=RTOL, atol=ATOL)
        torch.testing.assert_close(env.sea_cell_state.cpu(), s["cell"], rtol=RTOL, atol=ATOL)
    assert env.sea_hidden_state_per_env.shape == (2, 64, 12, 8) and env.sea_input.shape == (64 * 12, 1, 2)


@pytest.mark.gpu
def test_every_kernel_form_returns_the_same_bits():
    """elg_set_actuator_tuning: the row-per-thread forms (shared-memory weights staged before / after the grid-dependency wait,
    constant-bank weights), eight lanes per row and the unit-split forms evaluate the same expressions per hidden unit -- torques
    and states must agree bit for bit"""
    n = 1000
    env, cfg, spec, st, hf = make_anymal("anymal_c_rough", n, 5)
    g = torch.Generator().manual_seed(9)
    h = (torch.randn(2, n * 12, 8, generator=g) * 0.5).to(DEV)
    c = torch.randn(2, n * 12, 8, generator=g).to(DEV)
    a = (torch.randn(n, 12, generator=g) * 2).to(DEV)
    lib = _lib.load()
    outs = []
    try:
        for mode in (0, 1, 2, 3, 4, 5):
            lib.elg_set_actuator_tuning(mode)
            env.sea_hidden_state.copy_(h)
            env.sea_cell_state.copy_(c)
            t = env._compute_torques(a).clone()
            torch.cuda.synchronize()
            outs.append((t, env.sea_hidden_state.clone(), env.sea_cell_state.clone()))
    finally:
        lib.elg_set_actuator_tuning(0)
    for mode, o in enumerate(outs[1:], start=1):
        for x, y, what in zip(outs[0], o, ("torques", "hidden", "cell")):
            assert torch.equal(x, y), f"actuator kernel form {mode} differs from the default in {what}"


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 37, 4096])
def test_actuator_kernel_matches_oracle_ragged_and_full_size(n):
    w, _, _, _ = golden()
    ora = ActuatorNetOracle(w)
    env, cfg, spec, st, hf = make_anymal("anymal_c_rough", n, 3)
    g = torch.Generator().manual_seed(n)
    h = torch.randn(2, n * 12, 8, generator=g) * 0.5
    c = torch.randn(2, n * 12, 8, generator=g)
    a = torch.randn(n, 12, generator=g) * 2
    env.sea_hidden_state.copy_(h.to(DEV))
    env.sea_cell_state.copy_(c.to(DEV))
    t = env._compute_torques(a.to(DEV))
    torch.cuda.synchronize()
    tw, hw, cw = ora.compute_torques(a, cfg.control.action_scale, env.default_dof_pos.cpu(), env.dof_pos.cpu(), env.dof_vel.cpu(), h, c)
    torch.testing.assert_close(t.cpu(), tw, rtol=RTOL, atol=ATOL_TORQUE)
    torch.testing.assert_close(env.sea_hidden_state.cpu(), hw, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(env.sea_cell_state.cpu(), cw, rtol=RTOL, atol=ATOL)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
def test_anymal_step_runs_the_network_and_resets_clear_its_state(fused):
    n = 256
    env, cfg, spec, st, hf = make_anymal("anymal_c_rough", n, 5)
    env.fused_reset = fused
    cfg.domain_rand.push_robots = False
    env.episode_length_buf[:7] = 5000          # time-outs -> these envs reset this step
    obs, _, rew, reset, _ = env.step(torch.randn(n, 12, device=DEV))
    torch.cuda.synchronize()
    assert obs.shape == (n, env.num_obs) and torch.isfinite(obs).all() and torch.isfinite(rew).all()
    assert reset[:7].all()
    hs, cs = env.sea_hidden_state_per_env, env.sea_cell_state_per_env
    assert float(hs[:, reset].abs().max()) == 0.0 and float(cs[:, reset].abs().max()) == 0.0
    assert float(hs[:, ~reset].abs().max()) > 0.0      # decimation x network calls left their state in the other envs
    assert float(env.torques.abs().max()) > 0.0


@pytest.mark.gpu
def test_gait_scheduler_reward_inside_the_fused_step():
    """gait_scheduler.reward_foot_z_track (utils/gait_scheduler.py:74-81) on the previous step's feet and phases, then
    GaitScheduler.step (:63-72) -- evaluated in the step kernel -- against the oracle over three steps."""
    n = 512
    env, cfg, spec, st, hf = make_anymal("anymal_c_rough", n, 11, use_net=False, gait_scale=-1.0)
    ocfg = common.CASES["anymal_c_rough"][0]()
    ocfg.rewards.scales.gait_scheduler = -1.0
    ora = LeggedOracle(ocfg, spec, {k: v.clone() for k, v in st.items()}, hf, use_gait_scheduler=True)
    assert "gait_scheduler" in env.reward_scales and "gait_scheduler" in env._kernel_terms
    env.gait_prev_foot_z.copy_(ora.gait_foot_pos[:, :, 2].to(DEV))
    env.fused_reset = False
    g = torch.Generator().manual_seed(1)
    for s in range(3):
        u = torch.rand(n, env.num_obs, generator=g)
        ora.hot_step(noise_u=u)
        ora.gait_step()
        env.noise_u = u.to(DEV)
        env.torques = LeggedRobot_compute_torques(env)
        env._launch(_lib.PHASE_FUSED)
        env.reset_buf = env._reset_bool
        torch.cuda.synchronize()
        common.assert_state_close(common.snapshot(env), common.snapshot(ora), what=f"gait step {s}")
        torch.testing.assert_close(env.gait_idx.cpu(), ora.gait_idx, rtol=RTOL, atol=ATOL)
        torch.testing.assert_close(env.episode_sums["gait_scheduler"].cpu(), ora.episode_sums["gait_scheduler"], rtol=RTOL, atol=ATOL)


def LeggedRobot_compute_torques(env):
    from extended_legged_gym_b200.envs import LeggedRobot
    return LeggedRobot._compute_torques(env, env.actions).view(env.torques.shape)


@pytest.mark.gpu
def test_elspider_class_runs_the_network_on_18_dofs_and_resets_upside_down_envs():
    """ElSpider (envs/elspider_air/elspider.py:225-408): the same LSTMsea on [N * 18] rows, six feet, upside-down termination"""
    from extended_legged_gym_b200.envs import TASKS, ElSpider
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    assert TASKS["elspider_air_rough"][0] is ElSpider
    n = 130
    cfg, spec, st = common.make_case_state("elspider_air_rough", n, seed=6)
    cfg.env.num_envs = n
    cfg.control.use_actuator_network = True
    cfg.control.actuator_net_weights = GOLDEN
    cfg.domain_rand.push_robots = False
    hf = synthetic.make_height_field(seed=0)
    env = ElSpider(cfg, None, SyntheticSim(cfg, n, DEV, spec=spec, height_samples=hf, state={k: v.clone() for k, v in st.items()}), DEV, True)
    env.set_env_state(st)
    assert env.sea_hidden_state_per_env.shape == (2, n, 18, 8) and len(env.feet_indices) == 6 and env.num_obs == 253
    w, _, _, _ = golden()
    ora = ActuatorNetOracle(w)
    g = torch.Generator().manual_seed(2)
    a = torch.randn(n, 18, generator=g)
    h, c = env.sea_hidden_state.cpu().clone(), env.sea_cell_state.cpu().clone()
    t = env._compute_torques(a.to(DEV))
    torch.cuda.synchronize()
    tw, hw, cw = ora.compute_torques(a, cfg.control.action_scale, env.default_dof_pos.cpu(), env.dof_pos.cpu(), env.dof_vel.cpu(), h, c)
    torch.testing.assert_close(t.cpu(), tw, rtol=RTOL, atol=ATOL_TORQUE)
    torch.testing.assert_close(env.sea_hidden_state.cpu(), hw, rtol=RTOL, atol=ATOL)
    # a full step: envs whose projected gravity points up are reset although nothing touches the ground and no episode timed out
    env.contact_forces.zero_()
    env.episode_length_buf.zero_()
    env.root_states[:10, 3:7] = torch.tensor([1.0, 0.0, 0.0, 0.0], device=DEV)        # rolled by 180 degrees: on their backs
    q = env.root_states[:, 3:7].cpu().clone()
    obs, _, rew, reset, _ = env.step(torch.randn(n, 18, device=DEV))
    torch.cuda.synchronize()
    from oracle import torch_utils as tu
    up = tu.quat_rotate_inverse(q, torch.tensor([[0.0, 0.0, -1.0]]).repeat(n, 1))[:, 2] > 0
    assert torch.equal(reset.cpu(), up) and bool(up[:10].all()) and not bool(up.all())
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()


def test_gait_2_step_needs_its_feet():
    """the reward indexes feet 0..3 (0..5 in the hexapod form); fewer feet is an IndexError in the reference, an error code here"""
    import ctypes as C
    lib = _lib.load()
    d, p, b = _lib.ElgDims(), _lib.ElgStepParams(), _lib.ElgStepBuffers()
    d.num_envs, d.num_dof, d.num_bodies, d.num_feet, d.num_obs, d.num_commands = 8, 12, 13, 4, 48, 4
    p.reward_mask = 1 << _lib.TERM_ID["gait_2_step"]
    p.gait_2_step_hexapod = 1
    for f in _lib.ElgStepBuffers._fields_:
        setattr(b, f[0], 256)
    rc = lib.elg_post_physics_step(C.byref(d), C.byref(p), C.byref(b), _lib.PHASE_FUSED, None)
    assert rc == -1 and b"gait_2_step" in lib.elg_last_error()
