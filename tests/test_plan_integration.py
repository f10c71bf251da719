"""Kinematic state integration of the planning variant (SURVEY section 8f-4;
RobotPlanGradSampling._integrate_state_velocities / _sync_integration_to_sim, robot_plan_grad_sampling.py:103-225).

not gpu: the oracle against the golden vectors generated from the UNMODIFIED reference methods and (container only) against
those methods themselves; ABI checks.  gpu: elg_integrate_state_velocities against the golden vectors and against the oracle at
the rollout size of BASELINE config 5, the write-through-only mode, and the RobotPlanGradSampling class."""
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
from oracle import plan_oracle as po, ref_harness  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_plan_golden import CASES  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "plan_integration.npz")
DEV = "cuda:0"
RTOL, ATOL = 1e-5, 1e-6
# base-frame velocities are a rotation of a vector of norm <= sqrt(3) * max_base_*_vel by a quaternion that went through up to 7
# multiply / renormalise sub-steps: a component near zero carries the absolute error of the whole vector (cancellation), so
# the absolute tolerance of these two fields is 1e-6 x 2 x the velocity limit; everything else keeps 1e-6.
ATOL_KEY = {"base_lin_vel": 2 * 3.0 * 1e-6, "base_ang_vel": 2 * 2.0 * 1e-6}


def load_case(tag):
    z, c = np.load(GOLDEN), CASES[tag]
    o = po.make_state(c["n"], c["d"], c["seed"], c["method"], c["enforce"], c["max_step"])
    for k in po.KEYS:
        torch.testing.assert_close(getattr(o, k), torch.from_numpy(z[f"{tag}__in__{k}"]), rtol=0, atol=0)     # the seeded state is the fixture's
    want = {k: torch.from_numpy(z[f"{tag}__out__{k}"]) for k in po.KEYS}
    return c, o, torch.from_numpy(z[f"{tag}__state_vels"]), torch.from_numpy(z[f"{tag}__env_ids"]), want


def assert_state(got, want):
    for k in po.KEYS:
        torch.testing.assert_close(got[k], want[k], rtol=RTOL, atol=ATOL_KEY.get(k, ATOL), msg=lambda m, k=k: f"{k}: {m}")


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_matches_reference_fixture(tag):
    c, o, sv, idx, want = load_case(tag)
    po.integrate_state_velocities(o, sv, c["dt"], idx)
    po.sync_integration_to_sim(o, idx)
    assert_state(po.snapshot(o), want)


@pytest.mark.skipif(not ref_harness.available(), reason="the reference checkout is only present in the build container")
@pytest.mark.parametrize("method,enforce", [("euler", False), ("euler", True), ("rk4", True)])
def test_oracle_matches_live_reference_methods(method, enforce):
    ref_harness.install()
    from legged_gym.envs.batch_rollout.robot_plan_grad_sampling import RobotPlanGradSampling as Ref
    a, b = po.make_state(40, 12, 7, method, enforce, 0.004), po.make_state(40, 12, 7, method, enforce, 0.004)
    g = torch.Generator().manual_seed(3)
    idx = torch.randperm(40, generator=g)[:25]
    sv = torch.randn(25, 18, generator=g) * 4
    Ref._integrate_state_velocities(a, sv, 0.02, idx)
    Ref._sync_integration_to_sim(a, idx)
    po.integrate_state_velocities(b, sv, 0.02, idx)
    po.sync_integration_to_sim(b, idx)
    assert_state(po.snapshot(b), po.snapshot(a))


def test_plan_abi_argument_checks():
    lib = _lib.load()
    assert lib.elg_sizeof_plan_params() == C.sizeof(_lib.ElgPlanParams)
    assert lib.elg_sizeof_plan_buffers() == C.sizeof(_lib.ElgPlanBuffers)
    ok = _lib.ElgPlanParams(12, 0, 2, 0, 0.01, 3.0, 2.0, 10.0)
    buf = _lib.ElgPlanBuffers(*([16] * 11))
    assert lib.elg_integrate_state_velocities(None, C.byref(buf), 16, None, 4, None) == -4
    assert lib.elg_integrate_state_velocities(C.byref(ok), C.byref(_lib.ElgPlanBuffers()), 16, None, 4, None) == -4
    for bad in (_lib.ElgPlanParams(0, 0, 2, 0, 0.01, 3.0, 2.0, 10.0), _lib.ElgPlanParams(12, 2, 2, 0, 0.01, 3.0, 2.0, 10.0),
                _lib.ElgPlanParams(12, 0, 0, 0, 0.01, 3.0, 2.0, 10.0)):
        assert lib.elg_integrate_state_velocities(C.byref(bad), C.byref(buf), 16, None, 4, None) == -1
    assert lib.elg_integrate_state_velocities(C.byref(ok), C.byref(buf), 16, None, -1, None) == -1
    assert lib.elg_integrate_state_velocities(C.byref(ok), C.byref(buf), 16, None, 0, None) == 0            # nothing to do


# ---------------------------------------------------------------------------------------------------------------
def to_device(o):
    """the oracle's state object with every tensor on the GPU, plus the host mixin's methods"""
    from extended_legged_gym_b200.envs import KinematicStateIntegration

    class Dev(KinematicStateIntegration):
        pass
    d = Dev()
    for k, v in vars(o).items():
        setattr(d, k, v.to(DEV).contiguous() if isinstance(v, torch.Tensor) else v)
    d.device = DEV
    d.state_vel_dim = 6 + o.num_dof
    d.dof_pos = d.dof_state.view(o.total_num_envs, o.num_dof, 2)[..., 0]
    d.dof_vel = d.dof_state.view(o.total_num_envs, o.num_dof, 2)[..., 1]
    return d


def snapshot_cpu(d):
    torch.cuda.synchronize()
    return {k: getattr(d, k).cpu() for k in po.KEYS}


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_kernel_matches_reference_fixture(tag):
    c, o, sv, idx, want = load_case(tag)
    d = to_device(o)
    d._integrate_state_velocities(sv.to(DEV), c["dt"], idx.to(DEV))
    assert_state(snapshot_cpu(d), want)


@pytest.mark.gpu
@pytest.mark.parametrize("method,enforce,max_step", [("euler", False, 0.01), ("euler", True, 0.003), ("rk4", True, 0.02)])
def test_kernel_matches_oracle_rollout_rows_full_size(method, enforce, max_step):
    """BASELINE config 5 layout: 64 mains x 512 rollouts, the rollout rows integrated, the main rows untouched (bit-exact)"""
    m, r, dnum = 64, 512, 12
    n = m * (1 + r)
    o = po.make_state(n, dnum, 11, method, enforce, max_step)
    d = to_device(o)
    idx = torch.arange(n).view(m, 1 + r)[:, 1:].reshape(-1)
    g = torch.Generator().manual_seed(5)
    sv = torch.randn(len(idx), 6 + dnum, generator=g) * torch.tensor([2.5] * 3 + [1.5] * 3 + [6.0] * dnum)
    sv[::50, 3:6] = 0.0
    before = po.snapshot(o)
    d._integrate_state_velocities(sv.to(DEV), 0.02, idx.to(DEV))
    po.integrate_state_velocities(o, sv, 0.02, idx)
    po.sync_integration_to_sim(o, idx)
    got = snapshot_cpu(d)
    assert_state(got, po.snapshot(o))
    mains = torch.arange(0, n, 1 + r)
    for k in po.KEYS:
        assert torch.equal(got[k][mains], before[k][mains]), f"{k}: a main-env row changed"
    qn = got["integration_base_quat"].norm(dim=1)
    assert float((qn - 1).abs().max()) < 1e-6
    if enforce:
        lo, hi = (float(torch.tensor(v, dtype=torch.float32)) for v in (-0.6, 0.8))         # the limits as the fp32 the tensors hold
        assert float(got["integration_dof_pos"][idx].min()) >= lo and float(got["integration_dof_pos"][idx].max()) <= hi


@pytest.mark.gpu
def test_write_through_only_mode_and_all_rows():
    o = po.make_state(37, 18, 2)
    o.integration_base_lin_vel.normal_(generator=torch.Generator().manual_seed(1))
    o.integration_dof_vel.normal_(generator=torch.Generator().manual_seed(2))
    d = to_device(o)
    d._sync_integration_to_sim()                      # env_indices None: every row
    po.sync_integration_to_sim(o, torch.arange(37))
    got = snapshot_cpu(d)
    for k in po.KEYS:
        if k.startswith("integration_"):
            assert torch.equal(got[k], getattr(o, k)), f"{k} changed in write-through-only mode"
    assert_state(got, po.snapshot(o))
    d._sync_sim_to_integration(torch.tensor([3, 5], device=DEV))
    assert torch.equal(d.integration_dof_pos[3], d.dof_pos[3])


@pytest.mark.gpu
def test_plan_class_step_rollout_integrates_then_scores():
    from extended_legged_gym_b200.envs import RobotPlanGradSampling, RobotPlanGradSamplingCfg
    from extended_legged_gym_b200.envs.anymal_c.anymal_c_config import AnymalCRoughCfg
    from extended_legged_gym_b200.sim_backend import SyntheticSim

    class Cfg(AnymalCRoughCfg, RobotPlanGradSamplingCfg):
        class env(AnymalCRoughCfg.env):
            num_envs = 4
            rollout_envs = 3

        class domain_rand(AnymalCRoughCfg.domain_rand):
            rollout_envs_sync_pos_drift = 0.0
            push_robots = False

    cfg = Cfg()
    n = 4 * 4
    _, spec, st = common.make_case_state("anymal_c_rough", n, seed=8)
    hf = synthetic.make_height_field(seed=0)
    env = RobotPlanGradSampling(cfg, None, SyntheticSim(cfg, n, DEV, spec=spec, height_samples=hf, state=st), DEV, True)
    env.set_env_state(st)
    env._sync_sim_to_integration()
    env.add_noise = False
    torch.cuda.synchronize()
    g = torch.Generator().manual_seed(1)
    main_out = env.step(torch.randn(4, 18, generator=g).to(DEV))          # mains integrate, rollouts become copies, cache filled
    torch.cuda.synchronize()
    assert main_out[0].shape[0] == 4
    R1 = 4
    for k in ("integration_base_pos", "integration_dof_pos", "integration_base_quat"):
        t = getattr(env, k).view(4, R1, -1)
        assert torch.equal(t[:, 1:], t[:, :1].expand_as(t[:, 1:])), f"{k}: rollout rows are not copies of their main row"
    root0 = env.root_states.clone()
    sv = torch.randn(12, 18, generator=g)
    obs, _, rew, reset, _ = env.step_rollout(sv.to(DEV))
    torch.cuda.synchronize()
    assert obs.shape[0] == 12 and rew.shape == (12,)
    roll = env.rollout_env_indices.cpu()
    # the rollout rows moved by v * dt, the main rows are back where they were
    lin = sv[:, :3].clamp(-3, 3)
    torch.testing.assert_close(env.root_states.cpu()[roll, :3], root0.cpu()[roll, :3] + lin * env.dt, rtol=1e-5, atol=1e-5)
    assert torch.equal(env.root_states[env.main_env_indices], root0[env.main_env_indices])
    # the observation's base-frame velocity is the integrated one (obs_scales.lin_vel = 2)
    q = env.root_states.cpu()[roll, 3:7]
    from oracle import torch_utils as tu
    torch.testing.assert_close(obs.cpu()[:, 0:3], tu.quat_rotate_inverse(q, lin) * 2.0, rtol=1e-5, atol=1e-5)
    with pytest.raises(ValueError):
        env._integrate_state_velocities(torch.zeros(3, 5, device=DEV), env.dt, env.rollout_env_indices)
