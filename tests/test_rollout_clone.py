"""Main -> rollout state clone.
not-gpu: the oracle restatement against the fixture generated from the unmodified reference, and (container only)
against the live reference; ABI argument checks.  gpu: elg_clone_rows through RobotBatchRollout against the oracle,
bit-exact (the clone moves bytes; the drift is three individually rounded fp32 ops), and the full-size property."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402
from extended_legged_gym_b200 import _lib  # noqa: E402
from oracle import ref_harness, rollout_oracle as ro  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rollout_clone.npz")
DEV = "cuda:0"


def load_case(tag):
    z = np.load(GOLDEN)
    m, r, seed = (int(x) for x in z[f"{tag}__meta"])
    drift = float(z[f"{tag}__drift"][0])
    get = lambda st: {k: torch.from_numpy(z[f"{tag}__{st}__{k}"]) for k in ro.STATE_KEYS}
    idx = {k: torch.from_numpy(z[f"{tag}__idx__{k}"]) for k in ("main_env_indices", "rollout_env_indices", "rollout_to_main_map")}
    return m, r, seed, drift, torch.from_numpy(z[f"{tag}__drift_u"]), get("in"), {s: get(s) for s in ("sync", "perturbed", "restore")}, idx


def oracle_from(inputs, m, r):
    o = ro.make_rollout_state(m, r, seed=0)
    for k, v in inputs.items():
        getattr(o, k).copy_(v)
    ro.init_env_indices(o)
    return o


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_matches_reference_fixture(tag):
    m, r, seed, drift, u, inputs, stages, idx = load_case(tag)
    o = oracle_from(inputs, m, r)
    for k, v in idx.items():
        assert torch.equal(getattr(o, k), v), k
    ro.sync_main_to_rollout(o, drift, u)
    for k, v in stages["sync"].items():
        assert torch.equal(getattr(o, k), v), f"sync {k}"
    ro.cache_main_env_states(o)
    for k, v in stages["perturbed"].items():
        getattr(o, k).copy_(v)
    ro.restore_main_env_states(o)
    for k, v in stages["restore"].items():
        assert torch.equal(getattr(o, k), v), f"restore {k}"


@pytest.mark.reference
@pytest.mark.skipif(not ref_harness.available(), reason="needs /root/reference")
def test_oracle_matches_live_reference():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_rollout_golden as mg
    inputs, u, stages, idx = mg.reference_run(5, 9, 3, 0.02)
    o = oracle_from(inputs, 5, 9)
    ro.sync_main_to_rollout(o, 0.02, u)
    for k, v in stages["sync"].items():
        assert torch.equal(getattr(o, k), v), k


def test_clone_abi_argument_checks():
    import ctypes as C
    from extended_legged_gym_b200 import _lib
    lib = _lib.load()
    assert lib.elg_sizeof_clone_table() == C.sizeof(_lib.ElgCloneTable)
    assert lib.elg_clone_rows(None, 0, 0.0, None, 0, 0, None) == -4
    tb = _lib.ElgCloneTable()
    tb.num_fields = _lib.MAX_CLONE_FIELDS + 1
    assert lib.elg_clone_rows(C.byref(tb), 0, 0.0, None, 0, 0, None) == -1
    tb.num_fields, tb.num_main, tb.rollouts_per_main, tb.drift_field = 1, 2, 2, -1
    assert lib.elg_clone_rows(C.byref(tb), 0, 0.0, None, 0, 0, None) == -4          # NULL base
    assert lib.elg_clone_rows(C.byref(tb), 7, 0.0, None, 0, 0, None) == -1          # bad mode
    assert b"mode" in lib.elg_last_error()


# ---------------------------------------------------------------------------------------------- GPU
def make_rollout_env(m, r, case="anymal_c_rough", seed=0, drift=0.0):
    from extended_legged_gym_b200 import synthetic
    from extended_legged_gym_b200.envs import RobotBatchRollout
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    n = m * (1 + r)
    cfg, spec, st = common.make_case_state(case, n, seed=seed)
    cfg.env.num_envs, cfg.env.rollout_envs = m, r
    cfg.domain_rand.rollout_envs_sync_pos_drift = drift
    hf = synthetic.make_height_field(seed=0)
    sim = SyntheticSim(cfg, n, DEV, spec=spec, height_samples=hf, state=st)
    env = RobotBatchRollout(cfg, None, sim, DEV, True)
    env.set_env_state(st)
    return env, st


def load_into(env, tensors):
    for k, v in tensors.items():
        getattr(env, k).copy_(v.view(getattr(env, k).shape).to(DEV))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_clone_kernel_matches_reference_fixture(tag):
    m, r, seed, drift, u, inputs, stages, idx = load_case(tag)
    env, _ = make_rollout_env(m, r, drift=drift)
    for k, v in idx.items():
        assert torch.equal(getattr(env, k).cpu(), v), k
    load_into(env, inputs)
    env.drift_u = u.to(DEV)
    env._sync_main_to_rollout()
    torch.cuda.synchronize()
    for k, v in stages["sync"].items():
        assert torch.equal(getattr(env, k).cpu().view(v.shape), v), f"sync {k}"
    env._cache_main_env_states()
    load_into(env, stages["perturbed"])
    env._restore_main_env_states()
    torch.cuda.synchronize()
    for k, v in stages["restore"].items():
        assert torch.equal(getattr(env, k).cpu().view(v.shape), v), f"restore {k}"
    assert torch.equal(env.main_env_cache["dof_pos"].cpu(), stages["sync"]["dof_state"].view(-1, 12, 2)[idx["main_env_indices"], :, 0])


@pytest.mark.gpu
@pytest.mark.parametrize("m,r,drift", [(1, 1, 0.0), (5, 3, 0.1), (7, 33, 0.0), (64, 32, 0.03), (3, 0, 0.0)])
def test_clone_kernel_matches_oracle(m, r, drift):
    env, _ = make_rollout_env(m, r, drift=drift, seed=m)
    o = ro.make_rollout_state(m, r, seed=10 + m)
    ro.init_env_indices(o)
    load_into(env, ro.snapshot(o))
    u = torch.rand(m * r, 3, generator=torch.Generator().manual_seed(5))
    env.drift_u = u.to(DEV)
    ro.sync_main_to_rollout(o, drift, u)
    env._sync_main_to_rollout()
    torch.cuda.synchronize()
    for k, v in ro.snapshot(o).items():
        assert torch.equal(getattr(env, k).cpu().view(v.shape), v), k


@pytest.mark.gpu
@pytest.mark.parametrize("m,r,drift", [(9, 37, 0.05), (3, 130, 0.0), (2, 4, 0.2), (150, 8, 0.1)])
def test_clone_tma_path_equals_per_thread_stores(m, r, drift):
    """ELG_CLONE_SYNC writes through TMA bulk stores from replicated shared-memory tiles by default; the per-thread
    16-byte-store kernel (elg_set_clone_tuning(1)) must produce the same bytes, ragged heads / tails and drift included."""
    lib = _lib.load()
    outs = []
    for no_bulk in (0, 1):
        env, _ = make_rollout_env(m, r, drift=drift, seed=m)
        o = ro.make_rollout_state(m, r, seed=20 + m)
        ro.init_env_indices(o)
        load_into(env, ro.snapshot(o))
        env.drift_u = torch.rand(m * r, 3, generator=torch.Generator().manual_seed(6)).to(DEV)
        lib.elg_set_clone_tuning(no_bulk)
        try:
            env._sync_main_to_rollout()
            torch.cuda.synchronize()
        finally:
            lib.elg_set_clone_tuning(0)
        outs.append({k: getattr(env, k).cpu().clone() for k in ro.snapshot(o)})
        if no_bulk == 0:
            ro.sync_main_to_rollout(o, drift, env.drift_u.cpu())
            for k, v in ro.snapshot(o).items():
                assert torch.equal(outs[0][k].view(v.shape), v), k
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k


@pytest.mark.gpu
def test_clone_full_size_property_and_philox_drift():
    """BASELINE config 5: 64 mains x 512 rollouts.  Without drift every rollout row equals its main row, bit for bit;
    with in-kernel Philox drift only base_pos moves, by less than drift / 2, mean ~ 0."""
    m, r = 64, 512
    env, _ = make_rollout_env(m, r)
    env._sync_main_to_rollout()
    torch.cuda.synchronize()
    for name in ("root_states", "actions", "last_actions", "last_dof_vel", "last_root_vel", "base_lin_vel", "base_ang_vel",
                 "projected_gravity", "feet_air_time", "feet_contact_time", "last_contacts"):
        t = getattr(env, name).view(m, 1 + r, -1)
        assert torch.equal(t[:, 1:], t[:, :1].expand(-1, r, -1)), name
    d = env.dof_state.view(m, 1 + r, -1)
    assert torch.equal(d[:, 1:], d[:, :1].expand(-1, r, -1))
    env.cfg.domain_rand.rollout_envs_sync_pos_drift = 0.2
    env.drift_u = None
    env._sync_main_to_rollout()
    torch.cuda.synchronize()
    rs = env.root_states.view(m, 1 + r, 13)
    delta = rs[:, 1:, :3] - rs[:, :1, :3]
    assert torch.equal(rs[:, 1:, 3:], rs[:, :1, 3:].expand(-1, r, -1))
    assert float(delta.abs().max()) <= 0.1 + 1e-6 and abs(float(delta.mean())) < 2e-3 and float(delta.std()) > 0.05


@pytest.mark.gpu
def test_step_and_step_rollout_contract():
    """step(): main rows out, rollouts re-synced; step_rollout(): rollout rows out, main rows restored bit-exactly."""
    m, r = 6, 4
    env, st = make_rollout_env(m, r, seed=2)
    env.cfg.domain_rand.push_robots = False
    obs, priv, rew, reset, extras = env.step(torch.zeros(m, env.num_actions, device=DEV))
    assert obs.shape == (m, env.num_obs) and rew.shape == (m,) and reset.shape == (m,) and reset.dtype == torch.bool
    before = {k: getattr(env, k)[env.main_env_indices].clone() for k in ("root_states", "last_actions", "base_lin_vel", "feet_air_time",
                                                                          "base_lin_acc", "last_contacts")}
    a = torch.randn(m * r, env.num_actions, device=DEV)
    robs, _, rrew, rreset, _ = env.step_rollout(a)
    assert robs.shape == (m * r, env.num_obs) and rrew.shape == (m * r,)
    assert torch.equal(env.actions[env.rollout_env_indices], a.clamp(-100, 100))
    for k, v in before.items():
        assert torch.equal(getattr(env, k)[env.main_env_indices], v), k
    # rollout rows: last_actions follow the applied actions (robot_batch_rollout.py:814)
    assert torch.equal(env.last_actions[env.rollout_env_indices], env.actions[env.rollout_env_indices])
    with pytest.raises(ValueError):
        env.step_rollout(torch.zeros(3, env.num_actions, device=DEV))
