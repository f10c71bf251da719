"""Zero-sized inputs through the C ABI: every entry point of the hot path returns ELG_OK and touches nothing when it is handed no
envs / rays / points / rows (the reference's torch code degrades the same way: empty tensors flow through)."""
import copy
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402
from extended_legged_gym_b200 import _lib, synthetic  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _env(n=8):
    from extended_legged_gym_b200.envs import LeggedRobot
    from extended_legged_gym_b200.sim_backend import SyntheticSim
    cfg, spec, st = common.make_case_state("anymal_c_rough", n, seed=1)
    cfg.env.num_envs = n
    hf = synthetic.make_height_field(rows=60, cols=60, border=5, tile=25, seed=2)
    env = LeggedRobot(cfg, None, SyntheticSim(cfg, n, DEV, spec=spec, height_samples=hf, state=st), DEV, True)
    env.set_env_state(st)
    return env


def test_step_torques_reset_with_zero_envs_leave_everything_untouched():
    env = _env()
    lib = env._lib
    env._sync_native()
    before = {k: getattr(env, k).clone() for k in ("obs_buf", "rew_buf", "torques", "commands", "episode_length_buf", "last_actions")}
    dims0 = copy.copy(env._dims)
    dims0.num_envs = 0
    s = torch.cuda.current_stream().cuda_stream
    assert lib.elg_post_physics_step(C.byref(dims0), C.byref(env._params), C.byref(env._bufs), _lib.PHASE_FUSED, s) == 0
    assert lib.elg_compute_torques(C.byref(dims0), C.byref(env._params), env.actions.data_ptr(), env.dof_state.data_ptr(), env.last_dof_vel.data_ptr(),
                                   env.p_gains.data_ptr(), env.d_gains.data_ptr(), env.torque_limits.data_ptr(), env.default_dof_pos.data_ptr(),
                                   env.torques.data_ptr(), None, 0, s) == 0
    rp, rb = env._reset_native_synced()
    assert lib.elg_resample_commands(C.byref(dims0), C.byref(rp), env.episode_length_buf.data_ptr(), env.commands.data_ptr(), None, None, s) == 0
    assert lib.elg_reset_envs(C.byref(dims0), C.byref(rp), C.byref(env._params), C.byref(rb), s) == 0
    # an env-id list of length zero (a non-NULL pointer with num_ids == 0; NULL means "all envs" in this ABI, and torch hands out NULL
    # for an empty tensor, which is why the host classes return before the call when the index tensor is empty)
    ids1 = torch.zeros(1, dtype=torch.int64, device=DEV)
    ids = ids1[:0]
    assert lib.elg_compute_torques(C.byref(env._dims), C.byref(env._params), env.actions.data_ptr(), env.dof_state.data_ptr(), env.last_dof_vel.data_ptr(),
                                   env.p_gains.data_ptr(), env.d_gains.data_ptr(), env.torque_limits.data_ptr(), env.default_dof_pos.data_ptr(),
                                   env.torques.data_ptr(), ids1.data_ptr(), 0, s) == 0
    torch.cuda.synchronize()
    for k, v in before.items():
        assert torch.equal(getattr(env, k), v), f"{k} changed although no env was processed"
    env.reset_idx(ids)                                        # the host path on an empty index tensor (legged_robot.py:169-170)
    torch.cuda.synchronize()
    assert torch.equal(env.commands, before["commands"])


def test_mesh_queries_with_zero_rays_points_cameras():
    from extended_legged_gym_b200.utils.ray_caster import Mesh
    from oracle import mesh_oracle as mo
    v, t = mo.box_mesh()
    mesh = Mesh(v, t, DEV)
    lib = _lib.load()
    s = torch.cuda.current_stream().cuda_stream
    z3 = torch.zeros(0, 3, device=DEV)
    assert lib.elg_raycast(mesh.id, z3.data_ptr(), z3.data_ptr(), 0, 5.0, z3.data_ptr(), None, None, None, s) == 0
    zs = torch.zeros(0, device=DEV)
    assert lib.elg_sdf_query(mesh.id, z3.data_ptr(), 0, 1.0, 1.0e-3, zs.data_ptr(), z3.data_ptr(), None, None, s) == 0
    torch.cuda.synchronize()


def test_rollout_helpers_with_zero_rows():
    lib = _lib.load()
    s = torch.cuda.current_stream().cuda_stream
    a = torch.zeros(4, 12, device=DEV)
    assert lib.elg_rollout_actions(a.data_ptr(), 0, 4, 12, 0, 100.0, None, None, a.data_ptr(), s) == 0
    assert lib.elg_rollout_actions(a.data_ptr(), 4, 0, 12, 0, 100.0, None, None, a.data_ptr(), s) == 0
    tb = _lib.ElgCloneTable()
    tb.num_fields, tb.num_main, tb.rollouts_per_main, tb.drift_field = 0, 0, 0, -1
    assert lib.elg_clone_rows(C.byref(tb), 0, 0.0, None, 0, 0, s) == 0
    x = torch.zeros(0, 7, device=DEV)
    m = torch.zeros(7, device=DEV)
    cnt = torch.zeros(1, dtype=torch.int64, device=DEV)
    assert lib.elg_normalize_observations(0, 7, x.data_ptr(), m.data_ptr(), m.data_ptr(), m.data_ptr(), cnt.data_ptr(), 0.01, -1, 0, x.data_ptr(), None,
                                          None, None, None, None, s) == 0
    torch.cuda.synchronize()
    assert int(cnt) == 0
