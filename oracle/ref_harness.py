"""TEST INFRASTRUCTURE ONLY (oracle) -- container-only harness.

Imports the *unmodified* reference (``/root/reference/legged_gym``) with stub
modules standing in for the third-party binaries that are absent from this
image (isaacgym, warp, trimesh, matplotlib, cv2, traj_sampling, ...), so the
reference's own ``LeggedRobot`` / ``RobotBatchRollout`` methods can be bound to
a synthetic ``self`` and run on torch-CPU.  Used ONLY by
``tests/golden/make_golden.py`` (fixture generation) and by the not-gpu tests
that pin ``oracle/legged_oracle.py`` against the real reference.

``/root/reference`` does not exist on the GPU box, so nothing that runs there
may import this module; ``available()`` says whether it can be used.
"""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("ELG_REFERENCE_ROOT", "/root/reference")
_LG = os.path.join(REFERENCE_ROOT, "legged_gym")
_RL = os.path.join(REFERENCE_ROOT, "rsl_rl")

_STUB_ROOTS = ("isaacgym", "warp", "trimesh", "rtree", "cv2", "tensordict", "isaac_utils", "matplotlib",
               "mpl_toolkits", "traj_sampling", "pynput", "pygame")

_installed = False


def available() -> bool:
    return os.path.isdir(os.path.join(_LG, "legged_gym"))


class _Anything(mock.MagicMock):
    """MagicMock that is also usable as a decorator factory (``@wp.kernel``)."""

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k and isinstance(a[0], types.FunctionType):
            return a[0]
        return super().__call__(*a, **k)


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Resolves any module under an absent third-party root to a MagicMock module."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS and fullname != "isaacgym.torch_utils":
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Anything(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def install():
    """Install stubs + sys.path entries.  Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    from . import torch_utils as _tu

    sys.meta_path.insert(0, _StubFinder())
    importlib.import_module("isaacgym")
    # real restatement of torch_utils (star-imported by the reference)
    tu = types.ModuleType("isaacgym.torch_utils")
    for k in dir(_tu):
        if not k.startswith("_"):
            setattr(tu, k, getattr(_tu, k))
    tu.__all__ = [k for k in dir(_tu) if not k.startswith("_") and k not in ("np", "torch")]
    sys.modules["isaacgym.torch_utils"] = tu
    sys.modules["isaacgym"].torch_utils = tu
    for p in (_LG, _RL):
        if p not in sys.path:
            sys.path.insert(0, p)
    _installed = True


def legged_gym():
    install()
    import legged_gym  # noqa
    return legged_gym


def reference_classes():
    """Return the reference classes used by the golden generator."""
    install()
    from legged_gym.envs.base.legged_robot import LeggedRobot
    from legged_gym.envs.base.legged_robot_config import LeggedRobotCfg
    from legged_gym.envs.anymal_c.mixed_terrains.anymal_c_rough_config import AnymalCRoughCfg
    from legged_gym.envs.anymal_c.flat.anymal_c_flat_config import AnymalCFlatCfg
    from legged_gym.envs.a1.a1_config import A1RoughCfg
    from legged_gym.envs.go2.flat.go2_rough_config import Go2RoughCfg
    from legged_gym.envs.elspider_air.mixed_terrains.elspider_air_rough_config import ElSpiderAirRoughCfg
    from legged_gym.envs.elspider_air.elspider import ElSpider
    return dict(LeggedRobot=LeggedRobot, LeggedRobotCfg=LeggedRobotCfg, AnymalCRoughCfg=AnymalCRoughCfg,
                AnymalCFlatCfg=AnymalCFlatCfg, A1RoughCfg=A1RoughCfg, Go2RoughCfg=Go2RoughCfg,
                ElSpiderAirRoughCfg=ElSpiderAirRoughCfg, ElSpider=ElSpider)


# ---------------------------------------------------------------------------------------------
# Reference env on synthetic state: the reference's OWN LeggedRobot methods on a synthetic self
# ---------------------------------------------------------------------------------------------
def synthetic_terrain_origins(cfg):
    """[rows, cols, 3] tile centres, the data contract of Terrain.env_origins (utils/terrain.py)."""
    import torch
    nrow, ncol = cfg.terrain.num_rows, cfg.terrain.num_cols
    to = torch.zeros(nrow, ncol, 3)
    to[..., 0] = (torch.arange(nrow).float().view(-1, 1) + 0.5) * cfg.terrain.terrain_length
    to[..., 1] = (torch.arange(ncol).float().view(1, -1) + 0.5) * cfg.terrain.terrain_width
    return to


def make_reference_env(cfg, spec, state, height_samples=None, cls=None, terrain_origins=None):
    """Build an instance of the reference ``LeggedRobot`` (or subclass ``cls``) WITHOUT PhysX.

    ``__init__`` is skipped (it creates the simulator); instead the attributes ``_create_envs`` /
    ``BaseTask.__init__`` would set are filled from ``spec`` / ``state`` and the reference's own
    ``_parse_cfg``, ``_init_buffers`` and ``_prepare_reward_function`` run unmodified with a mocked
    ``gym`` whose ``acquire_*_tensor`` calls return the synthetic tensors.
    """
    import types as _types
    import numpy as _np
    import torch
    install()
    from legged_gym.envs.base.legged_robot import LeggedRobot
    import legged_gym.envs.base.legged_robot as _lrmod
    _lrmod.gymtorch.wrap_tensor = lambda t: t
    _lrmod.gymtorch.unwrap_tensor = lambda t: t

    cls = cls or LeggedRobot
    env = object.__new__(cls)
    N = state["root_states"].shape[0]
    cfg.env.num_envs = N
    env.cfg = cfg
    env.sim_params = _types.SimpleNamespace(dt=cfg.sim.dt, use_gpu_pipeline=False)
    env.height_samples = None
    env.debug_viz = False
    env.init_done = False
    env._parse_cfg(cfg)
    # --- BaseTask.__init__ (base_task.py:41-103) without the simulator
    env.gym = mock.MagicMock(name="gym")
    env.sim = None
    env.physics_engine = None
    env.sim_device = "cpu"
    env.sim_device_id = 0
    env.headless = True
    env.device = "cpu"
    env.graphics_device_id = -1
    env.num_envs = N
    env.num_obs = cfg.env.num_observations
    env.num_privileged_obs = cfg.env.num_privileged_obs
    env.num_actions = cfg.env.num_actions
    env.obs_buf = torch.zeros(N, env.num_obs)
    env.rew_buf = torch.zeros(N)
    env.reset_buf = torch.ones(N, dtype=torch.long)
    env.episode_length_buf = torch.zeros(N, dtype=torch.long)
    env.time_out_buf = torch.zeros(N, dtype=torch.bool)
    env.privileged_obs_buf = None
    env.extras = {}
    env.enable_viewer_sync = True
    env.viewer = None
    # --- create_sim / _create_envs facts
    env.up_axis_idx = 2
    env.num_dof = env.num_dofs = spec.num_dof
    env.num_bodies = spec.num_bodies
    env.dof_names = list(spec.dof_names)
    long = lambda idx: torch.tensor(idx, dtype=torch.long)
    env.feet_indices = long(spec.indices_matching(cfg.asset.foot_name))
    env.penalised_contact_indices = long(spec.indices_matching(cfg.asset.penalize_contacts_on))
    env.termination_contact_indices = long(spec.indices_matching(cfg.asset.terminate_after_contacts_on))
    # Isaac Gym hands a structured array with one record per DOF
    props = _np.zeros(spec.num_dof, dtype=[("lower", "f4"), ("upper", "f4"), ("velocity", "f4"), ("effort", "f4")])
    props["lower"], props["upper"] = spec.dof_lower, spec.dof_upper
    props["velocity"], props["effort"] = spec.dof_velocity, spec.dof_effort
    env._process_dof_props(props, 0)
    base_init = cfg.init_state.pos + cfg.init_state.rot + cfg.init_state.lin_vel + cfg.init_state.ang_vel
    env.base_init_state = torch.tensor(base_init, dtype=torch.float)
    mesh_type = cfg.terrain.mesh_type
    if mesh_type in ("heightfield", "trimesh", "confined_trimesh"):
        assert height_samples is not None
        env.height_samples = height_samples
        nrow, ncol = cfg.terrain.num_rows, cfg.terrain.num_cols
        if terrain_origins is None:
            terrain_origins = synthetic_terrain_origins(cfg)
        env.terrain = _types.SimpleNamespace(cfg=cfg.terrain, env_length=cfg.terrain.terrain_length,
                                             env_width=cfg.terrain.terrain_width, env_origins=terrain_origins.numpy())
    # the reference's own _get_env_origins (legged_robot.py:817-844); it draws terrain levels from the global RNG
    with torch.random.fork_rng():
        torch.manual_seed(1234)
        env._get_env_origins()
    # --- the reference's own _init_buffers with acquire_* returning the synthetic tensors
    env.gym.acquire_actor_root_state_tensor.return_value = state["root_states"]
    env.gym.acquire_dof_state_tensor.return_value = state["dof_state"]
    env.gym.acquire_net_contact_force_tensor.return_value = state["contact_forces"]
    env.gym.acquire_rigid_body_state_tensor.return_value = state["rigid_body_state"]
    from legged_gym.envs.base.legged_robot_rew_mixin import LeggedRobotRewMixin
    env.speed_min = 0.1          # LeggedRobotRewMixin.__init__ (legged_robot_rew_mixin.py:13)
    env._init_buffers()
    env._prepare_reward_function()
    env.init_done = True
    env.acc_ema = 0.9
    # --- histories / env-owned state from the synthetic set
    for k in ("actions", "last_actions", "last_dof_vel", "last_root_vel", "commands", "feet_air_time",
              "feet_contact_time", "last_contacts", "base_lin_acc", "base_ang_acc"):
        getattr(env, k)[:] = state[k]
    env.episode_length_buf[:] = state["episode_length_buf"]
    env.reset_buf = torch.zeros(N, dtype=torch.bool)
    return env
