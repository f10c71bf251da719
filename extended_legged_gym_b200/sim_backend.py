"""Stand-in for the simulator boundary.

In the reference the state tensors come from Isaac Gym (``gym.acquire_*_tensor`` wrapped by
``gymtorch.wrap_tensor``, envs/base/legged_robot.py:564-584) and PhysX advances them in
``gym.simulate``.  PhysX is outside this build's scope (BASELINE.json north_star), so the host
classes talk to a small backend object instead; ``SyntheticSim`` serves seeded synthetic state of
the right shapes/layouts, which is all the per-step path needs.  A maintainer integrating with
the real simulator passes a backend whose tensors are the ``gymtorch``-wrapped PhysX tensors
(INTEGRATION.md shows the adapter).
"""
from typing import Dict, Optional

import torch

from . import synthetic
from .envs.robot_specs import RobotSpec, get_robot_spec


class SimBackend:
    """Duck-typed protocol; every tensor lives on ``device`` and is updated in place by the simulator."""
    device: str
    num_envs: int
    spec: RobotSpec
    root_states: torch.Tensor        # [N,13]
    dof_state: torch.Tensor          # [N*D,2]
    contact_forces: torch.Tensor     # [N*B,3]
    rigid_body_state: torch.Tensor   # [N*B,13]
    height_samples: Optional[torch.Tensor] = None   # int16 [rows, cols]
    terrain_origins: Optional[torch.Tensor] = None  # [rows, cols, 3]
    dt: float = 0.005
    use_gpu_pipeline: bool = True
    capturable: bool = False         # True: simulate / refresh / set_* enqueue nothing that breaks a CUDA-graph capture

    def simulate(self) -> None: ...
    def refresh(self) -> None: ...
    def set_dof_actuation_force(self, torques: torch.Tensor) -> None: ...
    def set_dof_state_indexed(self, env_ids: torch.Tensor) -> None: ...
    def set_root_state_indexed(self, env_ids: torch.Tensor) -> None: ...
    def set_root_state(self) -> None: ...
    def set_dof_state(self) -> None: ...


class SyntheticSim(SimBackend):
    capturable = True
    def __init__(self, cfg, num_envs: Optional[int] = None, device: str = "cuda:0", seed: int = 0,
                 spec: Optional[RobotSpec] = None, height_samples: Optional[torch.Tensor] = None,
                 state: Optional[Dict[str, torch.Tensor]] = None, packed: bool = False):
        self.cfg = cfg
        self.device = device
        self.spec = spec or get_robot_spec(cfg.asset.name)
        self.num_envs = int(num_envs if num_envs is not None else cfg.env.num_envs)
        self.dt = cfg.sim.dt
        sp = self.spec
        if state is None:
            q0 = [cfg.init_state.default_joint_angles[n] for n in sp.dof_names]
            state = synthetic.make_state(self.num_envs, sp.num_dof, sp.num_bodies, sp.indices_matching(cfg.asset.foot_name),
                                         sp.indices_matching(cfg.asset.penalize_contacts_on),
                                         sp.indices_matching(cfg.asset.terminate_after_contacts_on), q0, sp.foot_offsets,
                                         num_commands=cfg.commands.num_commands, seed=seed)
        self.initial_state = state      # CPU copy (histories included) for the env to seed itself from
        if packed:
            # the four simulator tensors (+ one [N, D] action block) as views of ONE device block, 256-byte aligned slices: a host
            # that keeps the simulator state elsewhere refreshes all of it with a single copy (`state_block.copy_(pinned_block)`)
            names = ("root_states", "dof_state", "contact_forces", "rigid_body_state")
            shapes = {k: tuple(state[k].shape) for k in names}
            shapes["actions"] = (self.num_envs, sp.num_dof)
            off, self.state_slices = 0, {}
            for k, shp in shapes.items():
                n = int(torch.tensor(shp).prod())
                self.state_slices[k] = (off, n, shp)
                off += (n + 63) // 64 * 64
            self.state_block = torch.zeros(off, dtype=torch.float, device=device)
            for k, (o, n, shp) in self.state_slices.items():
                v = self.state_block[o:o + n].view(shp)
                if k in state:
                    v.copy_(state[k])
                setattr(self, "actions_in" if k == "actions" else k, v)
        else:
            self.root_states = state["root_states"].to(device).contiguous()
            self.dof_state = state["dof_state"].to(device).contiguous()
            self.contact_forces = state["contact_forces"].to(device).contiguous()
            self.rigid_body_state = state["rigid_body_state"].to(device).contiguous()
        needs_hf = cfg.terrain.mesh_type in ("heightfield", "trimesh", "confined_trimesh")
        if needs_hf:
            if height_samples is None:
                rows = int(cfg.terrain.num_rows * cfg.terrain.terrain_length / cfg.terrain.horizontal_scale) + \
                    2 * int(cfg.terrain.border_size / cfg.terrain.horizontal_scale)
                cols = int(cfg.terrain.num_cols * cfg.terrain.terrain_width / cfg.terrain.horizontal_scale) + \
                    2 * int(cfg.terrain.border_size / cfg.terrain.horizontal_scale)
                height_samples = synthetic.make_height_field(rows, cols, int(cfg.terrain.border_size / cfg.terrain.horizontal_scale),
                                                             int(cfg.terrain.terrain_length / cfg.terrain.horizontal_scale), seed=seed)
            self.height_samples = height_samples.to(device).contiguous()
            nrow, ncol = cfg.terrain.num_rows, cfg.terrain.num_cols
            to = torch.zeros(nrow, ncol, 3)
            to[..., 0] = (torch.arange(nrow).float().view(-1, 1) + 0.5) * cfg.terrain.terrain_length
            to[..., 1] = (torch.arange(ncol).float().view(1, -1) + 0.5) * cfg.terrain.terrain_width
            self.terrain_origins = to.to(device)
        self.applied_torques = None

    # PhysX would integrate here; the synthetic backend leaves the state as is
    def simulate(self) -> None:
        pass

    def refresh(self) -> None:
        pass

    def set_dof_actuation_force(self, torques: torch.Tensor) -> None:
        self.applied_torques = torques

    def set_dof_state_indexed(self, env_ids) -> None:
        pass

    def set_root_state_indexed(self, env_ids) -> None:
        pass

    def set_dof_state(self) -> None:
        pass

    def set_root_state(self) -> None:
        pass
